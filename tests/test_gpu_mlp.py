"""GPU parity of the fused positional-encoding + MLP kernels (tensor-core fp16/fp32-acc and CUDA-core
fp32) against the reference goldens and the oracle.  Tolerances: fp32 SIMT 2e-5 on raw outputs;
tensor-core path: rendered RGB/depth <= 1e-4 abs (north_star), raw <= 1e-3."""
import numpy as np
import pytest
import torch

import neuman_b200 as nb
from neuman_b200 import _lib, ops
from oracle import neuman_oracle as no
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda"
MODES = {"simt": _lib.NM_MLP_SIMT_F32, "tc": _lib.NM_MLP_TC_F16}
RAW_TOL = {"simt": 2e-5, "tc": 1e-3}


@pytest.fixture(scope="module")
def nets():
    return tuple(n.to(DEV) for n in util.product_nets())


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_net_forward_golden(nets, mode):
    g = util.golden("stages.npz")
    pts, views = torch.from_numpy(g["n_pts"]).to(DEV), torch.from_numpy(g["n_views"]).to(DEV)
    for net, key in zip(nets, ("n_coarse", "n_fine", "n_human")):
        y = ops.joiner_forward(net, pts, views, mode=MODES[mode]).cpu().numpy()
        err = np.abs(y - g[key]).max()
        assert err < RAW_TOL[mode], (key, err)


@pytest.mark.parametrize("mode", ["simt", "tc"])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 255, 257, 1000, 70000])
def test_ragged_sizes(nets, mode, n):
    torch.manual_seed(n)
    pts, views = torch.randn(n, 3), torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    net = nets[2] if n % 2 else nets[0]
    with torch.no_grad():
        ref = no.net_forward(util.oracle_params(net.cpu() if False else net.to("cpu")), pts, views)
    net.to(DEV)
    y = ops.joiner_forward(net, pts.to(DEV), views.to(DEV), mode=MODES[mode]).cpu()
    assert y.shape == (n, 4)
    assert (y - ref).abs().max() < RAW_TOL[mode]


@pytest.mark.parametrize("mode", ["simt", "tc"])
def test_rays_mode_equals_pts_mode(nets, mode):
    torch.manual_seed(5)
    R, S = 300, 96
    o, d = torch.randn(R, 3).to(DEV), torch.nn.functional.normalize(torch.randn(R, 3), dim=-1).to(DEV)
    pts, dirs, z = nb.ray_to_samples({"origin": o, "direction": d, "near": torch.zeros(R, 1, device=DEV),
                                      "far": torch.full((R, 1), 3.0, device=DEV)}, S)
    a = ops.mlp_forward_rays(nets[0], o, d, z, mode=MODES[mode])
    b = ops.joiner_forward(nets[0], pts, dirs, mode=MODES[mode])
    assert torch.equal(a, b)                                       # same arithmetic, bit-identical


def test_tc_render_tolerance(nets):
    """End-to-end effect of fp16 operands: composite of TC raw vs fp32 oracle raw on the same samples
    stays within the 1e-4 abs gate on rgb and depth (and the SIMT path within 1e-5)."""
    torch.manual_seed(11)
    R, S = 512, 128
    o, d = torch.randn(R, 3) * 0.3, torch.nn.functional.normalize(torch.randn(R, 3), dim=-1)
    pts, dirs, z = no.ray_to_samples(o, d, torch.zeros(R, 1), torch.full((R, 1), 3.14), S)
    net = nets[0]
    with torch.no_grad():
        net.to("cpu")
        raw_ref = no.net_forward(util.oracle_params(net), pts, dirs)
        net.to(DEV)
    rgb_ref, _, _, _, dep_ref = no.raw2outputs(raw_ref, z, d)
    for mode, tol in (("simt", 1e-5), ("tc", 1e-4)):
        raw = ops.mlp_forward_rays(net, o.to(DEV), d.to(DEV), z.to(DEV), mode=MODES[mode]).cpu()
        rgb, _, _, _, dep = no.raw2outputs(raw, z, d)
        assert (rgb - rgb_ref).abs().max() < tol and (dep - dep_ref).abs().max() < tol, mode


def test_repack_on_weight_update(nets):
    net = nets[1]
    pts, views = torch.randn(64, 3, device=DEV), torch.randn(64, 3, device=DEV)
    a = net(pts, views)
    with torch.no_grad():
        net.nerf.rgb_linear.bias.add_(1.0)
    b = net(pts, views)
    assert torch.allclose(b[:, :3], a[:, :3] + 1.0, atol=1e-5) and torch.equal(a[:, 3], b[:, 3])
    with torch.no_grad():
        net.nerf.rgb_linear.bias.sub_(1.0)


def test_c_abi_views_per_ray(nets):
    """nm_mlp_forward with views_per_ray = S (one direction row per ray, include/neuman_b200.h) equals the
    per-sample-views call; called straight through ctypes."""
    import ctypes as C
    from neuman_b200._lib import Context
    torch.manual_seed(2)
    R, S = 40, 24
    pts = torch.randn(R * S, 3, device=DEV)
    dirs = torch.nn.functional.normalize(torch.randn(R, 3, device=DEV), dim=-1)
    ctx = Context.get(0)
    slot = ops.net_slot(nets[0], ctx)
    for mode in MODES.values():
        a = torch.empty(R * S, 4, device=DEV)
        ctx.check(ctx.lib.nm_mlp_forward(ctx.h, slot, mode, pts.data_ptr(), dirs.data_ptr(), R * S, S, a.data_ptr(),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        b = ops.joiner_forward(nets[0], pts, dirs[:, None, :].expand(-1, S, -1).reshape(-1, 3), mode=mode)
        assert torch.equal(a, b)
        # error paths: n not a multiple of views_per_ray, unpacked slot
        rc = ctx.lib.nm_mlp_forward(ctx.h, slot, mode, pts.data_ptr(), dirs.data_ptr(), R * S - 1, S, a.data_ptr(), None)
        assert rc == -1 and b"multiple" in ctx.lib.nm_last_error(ctx.h)
    rc = ctx.lib.nm_mlp_forward(ctx.h, 15, 0, pts.data_ptr(), dirs.data_ptr(), 8, 0, a.data_ptr(), None)
    assert rc == -4


def test_large_weights_stay_finite(nets):
    """fp16 operands: 3x larger weights (activations grow ~3^8) must not overflow to inf/nan and the error
    must grow no faster than the oracle's own fp32 noise floor allows (documented in DESIGN.md §3)."""
    import copy
    net = copy.deepcopy(nets[0]).to("cpu")
    with torch.no_grad():
        for p in net.nerf.pts_linears.parameters():
            p.mul_(1.6)
    torch.manual_seed(4)
    pts, views = torch.randn(4096, 3), torch.nn.functional.normalize(torch.randn(4096, 3), dim=-1)
    with torch.no_grad():
        ref = no.net_forward(util.oracle_params(net), pts, views)
    net.to(DEV)
    y = ops.joiner_forward(net, pts.to(DEV), views.to(DEV), mode=MODES["tc"]).cpu()
    assert torch.isfinite(y).all()
    rel = (y - ref).abs().max() / ref.abs().max()
    assert rel < 2e-3, rel


def _emul_floor(net_cpu, pts, views):
    """|fp32 oracle - oracle with 11-bit matmul operands| on these inputs: what the operand precision alone does."""
    p = util.oracle_params(net_cpu)
    with torch.no_grad():
        ref = no.net_forward(p, pts, views)
        with no.precision(operands="f16"):
            emu = no.net_forward(p, pts, views)
    return ref, float((emu - ref).abs().max())


@pytest.mark.parametrize("kind", ["posenc", "rotate"])
@pytest.mark.parametrize("xmax", [4.0, 30.0, 100.0])
def test_large_coordinates(kind, xmax):
    """|x| up to 100 for both encodings: the tensor-core path stays finite, its error stays within 3x what 11-bit operands do
    to the oracle on the same inputs (the raw coordinates are input channels: their fp16 rounding grows with |x|), the fp32
    mode within 2e-5 relative, and no range flag is raised (100 << 65504)."""
    import copy
    torch.manual_seed(int(xmax))
    net, _ = util.scenes.seed_nets(nb.build_nerf, nb.default_opt(use_cuda=False, posenc=kind), 7)
    n = 2048
    pts = (torch.rand(n, 3) * 2 - 1) * xmax
    views = torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    ref, floor = _emul_floor(copy.deepcopy(net), pts, views)
    net = net.to(DEV)
    ctx = _lib.Context.get(0)
    ctx.range_check()                                            # clear
    y = ops.joiner_forward(net, pts.to(DEV), views.to(DEV), mode=MODES["tc"]).cpu()
    assert torch.isfinite(y).all()
    assert (y - ref).abs().max() <= 3 * floor + 1e-4, (float((y - ref).abs().max()), floor)
    ctx.range_check()                                            # raises if the flag were set
    y32 = ops.joiner_forward(net, pts.to(DEV), views.to(DEV), mode=MODES["simt"]).cpu()
    assert (y32 - ref).abs().max() <= 2e-5 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("scale", [3.0, 10.0])
def test_scaled_weights_and_the_range_flag(scale):
    """Weights x3 (hidden activations grow to ~1e3: still in range) and x10 (1e8: beyond fp16).  In range, the error is that
    of the operand precision (relative to the output scale); out of range the packs saturate -- finite outputs -- and
    nm_range_status reports NM_ERR_RANGE, once (the flag is cleared by the call), while the fp32 mode is unaffected."""
    import copy
    torch.manual_seed(3)
    net, _ = util.scenes.seed_nets(nb.build_nerf, nb.default_opt(use_cuda=False), 11)
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("weight"):
                p.mul_(scale)
    n = 3000
    pts, views = torch.randn(n, 3), torch.nn.functional.normalize(torch.randn(n, 3), dim=-1)
    ref, floor = _emul_floor(copy.deepcopy(net), pts, views)
    net = net.to(DEV)
    ctx = _lib.Context.get(0)
    ctx.range_check()
    y = ops.joiner_forward(net, pts.to(DEV), views.to(DEV), mode=MODES["tc"]).cpu()
    assert torch.isfinite(y).all()
    hidden_max = float(ref.abs().max())
    if scale <= 3.0:
        assert (y - ref).abs().max() <= 3 * floor + 1e-4 * max(1.0, hidden_max), (float((y - ref).abs().max()), floor)
        ctx.range_check()
    else:
        with pytest.raises(_lib.NmError, match="range"):
            ctx.range_check()
        ctx.range_check()                                        # cleared by the failing call
        y32 = ops.joiner_forward(net, pts.to(DEV), views.to(DEV), mode=MODES["simt"]).cpu()
        assert torch.isfinite(y32).all() and (y32 - ref).abs().max() <= 1e-4 * max(1.0, hidden_max)
