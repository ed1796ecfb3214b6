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
