"""GPU parity at the sizes BASELINE.json states for configurations 2-5: the full frame is rendered at the
stated resolution and sample counts and a 64x64 block of it (4096 rays straddling a body silhouette) is compared with the
output of the UNMODIFIED reference on exactly those rays (tests/golden/fullsize.npz, tools/make_golden_fullsize.py).

Gates.  north_star: rgb / depth within 1e-4 abs of the reference path.  Two measured noise floors bound what that can
mean on a given configuration, both stored next to the golden (per ray, same rays):
  floor64 = |reference fp32 algorithm - the same algorithm in float64|: the reference's own rounding noise.  The hybrid
            renderers sort background and human samples by depth, so an ulp moves a sample across another one and changes
            the pixel by O(1e-4..1e-3); the fp32 reference is only defined up to that.
  floor16 = |fp32 algorithm - fp32 algorithm with the nets' matmul operands rounded to 11 significand bits|: what any
            tensor-core evaluation (tcgen05 kind::f16 or kind::tf32) does to the result, independent of the kernel.
The fp32 CUDA-core mode (NM_MLP_SIMT_F32) is held to max(1e-4, K64 * floor64); the tensor-core mode (the default and the
benchmarked one) to max(1e-4, K16 * max(floor16, floor64)), with K = 2 on the 99.5th percentile and K = 8 on the maximum
(the floors are ONE realisation of the rounding noise, not a bound: the tails are sample flips across a discontinuity --
depth-sorted merges, nearest-triangle changes of the posed warp -- whose maxima over 4096 rays vary by several x between
realisations; the percentile gate is the tight one).  Rays the oracle proves ill-conditioned (an actor's
|far - near| < 1e-3: hit/miss flips under 1-ulp changes and delta_last = 1e10 turns that into O(1)) are excluded.
"""
import numpy as np
import pytest
import torch

import neuman_b200 as nb
from neuman_b200 import _lib, render, synthetic
from oracle import scenes, synth_smpl
from tests import util

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4


@pytest.fixture(scope="module")
def gold():
    return util.golden("fullsize.npz")


@pytest.fixture(scope="module")
def nets():
    return tuple(n.to(DEV) for n in util.product_nets())


@pytest.fixture(scope="module")
def human():
    return util.product_human_model(DEV)


def block(frame, g, name, C):
    x0, y0, w, h = (int(v) for v in g[f"{name}_window"])
    cfg = synthetic.FULLSIZE[name]
    return frame.reshape(cfg["H"], cfg["W"], C)[y0:y0 + h, x0:x0 + w].reshape(h * w, C)


def gate(err, floor_map, graz, what, k_p=2.0, k_max=8.0):
    """err, floor_map: per-ray [4096]; returns the report and asserts the percentile / maximum gates."""
    ok = ~graz.reshape(-1)
    e, f = err[ok], floor_map.reshape(-1)[ok]
    p995, fmax = float(np.percentile(f, 99.5)), float(f.max())
    rep = dict(what=what, err_max=float(e.max()), err_p995=float(np.percentile(e, 99.5)), floor_max=fmax, floor_p995=p995,
               frac_gt_tol=float((e > TOL).mean()))
    assert rep["err_p995"] <= max(TOL, k_p * p995), rep
    assert rep["err_max"] <= max(TOL, k_max * fmax), rep
    return rep


def floors(g, name, plane, mode, variant=""):
    f64 = g[f"{name}_{variant}floor64_{plane}_map"]
    if mode == "simt":
        return f64
    return np.maximum(f64, g[f"{name}_{variant}floor16_{plane}_map"])


def bodies_of(name):
    return [synth_smpl.random_body(seed=a["seed"], scale=a["scale"], center=a["center"]) for a in synthetic.FULLSIZE[name]["actors"]]


@pytest.fixture(params=["tc", "simt"])
def mode(request, monkeypatch):
    monkeypatch.setenv("NEUMAN_MLP_MODE", request.param)
    return request.param


def test_checksums(gold, nets, human):
    """The nets here are the ones the reference rendered the goldens with."""
    s = gold["net_sums"]
    got = [scenes.net_checksum(nets[0]), scenes.net_checksum(nets[1]), scenes.net_checksum(human.coarse_bkg_net),
           scenes.net_checksum(human.fine_bkg_net), scenes.net_checksum(human.coarse_human_net)]
    assert np.allclose(got, s, rtol=1e-6)


def test_cfg2_vanilla_1280x720_64_128(gold, nets, mode):
    c = synthetic.FULLSIZE["cfg2"]
    K, c2w = synthetic.fullsize_camera("cfg2")
    cap = nb.SimpleCapture(K, c2w, c["H"], c["W"], c["near"], c["far"])
    if mode == "simt":      # the fp32 CUDA-core mode is ~40x slower: render the rows of the block only
        x0, y0, w, h = (int(v) for v in gold["cfg2_window"])
        rgb, dep = render.render_vanilla_range(nets[0], cap, nets[1], c["S"], c["N"], pix0=y0 * c["W"], n=h * c["W"], host_out=True)
        rgb = rgb.numpy().reshape(h, c["W"], 3)[:, x0:x0 + w].reshape(-1, 3)
        dep = dep.numpy().reshape(h, c["W"])[:, x0:x0 + w].reshape(-1)
    else:
        rgb, dep = nb.render_vanilla(nets[0], cap, fine_net=nets[1], samples_per_ray=c["S"], importance_samples_per_ray=c["N"],
                                     return_depth=True)
        rgb, dep = block(rgb, gold, "cfg2", 3), block(dep, gold, "cfg2", 1)[:, 0]
    graz = gold["cfg2_grazing"]
    gate(np.abs(rgb - gold["cfg2_rgb"].reshape(-1, 3)).max(-1), floors(gold, "cfg2", "rgb", mode), graz, f"cfg2 rgb {mode}")
    gate(np.abs(dep - gold["cfg2_depth"].reshape(-1)), floors(gold, "cfg2", "depth", mode), graz, f"cfg2 depth {mode}")
    assert abs(round(util.psnr(rgb, 0.5 * np.ones_like(rgb)), 2) - round(util.psnr(gold["cfg2_rgb"].reshape(-1, 3), 0.5 * np.ones_like(rgb)), 2)) <= 0.01


def _rows_of_block(g, name):
    x0, y0, w, h = (int(v) for v in g[f"{name}_window"])
    W = synthetic.FULLSIZE[name]["W"]
    return x0, y0, w, h, y0 * W, h * W


@pytest.mark.parametrize("can", [1, 0])
def test_cfg3_human_512x512_128(gold, human, mode, can):
    c = synthetic.FULLSIZE["cfg3"]
    K, c2w = synthetic.fullsize_camera("cfg3")
    cap = nb.SimpleCapture(K, c2w, c["H"], c["W"])
    b = bodies_of("cfg3")[0]
    geo = float(gold["cfg3_geo"])
    if mode == "simt":
        x0, y0, w, h, p0, n = _rows_of_block(gold, "cfg3")
        r, d, a = render.render_smpl_nerf_range(human, cap, b["verts"], b["faces"], b["Ts"], c["S"], True, bool(can), geo, 1.0,
                                                pix0=p0, n=n, host_out=True)
        cut = lambda t, C: t.numpy().reshape(h, c["W"], C)[:, x0:x0 + w].reshape(-1, C)
        r, d, a = cut(r, 3), cut(d, 1)[:, 0], cut(a, 1)[:, 0]
    else:
        r, d, a = nb.render_smpl_nerf(human, cap, b["verts"], b["faces"], b["Ts"], samples_per_ray=c["S"], render_can=bool(can),
                                      geo_threshold=geo, return_depth=True, return_mask=True)
        r, d, a = block(r, gold, "cfg3", 3), block(d, gold, "cfg3", 1)[:, 0], block(a, gold, "cfg3", 1)[:, 0]
    graz = gold["cfg3_grazing"]
    v = "" if can else "posed_"       # the posed render's floors are measured on the posed render (medial-axis flips of the warp)
    gate(np.abs(r - gold[f"cfg3_can{can}_rgb"].reshape(-1, 3)).max(-1), floors(gold, "cfg3", "rgb", mode, v), graz, f"cfg3 can={can} rgb {mode}")
    gate(np.abs(d - gold[f"cfg3_can{can}_depth"].reshape(-1)), floors(gold, "cfg3", "depth", mode, v), graz, f"cfg3 can={can} depth {mode}")
    gate(np.abs(a - gold[f"cfg3_can{can}_acc"].reshape(-1)), floors(gold, "cfg3", "acc", mode, v), graz, f"cfg3 can={can} acc {mode}")
    hit = gold["cfg3_hit"].reshape(-1)
    assert 0.2 < hit.mean() < 0.8                                   # the block straddles the silhouette


def test_cfg4_hybrid_1280x720_128_128(gold, human, mode):
    c = synthetic.FULLSIZE["cfg4"]
    K, c2w = synthetic.fullsize_camera("cfg4")
    cap = nb.SimpleCapture(K, c2w, c["H"], c["W"], c["near"], c["far"])
    b = bodies_of("cfg4")[0]
    geo = float(gold["cfg4_geo"])
    if mode == "simt":
        x0, y0, w, h, p0, n = _rows_of_block(gold, "cfg4")
        r, d, _ = render.render_hybrid_nerf_range(human, cap, b["verts"], b["faces"], b["Ts"], c["S"], c["N"], True, geo, pix0=p0, n=n,
                                                  host_out=True)
        cut = lambda t, C: t.numpy().reshape(h, c["W"], C)[:, x0:x0 + w].reshape(-1, C)
        r, d = cut(r, 3), cut(d, 1)[:, 0]
    else:
        r, d = nb.render_hybrid_nerf(human, cap, b["verts"], b["faces"], b["Ts"], samples_per_ray=c["S"],
                                     importance_samples_per_ray=c["N"], geo_threshold=geo, return_depth=True)
        r, d = block(r, gold, "cfg4", 3), block(d, gold, "cfg4", 1)[:, 0]
    graz = gold["cfg4_grazing"]
    gate(np.abs(r - gold["cfg4_rgb"].reshape(-1, 3)).max(-1), floors(gold, "cfg4", "rgb", mode), graz, f"cfg4 rgb {mode}")
    gate(np.abs(d - gold["cfg4_depth"].reshape(-1)), floors(gold, "cfg4", "depth", mode), graz, f"cfg4 depth {mode}")
    assert 0.2 < gold["cfg4_hit"].mean() < 0.8


def test_cfg5_three_actors_1280x720_128_128(gold, human, mode):
    c = synthetic.FULLSIZE["cfg5"]
    K, c2w = synthetic.fullsize_camera("cfg5")
    cap = nb.SimpleCapture(K, c2w, c["H"], c["W"], c["near"], c["far"])
    bs = bodies_of("cfg5")
    geo = float(gold["cfg5_geo"])
    args = ([human] * 3, [b["verts"] for b in bs], [b["faces"] for b in bs], [b["Ts"] for b in bs])
    if mode == "simt":
        x0, y0, w, h, p0, n = _rows_of_block(gold, "cfg5")
        r, d = nb.render_hybrid_nerf_multi_persons(human, cap, *args, samples_per_ray=c["S"], importance_samples_per_ray=c["N"],
                                                   geo_threshold=geo, return_depth=True, pix0=p0, n=n)
        r = r.reshape(h, c["W"], 3)[:, x0:x0 + w].reshape(-1, 3)
        d = d.reshape(h, c["W"])[:, x0:x0 + w].reshape(-1)
    else:
        r, d = nb.render_hybrid_nerf_multi_persons(human, cap, *args, samples_per_ray=c["S"], importance_samples_per_ray=c["N"],
                                                   geo_threshold=geo, return_depth=True)
        r, d = block(r, gold, "cfg5", 3), block(d, gold, "cfg5", 1)[:, 0]
    graz = gold["cfg5_grazing"]
    gate(np.abs(r - gold["cfg5_rgb"].reshape(-1, 3)).max(-1), floors(gold, "cfg5", "rgb", mode), graz, f"cfg5 rgb {mode}")
    gate(np.abs(d - gold["cfg5_depth"].reshape(-1)), floors(gold, "cfg5", "depth", mode), graz, f"cfg5 depth {mode}")


def test_pixel_lists_match_ranges(nets, human):
    """A rank's shard given as a pixel list (interleaved 16x16 tiles) renders exactly what the row-major
    range renders for the same pixels, for every driver; the one-gather reassembly puts every pixel back."""
    from neuman_b200 import sharding
    H, W = 72, 100                                                  # edge tiles are clipped (100 = 6*16 + 4, 72 = 4*16 + 8)
    K, c2w = scenes.camera(H, W, focal=90.0, seed=0)
    cap = nb.SimpleCapture(K, c2w, H, W, 0.0, 3.14)
    b1, _ = util.bodies()
    geo = b1["geo_threshold"]
    full_v = render.render_vanilla_range(nets[0], cap, nets[1], 32, 32, host_out=False)
    full_h = render.render_hybrid_nerf_range(human, cap, b1["verts"], b1["faces"], b1["Ts"], 32, 32, True, geo, host_out=False)
    full_s = render.render_smpl_nerf_range(human, cap, b1["verts"], b1["faces"], b1["Ts"], 32, True, False, geo, 1.0, host_out=False)
    seen = torch.zeros(H * W, dtype=torch.int32, device=DEV)
    for world in (1, 3):
        seen.zero_()
        for rank in range(world):
            part = sharding.TilePartition(H, W, rank, world, device=DEV)
            seen[part.pixels.long()] += 1
            idx = part.pixels.long()
            rgb, dep, _ = part.buffers(with_acc=False)
            render.render_vanilla_range(nets[0], cap, nets[1], 32, 32, pixels=part.pixels, host_out=False, out=(rgb, dep))
            assert torch.equal(rgb, full_v[0][idx]) and torch.equal(dep, full_v[1][idx])
            bufs = part.buffers()
            render.render_hybrid_nerf_range(human, cap, b1["verts"], b1["faces"], b1["Ts"], 32, 32, True, geo, pixels=part.pixels,
                                            host_out=False, out=bufs)
            for a, f in zip(bufs, full_h):
                assert torch.equal(a, f[idx])
            r, d, a = render.render_smpl_nerf_range(human, cap, b1["verts"], b1["faces"], b1["Ts"], 32, True, False, geo, 1.0,
                                                    pixels=part.pixels, host_out=False)
            assert torch.equal(r, full_s[0][idx]) and torch.equal(d, full_s[1][idx]) and torch.equal(a, full_s[2][idx])
            if world == 1:
                fr = part.gather()                                   # world 1: the un-permute kernel alone
                assert torch.equal(fr[0], full_h[0]) and torch.equal(fr[1], full_h[1]) and torch.equal(fr[2], full_h[2])
        assert bool((seen == 1).all())                               # every pixel in exactly one shard
