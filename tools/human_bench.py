"""Full-size runs of the human renderers (BASELINE.json configs 3-5) with timing and an oracle check on
a ray subsample.  python tools/human_bench.py [cfg3 cfg4 cfg5] [--check]

cfg3: render_smpl_nerf 512x512 S=128 (canonical = no warp, and posed = warp)
cfg4: render_hybrid_nerf 1280x720 128+128 + one actor
cfg5: render_hybrid_nerf_multi_persons 1280x720 128+128 + three actors
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neuman_b200 as nb                                  # noqa: E402
from neuman_b200 import render, synthetic                 # noqa: E402
from neuman_b200._lib import Context                      # noqa: E402
from oracle import synth_smpl                             # noqa: E402  (synthetic SMPL-shaped body only)

which = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg3", "cfg4", "cfg5"]
check = "--check" in sys.argv
dev = torch.device("cuda", 0)
torch.manual_seed(1)
model = nb.HumanNeRF(nb.default_opt(use_cuda=False))
synthetic.boost_density(model.coarse_human_net)
model = model.to(dev)
ctx = Context.get(0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile(True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        prof = ctx.profile_read()
        ctx.profile(False)
        ts.append((e0.elapsed_time(e1), prof["mlp_ms"]))
    ms = float(np.median([t[0] for t in ts]))
    mlp = float(np.median([t[1] for t in ts]))
    return ms, mlp, out


def body(seed, center, scale=0.4):
    return synth_smpl.random_body(seed=seed, scale=scale, center=center)


res = {}
if "cfg3" in which:
    H = W = 512
    b = body(1, (0.0, 0.0, 0.0))
    # canonical camera: CANONICAL_ZOOM_FACTOR * W focal, distance 3 (utils/constant.py:12-13), body scaled to fill
    K, c2w = synthetic.camera(H, W, focal=1000 / 1280 * W * 3.0, seed=0, eye=(0.0, -0.05, -3.0), yaw=0.0)
    cap = nb.SimpleCapture(K, c2w, H, W)
    for can in (True, False):
        fn = lambda: render.render_smpl_nerf_range(model, cap, b["verts"], b["faces"], b["Ts"], 128, True, can,
                                                   b["geo_threshold"], 1.0, host_out=False)
        ms, mlp, out = timed(fn)
        st = ctx.render_stats()
        res[f"cfg3_can{int(can)}"] = {"ms": ms, "mlp_ms": mlp, "rays": H * W, "hit_rays": st["hit_rays"],
                                      "Mrays_s": H * W / ms / 1e3, "mlp_evals": st["mlp_evals"]}
if "cfg4" in which or "cfg5" in which:
    H, W = 720, 1280
    K, c2w = synthetic.camera(H, W, seed=1)
    cap = nb.SimpleCapture(K, c2w, H, W, 0.0, 3.14)
    bodies = [body(1, (0.1, 0.0, 0.3), 0.45), body(4, (-0.35, 0.0, 0.6), 0.45), body(7, (0.55, 0.05, 0.9), 0.45)]
    geo = bodies[0]["geo_threshold"]
    if "cfg4" in which:
        b = bodies[0]
        fn = lambda: render.render_hybrid_nerf_range(model, cap, b["verts"], b["faces"], b["Ts"], 128, 128, True, geo, host_out=False)
        ms, mlp, out4 = timed(fn)
        st = ctx.render_stats()
        res["cfg4"] = {"ms": ms, "mlp_ms": mlp, "rays": H * W, "hit_rays": st["hit_rays"], "Mrays_s": H * W / ms / 1e3,
                       "mlp_evals": st["mlp_evals"]}
    if "cfg5" in which:
        fn = lambda: render._hybrid(model, [model] * 3, cap, [x["verts"] for x in bodies], [x["faces"] for x in bodies],
                                    [x["Ts"] for x in bodies], 128, 128, True, geo, True, 0, None, False, render.CHUNK)
        ms, mlp, out5 = timed(fn)
        st = ctx.render_stats()
        res["cfg5"] = {"ms": ms, "mlp_ms": mlp, "rays": H * W, "hit_rays_sum": st["hit_rays"], "Mrays_s": H * W / ms / 1e3,
                       "mlp_evals": st["mlp_evals"]}
    if check and "cfg4" in which:
        # oracle on a 256-ray subsample that straddles the body silhouette
        from oracle import neuman_oracle as no
        rgb = out4[0].cpu().numpy()
        acc = out4[2].cpu().numpy()
        hit = np.flatnonzero(acc > 0)
        idx = np.sort(np.concatenate([hit[:: max(1, len(hit) // 160)][:160], np.arange(0, H * W, H * W // 96)[:96]]))
        m = model.to("cpu")
        hp = no.net_params_from_joiner(m.coarse_human_net)
        cb, fb = no.net_params_from_joiner(m.coarse_bkg_net), no.net_params_from_joiner(m.fine_bkg_net)
        t0 = time.time()
        r_o, d_o, a_o = no.render_hybrid_nerf(cb, fb, hp, K, c2w, H, W, 0.0, 3.14, b["verts"], b["faces"], b["Ts"],
                                              samples_per_ray=128, importance_samples_per_ray=128, geo_threshold=geo,
                                              ray_subset=idx)
        err = np.abs(rgb[idx] - r_o).max(-1)
        res["cfg4_check"] = {"n": int(len(idx)), "hit_in_sample": int((a_o > 0).sum()), "max_abs_rgb": float(err.max()),
                             "frac_gt_1e-4": float((err > 1e-4).mean()), "oracle_s": time.time() - t0}
        model.to(dev)
print(json.dumps(res, indent=1))
