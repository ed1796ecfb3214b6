"""End-to-end background-NeRF training on the device: python tools/train_demo.py [steps] [rays]
A seeded "teacher" NeRF renders a few synthetic views (render_vanilla); a freshly initialised student is trained on
them with the whole loop on the GPU: BackgroundRayBatcher (datasets/background_rays.py) -> train_batch
(trainers/vanilla_nerf_trainer.py:206-223) with Adam.  Prints the loss trajectory, steps/s including batch
production, and the PSNR of a held-out view rendered by the student against the teacher's image."""
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neuman_b200 as nb                                 # noqa: E402
from neuman_b200 import data as nd, synthetic            # noqa: E402
from neuman_b200 import train as nt                      # noqa: E402
from neuman_b200.render import SimpleCapture             # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
RAYS = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
H = W = 96
S, N = 64, 64

opt = nb.default_opt(samples_per_ray=S, importance_samples_per_ray=N, perturb=1.0, raw_noise_std=0.0, rays_per_batch=RAYS,
                     use_fused_depth=False, ablate_nerft=False)
t_coarse, t_fine = synthetic.seed_nets(nb.build_nerf, opt, 11)
for n_ in (t_coarse, t_fine):
    synthetic.boost_density(n_, gain=20.0, bias=0.5)
    with torch.no_grad():                                 # default-init colours are nearly constant: add contrast
        n_.nerf.rgb_linear.weight.mul_(25.0)
        n_.nerf.pts_linears[0].weight.mul_(3.0)


def make_cap(k):
    K, c2w = synthetic.camera(H, W, focal=110.0, seed=k, yaw=-0.5 + 0.25 * k)
    return SimpleCapture(K, c2w, H, W, near=0.5, far=3.5)


def teacher_image(cap):
    with torch.no_grad():
        img = nb.render_vanilla(t_coarse, cap, t_fine, samples_per_ray=S, importance_samples_per_ray=N)
    return np.asarray(img)


caps = [make_cap(k) for k in range(5)]
for k, cap in enumerate(caps):
    img = teacher_image(cap)
    cap.image = img if img.dtype == np.uint8 else (np.clip(img, 0, 1) * 255).astype(np.uint8)
    cap.depth_map = np.ones((H, W), np.float32)
    cap.mask = np.zeros((H, W), np.uint8)
    cap.binary_mask = cap.mask
    cap.frame_id = {'frame_id': k, 'total_frames': len(caps)}
train_caps, held_out = caps[:4], caps[4]
print("teacher image stats: mean %.3f std %.3f" % (caps[0].image.mean() / 255, caps[0].image.std() / 255), flush=True)

torch.manual_seed(0)
np.random.seed(0)
coarse, fine = nb.build_nerf(opt)
for n_ in (coarse, fine):
    synthetic.boost_density(n_, gain=4.0, bias=0.1)      # avoid the dead-density re-init branch at step 0
optim = torch.optim.Adam(list(coarse.parameters()) + list(fine.parameters()), lr=5e-4)
batcher = nd.BackgroundRayBatcher(opt, train_caps)


def psnr_held_out():
    with torch.no_grad():
        img = np.asarray(nb.render_vanilla(coarse, held_out, fine, samples_per_ray=S, importance_samples_per_ray=N))
    a = img.astype(np.float64) / (255.0 if img.dtype == np.uint8 else 1.0)
    b = held_out.image.astype(np.float64) / 255.0
    return float(-10 * np.log10(((a - b) ** 2).mean()))


print(json.dumps({"step": 0, "held_out_psnr_db": round(psnr_held_out(), 2)}), flush=True)
losses = []
torch.cuda.synchronize()
t0 = time.time()
for it in range(STEPS):
    loss = nt.train_batch(coarse, fine, optim, batcher(), opt, iteration=it, check_bad_weights=False)
    if it % max(1, STEPS // 10) == 0 or it == STEPS - 1:
        losses.append((it, float(loss)))
torch.cuda.synchronize()
dt = time.time() - t0
print(json.dumps({"steps": STEPS, "rays_per_step": RAYS, "samples": [S, S + N], "wall_s": round(dt, 2),
                  "steps_per_s": round(STEPS / dt, 1), "rays_per_s": round(STEPS * RAYS / dt),
                  "loss": [(i, round(l, 5)) for i, l in losses], "held_out_psnr_db": round(psnr_held_out(), 2)}), flush=True)
