"""Training-step timing: python tools/train_bench.py [rays] [steps]
One step = trainers/vanilla_nerf_trainer.py:206-223 (loss_func + backward + Adam) at the reference's
defaults (rays_per_batch 2048, 128 + 128 samples, perturb 1, raw_noise_std 1).
Arms: this repo's CUDA path (neuman_b200.train) and a plain PyTorch restatement of the same step on the
same GPU (nn.Linear modules, torch ops for sampling / compositing; fp32 matmuls and TF32 matmuls).
Prints one JSON object per arm."""
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neuman_b200 as nb                      # noqa: E402
from neuman_b200 import train as nt           # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 20
S, N = 128, 128
dev = "cuda"
opt = nb.default_opt(perturb=1.0, raw_noise_std=1.0)
torch.manual_seed(0)
batch = dict(origin=torch.randn(R, 3, device=dev) * 0.1,
             direction=F.normalize(torch.randn(R, 3, device=dev), dim=-1),
             near=torch.full((R,), 0.5, device=dev), far=torch.full((R,), 4.0, device=dev),
             color=torch.rand(R, 3, device=dev))


# ---- plain PyTorch arm ---------------------------------------------------------------------------
class TorchNeRF(nn.Module):
    def __init__(self):
        super().__init__()
        self.pts = nn.ModuleList([nn.Linear(63, 256)] + [nn.Linear(256 + (63 if i == 4 else 0), 256) for i in range(7)])
        self.views = nn.Linear(256 + 27, 128)
        self.feature, self.alpha, self.rgb = nn.Linear(256, 256), nn.Linear(256, 1), nn.Linear(128, 3)

    @staticmethod
    def pe(x, nf, maxf):
        fr = 2.0 ** torch.linspace(0, maxf, nf, device=x.device)
        parts = [x]
        for f in fr:
            parts += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(parts, -1)

    def forward(self, pts, dirs):
        e, v = self.pe(pts, 10, 9), self.pe(dirs, 4, 3)
        h = e
        for i, l in enumerate(self.pts):
            h = F.relu(l(h))
            if i == 4:
                h = torch.cat([e, h], -1)
        a = self.alpha(h)
        h = F.relu(self.views(torch.cat([self.feature(h), v], -1)))
        return torch.cat([self.rgb(h), a], -1)


def t_raw2outputs(raw, z, d, noise_std):
    dz = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], -1) * d.norm(dim=-1, keepdim=True)
    sig = raw[..., 3] + torch.randn_like(raw[..., 3]) * noise_std
    alpha = 1. - torch.exp(-F.relu(sig) * dz)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T
    rgb = (w[..., None] * torch.sigmoid(raw[..., :3])).sum(-2)
    return rgb + (1. - w.sum(-1, keepdim=True)), w


def t_sample_pdf(bins, weights, n):
    weights = weights + 1e-5
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    u = torch.linspace(0., 1., n, device=bins.device).expand(bins.shape[0], n).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = (inds - 1).clamp_min(0), inds.clamp_max(cdf.shape[-1] - 1)
    g = torch.stack([below, above], -1)
    ms = [g.shape[0], g.shape[1], cdf.shape[-1]]
    cg = torch.gather(cdf[:, None].expand(ms), 2, g)
    bg = torch.gather(bins[:, None].expand(ms), 2, g)
    den = cg[..., 1] - cg[..., 0]
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    return bg[..., 0] + (u - cg[..., 0]) / den * (bg[..., 1] - bg[..., 0])


def torch_step(coarse, fine, optim):
    optim.zero_grad()
    o, d = batch['origin'], batch['direction']
    t = torch.linspace(0., 1., S, device=dev)
    z = (batch['near'][:, None] * (1 - t) + batch['far'][:, None] * t)
    mid = .5 * (z[..., 1:] + z[..., :-1])
    up, lo = torch.cat([mid, z[..., -1:]], -1), torch.cat([z[..., :1], mid], -1)
    z = lo + (up - lo) * torch.rand_like(z)
    pts = o[:, None] + d[:, None] * z[..., None]
    dd = d[:, None].expand_as(pts)
    rgb, w = t_raw2outputs(coarse(pts, dd), z, d, 1.0)
    loss = F.mse_loss(rgb, batch['color'])
    zm = .5 * (z[..., 1:] + z[..., :-1])
    zs = t_sample_pdf(zm, w[..., 1:-1].detach(), N).detach()
    Fz = torch.sort(torch.cat([z, zs], -1), -1)[0]
    Fp = o[:, None] + d[:, None] * Fz[..., None]
    Frgb, _ = t_raw2outputs(fine(Fp, d[:, None].expand_as(Fp)), Fz, d, 1.0)
    loss = loss + F.mse_loss(Frgb, batch['color'])
    loss.backward()
    optim.step()
    return loss.detach()


def timeit(fn, label, extra):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        last = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    flops = 3 * 1186816 * R * (S + S + N)            # fwd + dX + dW, coarse S and fine S+N evaluations
    print(json.dumps(dict(arm=label, rays=R, samples=[S, S + N], ms_per_step=round(ms, 3), rays_per_s=round(R / ms * 1e3),
                          model_tflops=round(flops / ms / 1e9, 1), loss=float(last), **extra)), flush=True)
    return ms


coarse, fine = nb.build_nerf(nb.default_opt())
optim = torch.optim.Adam(list(coarse.parameters()) + list(fine.parameters()), lr=5e-4)
ms_ours = timeit(lambda: nt.train_batch(coarse, fine, optim, batch, opt, check_bad_weights=False), "neuman_b200", {})
# per-phase split of our step
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
optim.zero_grad()
ev[0].record()
ls = nt.vanilla_loss_func(coarse, fine, batch, opt, check_bad_weights=False)
ev[1].record()
(ls[0] + ls[2]).backward()
ev[2].record()
optim.step()
ev[3].record()
torch.cuda.synchronize()
print(json.dumps(dict(arm="neuman_b200 phases", forward_ms=round(ev[0].elapsed_time(ev[1]), 3),
                      backward_ms=round(ev[1].elapsed_time(ev[2]), 3), adam_ms=round(ev[2].elapsed_time(ev[3]), 3))), flush=True)

if os.environ.get("TRAIN_BENCH_PROFILE"):
    # kernel-level split of our step (CUPTI through torch.profiler; times are not bench values)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            nt.train_batch(coarse, fine, optim, batch, opt, check_bad_weights=False)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90), flush=True)
    sys.exit(0)

for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    tc, tf = TorchNeRF().to(dev), TorchNeRF().to(dev)
    to = torch.optim.Adam(list(tc.parameters()) + list(tf.parameters()), lr=5e-4)
    ms_t = timeit(lambda: torch_step(tc, tf, to), "pytorch " + ("tf32" if tf32 else "fp32"), {"speedup_of_neuman_b200": None})
    print(json.dumps(dict(arm="ratio", torch=("tf32" if tf32 else "fp32"), speedup=round(ms_t / ms_ours, 2))), flush=True)
    del tc, tf, to
