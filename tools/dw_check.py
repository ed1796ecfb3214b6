"""GPU diagnostic for k_dw_gemm: python tools/dw_check.py [n]  -- per-item error against torch fp32 matmuls and timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuman_b200 import ops                   # noqa: E402
from neuman_b200.ops import _p       # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
torch.manual_seed(0)
dev = "cuda"
h = dict(device=dev, dtype=torch.float16)
g_pre = (torch.randn(8, n, 256, device=dev) * 0.5).half()
g_f = (torch.randn(n, 256, device=dev) * 0.5).half()
g_v = (torch.randn(n, 128, device=dev) * 0.5).half()
sx = torch.relu(torch.randn(8, n, 256, device=dev)).half()
sf = torch.randn(n, 256, device=dev).half()
ctx = ops._ctx_for(g_f)
out = torch.full((9, 256, 256), float("nan"), device=dev)
bias = torch.full((9, 256), float("nan"), device=dev)
ctx.check(ctx.lib.nm_dw_gemm(ctx.h, _p(g_pre), _p(g_f), _p(g_v), _p(sx), _p(sf), n, _p(out), _p(bias), ctx.stream()))
torch.cuda.synchronize()
ref = torch.zeros(9, 256, 256, device=dev)
for k in range(7):
    ref[k] = g_pre[k + 1].float().t() @ sx[k].float()
ref[7] = g_f.float().t() @ sx[7].float()
ref[8, :128] = g_v.float().t() @ sf.float()
for k in range(9):
    d = (out[k] - ref[k]).abs()
    print(f"item {k}: rel {float((out[k] - ref[k]).norm() / ref[k].norm()):.3e} max abs {float(d.max()):.3e} "
          f"ref max {float(ref[k].abs().max()):.3e} nan {int(torch.isnan(out[k]).sum())}", flush=True)
bref = torch.zeros(9, 256, device=dev)
bref[:7] = g_pre[1:].float().sum(1)
bref[7] = g_f.float().sum(0)
bref[8, :128] = g_v.float().sum(0)
print("bias max abs err", float((bias - bref).abs().max()), "ref max", float(bref.abs().max()), "nan", int(torch.isnan(bias).sum()))
if n >= 100000:
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.check(ctx.lib.nm_dw_gemm(ctx.h, _p(g_pre), _p(g_f), _p(g_v), _p(sx), _p(sf), n, _p(out), _p(bias), ctx.stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        gb = n * (8 * 1024 + 768) / 1e9
        print(f"n={n}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  {2 * n * 256 * (8 * 256 + 128) / ms / 1e9:.0f} TFLOP/s")
