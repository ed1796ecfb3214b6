"""Timeline of the MLP kernel's MMA/epilogue hand-offs (debug): NEUMAN_TC_TRACE is set to a device buffer."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neuman_b200 as nb
from neuman_b200 import ops, synthetic
buf = torch.zeros(2 * 2 * 4 * 256, dtype=torch.int64, device="cuda")
MODE = sys.argv[1] if len(sys.argv) > 1 else "inference"      # inference | train (training forward) | bwd (backward chain)
os.environ["NEUMAN_TC_TRACE_BWD" if MODE == "bwd" else "NEUMAN_TC_TRACE"] = hex(buf.data_ptr())
coarse, _ = synthetic.seed_nets(nb.build_nerf, nb.default_opt(use_cuda=False), 1)
coarse.cuda()
R, S = 32768, 128
o = torch.randn(R, 3, device="cuda") * 0.3
d = torch.nn.functional.normalize(torch.randn(R, 3, device="cuda"), dim=-1)
z = torch.linspace(0, 3.14, S, device="cuda")[None].repeat(R, 1).contiguous()
TRAIN = MODE in ("train", "bwd")
def run():
    if TRAIN:
        pts = (o[:, None] + d[:, None] * z[..., None]).reshape(-1, 3)
        raw = coarse(pts, d[:, None].expand(R, S, 3).reshape(-1, 3))
        if MODE == "bwd":
            coarse.zero_grad()
            (raw * torch.randn_like(raw)).sum().backward()
    else:
        ops.mlp_forward_rays(coarse, o, d, z)
    torch.cuda.synchronize()
run()
buf.zero_()
run()
t = buf.cpu().numpy().reshape(2, 2, 4, 256)
m0, m1 = t[0, 0, 0], t[0, 0, 1]            # leader MMA thread: ready seen / issued+committed (tile 0)
e0, e1, e2 = t[0, 1, 0], t[0, 1, 1], t[0, 1, 2]   # leader epilogue: acc ready / drained / published
p0, p1, p2 = t[1, 1, 0], t[1, 1, 1], t[1, 1, 2]   # peer epilogue
sl = slice(22, 200)
NSTEP = 9 if MODE == "bwd" else 11
steps = np.arange(256)[sl] % NSTEP
def stat(name, x):
    x = x[sl]
    print(f"{name:58s} mean {x.mean():8.0f}  p10 {np.percentile(x,10):7.0f}  p90 {np.percentile(x,90):7.0f}")
    for s in range(NSTEP):
        print(f"      step {s:2d}: {x[steps == s].mean():8.0f}", end="")
    print()
stat("MMA: operand-ready seen -> step issued+committed", m1 - m0)
stat("MMA commit -> leader epilogue sees accumulator (incl. exec)", e0 - m1)
stat("leader epilogue: accumulator seen -> drained", e1 - e0)
stat("leader epilogue: publish (fences + arrive)", e2 - e1)
stat("leader publish -> MMA thread sees operand ready (next step)", m0[1:][sl] - e2[:-1][sl] if False else (np.roll(m0, -1) - e2))
stat("peer epilogue: accumulator seen -> drained", p1 - p0)
stat("peer epilogue: publish", p2 - p1)
stat("step period (MMA ready seen, step to step)", np.roll(m0, -1) - m0)
