"""One human-trainer step (bench.py's `human_train_step` workload: 2048 rays x 128 samples, offset net included, backward,
no optimiser) between cudaProfilerStart / Stop, for an ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/r02_human_step_launches.csv python tools/human_train_step_once.py
    python tools/human_train_step_once.py --summarize gpurun_out/r02_human_step_launches.csv > profiles/r02_human_step_launches.md
"""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarize(path):
    rows = []
    with open(path, newline="") as fp:
        lines = [l for l in fp if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
            rows.append((r["Kernel Name"].split("(")[0][:90], v))
    agg = {}
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in rows)
    print("# ncu launch list of ONE human-trainer step (forward + backward, 2048 rays x 128 samples, offset net included)\n")
    print("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` around one step of "
          "`tools/human_train_step_once.py`; per-launch times are cold-cache and serialised: shares, not a step time "
          f"({len(rows)} launches, {tot:.3f} ms summed).\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {v:.4f} | {100 * v / tot:.2f} % |")


def main():
    import numpy as np
    import torch
    import neuman_b200 as nb
    from neuman_b200 import train as nt
    from neuman_b200.synthetic import make_model
    dev = "cuda"
    R, S = 2048, 128
    rng = np.random.RandomState(0)
    pose, betas = rng.normal(0, 0.3, (1, 72)).astype(np.float32), rng.normal(0, 1.0, (1, 10)).astype(np.float32)
    align = np.eye(4, dtype=np.float32)
    align[3, :3] = (0.3, -0.1, 2.0)
    opt = nb.default_opt(use_cuda=True, num_offset_nets=1, offset_scale=0.02, offset_scale_type='tanh', samples_per_ray=S)
    torch.manual_seed(0)
    model = make_model(0)
    net = nb.HumanNeRF(opt, poses=pose, betas=betas, alignments=align[None], scale=0.4, smpl_model=model)
    faces = np.ascontiguousarray(model["f"][:, :3].astype(np.int64))
    with torch.no_grad():
        V0 = net.vertex_forward(0)[0][0]
    eye = V0.mean(0) + torch.tensor([0.0, 0.0, -2.0], device=dev)
    d = V0[torch.randint(0, V0.shape[0], (R,), device=dev)] + 0.02 * torch.randn(R, 3, device=dev) - eye
    dist = d.norm(dim=1, keepdim=True)
    batch = {'origin': eye[None].repeat(R, 1), 'direction': d / dist, 'human_near': dist - 0.2, 'human_far': dist + 0.2,
             'cur_view_f': 3 / 11, 'cap_id': 0}

    def step():
        net.zero_grad()
        out = nt.eval_human_samples(net, batch, opt, faces, offset_net=net.offset_nets[0])
        out[5].square().mean().backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarize":
        summarize(sys.argv[2])
    else:
        main()
