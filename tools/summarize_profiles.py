"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tools/summarize_profiles.py r01
"""
import csv
import collections
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list of one bench run (cold-cache, serialised: compare SHARES, not absolutes) ----
lc = os.path.join(G, "launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        ns = float(r[14])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_bench_launches.md"), "w") as f:
        f.write(f"# ncu launch list, `python bench.py --steps 1 --warmup 3 --no-cpu-baseline` ({len(rows)} launches captured)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 700`; per-launch times are cold-cache and serialised: shares only.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns/1e6:.3f} | {100*ns/tot:.2f} % |\n")
    print("wrote launches summary", len(rows))

# ---- full capture of the MLP kernel ----
rep = os.path.join(G, "prof_mlp.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
            "launch__cluster_dim_x", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    with open(os.path.join(P, f"{tag}_mlp_tc_ncu.md"), "w") as f:
        f.write("# `ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 8 -c 1` on `tools/tc_check.py tc` (4 194 304 evaluations, pair mode)\n\n")
        f.write("| metric | value | unit |\n|---|---|---|\n")
        for w in want:
            if w in d:
                f.write(f"| {w} | {d[w][1]} | {d[w][0]} |\n")
        # stall summary from the source page
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        h = rows[1]
        ix = {k: i for i, k in enumerate(h)}
        data = rows[2:]
        tot = sum(int(r[ix["# Samples"]]) for r in data)
        f.write(f"\n## Hottest SASS instructions (warp-stall samples, total {tot})\n\n| samples | executed | instruction | top stall |\n|---|---|---|---|\n")
        stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:15]:
            st = max(((int(r[ix[c]]), c) for c in stall_cols))
            f.write(f"| {r[ix['# Samples']]} | {r[ix['Instructions Executed']]} | `{r[ix['Source']].strip()[:70]}` | {st[1]} |\n")
        mn = collections.Counter()
        for r in data:
            op = r[ix["Source"]].strip()
            for key in ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "STS.128", "LDS.128", "SYNCS"):
                if key in op:
                    mn[key] += int(r[ix["Instructions Executed"]])
        f.write("\n## Blackwell-native instruction counts (executed warp-instructions)\n\n")
        for k, v in mn.most_common():
            f.write(f"* `{k}`: {v}\n")
    print("wrote mlp summary")
