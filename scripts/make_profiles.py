#!/usr/bin/env python3
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/make_profiles.py <tag>      (tag as passed to scripts/gpu_check.sh)
Writes profiles/<tag>_launches.txt (per-kernel totals of the ncu launch list), profiles/<tag>_full.txt
(key metrics per captured kernel from the --set full report) and copies the bench JSON lines.
"""
import collections, csv, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']

lf = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lf):
    rows = list(csv.reader(open(lf)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
        agg[r[ki][:70]][0] += 1
        agg[r[ki][:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{tag}_launches.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n"
                f"# source: gpurun_out/launches_{tag}.csv, command: python bench.py --steps 2 --warmup 1 --cg-iters 20 --no-cpu-baseline --no-e2e\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:70s} n={v[0]:5d} total_ms={v[1] / 1e3:10.3f} avg_us={v[1] / v[0]:10.1f} share={v[1] / tot:.4f}\n")

rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, f"{tag}_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source: gpurun_out/prof_{tag}.ncu-rep (not tracked)\n")
        seen = set()
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')][:80]
            if name in seen:
                continue
            seen.add(name)
            f.write(f"--- {name}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"    {k:82s} {r[i]} {units[i]}\n")
    # DRAM bytes per launch of the two dominant kernels -> profiles/traffic.json (bench.py roofline.traffic)
    import json
    tj = os.path.join(P, "traffic.json")
    traffic = json.load(open(tj)) if os.path.exists(tj) else {}      # kernels absent from this capture keep their entry
    src = traffic.get("sources", {})
    conv = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ir, iw, ik = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum'), hdr.index('Kernel Name')
    best = {}
    for r in rows[2:]:
        for key in ("spmv_block_kernel", "spmv_stream_kernel", "lspace_gather_kernel", "lspace_cluster_kernel"):
            if key in r[ik]:
                b = float(r[ir].replace(',', '')) * conv.get(units[ir], 1.0) + float(r[iw].replace(',', '')) * conv.get(units[iw], 1.0)
                best.setdefault(key, []).append(b)
    for key, v in best.items():
        v.sort()
        traffic[key] = v[len(v) // 2]          # median launch
        src[key] = tag
    traffic["source"] = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch at the 1M-hex config (capture per kernel: sources)"
    traffic["sources"] = src
    if best:
        json.dump(traffic, open(tj, "w"), indent=1)
for fn in (f"bench_{tag}.json", f"bench_ref_{tag}.json"):
    if os.path.exists(os.path.join(G, fn)):
        shutil.copy(os.path.join(G, fn), os.path.join(P, fn))
print("profiles written for", tag)
