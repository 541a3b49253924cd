#!/usr/bin/env python
"""Turns the ncu outputs a gpurun call left in gpurun_out/ into the small text
summaries committed under profiles/ (run here, on the CPU box).
    python scripts/summarize_profiles.py r1
"""
import collections
import csv
import io
import json
import os
import shutil
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)


def launches(tag):
    src = os.path.join(G, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(P, f"{tag}_launches.csv"))
    rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
    agg = collections.defaultdict(list)
    for r in rows:
        agg[(r["Kernel Name"].split("(")[0].replace("void ", "").replace("eskf::<unnamed>::", "")
             .replace("unnamed>::", ""), r["Grid Size"])].append(us(r))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, f"{tag}_launch_summary.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of a short `bench.py` run\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare "
                "SHARES, not absolutes). Frame-phase kernels have small grids; (444|592|1184)-CTA rows, when "
                "present, are the dense roofline leg and its 10M-point map build.\n\n"
                "| kernel | grid | launches | median us | total us | share |\n|---|---|---|---|---|---|\n")
        for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {k} | {g} | {len(v)} | {statistics.median(v):.1f} | {sum(v):.0f} | {sum(v)/tot:.3f} |\n")


def ncu_rep(tag, name, title):
    rep = os.path.join(G, f"{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__warps_eligible.avg.per_cycle_active"]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    with open(os.path.join(P, f"{tag}_{name}_ncu.md"), "w") as f:
        f.write(f"# {tag}: {title}\n\n`ncu --set full --clock-control none --import-source on`, one launch "
                "(numbers under the profiler are not bench values).\n\n| metric | unit | value |\n|---|---|---|\n")
        for h, u, v in zip(hdr, units, vals):
            if h in want:
                f.write(f"| {h} | {u} | {v} |\n")
        if len(srows) > 2:
            sh = srows[1]
            ix = {h: i for i, h in enumerate(sh)}
            data = srows[2:]

            def g(r, k):
                try:
                    return float(r[ix[k]])
                except Exception:
                    return 0.0
            tot = sum(g(r, "# Samples") for r in data) or 1.0
            agg = {}
            for r in data:
                for k in sh:
                    if k.startswith("stall_") and "Not Issued" not in k:
                        agg[k] = agg.get(k, 0.0) + g(r, k)
            f.write("\n## warp-stall samples (all lines)\n\n| stall | share |\n|---|---|\n")
            for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
                f.write(f"| {k} | {v/tot:.3f} |\n")
            f.write("\n## hottest SASS lines\n\n| SASS | samples share | executed | top stall |\n|---|---|---|---|\n")
            for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:14]:
                st = {k: g(r, k) for k in sh if k.startswith("stall_") and "Not Issued" not in k}
                top = max(st.items(), key=lambda kv: kv[1])[0]
                f.write(f"| `{r[ix['Source']].strip()[:60]}` | {g(r,'# Samples')/tot:.3f} | "
                        f"{g(r,'Instructions Executed'):.0f} | {top} |\n")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(P, exist_ok=True)
    launches(tag)
    ncu_rep(tag, "prof_align", "align_kernel on the dense config (2M pts vs 10M-pt map @0.1 m, 10 GN iterations)")
    ncu_rep(tag, "prof_knn", "knn_cov_kernel on a 64k-point frame")
    ncu_rep(tag, "prof_knns", "knn_search_kernel on a 64k-point frame (0.3 m voxels, ~20k kept points)")
    ncu_rep(tag, "prof_knnf", "knn_finish_kernel on a 64k-point frame")
    ncu_rep(tag, "prof_vox", "voxelize_kernel on a 64k-point frame")
    ncu_rep(tag, "prof_voxcl", "voxelize_cluster_kernel (one 16-CTA cluster, DSMEM sort) on a 64k-point frame")
    ncu_rep(tag, "prof_ins", "insert_runs_kernel on one frame")
    ncu_rep(tag, "prof_fold", "fold_lists_kernel on one frame (sort-free map insert)")
    for n in (f"bench_{tag}.json", f"bench_ref_{tag}.json"):
        s = os.path.join(G, n)
        if os.path.exists(s):
            with open(s) as fh:
                line = fh.read().strip().splitlines()[-1]
            with open(os.path.join(P, n), "w") as fh:
                json.dump(json.loads(line), fh, indent=1)
                fh.write("\n")


if __name__ == "__main__":
    main()
