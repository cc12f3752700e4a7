"""Turns the raw ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python profiles/summarize.py r1      (needs ncu on PATH for the --page raw export of the .ncu-rep)"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")

# ---- launch list ---------------------------------------------------------------------------------
path = os.path.join(src, f"{tag}_launches.csv")
if os.path.exists(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rd:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[idx["Metric Value"]].replace(",", ""))
        u = r[idx["Metric Unit"]]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        a = agg.setdefault(r[idx["Kernel Name"]], [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if "at::" not in k and "elementwise" not in k}      # (capture.sh already filters on the library's kernel names)
    tot = sum(v[1] for v in ours.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none (first launches of the library's kernels)  python bench.py --steps 2 --warmup 1 --no-cpu   (SRPS_NO_GRAPH=1)",
           "# 4096x4096 HR, sf=4, 32 images; cold-cache serialised launch times: compare SHARES, not absolutes.",
           "# kernels of this library only (torch's synthetic-scene generation omitted); shares within the library",
           "%-72s %6s %12s %7s %10s" % ("kernel", "count", "total_us", "share", "avg_us")]
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append("%-72s %6d %12.1f %6.1f%% %10.2f" % (k[:72], n, t, 100 * t / tot, t / n))
    open(os.path.join(dst, f"{tag}_launches_summary.txt"), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(path, os.path.join(dst, f"{tag}_launches.csv"))
    print("\n".join(out))

# ---- full captures -----------------------------------------------------------------------------------
# profiles/capture.sh leaves the raw pages of its three `ncu --set full` captures in gpurun_out/<tag>_{fused,stack,tma}_raw.csv
# (round 1: the .ncu-rep files themselves, <tag>_final.ncu-rep / <tag>_full.ncu-rep, exported here)
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
out = [f"# ncu --set full --clock-control none --import-source on, SRPS_NO_GRAPH=1 bench.py --steps 2 --warmup 1 --no-cpu --no-extras (4096x4096, sf=4, 32 images), {tag}",
       "# one launch per kernel (the first captured launch of each; for the CG kernels a pass in the middle of a solve)", ""]
seen = set()
traffic = {}
sources = []
for name in ("fused", "stack", "tma", "pf_4k", "pf_n8sim", "pf_slab8"):
    path = os.path.join(src, f"{tag}_{name}_raw.csv")
    if os.path.exists(path):
        sources.append((f"{tag}_{name}_raw.csv", open(path, errors="ignore").read()))
for name in (f"{tag}_final.ncu-rep", f"{tag}_full.ncu-rep"):
    rep = os.path.join(src, name)
    if os.path.exists(rep):
        sources.append((name, subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout))
for name, raw in sources:
    rows = [r for r in csv.reader(raw.splitlines()) if r]
    start = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if start is None:
        continue
    hdr, units = rows[start], rows[start + 1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[start + 2:]:
        if len(r) < len(hdr):
            continue
        k = r[idx["Kernel Name"]]
        if "_pf_" in name:          # the same kernel on three workloads: keep them apart
            k += {"pf_4k": "  @4096x4096", "pf_n8sim": "  @4096x1024 (4-GPU slab)", "pf_slab8": "  @4096x512 (8-GPU slab)"}[name[len(tag) + 1:-len("_raw.csv")]]
        if k in seen:
            continue
        seen.add(k)
        out.append(f"## {k}    [{name}]")
        for w in want:
            if w in idx:
                out.append("    %-86s %s %s" % (w, r[idx[w]], units[idx[w]]))
        out.append("")

        def nbytes(metric):
            v, u = float(r[idx[metric]].replace(",", "")), units[idx[metric]]
            return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic[k] = int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"))
if seen:
    open(os.path.join(dst, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(out))
    print("wrote", f"{tag}_ncu_full_summary.txt", len(seen), "kernels")
    import json
    key = {"cg_persistent_fused_kernel<4, 1, 3, 128>(PersistentArgs)  @4096x4096": "cg_persistent_fused", "cg_fused_kernel<4, 0": "cg_fused", "cg_fused_tma_kernel<4": "cg_fused_tma", "stencil_strip_kernel<0, 4>": "cg_operator",
           "cg_update_kernel": "cg_update", "stack_project_kernel<1": "project_pass", "lighting_reduce_kernel": "lighting_pass"}
    tpath = os.path.join(dst, "traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t4k = dict(old.get("4k", {}))
    t4k["source"] = f"profiles/{tag}_ncu_full_summary.txt (entries not re-captured this round keep the value of the round before)"
    for k, v in traffic.items():
        for pat, nm in key.items():
            if pat in k:
                t4k[nm] = v
    json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures summarised under "
                           "profiles/ (4096x4096 sf=4 n=32, one GPU); read by bench.py -> roofline.traffic",
               "4k": t4k}, open(tpath, "w"), indent=1)
    print(json.dumps(t4k))
# the exported details / top-stall source lines travel as they are
import shutil
for name in ("fused_details.txt", "tma_details.txt", "stack_details.txt", "fused_source_top.csv", "tma_source_top.csv",
             "pf_4k_details.txt", "pf_n8sim_details.txt", "pf_slab8_details.txt", "pf_4k_source_top.csv", "pf_n8sim_source_top.csv", "pf_slab8_source_top.csv"):
    p = os.path.join(src, f"{tag}_{name}")
    if os.path.exists(p) and os.path.getsize(p) > 0:
        shutil.copy(p, os.path.join(dst, f"{tag}_ncu_{name}"))
