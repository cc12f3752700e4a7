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
    ours = {k: v for k, v in agg.items() if "srps::" in k or "light_consts" in k}
    tot = sum(v[1] for v in ours.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 600  python bench.py --steps 2 --warmup 1 --no-cpu   (SRPS_NO_GRAPH=1)",
           "# 4096x4096 HR, sf=4, 32 images; cold-cache serialised launch times: compare SHARES, not absolutes.",
           "# kernels of this library only (the first 600 launches also contain torch's synthetic-scene generation, omitted); shares within the library",
           "%-72s %6s %12s %7s %10s" % ("kernel", "count", "total_us", "share", "avg_us")]
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append("%-72s %6d %12.1f %6.1f%% %10.2f" % (k[:72], n, t, 100 * t / tot, t / n))
    open(os.path.join(dst, f"{tag}_launches_summary.txt"), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(path, os.path.join(dst, f"{tag}_launches.csv"))
    print("\n".join(out))

# ---- full captures -----------------------------------------------------------------------------------
# gpurun_out/<tag>_final.ncu-rep : the default path (lighting pass, stack projection pass, fused CG pass)
# gpurun_out/<tag>_full.ncu-rep  : the two-kernel CG form (SRPS_CG=graph: warp-strip operator + update)
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
out = [f"# ncu --set full --clock-control none --import-source on, SRPS_NO_GRAPH=1 bench.py --steps 1 --no-cpu (4096x4096, sf=4, 32 images), {tag}",
       "# one launch per kernel (for cg_fused_kernel the first launch that is not the <.., true> first-pass variant)", ""]
seen = set()
traffic = {}
for name in (f"{tag}_final.ncu-rep", f"{tag}_full.ncu-rep"):
    rep = os.path.join(src, name)
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        k = r[idx["Kernel Name"]]
        if k in seen:
            continue
        seen.add(k)
        out.append(f"## {k}    [{name}]")
        for w in want:
            if w in idx:
                out.append("    %-78s %s %s" % (w, r[idx[w]], units[idx[w]]))
        out.append("")

        def nbytes(metric):
            v, u = float(r[idx[metric]].replace(",", "")), units[idx[metric]]
            return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic[k] = int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"))
if seen:
    open(os.path.join(dst, f"{tag}_ncu_full_summary.txt"), "w").write("\n".join(out))
    print("wrote", f"{tag}_ncu_full_summary.txt", len(seen), "kernels")
    import json
    key = {"cg_fused_kernel<4, 0>": "cg_fused", "stencil_strip_kernel<0, 4>": "cg_operator", "cg_update_kernel": "cg_update",
           "stack_project_kernel<1>": "project_pass", "lighting_reduce_kernel": "lighting_pass"}
    t4k = {"source": f"profiles/{tag}_ncu_full_summary.txt"}
    for k, v in traffic.items():
        for pat, name in key.items():
            if pat in k:
                t4k[name] = v
    json.dump({"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full captures summarised in "
                           f"profiles/{tag}_ncu_full_summary.txt (4096x4096 sf=4 n=32); read by bench.py -> roofline.traffic",
               "4k": t4k}, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    print(json.dumps(t4k))
