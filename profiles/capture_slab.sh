#!/bin/bash
# Run ON the GPU box (gpurun, ONE GPU): full ncu capture of the persistent fused CG kernel (the default driver: the whole
# solve is one cooperative launch) on config 4 itself and on its per-GPU slabs at 4 and 8 GPUs (bench.py workloads
# n8sim / slab8 on one GPU: no exchange).  `python profiles/summarize.py r2` turns the raw pages into the tracked summaries.
tag=${1:-r2}
mkdir -p gpurun_out
for w in ${2:-4k slab8 n8sim}; do
    SRPS_CG=persistent_fused ncu --set full --clock-control none --import-source on -k regex:"cg_persistent_fused_kernel" -s 2 -c 1 -f -o /tmp/${tag}_pf_$w \
        python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/${tag}_under_ncu_pf_$w.log 2>&1
    ncu -i /tmp/${tag}_pf_$w.ncu-rep --page details > gpurun_out/${tag}_pf_${w}_details.txt 2>/dev/null
    ncu -i /tmp/${tag}_pf_$w.ncu-rep --page raw --csv > gpurun_out/${tag}_pf_${w}_raw.csv 2>/dev/null
    ncu -i /tmp/${tag}_pf_$w.ncu-rep --page source --csv > /tmp/${tag}_pf_${w}_source.csv 2>/dev/null
    python - /tmp/${tag}_pf_${w}_source.csv gpurun_out/${tag}_pf_${w}_source_top.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next((i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r)), None)
if hdr is None:
    open(sys.argv[2], "w").write("no source page\n"); sys.exit(0)
h = rows[hdr]
col = next(j for j, c in enumerate(h) if c.startswith("# Samples") or "Sampling (All" in c)
body = [r for r in rows[hdr + 1:] if len(r) == len(h)]
def val(r):
    try: return float(r[col].replace(",", ""))
    except Exception: return 0.0
body.sort(key=val, reverse=True)
w = csv.writer(open(sys.argv[2], "w")); w.writerow(h); w.writerows(body[:80])
PY
done
ls -la gpurun_out/${tag}_pf_*
