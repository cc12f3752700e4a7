#!/bin/bash
# Run ON the GPU box (gpurun, ONE GPU): ncu launch list + full captures of the hot kernels for round $1 (default r2).
# The .ncu-rep files stay on the box (/tmp: gpurun_out is limited to 64 MiB); their raw / details / source pages are
# exported to gpurun_out/ and summarised HERE (no GPU needed) by `python profiles/summarize.py r2`.
tag=${1:-r2}
mkdir -p gpurun_out
export SRPS_NO_GRAPH=1     # one launch per CG pass, so that ncu sees and serialises them
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cg_|stencil|lighting|stack_project|normals_energy|energy_depth|halo_push|light_consts|scatter|gather|fill_|mask_|lr_|line_|scan_" -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_under_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cg_fused_kernel" -s 130 -c 2 -f -o /tmp/${tag}_fused $B > gpurun_out/${tag}_under_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lighting_reduce_kernel|stack_project_kernel|normals_energy_kernel|stencil_kernel" -s 4 -c 4 -f -o /tmp/${tag}_stack $B > gpurun_out/${tag}_under_ncu_stack.log 2>&1
SRPS_CG=fused_tma ncu --set full --clock-control none --import-source on -k regex:"cg_fused_tma_kernel" -s 130 -c 2 -f -o /tmp/${tag}_tma $B > gpurun_out/${tag}_under_ncu_tma.log 2>&1
for r in fused stack tma; do
    ncu -i /tmp/${tag}_$r.ncu-rep --page raw --csv > gpurun_out/${tag}_${r}_raw.csv 2>/dev/null
    ncu -i /tmp/${tag}_$r.ncu-rep --page details > gpurun_out/${tag}_${r}_details.txt 2>/dev/null
done
# per-instruction stall samples of the two CG kernels (first captured launch): top lines only
for r in fused tma; do
    ncu -i /tmp/${tag}_$r.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > /tmp/${tag}_${r}_source.csv 2>/dev/null
    python - /tmp/${tag}_${r}_source.csv gpurun_out/${tag}_${r}_source_top.csv <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = i
        break
if hdr is None:
    open(sys.argv[2], "w").write("no source page\n")
    sys.exit(0)
h = rows[hdr]
col = next(j for j, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
body = [r for r in rows[hdr + 1:] if len(r) == len(h)]
def val(r):
    try:
        return float(r[col].replace(",", ""))
    except Exception:
        return 0.0
body.sort(key=val, reverse=True)
w = csv.writer(open(sys.argv[2], "w"))
w.writerow(h)
w.writerows(body[:60])
PY
done
ls -la gpurun_out/${tag}_*
