#!/bin/bash
# Run ON the GPU box (gpurun, ONE GPU): ncu launch list + full captures of the hot kernels for round $1 (default r2).
# Outputs land in gpurun_out/ and are summarised HERE (no GPU needed) by `python profiles/summarize.py r2`.
tag=${1:-r2}
mkdir -p gpurun_out
export SRPS_NO_GRAPH=1     # one launch per CG pass, so that ncu sees and serialises them
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"srps|light_consts" -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_under_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"cg_fused_kernel" -s 130 -c 2 -f -o gpurun_out/${tag}_fused $B > gpurun_out/${tag}_under_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lighting_reduce_kernel|stack_project_kernel|normals_energy_kernel|stencil_kernel" -s 4 -c 4 -f -o gpurun_out/${tag}_stack $B > gpurun_out/${tag}_under_ncu_stack.log 2>&1
SRPS_CG=fused_tma ncu --set full --clock-control none --import-source on -k regex:"cg_fused_tma_kernel" -s 130 -c 2 -f -o gpurun_out/${tag}_tma $B > gpurun_out/${tag}_under_ncu_tma.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
