#!/bin/bash
# launch list only (see capture.sh)
tag=${1:-r2}
mkdir -p gpurun_out
export SRPS_NO_GRAPH=1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"cg_|stencil|lighting|stack_project|normals_energy|energy_depth|halo_push|light_consts|scatter|gather|fill_|mask_|lr_|line_|scan_" -c 900 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/${tag}_under_ncu_launches.log 2>&1
wc -l gpurun_out/${tag}_launches.csv
