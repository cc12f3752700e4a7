#!/bin/bash
# A/B helper: bench one workload under several environment settings, one summary line each.
#   tests/ab_env.sh <workload> "<ENV=.. ENV=..>" ["<ENV=..>" ...]
w=$1; shift
for e in "$@"; do
  env $e SRPS_VERBOSE=1 python bench.py --workload $w --no-extras --no-cpu 2>/tmp/ab_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]);p=d['phases_ms']
print('$w [$e]', round(d['ms_per_step'],4), 'cg', round(p['ms_depth_cg'],4), 'light', round(p['ms_lighting'],4), 'alb', round(p['ms_albedo'],4), d['roofline']['cg_driver'])"
  grep -m1 "L2 persisting" /tmp/ab_err.log
done
