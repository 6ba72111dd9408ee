#!/bin/bash
# usage: scripts/knob_sweep.sh [bench args]  -- field_main time under the tuning knobs (measurement helper)
run() { printf "%-48s" "$1"; env $1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e "${@:2}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['step_kernels_ms']
print('step %.4f ms  main %.4f  wgrad %.4f' % (d['ms_per_step'], k['field_main'], k['field_wgrad']))"; }
for knobs in "X=0" "INF_PREFETCH=0" "INF_PREFETCH=4" "INF_CLUSTER=1" "INF_PREFETCH=0 INF_CLUSTER=1" "INF_CARVEOUT=100"; do
  run "$knobs" "$@"
done
