#!/bin/bash
# usage: scripts/sweep.sh  -- prints kernel times for a few configurations (measurement helper)
run() { echo "== $*"; env "${@:2}" python bench.py --steps 8 --warmup 3 --no-cpu-baseline $1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']; print('  rays/s %.1fM  step %.3f ms  main %.3f wgrad %.3f reduce %.3f  e2e %.1fM  b4096 %s' % (d['value']/1e6, d['ms_per_step'], r['step_kernels_ms']['field_main'], r['step_kernels_ms']['field_wgrad'], r['step_kernels_ms']['field_reduce'], d['e2e']['value']/1e6, d.get('config_batch_4096',{}).get('ms_per_step')))
"; }
run "--grid 224" INF_PREFETCH=0
run "--grid 224" INF_PREFETCH=3
run "--grid 100" INF_PREFETCH=0
run "--grid 224 --table-dtype f16" INF_PREFETCH=0
run "--grid 100 --table-dtype f16" INF_PREFETCH=0
run "--grid 224 --mode infer" INF_PREFETCH=0
run "--grid 100 --mode infer" INF_PREFETCH=0
run "--grid 640 --batch 131072" INF_PREFETCH=0
