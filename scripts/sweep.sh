#!/bin/bash
# usage: scripts/sweep.sh  -- prints kernel times for a few configurations (measurement helper)
run() { echo "== $*"; python bench.py --steps 8 --warmup 3 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']; print('  rays/s %.1fM  step %.3f ms  main %.3f wgrad %.3f reduce %.3f  e2e %.1fM  b4096 %s' % (d['value']/1e6, d['ms_per_step'], r['step_kernels_ms']['field_main'], r['step_kernels_ms']['field_wgrad'], r['step_kernels_ms']['field_reduce'], d['e2e']['value']/1e6, d.get('config_batch_4096',{}).get('ms_per_step')))
"; }
run --grid 224
run --grid 100
run --grid 224 --table-dtype f16
run --grid 100 --table-dtype f16
run --grid 224 --mode infer
run --grid 100 --mode infer
run --grid 224 --mode infer --table-dtype f16
