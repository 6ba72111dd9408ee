#!/bin/bash
# measurement helper: INF_DEBUG knob sweep (results with INF_DEBUG != 0 are numerically wrong on purpose)
run() { echo "== INF_DEBUG=$1 ${@:2}"; INF_DEBUG=$1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline ${@:2} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']; print('  rays/s %.1fM  step %.3f ms  main %.3f wgrad %.3f reduce %.3f' % (d['value']/1e6, d['ms_per_step'], r['step_kernels_ms']['field_main'], r['step_kernels_ms']['field_wgrad'], r['step_kernels_ms']['field_reduce']))
"; }
for d in 0 1 2 3 4 7; do run $d --grid 224; done
for d in 0 3 7; do run $d --grid 100 --table-dtype f16; done
run 0 --grid 100 --table-dtype f16 --mode infer
