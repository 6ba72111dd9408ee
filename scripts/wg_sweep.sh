#!/bin/bash
# measurement helper: wgrad cluster width (dz half-panels multicast within clusters of 1 / 2 / 4 big-slice CTAs)
q() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d.get('step_kernels_ms') or d['roofline']['step_kernels_ms']
print('step %.4f ms  main %.4f  wgrad %.4f' % (d['ms_per_step'], k['field_main'], k['field_wgrad']))"; }
for cw in 1 2 4; do echo -n "INF_WGRAD_CLUSTER=$cw: "; INF_WGRAD_CLUSTER=$cw python bench.py --warmup 5 --no-cpu-baseline --steps 40 --skip-e2e "$@" 2>/dev/null | tail -1 | q; done
