#!/bin/bash
# measurement helper: training step under INF_PREFETCH=0/2 at three table sizes
q() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d.get('step_kernels_ms') or d['roofline']['step_kernels_ms']
print('step %.4f ms  main %.4f  wgrad %.4f' % (d['ms_per_step'], k['field_main'], k['field_wgrad']))"; }
for pf in 0 2; do for g in 100 224 632; do echo -n "train prefetch $pf grid $g: "; INF_PREFETCH=$pf python bench.py --warmup 5 --no-cpu-baseline --steps 40 --skip-e2e --grid $g "$@" 2>/dev/null | tail -1 | q; done; done
