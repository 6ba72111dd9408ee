#!/bin/bash
# measurement helper: stash bulk copies with / without the L2 evict_first hint (INF_DEBUG 64 = stores, 128 = wgrad loads)
q() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d.get('step_kernels_ms') or d['roofline']['step_kernels_ms']
print('step %.4f ms  main %.4f  wgrad %.4f' % (d['ms_per_step'], k['field_main'], k['field_wgrad']))"; }
for dbg in 0 64 128 192; do for g in 224 632; do echo -n "INF_DEBUG=$dbg grid $g: "; INF_DEBUG=$dbg python bench.py --warmup 5 --no-cpu-baseline --steps 40 --skip-e2e --grid $g "$@" 2>/dev/null | tail -1 | q; done; done
