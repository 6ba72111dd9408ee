#!/bin/bash
# usage: scripts/variant_sweep.sh [bench args] -- times field_main for every build/variants/libinf_field_v*.so (measurement helper;
# overwrites the in-tree library of THIS checkout copy, so run it on the scratch copy gpurun makes)
for lib in build/variants/libinf_field_v*.so; do
  cp "$lib" intrinsic-neural-fields_b200/libinf_field.so
  printf "%-44s" "$(basename $lib)"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['step_kernels_ms']
print('step %.4f ms  main %.4f' % (d['ms_per_step'], k['field_main']))"
done
