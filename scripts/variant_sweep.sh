#!/bin/bash
# measurement helper: the training step (profiling run: per-stage device times) with each library variant in build/variants/
L=intrinsic-neural-fields_b200/libinf_field.so
cp $L /tmp/lib_default.so
for v in build/variants/*.so /tmp/lib_default.so; do
  cp $v $L
  for slots in 1; do
    echo "== $(basename $v) train vertex slots=$slots"
    INF_RAYS_SLOTS=$slots timeout 300 python bench.py --mode train --path vertex --steps 20 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['ms_per_step'], 4), {k: round(v * 1e3, 1) for k, v in d['step_kernels_ms'].items()})"
  done
done
cp /tmp/lib_default.so $L
