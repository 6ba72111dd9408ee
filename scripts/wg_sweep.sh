#!/bin/bash
# measurement helper: field_wgrad under its launch knobs on the vertex path (per-stage device times of the training step)
for knob in ${KNOBS:-"INF_WGRAD_BIGW=0.85" "INF_WGRAD_BIGW=1.0" "INF_WGRAD_BIGW=1.15" "INF_WGRAD_BIGW=1.3" "INF_WGRAD_CLUSTER=4" "INF_WGRAD_CLUSTER=1"}; do
  echo "== $knob"
  env $knob timeout 120 python bench.py --mode train --path vertex --steps 20 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['ms_per_step'], 4), {k: round(v * 1e3, 1) for k, v in d['step_kernels_ms'].items()})"
done
