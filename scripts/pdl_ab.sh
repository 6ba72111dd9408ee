#!/bin/bash
# measurement helper: programmatic dependent launch on / off -- whole-step times without per-stage events (bench.py's
# default resident timing), training and inference, plus the shipped 4096-ray batch
for pdl in ${PDL_MODES:-0 2 1}; do
  for mode in train infer; do
    echo "== INF_PDL=$pdl $mode"
    INF_PDL=$pdl INF_RAYS_SLOTS=1 timeout 300 python bench.py --mode $mode --path vertex --steps 50 --warmup 5 --skip-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['ms_per_step'], 4), round(d.get('rays_per_sec', d.get('value', 0)) / 1e6, 1), 'M rays/s')"
  done
  INF_PDL=$pdl python scripts/small_batch.py 4096 2>&1 | tail -1
done
