#!/bin/bash
# round-2 measurement sweep (profiling runs: resident steps only, per-stage device times)
mkdir -p gpurun_out
for slots in 2 1; do
for mode in train infer; do
    echo "== $mode vertex slots=$slots"
    INF_RAYS_SLOTS=$slots timeout 300 python bench.py --mode $mode --path vertex --steps 20 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1
done
done
