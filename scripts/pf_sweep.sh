#!/bin/bash
run() { echo "== INF_KEEP_MB=$1 ${@:2}"; INF_KEEP_MB=$1 python bench.py --steps 20 --warmup 3 --skip-e2e ${@:2} 2>&1 | tail -1; }
for k in 0 40 80 120 1000; do run $k --grid 224; done
for k in 0 80 1000; do run $k --grid 640; done
for k in 0 80 1000; do run $k --grid 224 --mode infer; done
