#!/bin/bash
run() { echo "== INF_PREFETCH_WGRAD=$1 ${@:2}"; INF_PREFETCH_WGRAD=$1 python bench.py --steps 8 --warmup 3 --skip-e2e ${@:2} 2>&1 | tail -1; }
for pf in 0 3 6 12 24; do run $pf --grid 224; done
