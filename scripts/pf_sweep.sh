#!/bin/bash
run() { echo "== INF_PREFETCH=$1 ${@:2}"; INF_PREFETCH=$1 python bench.py --steps 8 --warmup 3 --skip-e2e ${@:2} 2>&1 | tail -1; }
for pf in 0 2 4 8; do run $pf --grid 224; run $pf --grid 224 --mode infer; done
run 4 --grid 224 --table-dtype f16; run 0 --grid 224 --table-dtype f16; run 4 --grid 224 --table-dtype f16 --mode infer
run 4 --grid 640; run 0 --grid 640
