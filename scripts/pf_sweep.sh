#!/bin/bash
run() { echo "== INF_CLUSTER=$1 ${@:2}"; INF_CLUSTER=$1 python bench.py --steps 20 --warmup 3 --skip-e2e ${@:2} 2>&1 | tail -1; }
for c in 1 2; do run $c --grid 224; run $c --grid 224 --mode infer; run $c --grid 224 --table-dtype f16; run $c --grid 224 --table-dtype f16 --mode infer; run $c --grid 100 --mode infer; run $c --grid 100 --table-dtype f16 --mode infer; done
