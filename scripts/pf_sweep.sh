#!/bin/bash
run() { echo "== ${@}"; python bench.py --steps 20 --warmup 3 --skip-e2e ${@} 2>&1 | tail -1; }
run --grid 224 --mode infer; run --grid 224 --table-dtype f16 --mode infer; run --grid 100 --mode infer; run --grid 100 --table-dtype f16 --mode infer; run --grid 640 --mode infer
