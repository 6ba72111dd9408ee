#!/bin/bash
source <(sed -n '1,12p' scripts/sweep.sh)
run "--grid 224"
run "--grid 100"
run "--grid 224 --table-dtype f16"
run "--grid 224 --mode infer"
run "--grid 100 --mode infer"
