run() { printf "%-48s" "$1"; env $1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-e2e "${@:2}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['step_kernels_ms']
print('step %.4f ms  main %.4f' % (d['ms_per_step'], k['field_main']))"; }
for knobs in "INF_CLUSTER=1" "INF_CLUSTER=1 INF_DEBUG=16" "INF_CLUSTER=1 INF_DEBUG=32" "INF_CLUSTER=1 INF_DEBUG=48"; do run "$knobs" "$@"; done
