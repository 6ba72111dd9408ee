#!/bin/bash
# usage: scripts/config_sweep.sh OUT.jsonl -- the occupancy curve of config 2 (B = 4096 ... 1 048 576 rays per step) and the
# gather-bound config 4 (V ~ 400 k rows, 1.64 GB fp32 table): one line per run = {"args": ..., + bench.py's profiling JSON}
# (resident inputs, no e2e arm -- these are kernel measurements, not bench results).
out=${1:-gpurun_out/config_sweep.jsonl}
: > "$out"
run() {
  echo "== $*" >&2
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --skip-e2e "$@" 2>/dev/null | tail -1 | ARGS="$*" python -c "
import json, os, sys
d = json.loads(sys.stdin.read()); d['args'] = os.environ['ARGS']; print(json.dumps(d))" >> "$out"
}
for b in 4096 32768 131072 262144 1048576; do run --batch $b; done
for b in 131072 1048576; do run --grid 632 --batch $b; done
run --grid 632 --batch 131072 --mode infer
run --batch 131072 --mode infer
run --batch 1048576 --mode infer
run --batch 131072 --path ray
run --batch 131072 --table-dtype f16
python - "$out" <<'PY'
import json, re, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    m = re.search(r"--batch (\d+)", d["args"])
    b = int(m.group(1)) if m else 131072
    k = d["step_kernels_ms"]
    print("%-44s %-6s %8.1f M rays/s  %.4f ms/step  project %.3f  rays/main %.3f  gather %.3f  wgrad %.3f  reduce %.3f" % (
        d["args"], "vertex" if d.get("vertex_path") else "ray", b / d["ms_per_step"] / 1e3, d["ms_per_step"], k["vertex_project"],
        k["rays_fwd_bwd"], k["vertex_gather"], k["wgrad"], k["reduce_adam"]))
PY
