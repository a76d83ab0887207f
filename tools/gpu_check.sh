#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_check.sh <tag> [pytest-args]
# runs the GPU parity tests, then the default bench, and leaves the JSON line in gpurun_out/bench_<tag>.json
tag=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -5
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.3f stages %s e2e %.4g"%(d["value"],d["ms_per_step"],d["roofline"]["stage_ms_per_step"],d["e2e"]["value"]))
PY
tail -3 gpurun_out/bench_${tag}.err
