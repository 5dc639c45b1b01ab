#!/bin/bash
# Dev helper run on the GPU box: parity tests, a short bench, then the per-phase cycle breakdown.
tag=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/${tag}_bench.log").read().strip().splitlines()[-1])
    print("bench", round(j["value"]), "samples/s", round(j["ms_per_step"],3), "ms/step  e2e", round(j["e2e"]["value"]))
    for k in j["kernels"]: print("   ", k["name"], k["avg_ms"])
except Exception as e:
    print("bench ERR", e); print(open("gpurun_out/${tag}_bench.err").read()[-2000:])
PY
if [ "$2" == "phase" ]; then
  python ancde_b200/build.py --force -DANCDE_PHASE_TIMING > /dev/null 2>&1
  python tools/phase4.py sepsis 2>&1 | tail -24
fi
