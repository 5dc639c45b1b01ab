#!/bin/bash
# Dev helper (GPU box): the whole GPU suite, smoke, and the default bench.
python -m pytest tests -m gpu -q > gpurun_out/full_tests.log 2>&1; tail -3 gpurun_out/full_tests.log
python __graft_entry__.py smoke > gpurun_out/full_smoke.log 2>&1; tail -1 gpurun_out/full_smoke.log
python bench.py > gpurun_out/full_bench.log 2> gpurun_out/full_bench.err; tail -2 gpurun_out/full_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/full_bench.log').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['roofline']['frac'], d['roofline']['step_frac'])
for o in d['other_configs']: print(' ', o['workload'], o.get('adjoint'), round(o['value']), round(o['ms_per_step'],3), o.get('scaling'))
PY
