#!/bin/bash
# Dev helper (GPU box): the default bench line, summarised.
python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'h2d', d['e2e']['h2d_bytes_per_step'], 'frac', d['roofline']['frac'], d['roofline']['step_frac'])
PY
