#!/bin/bash
# Dev helper (run on the GPU box): parity tests for the NCDE path + a short bench summary.
timeout 300 python -m pytest tests/test_gpu_ncde.py -m gpu -q -x 2>&1 | tail -4
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" 2>gpurun_out/bench_q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('samples/s %.0f  ms/step %.2f  step_frac %.4f  e2e %.0f' % (d['value'], d['ms_per_step'], d['roofline']['step_frac'], d['e2e']['value']))
tot=0
for k in d['kernels']:
    tot+=k['avg_ms']*k['launches']/5
    print('  %-22s %8.4f ms  %s' % (k['name'], k['avg_ms'], k.get('frac_of_fp32_peak')))
print('  kernel ms per step %.2f' % tot)
"
tail -3 gpurun_out/bench_q.err
