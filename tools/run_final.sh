#!/bin/bash
# Dev helper (GPU box): the whole GPU suite, smoke, the default bench, then the UEA launch list and set-full capture for profiles/.
python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; tail -1 gpurun_out/final_bench.err
python bench.py --impl reference > gpurun_out/final_ref.log 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.log
python tools/one_step.py uea 3 > gpurun_out/final_uea.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/uea_launches.csv python tools/one_step.py uea 3 > gpurun_out/uea_ncu1.log 2>&1
python tools/one_step.py uea 1 > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -c 60 -o gpurun_out/uea_full python tools/one_step.py uea 1 > gpurun_out/uea_ncu2.log 2>&1; tail -1 gpurun_out/uea_ncu2.log
