#!/bin/bash
# Dev helper (GPU box): default bench, then the UEA launch list and set-full capture (whole step, no source) for profiles/.
python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; tail -1 gpurun_out/final_bench.err
python tools/one_step.py uea 3 > gpurun_out/final_uea.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/uea_launches.csv python tools/one_step.py uea 3 > gpurun_out/uea_ncu1.log 2>&1
python tools/one_step.py uea 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none -c 40 -o gpurun_out/uea_full python tools/one_step.py uea 1 > gpurun_out/uea_ncu2.log 2>&1; tail -1 gpurun_out/uea_ncu2.log
ls -la gpurun_out
