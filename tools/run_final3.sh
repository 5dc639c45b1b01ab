#!/bin/bash
# Dev helper (GPU box): default bench; UEA launch list + set-full capture summarised ON THE BOX (the report itself is > 64 MiB).
python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; tail -1 gpurun_out/final_bench.err
python tools/one_step.py uea 3 > gpurun_out/final_uea.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/uea_launches.csv python tools/one_step.py uea 3 > gpurun_out/uea_ncu1.log 2>&1
python tools/one_step.py uea 1 > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -c 30 -o /tmp/uea_full python tools/one_step.py uea 1 > gpurun_out/uea_ncu2.log 2>&1; tail -1 gpurun_out/uea_ncu2.log
python tools/make_profiles.py r02_lane gpurun_out/uea_launches.csv /tmp/uea_full.ncu-rep "python tools/one_step.py uea 3 / 1   (eager training steps, UEA CharacterTrajectories shape, B = 32, 724 stages per network: the lane engine)"
cp profiles/r02_lane_* gpurun_out/
ls -la gpurun_out | tail -8
