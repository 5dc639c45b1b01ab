#!/bin/bash
# Dev helper (GPU box): whole GPU suite, smoke, default bench; then the Mujoco (B = 1024, lane engine, tiles of 8) launch list and a
# set-full capture of its lane kernels, summarised ON THE BOX (the reports exceed the 64 MiB that travel back).
python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1; tail -2 gpurun_out/final_tests.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.log 2> gpurun_out/final_bench.err; tail -1 gpurun_out/final_bench.err
python tools/one_step.py mujoco 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/mujoco_launches.csv python tools/one_step.py mujoco 3 > gpurun_out/mj_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:ncde_lane -c 4 -o /tmp/mj_full python tools/one_step.py mujoco 1 > gpurun_out/mj_ncu2.log 2>&1; tail -1 gpurun_out/mj_ncu2.log
python tools/make_profiles.py r02_lane_mujoco gpurun_out/mujoco_launches.csv /tmp/mj_full.ncu-rep "python tools/one_step.py mujoco 3 / 1   (eager training steps, Mujoco shape, B = 1024: the lane engine with tiles of 8 samples)" 2>&1 | tail -1
cp profiles/r02_lane_mujoco_* gpurun_out/ 2>/dev/null
ls gpurun_out | head -20
