#!/bin/bash
# Dev helper (GPU box): the whole GPU suite, the default bench, and the lane engine forced on Mujoco (B = 1024).
python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1; tail -3 gpurun_out/full_tests.log
python bench.py > gpurun_out/full_bench.log 2> gpurun_out/full_bench.err; tail -2 gpurun_out/full_bench.err
python tools/show_bench.py gpurun_out/full_bench.log | head -3
for e in 0 4; do
  ANCDE_ENGINE=$e python bench.py --workload mujoco --no-other-configs --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/mj_$e.log 2> gpurun_out/mj_$e.err
  echo "mujoco ANCDE_ENGINE=$e"; python tools/show_bench.py gpurun_out/mj_$e.log 2>&1 | head -5; tail -2 gpurun_out/mj_$e.err
done
