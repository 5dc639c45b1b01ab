#!/bin/bash
# Dev helper (GPU box): lane-engine parity tests, then the small-network benches at their BASELINE batch sizes.
timeout 900 python -m pytest tests/test_gpu_lane.py -x -q > gpurun_out/lane_tests.log 2>&1; tail -2 gpurun_out/lane_tests.log
for w in uea stock mujoco; do
  timeout 300 python bench.py --workload $w --no-other-configs --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/lane_$w.log 2> gpurun_out/lane_$w.err
  python tools/show_bench.py gpurun_out/lane_$w.log 2>&1 | head -5
done
