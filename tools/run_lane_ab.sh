#!/bin/bash
# Dev helper (GPU box): A/B of the lane engine's unit-splitting threshold (LANE_SPLIT) on the small shapes.
timeout 600 python -m pytest tests/test_gpu_lane.py -x -q 2>&1 | tail -2
for v in 12 4 2; do
  python ancde_b200/build.py --force -DLANE_SPLIT=$v > /dev/null 2>&1
  for w in uea stock; do
    timeout 300 python bench.py --workload $w --no-other-configs --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ab_${w}_$v.log 2> gpurun_out/ab_${w}_$v.err
    echo "LANE_SPLIT=$v $w"; python tools/show_bench.py gpurun_out/ab_${w}_$v.log 2>&1 | head -5
  done
done
