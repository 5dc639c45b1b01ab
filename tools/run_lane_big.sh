#!/bin/bash
# Dev helper (GPU box): lane tests, then the automatic choice (lane, tiles of 8) against the streamed engine at B = 1024 on the small shapes.
timeout 900 python -m pytest tests/test_gpu_lane.py tests/test_gpu_umma.py -x -q 2>&1 | tail -3
for w in mujoco stock uea; do
  for e in 0 1; do
    ANCDE_ENGINE=$e timeout 300 python bench.py --workload $w --batch 1024 --no-other-configs --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/big_${w}_$e.log 2> gpurun_out/big_${w}_$e.err
    echo "$w B=1024 ANCDE_ENGINE=$e"; python tools/show_bench.py gpurun_out/big_${w}_$e.log 2>/dev/null | head -9; tail -1 gpurun_out/big_${w}_$e.err
  done
done
