#!/bin/bash
# Dev helper (GPU box): lane engine with tiles of 8 samples forced on the small batches (tcgen05 weight-gradient kernels) against tiles of 1.
for w in uea stock; do
  for t in 8; do
    ANCDE_ENGINE=4 ANCDE_SAMPLES_PER_CTA=$t timeout 300 python bench.py --workload $w --no-other-configs --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/ts_${w}_$t.log 2> gpurun_out/ts_${w}_$t.err
    echo "$w tiles of $t"; python tools/show_bench.py gpurun_out/ts_${w}_$t.log 2>/dev/null | head -10; tail -1 gpurun_out/ts_${w}_$t.err
  done
done
