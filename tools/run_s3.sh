python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; tail -3 gpurun_out/s3_tests.log
( time python bench.py ) > gpurun_out/s3_bench.log 2> gpurun_out/s3_bench.err
python bench.py --workload sepsis_oi --no-other-configs --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/s3_oi.log 2> gpurun_out/s3_oi.err
python bench.py --workload uea --no-other-configs --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/s3_uea.log 2> gpurun_out/s3_uea.err
tail -3 gpurun_out/s3_bench.err
