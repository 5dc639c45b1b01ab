#!/bin/bash
# Dev helper (N-GPU box, NG=2|4|8, default 2): the NCCL sharded-gradient test (2 GPUs) and the weak-scaling bench line at NG GPUs.
[ "${NG:-2}" == "2" ] && timeout 600 python -m pytest tests/test_gpu_ncde.py -m gpu -q -k two_gpu 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus ${NG:-2} --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err
tail -2 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.log').read().strip().splitlines()[-1])
import os
print(os.environ.get('NG', '2'), 'GPUs:', round(d['value']), 'samples/s', round(d['ms_per_step'],3), 'ms/step; e2e', round(d['e2e']['value']))
for o in d.get('other_configs', []): print('  ', o['workload'], o.get('adjoint'), o.get('scaling'), round(o['value']), round(o['ms_per_step'],3))
PY
