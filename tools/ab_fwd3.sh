timeout 900 python -m pytest tests/test_gpu_ncde.py tests/test_gpu_bwd4.py tests/test_gpu_umma.py -m gpu -x -q 2>&1 | tail -4
for v in "A=1"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-other-configs 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], [(k['name'],k['avg_ms']) for k in d['kernels'][:8]])"
done
