# round 2, call 8: zero-copy per-call path, LMP kernel, averaging kernel: tests + default bench (with per_call)
set -x
O=gpurun_out/r2_call8
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -12
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_call8/bench_default.json'))
print('per_call', d.get('per_call'))
print({k:(v.get('kernel_ms'),v.get('post_kernels_ms'),v.get('frac')) for k,v in d['configs'].items()})
print('cpu', d.get('cpu_baseline',{}).get('value'))
PY
