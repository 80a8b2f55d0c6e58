set -x
O=gpurun_out/r2_call11
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|^E  " $O/pytest_gpu.log | tail -12
timeout 300 python bench.py --workload c2 --steps 50 --no-cpu --no-configs > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --workload odd --steps 20 --no-cpu --no-configs --no-e2e > $O/bench_odd.json 2> $O/bench_odd.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_call11/bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, r['kernel'], 'kernel_ms %.4f frac %.3f step_ms %.4f post %.4f e2e %s'%(r['kernel_ms'], r['frac'], d['ms_per_step'], r['post_kernels_ms'], (d.get('e2e') or {}).get('value')))
    except Exception as e: print(f,'ERR',e)
PY
