set -x
O=gpurun_out/r2_call14
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  " $O/pytest_gpu.log | tail -6
for w in metric c2 c3; do for pref in 0 5; do
  timeout 300 python bench.py --workload $w --steps 50 --no-cpu --no-configs --no-e2e --kernel-pref $pref > $O/bench_${w}_pref$pref.json 2>> $O/err.log
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_call14/bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, r['kernel'], 'kernel_ms %.4f frac %.3f tf %.1f'%(r['kernel_ms'], r['frac'], r['fp32_tflops_5nlogn']))
    except Exception as e: print(f,'ERR',e)
PY
