# round 2, call 4: the 32-points-per-thread kernel: tests, C4 / N=32768 benches against the ring kernel
set -x
O=gpurun_out/r2_call4
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -12
for pref in 0 2; do
  timeout 300 python bench.py --workload c4 --steps 50 --no-cpu --no-configs --no-e2e --kernel-pref $pref > $O/bench_c4_pref$pref.json 2> $O/bench_c4_pref$pref.err
done
timeout 300 python bench.py --workload c4 --steps 50 --no-cpu --no-configs --no-e2e --no-submean > $O/bench_c4_nosub.json 2> $O/bench_c4_nosub.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_call4/bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, r['kernel'], 'kernel_ms %.4f frac %.3f'%(r['kernel_ms'], r['frac']))
    except Exception as e: print(f,'ERR',e)
PY
