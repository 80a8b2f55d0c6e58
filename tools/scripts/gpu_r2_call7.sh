# round 2, call 7: tests (FFTW layout library, big multitaper kernel, AGC), C5 / default benches
set -x
O=gpurun_out/r2_call7
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -12
for pref in 0 1; do
timeout 300 python bench.py --workload c5 --steps 10 --no-cpu --no-configs --no-e2e --kernel-pref $pref > $O/bench_c5_pref$pref.json 2> $O/bench_c5_pref$pref.err
done
timeout 900 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_call7/bench_c5*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, r['kernel'], 'kernel_ms %.4f frac %.3f tflops %.1f'%(r['kernel_ms'], r['frac'], r['fp32_tflops_5nlogn']))
    except Exception as e: print(f,'ERR',e)
d=json.load(open('gpurun_out/r2_call7/bench_default.json'))
e=d['e2e']
for k in ('pcm16_in_u8_out','pcm16_in_u8_out_autoscale'): print(k, e[k]['value'], e[k]['ms_per_step'])
print({k:(v.get('kernel'),v.get('kernel_ms'),v.get('frac'),v.get('fp32_tflops_5nlogn')) for k,v in d['configs'].items()})
PY
