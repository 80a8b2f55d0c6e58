set -x
O=gpurun_out/r2_ramp
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
(time python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -4 $O/bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_ramp/bench_default.json'))
e=d['e2e']
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',e['value'], e.get('fraction_of_copy_ceiling'))
for k in ['pcm16_input','pcm16_in_u8_out','pcm16_in_u8_out_autoscale']:
    print(k, e[k]['value'], e[k]['ms_per_step'])
print({k:(round(v.get('value',0)), round(v.get('kernel_ms',0),3), round(v.get('post_kernels_ms',0),3), round(v.get('e2e',{}).get('value',0))) for k,v in d['configs'].items()})
print(d['per_call'])
PY
