set -x
O=gpurun_out/r2_call17
mkdir -p $O
for v in default ns4; do
  if [ $v = default ]; then unset GLFER_B200_LIB; else export GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200_$v.so; fi
  timeout 300 python bench.py --steps 20 --no-cpu --no-configs > $O/bench_$v.json 2>> $O/err.log
done
unset GLFER_B200_LIB
python - <<'PY'
import json
for v in ('default','ns4'):
    d=json.load(open('gpurun_out/r2_call17/bench_%s.json'%v)); e=d['e2e']
    print(v, 'float %.4g pcm %.4g u8 %.4g auto %.4g'%(e['value'], e['pcm16_input']['value'], e['pcm16_in_u8_out']['value'], e['pcm16_in_u8_out_autoscale']['value']))
PY
