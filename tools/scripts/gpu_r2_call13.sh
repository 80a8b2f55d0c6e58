set -x
O=gpurun_out/r2_call13
mkdir -p $O
for v in default notw2; do
  if [ $v = default ]; then unset GLFER_B200_LIB; else export GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200_$v.so; fi
  for i in 1 2; do
  timeout 300 python bench.py --workload c4 --steps 50 --no-cpu --no-configs --no-e2e > $O/bench_c4_${v}_$i.json 2> $O/err.log
  done
  timeout 300 python bench.py --workload c5 --steps 10 --no-cpu --no-configs --no-e2e > $O/bench_c5_$v.json 2>> $O/err.log
done
unset GLFER_B200_LIB
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_call13/bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, r['kernel'], 'kernel_ms %.4f frac %.3f'%(r['kernel_ms'], r['frac']))
    except Exception as e: print(f,'ERR',e)
PY
