set -x
O=gpurun_out/r2_tmem
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or strict or subrange or windows or fuzz or random or shard or smoke" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for lib in "" _tt0 _tt8 "" _tt0 _tt8; do
  for w in metric c1 c2; do
  GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200$lib.so python bench.py --workload $w --steps 100 --warmup 5 --no-configs --no-e2e > $O/$w$lib.json 2> $O/$w$lib.err
  python - $O/$w$lib.json "$w lib$lib" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("AB", sys.argv[2], r["kernel_ms"], r["frac"])
PY
  done
done
