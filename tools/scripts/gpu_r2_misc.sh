set -x
O=gpurun_out/r2_misc
mkdir -p $O
for lib in "" _old "" _old "" _old; do
  GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200$lib.so python bench.py --workload c4 --steps 40 --warmup 5 --no-configs --no-e2e > $O/c4$lib.json 2> $O/c4$lib.err
  python - $O/c4$lib.json "c4 lib$lib" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("AB", sys.argv[2], r["kernel_ms"], r["frac"])
PY
done
