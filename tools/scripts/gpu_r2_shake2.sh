set -x
O=gpurun_out/r2_shake2
mkdir -p $O
for s in 0 2 3 4 8000 0 2 3 4; do
  python bench.py --workload c4 --steps 40 --warmup 5 --no-configs --no-e2e --stagger $s > $O/c4_s$s.json 2> $O/c4_s$s.err
  python - $O/c4_s$s.json $s <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("c4 stagger", sys.argv[2], r["kernel_ms"], r["frac"])
PY
done
