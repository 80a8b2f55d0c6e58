set -x
O=gpurun_out/r2_stagger
mkdir -p $O
for s in 0 400 800 1200 1457 2000 3000 4371 0; do
  python bench.py --workload metric --steps 100 --warmup 5 --no-configs --no-e2e --stagger $s > $O/metric_s$s.json 2> $O/metric_s$s.err
  python - $O/metric_s$s.json $s <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("metric stagger", sys.argv[2], r["kernel_ms"], r["frac"])
PY
done
for s in 0 6000 6500 7000 7500 8000 9000 0; do
  python bench.py --workload c4 --steps 40 --warmup 5 --no-configs --no-e2e --stagger $s > $O/c4_s$s.json 2> $O/c4_s$s.err
  python - $O/c4_s$s.json $s <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("c4 stagger", sys.argv[2], r["kernel_ms"], r["frac"])
PY
done
