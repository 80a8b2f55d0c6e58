set -x
O=gpurun_out/r2_stagger3
mkdir -p $O
# metric ring kernel: delay = (arrival rank on the SM mod m) x cycles; argument = (m << 20) | cycles
for spec in "0 0" "2 4371" "2 3000" "2 5500" "3 2914" "3 2000" "6 1457" "6 1000" "2 2000" "0 0"; do
  set -- $spec
  v=$(( ($1 << 20) | $2 ))
  python bench.py --workload metric --steps 100 --warmup 5 --no-configs --no-e2e --stagger $v > $O/metric_m$1_c$2.json 2> $O/metric_m$1_c$2.err
  python - $O/metric_m$1_c$2.json "$spec" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("metric mod/cycles", sys.argv[2], r["kernel_ms"], r["frac"])
PY
done
