set -x
O=gpurun_out/r2_ring2
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "antiphase" > $O/pytest_ring2.log 2>&1; tail -8 $O/pytest_ring2.log
for spec in "0 0" "0 1" "7 0" "0 0" "0 1" "7 0"; do
  set -- $spec
  for w in metric c2; do
  python bench.py --workload $w --steps 100 --warmup 5 --no-configs --no-e2e --kernel-pref $1 --stagger $2 > $O/${w}_p$1_s$2.json 2> $O/${w}_p$1_s$2.err
  python - $O/${w}_p$1_s$2.json "$w pref $1 stagger $2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("AB", sys.argv[2], r["kernel_ms"], r["frac"])
PY
  done
done
