set -x
O=gpurun_out/r2_ring2b
mkdir -p $O
for spec in "0 0" "0 1" "7 0" "0 0" "7 0"; do
  set -- $spec
  for w in metric; do
  python bench.py --workload $w --steps 100 --warmup 5 --no-configs --no-e2e --kernel-pref $1 --stagger $2 > $O/${w}_p$1_s$2.json 2> $O/${w}_p$1_s$2.err
  python - $O/${w}_p$1_s$2.json "$w pref $1 stagger $2" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("AB", sys.argv[2], r["kernel_ms"], r["frac"])
PY
  done
done
for s in 0 1 0 1; do
  python bench.py --workload c4 --steps 40 --warmup 5 --no-configs --no-e2e --stagger $s > $O/c4_s$s.json 2> $O/c4_s$s.err
  python - $O/c4_s$s.json "c4 stagger $s" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("AB", sys.argv[2], r["kernel_ms"], r["frac"])
PY
done
