set -x
O=gpurun_out/r2_ahead
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "ahead" > $O/pytest_ahead.log 2>&1; tail -8 $O/pytest_ahead.log
for pref in 0 7 0 7; do
  for w in metric c2; do
  python bench.py --workload $w --steps 100 --warmup 5 --no-configs --no-e2e --kernel-pref $pref > $O/${w}_pref$pref.json 2> $O/${w}_pref$pref.err
  python - $O/${w}_pref$pref.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print(sys.argv[1], r["kernel"][:60], r["kernel_ms"], r["frac"])
PY
  done
done
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
