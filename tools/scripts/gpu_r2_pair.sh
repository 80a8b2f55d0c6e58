set -x
O=gpurun_out/r2_pair
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "big or shard or smoke or 24" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for pref in 0 6 0 6; do
  python bench.py --workload c4 --steps 40 --warmup 5 --no-configs --kernel-pref $pref > $O/c4_pref$pref.json 2> $O/c4_pref$pref.err
  python - $O/c4_pref$pref.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print(sys.argv[1], r["kernel"], r["kernel_ms"], r["frac"])
PY
done
