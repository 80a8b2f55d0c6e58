set -x
O=gpurun_out/r2_tmem
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "big or multitaper or mtm or c5 or golden or fuzz or random or shard" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for i in 1 2; do
  python bench.py --workload c5 --steps 20 --warmup 3 --no-configs --no-e2e > $O/c5.json 2> $O/c5.err
  python - $O/c5.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("AB c5", r["kernel_ms"], r.get("fp32_tflops_5nlogn"))
PY
done
