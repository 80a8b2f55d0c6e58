set -x
O=gpurun_out/r2_lmp
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "lmp or random_configuration or headless" > $O/pytest.log 2>&1; tail -15 $O/pytest.log
python bench.py --workload lmp --steps 20 --warmup 3 --no-configs --no-e2e > $O/lmp.json 2> $O/lmp.err
python - $O/lmp.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print(d["value"], d["ms_per_step"], r["kernel_ms"], r.get("post_kernels_ms"))
PY
