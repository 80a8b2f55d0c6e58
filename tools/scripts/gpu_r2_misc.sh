set -x
O=gpurun_out/r2_misc
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "wav or headless" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py --no-configs --no-e2e --steps 50 > $O/b.json 2> $O/b.err; tail -3 $O/b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_misc/b.json')); print(json.dumps(d['per_call'], indent=1))
PY
