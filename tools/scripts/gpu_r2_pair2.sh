set -x
O=gpurun_out/r2_pair2
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "big or shard or smoke or 24 or golden" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for w in c4 c5; do
  python bench.py --workload $w --steps 30 --warmup 5 --no-configs > $O/$w.json 2> $O/$w.err
  python - $O/$w.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print(sys.argv[1], r["kernel"], r["kernel_ms"], r["frac"], r.get("fp32_tflops_5nlogn"))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_big -s 2 -c 1 -o $O/prof_pair python bench.py --workload c4 --steps 2 --warmup 1 --no-configs --no-e2e > $O/ncu.log 2>&1; tail -3 $O/ncu.log
timeout 600 python tools/scripts/e2e_c2_probe.py > $O/c2_probe.log 2>&1; cat $O/c2_probe.log
