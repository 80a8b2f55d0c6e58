set -x
O=gpurun_out/r2_stage
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 600 python tools/scripts/e2e_c2_probe.py > $O/c2_probe.log 2>&1; cat $O/c2_probe.log
python bench.py --workload c2 --steps 20 --warmup 3 --no-configs > $O/c2.json 2> $O/c2.err
python - $O/c2.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("fraction_of_copy_ceiling"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lmp_kernel -s 1 -c 1 -o $O/prof_lmp python bench.py --workload lmp --steps 2 --warmup 1 --no-configs --no-e2e > $O/ncu.log 2>&1; tail -3 $O/ncu.log
