# final measurement of round 2 on one GPU: tests, smoke, the driver's bench line (headline + configs block +
# per-call + CPU baselines), the reference arm, every workload on its own, launch list of the bench command
set -x
TAG=${TAG:-r2_final}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt
nproc >> $O/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
(time python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -4 $O/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for w in c1 c2 c3 c4 c5 lmp odd; do
  python bench.py --workload $w --steps 30 --warmup 3 --no-configs > $O/bench_$w.json 2> $O/bench_$w.err
done
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
r=d.get("roofline",{})
print(" value=%.4g ms=%.4g kernel=%s kernel_ms=%s frac=%s e2e=%s cpu=%s" % (d["value"], d["ms_per_step"], r.get("kernel"), r.get("kernel_ms"), r.get("frac"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value")))
PY
done
