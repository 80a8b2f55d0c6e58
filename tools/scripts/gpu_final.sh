# final measurement of a round on one GPU: tests, the driver's bench line (with e2e and the CPU
# baseline), the reference arm, the other workloads, launch list + one ncu --set full capture
set -x
TAG=${TAG:-final}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py > $O/bench_metric.json 2> $O/bench_metric.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
for w in c1 c2 c3 c4 c5 lmp; do
  python bench.py --workload $w --steps 50 --warmup 3 > $O/bench_$w.json 2> $O/bench_$w.err
done
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv $CMD > $O/ncu1.log 2>&1
$CMD > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_ -s 3 -c 1 -o $O/prof $CMD > $O/ncu2.log 2>&1
for f in $O/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
r=d.get("roofline",{})
print(" value=%.4g ms=%.4g kernel=%s kernel_ms=%s frac=%s e2e=%s cpu=%s" % (d["value"], d["ms_per_step"], r.get("kernel"), r.get("kernel_ms"), r.get("frac"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value")))
PY
done
