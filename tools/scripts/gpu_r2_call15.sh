set -x
O=gpurun_out/r2_call15
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  " $O/pytest_gpu.log | tail -6
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --kernel-pref 5 --seconds 600"
$CMD > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_big -s 3 -c 1 -o $O/prof_big_n4096 $CMD > $O/ncu.log 2>&1
tail -2 $O/ncu.log
