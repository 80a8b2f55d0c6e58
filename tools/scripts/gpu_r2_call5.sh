# round 2, call 5: tests (radix-select floor statistics, AGC chain, fft_do_batch), default bench, and one
# ncu --set full capture of the 32-points-per-thread kernel on the C4 workload
set -x
O=gpurun_out/r2_call5
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -12
timeout 900 python bench.py --no-cpu > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
CMD="python bench.py --workload c4 --seconds 1800 --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs"
$CMD > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_big -s 3 -c 1 -o $O/prof_big $CMD > $O/ncu.log 2>&1
tail -3 $O/ncu.log
