# round 2, call 2: GPU test suite with the display pinning / fused levels, bench with the u8 e2e leg
set -x
O=gpurun_out/r2_call2
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -40 $O/pytest_gpu.log
for w in metric c1 c4; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err
  tail -2 $O/bench_$w.err
done
