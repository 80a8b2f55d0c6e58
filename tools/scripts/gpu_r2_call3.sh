# round 2, call 3: tests, the restructured default bench line (configs block, strong C4, per-call, copy ceiling),
# the warp-per-frame family re-measured, C2 with band-only averaged rows
set -x
O=gpurun_out/r2_call3
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED" $O/pytest_gpu.log | tail -12
(time timeout 900 python bench.py) > $O/bench_default.json 2> $O/bench_default.err; tail -4 $O/bench_default.err
timeout 300 python bench.py --workload c2 --steps 50 --no-cpu --no-configs > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --steps 50 --no-cpu --no-configs --no-e2e --kernel-pref 3 > $O/bench_wpf.json 2> $O/bench_wpf.err
timeout 300 python bench.py --steps 50 --no-cpu --no-configs --no-e2e --kernel-pref 4 > $O/bench_pair.json 2> $O/bench_pair.err
