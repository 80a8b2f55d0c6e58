# round 2, call 12: ncu --set full of the 32-point kernel with the TMA landing (C4 workload)
set -x
O=gpurun_out/r2_call12
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -k "fused_averaging or big_frame" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED" $O/pytest_gpu.log | tail -3
CMD="python bench.py --workload c4 --seconds 1800 --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs"
$CMD > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_big -s 3 -c 1 -o $O/prof_big_tma $CMD > $O/ncu.log 2>&1
tail -2 $O/ncu.log
