# round 2, call 1: the whole GPU test suite (with the new parity tests), memcheck over every kernel
# family, and this round's baseline numbers of the round-1 kernels on today's box
set -x
O=gpurun_out/r2_call1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu.txt
nproc > $O/nproc.txt; free -g >> $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
timeout 300 python tools/sanitize_driver.py > $O/sanitize_plain.log 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_driver.py --quick > $O/memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/memcheck.log
tail -5 $O/memcheck.log
for w in metric c2 c4; do
  timeout 600 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err
done
