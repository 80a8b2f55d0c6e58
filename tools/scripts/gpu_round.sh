# one GPU round: parity tests on the default library, then the metric bench for every variant
# library listed in $VARIANTS (glfer_b200/libglfer_b200_<v>.so) and for the default library
set -x
TAG=${TAG:-r2}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$TAG/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$TAG/pytest_gpu.log
tail -3 gpurun_out/$TAG/pytest_gpu.log
for v in $VARIANTS; do
  GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200_$v.so python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu > gpurun_out/$TAG/bench_$v.json 2> gpurun_out/$TAG/bench_$v.err
done
python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu > gpurun_out/$TAG/bench_default.json 2> gpurun_out/$TAG/bench_default.err
python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu > gpurun_out/$TAG/bench_default2.json 2> gpurun_out/$TAG/bench_default2.err
for w in ${WORKLOADS:-c2 c3 c4 c5}; do
  python bench.py --workload $w --steps 50 --warmup 3 --no-e2e --no-cpu > gpurun_out/$TAG/bench_$w.json 2> gpurun_out/$TAG/bench_$w.err
done
for f in gpurun_out/$TAG/bench_*.json; do echo $f; grep -h -o '"kernel_ms": [0-9.]*\|"frac": [0-9.]*' $f | tr '\n' ' '; echo; done
