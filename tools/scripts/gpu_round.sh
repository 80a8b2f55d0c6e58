set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log
tail -5 gpurun_out/r2/pytest_gpu.log
for v in pk1 pk2; do
  GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200_$v.so python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2/bench_$v.json 2> gpurun_out/r2/bench_$v.err
done
python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2/bench_pk3.json 2> gpurun_out/r2/bench_pk3.err
python bench.py --steps 200 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2/bench_pk3b.json 2> gpurun_out/r2/bench_pk3b.err
for w in c2 c3 c4 c5; do
  python bench.py --workload $w --steps 50 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2/bench_$w.json 2> gpurun_out/r2/bench_$w.err
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2 tools/microbench/ffma2.cu && /tmp/ffma2 > gpurun_out/r2/ffma2.txt 2>&1
grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"kernel_ms": [0-9.]*' gpurun_out/r2/bench_*.json
