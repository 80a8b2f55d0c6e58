# round 2, call 9: the CHECKED library (device-side bounds assertions on every exchange index, bulk-copy source
# and row store; compute-sanitizer is closed on this pool) over every kernel family and the whole parity suite;
# LMP bench
set -x
O=gpurun_out/r2_call9
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" $O/pytest_gpu.log | tail -5
export GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200_checked.so
timeout 600 python tools/sanitize_driver.py > $O/checked_driver.log 2>&1; echo "checked driver rc=$?" >> $O/checked_driver.log
tail -4 $O/checked_driver.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_gpu_fuzz.py -m gpu -q > $O/checked_pytest.log 2>&1; echo "checked pytest rc=$?" >> $O/checked_pytest.log
grep -E "passed|failed|FAILED|GLB_CHECK" $O/checked_pytest.log | tail -8
unset GLFER_B200_LIB
timeout 300 python bench.py --workload lmp --steps 20 --no-cpu --no-configs --no-e2e > $O/bench_lmp.json 2> $O/bench_lmp.err
python -c "
import json; d=json.load(open('$O/bench_lmp.json')); r=d['roofline']; print('lmp kernel_ms', r['kernel_ms'], 'post', r['post_kernels_ms'])"
