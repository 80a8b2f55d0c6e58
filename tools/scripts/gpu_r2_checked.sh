set -x
O=gpurun_out/r2_checked
mkdir -p $O
export GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200_checked.so
timeout 600 python tools/sanitize_driver.py > $O/checked_driver.log 2>&1; echo "checked driver rc=$?" >> $O/checked_driver.log
tail -4 $O/checked_driver.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_r2.py tests/test_gpu_fuzz.py -m gpu -q > $O/checked_pytest.log 2>&1; echo "checked pytest rc=$?" >> $O/checked_pytest.log
grep -E "passed|failed|FAILED|GLB_CHECK" $O/checked_pytest.log | tail -8
