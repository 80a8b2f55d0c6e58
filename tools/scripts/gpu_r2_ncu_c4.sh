set -x
O=gpurun_out/r2_ncu_c4
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_big -s 2 -c 1 -o $O/prof_shake python bench.py --workload c4 --steps 2 --warmup 1 --no-configs --no-e2e > $O/ncu.log 2>&1; tail -3 $O/ncu.log
