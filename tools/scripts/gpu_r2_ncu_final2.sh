set -x
O=gpurun_out/r2_ncu_final
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_ring -s 2 -c 1 -o $O/prof_metric python bench.py --workload metric --steps 2 --warmup 1 --no-configs --no-e2e > $O/ncu_metric.log 2>&1; tail -2 $O/ncu_metric.log | cut -c1-200
