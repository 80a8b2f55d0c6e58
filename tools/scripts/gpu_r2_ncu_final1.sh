set -x
O=gpurun_out/r2_ncu_final
mkdir -p $O
python bench.py --steps 2 --warmup 1 > $O/b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1; tail -2 $O/ncu_launches.log | cut -c1-200
