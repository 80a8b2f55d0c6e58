# launch lists (per-kernel device time, ncu --metrics gpu__time_duration.sum) of the workloads whose
# step has more than one kernel; shares only (cold-cache, serialised)
set -x
O=gpurun_out/${TAG:-lists}; mkdir -p $O
for w in c2 lmp odd; do
  CMD="python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu"
  $CMD > $O/plain_$w.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_$w.csv $CMD > $O/ncu_$w.log 2>&1
done
ls -la $O
