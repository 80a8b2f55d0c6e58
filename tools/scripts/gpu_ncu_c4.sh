set -x
O=gpurun_out/r41; mkdir -p $O
CMD="python bench.py --workload c4 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gram_ -s 3 -c 1 -o $O/prof_c4 $CMD > $O/ncu.log 2>&1
ls -la $O
