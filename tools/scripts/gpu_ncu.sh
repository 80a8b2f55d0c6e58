# ncu capture of the metric kernel (run under gpurun): plain run first, then launch list, then --set full
set -x
TAG=${1:-r2a}
mkdir -p gpurun_out/$TAG
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/$TAG/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/$TAG/launches.csv $CMD > gpurun_out/$TAG/ncu1.log 2>&1
$CMD > gpurun_out/$TAG/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-gram_} -s 3 -c 1 -o gpurun_out/$TAG/prof $CMD > gpurun_out/$TAG/ncu2.log 2>&1
ls -la gpurun_out/$TAG
