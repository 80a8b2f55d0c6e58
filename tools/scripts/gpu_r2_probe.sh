set -x
O=gpurun_out/r2_probe
mkdir -p $O
timeout 600 python tools/scripts/e2e_float_probe.py > $O/float_probe.log 2>&1; cat $O/float_probe.log
