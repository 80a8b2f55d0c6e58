set -x
O=gpurun_out/r2_call18
mkdir -p $O
timeout 600 python tools/scripts/e2e_display_probe.py > $O/display_probe.log 2>&1
cat $O/display_probe.log
