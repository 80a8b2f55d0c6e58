set -x
O=gpurun_out/r2_disp
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "floor or display or level or palette or smoke or headless" > $O/pytest.log 2>&1; tail -15 $O/pytest.log
timeout 600 python tools/scripts/e2e_display_probe.py > $O/display_probe.log 2>&1; cat $O/display_probe.log
for ns in 3 4; do GLFER_B200_LIB=$PWD/glfer_b200/libglfer_b200_ns$ns.so timeout 600 python tools/scripts/e2e_display_probe.py > $O/display_probe_ns$ns.log 2>&1; cat $O/display_probe_ns$ns.log; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_display.csv python tools/scripts/e2e_display_probe.py --once > $O/ncu.log 2>&1; tail -2 $O/ncu.log
