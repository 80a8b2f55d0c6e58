# round 2, call 6: F-test, TMA landing in the big kernel; C4 benches; launch list of the display path
set -x
O=gpurun_out/r2_call6
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|ftest rel" $O/pytest_gpu.log | tail -12
timeout 300 python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k ftest > $O/pytest_ftest.log 2>&1; grep "ftest rel" $O/pytest_ftest.log
timeout 300 python bench.py --workload c4 --steps 50 --no-cpu --no-configs --no-e2e > $O/bench_c4.json 2> $O/bench_c4.err
timeout 300 python bench.py --workload c4 --steps 50 --no-cpu --no-configs --no-e2e --no-submean > $O/bench_c4_nosub.json 2> $O/bench_c4_nosub.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_call6/bench_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']; print(f, r['kernel'], 'kernel_ms %.4f frac %.3f'%(r['kernel_ms'], r['frac']))
    except Exception as e: print(f,'ERR',e)
PY
cat > /tmp/disp.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from glfer_b200 import api, synth
x = synth.tiled_stream(48000 * 600, fs=48000, block_s=20.0)
pcm = np.rint(x * 32768.0).astype(np.int16)
p = api.GramPlan(n=4096, window_type=0, overlap=0.5, sub_mean=True)
for i in range(2):
    p.run_display(pcm, log_scale=True, autoscale=True)
p2 = api.GramPlan(n=4096, window_type=7, overlap=0.75, sub_mean=True, avg_mode=2, avg_depth=4, avg_minbin=34, avg_maxbin=102, avg_band_only=True)
for i in range(2):
    p2.run(x)
p3 = api.GramPlan(n=4096, mode=3, overlap=0.5, sub_mean=True, lmp_av=4)
p3.run(x)
PY
python /tmp/disp.py > $O/disp_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_display.csv python /tmp/disp.py > $O/ncu_disp.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2_call6/launches_display.csv')))
hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        k=d['Kernel Name'][:60]; agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-62s n=%4d total %.3f ms"%(k,n,t/1e6))
PY
