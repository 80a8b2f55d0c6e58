# round 2, call 16: tests, C2 with the pipelined window sums, display launch list, launch list of the bench command
set -x
O=gpurun_out/r2_call16
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
grep -E "passed|failed|FAILED|^E  " $O/pytest_gpu.log | tail -6
timeout 300 python bench.py --workload c2 --steps 50 --no-cpu --no-configs --no-e2e > $O/bench_c2.json 2> $O/err.log
python -c "
import json; d=json.load(open('$O/bench_c2.json')); r=d['roofline']; print('c2 kernel_ms', r['kernel_ms'], 'post', r['post_kernels_ms'], 'step', d['ms_per_step'])"
cat > /tmp/disp.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from glfer_b200 import api, synth
x = synth.tiled_stream(48000 * 600, fs=48000, block_s=20.0)
pcm = np.rint(x * 32768.0).astype(np.int16)
p = api.GramPlan(n=4096, window_type=0, overlap=0.5, sub_mean=True)
for i in range(2):
    p.run_display(pcm, log_scale=True, autoscale=True)
PY
python /tmp/disp.py > $O/disp_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_display.csv python /tmp/disp.py > $O/ncu_disp.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > $O/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv $CMD > $O/ncu_bench.log 2>&1
python - <<'PY'
import csv,collections
for name in ('launches_display','launches_bench'):
    rows=list(csv.reader(open('gpurun_out/r2_call16/%s.csv'%name)))
    hdr=None; agg=collections.defaultdict(lambda:[0,0.0])
    for r in rows:
        if 'Kernel Name' in r: hdr=r; continue
        if hdr and len(r)==len(hdr):
            d=dict(zip(hdr,r))
            try: v=float(d['Metric Value'].replace(',',''))
            except: continue
            k=d['Kernel Name'][:70]; agg[k][0]+=1; agg[k][1]+=v
    print(name)
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("  %-72s n=%4d total %.3f ms"%(k,n,t/1e6))
PY
