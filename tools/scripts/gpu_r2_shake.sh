set -x
O=gpurun_out/r2_shake
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "big or shard or smoke or 24 or golden" > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for s in 0 1 8000 0 1 8000; do
  python bench.py --workload c4 --steps 40 --warmup 5 --no-configs --no-e2e --stagger $s > $O/c4_s$s.json 2> $O/c4_s$s.err
  python - $O/c4_s$s.json $s <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]; print("c4 stagger", sys.argv[2], r["kernel_ms"], r["frac"])
PY
done
python - <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from glfer_b200 import api, synth
FS = 48000
x = synth.tiled_stream(3 * 3600 * FS, fs=FS, block_s=20.0)
for kw in (dict(n=16384, window_type=7, overlap=0.75, sub_mean=True), dict(n=16384, window_type=0, overlap=0.0, sub_mean=True),
           dict(n=16384, window_type=0, overlap=0.5, sub_mean=False)):
    for st in (0, 1, 0, 1):
        api.set_stagger_cycles(st)
        p = api.GramPlan(**kw)
        nf = p.num_frames(len(x))
        p.stage(x)
        for _ in range(3):
            p.exec(0, nf)
        p.sync()
        ms = [p.exec(0, nf, timed=True) for _ in range(20)]
        print(kw["overlap"], kw["sub_mean"], "stagger", st, "frames", nf, "ms %.4f" % (sum(ms) / len(ms)), flush=True)
        p.close()
PY
