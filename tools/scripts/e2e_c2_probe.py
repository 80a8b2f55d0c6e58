"""Where does the C2 (Kaiser 75 % + plain averaging) end-to-end path spend its time?  glfer_gram_run with pinned
buffers, outputs switched on one by one, several chunk sizes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from glfer_b200 import api, synth      # noqa: E402

FS = 48000
xs = synth.tiled_stream(3600 * FS, fs=FS, block_s=20.0)
x = api.pinned_empty((len(xs),), np.float32)
x[:] = xs
KW = dict(n=4096, window_type=7, overlap=0.75, sub_mean=True)
AVG = dict(avg_mode=2, avg_depth=4, avg_minbin=34, avg_maxbin=102, avg_band_only=True)
nf = len(x) // 1024
psd = api.pinned_empty((nf, 2049), np.float32)
avg = api.pinned_empty((nf, 68), np.float32)
ret = api.pinned_empty((nf,), np.float64)
pk = api.pinned_empty((nf,), np.int32)
var = api.pinned_empty((nf,), np.float64)


def timed(label, kw, out, want_psd=True):
    p = api.GramPlan(**kw)
    p.run(x, out=dict(out), want_psd=want_psd)
    t0 = time.perf_counter()
    for _ in range(3):
        p.run(x, out=dict(out), want_psd=want_psd)
    dt = (time.perf_counter() - t0) / 3
    print(f"{label:60s} {1e3 * dt:8.2f} ms per step  {nf / dt:.3e} frames/s", flush=True)
    p.close()


full = {"psd": psd, "avg": avg, "ret": ret, "peakbin": pk, "variance": var}
timed("no averaging, psd rows only", KW, {"psd": psd})
timed("averaging, every output pinned and preallocated", {**KW, **AVG}, full)
timed("averaging, scalars allocated per call (bench standalone)", {**KW, **AVG}, {"psd": psd, "avg": avg})
timed("averaging, no psd rows (avg rows + scalars only)", {**KW, **AVG}, {"psd": None, "avg": avg, "ret": ret, "peakbin": pk, "variance": var}, want_psd=False)
timed("averaging, no peakbin", {**KW, **AVG}, {"psd": psd, "avg": avg, "ret": ret, "peakbin": None, "variance": var})
timed("averaging, rows only (no scalars)", {**KW, **AVG}, {"psd": psd, "avg": avg, "ret": None, "peakbin": None, "variance": None})
for mib in (8, 128):
    os.environ["GLFER_B200_CHUNK_MIB"] = str(mib)
    timed(f"averaging, all pinned, chunk {mib} MiB", {**KW, **AVG}, full)
