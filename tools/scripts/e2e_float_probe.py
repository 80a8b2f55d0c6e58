"""glfer_gram_run (float32 in, float32 rows out, pinned buffers) against the chunk size, and the copies alone."""
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
nf = len(x) // 2048
psd = api.pinned_empty((nf, 2049), np.float32)
for mib in (8, 16, 32, 64, 128, 256):
    os.environ["GLFER_B200_CHUNK_MIB"] = str(mib)
    p = api.GramPlan(n=4096, window_type=0, overlap=0.5, sub_mean=True)
    p.run(x, out={"psd": psd})
    t0 = time.perf_counter()
    for _ in range(5):
        p.run(x, out={"psd": psd})
    dt = (time.perf_counter() - t0) / 5
    print(f"chunk {mib:4d} MiB  {1e3 * dt:8.2f} ms per step  {nf / dt:.3e} frames/s", flush=True)
    p.close()
