"""Where does the autoscale display path spend its time end to end?  Times glfer_gram_run_display_pcm16 for
several chunk sizes and with stages switched off one by one (levels only / no range download)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from glfer_b200 import api, synth      # noqa: E402

FS = 48000
x = synth.tiled_stream(3600 * FS, fs=FS, block_s=20.0)
pcm = api.pinned_empty((len(x),), np.int16)
pcm[:] = np.rint(x * 32768.0).astype(np.int16)
nf = len(x) // 2048
lev = api.pinned_empty((nf, 2049), np.uint8)


def timed(label, **kw):
    p = api.GramPlan(n=4096, window_type=0, overlap=0.5, sub_mean=True)
    p.run_display(pcm, out={"levels": lev}, want_range=False, **kw)
    t0 = time.perf_counter()
    for _ in range(4):
        p.run_display(pcm, out={"levels": lev}, want_range=False, **kw)
    dt = (time.perf_counter() - t0) / 4
    print(f"{label:40s} {1e3 * dt:8.2f} ms per step  {nf / dt:.3e} frames/s", flush=True)
    p.close()


if "--once" in sys.argv:                 # launch list under ncu: one autoscale run
    p = api.GramPlan(n=4096, window_type=0, overlap=0.5, sub_mean=True)
    p.run_display(pcm[: 40 * 2048 * 4096], out={"levels": lev[: 40 * 4096 - 1]}, log_scale=True, autoscale=True)
    sys.exit(0)
for mib in (32, 128):
    os.environ["GLFER_B200_CHUNK_MIB"] = str(mib)
    timed(f"autoscale, chunk {mib} MiB", log_scale=True, autoscale=True)
    timed(f"fixed range (fused), chunk {mib} MiB", log_scale=True, autoscale=False)
os.environ["GLFER_B200_CHUNK_MIB"] = "32"
api.set_fused_levels(False)
timed("fixed range, two-pass, chunk 32 MiB", log_scale=True, autoscale=False)
