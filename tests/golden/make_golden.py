"""Generates tests/golden/*.npz from the UNMODIFIED reference sources
(oracle/_ref/libglfer_ref_f64.so, built by `make -C oracle`; needs /root/reference, so it
runs in the build container only).  Inputs are stored with the outputs so the fixtures
are self-contained on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from glfer_b200 import synth            # noqa: E402
from oracle import ref_lib as R         # noqa: E402


def main():
    assert R.available("f64"), "run `make -C oracle` first"
    fs = 8000
    pcm = synth.qrss_stream_int16(fs * 5, fs=fs, seed=0xC1, dot_s=0.25)
    x = synth.pcm16_to_float(pcm)

    # C1: N=1024 Hanning, 50 % overlap on an 8 kHz 16-bit stream (sub_mean on = glfer default)
    out = {"pcm": pcm, "fs": np.int32(fs)}
    out["c1_rows"] = R.periodogram(x, 1024, 0, 0.5, True)
    out["c1_rows_nomean"] = R.periodogram(x, 1024, 0, 0.5, False)
    # every window type at N=512, 50 %
    out["win_rows"] = np.stack([R.periodogram(x[:8192], 512, t, 0.5, True) for t in range(8)])
    out["windows_1024"] = np.stack([R.window(1024, t) for t in range(8)])
    # C2 shape: N=4096 Kaiser 75 % + plain averaging over the default band at fs
    rows = R.periodogram(x, 4096, 7, 0.75, True)
    out["c2_rows"] = rows
    binsize = np.float32(fs) / np.float32(4096)
    mn, mx = int(np.float32(400.0) / binsize), int(np.float32(1200.0) / binsize)
    out["c2_band"] = np.array([mn, mx], dtype=np.int32)
    for mode, name in ((2, "plain"), (1, "sumavg"), (3, "sumextreme")):
        a, ret, pk, var = R.avg(mode, rows, 4096, 4, mn, mx, 0, nbins_out=2049)
        out[f"c2_avg_{name}"] = a
        out[f"c2_ret_{name}"] = ret
        out[f"c2_pk_{name}"] = pk
        out[f"c2_var_{name}"] = var
    # C3 shape: multitaper N=1024, kmax=7, NW=4, 50 %
    out["c3_rows"] = R.mtm(x, 1024, 0.5, 4.0, 7, True)
    tap, lam = R.dpss(1024, 4.0, 7)
    out["c3_lambda"] = lam
    out["c3_tapers"] = tap.astype(np.float32)
    # LMP (lmp.c): ring of 4 rectangular periodograms at N=1024, 50 %; ring of 7 at N=512, 75 %
    out["lmp_rows"] = R.lmp(x, 1024, 0.5, 4, True)
    out["lmp_rows_n512"] = R.lmp(x[:12000], 512, 0.75, 7, False)
    # odd hop (overlap 0.9 -> hop 102 at N=1024), RA9MB and limiter
    out["odd_rows"] = R.periodogram(x[:20000], 1024, 1, 0.9, True)
    out["preop_rows"] = R.periodogram(x[:20000], 1024, 0, 0.5, True, a=0.01, limiter=1)
    # known-answer: x = sin(2 pi i / 8), N=1024, 50 %
    i = np.arange(4096)
    xs = np.sin(2 * np.pi * i / 8).astype(np.float32)
    out["kat_sine_hann"] = R.periodogram(xs, 1024, 0, 0.5, False)[:2, 128]
    out["kat_sine_rect"] = R.periodogram(xs, 1024, 5, 0.5, False)[:2, 128]
    np.savez_compressed(os.path.join(HERE, "glfer_ref_f64.npz"), **out)
    print("wrote", os.path.join(HERE, "glfer_ref_f64.npz"), {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
