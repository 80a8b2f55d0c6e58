"""Round-2 fixtures, generated from the UNMODIFIED reference sources exactly like make_golden.py
(oracle/_ref/libglfer_ref_f64.so; needs /root/reference, so it runs in the build container only):
the BASELINE configurations at their own FFT sizes, the GUI's autoscale-off sequence
(glfer.first_buffer never cleared) and the harmonic F-test mtm_do leaves in its file-static buffer.

    python tests/golden/make_golden_r2.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from glfer_b200 import synth            # noqa: E402
from oracle import ref_lib as R         # noqa: E402


# keyword arguments of oracle.ref_gui.draw_rows (scale_type: glfer.h:43 SCALE_LIN, _LIN_MAX0, _LOG, _LOG_MAX0;
# averaging: glfer.h:55); the tests rebuild the same cases through the product and the restatement
DISPLAY_CASES = {
    "log_auto": dict(scale_type=2, autoscale=True, overlap=0.5),
    "log_fixed": dict(scale_type=2, autoscale=False, max_level_db=-30.0, min_level_db=-70.0, overlap=0.5),
    "lin_auto_thr10": dict(scale_type=0, autoscale=True, thr_level=10.0, overlap=0.5),
    "lin_fixed": dict(scale_type=0, autoscale=False, max_level_db=-30.0, min_level_db=-70.0, overlap=0.5),
    "log_auto_thr25": dict(scale_type=2, autoscale=True, thr_level=25.0, overlap=0.5),
    "log_fixed_bad_range": dict(scale_type=2, autoscale=False, max_level_db=-60.0, min_level_db=-40.0, overlap=0.5),
    "log_plain_avg": dict(scale_type=2, autoscale=True, overlap=0.5, averaging=2, avgsamples=4, sample_rate=8000,
                          data_block_size=1024),
    "logmax0_sumavg": dict(scale_type=3, autoscale=True, overlap=0.5, averaging=1, avgsamples=4, sample_rate=8000,
                           data_block_size=1024),
    "linmax0_sumextreme": dict(scale_type=1, autoscale=False, max_level_db=0.0, min_level_db=-30.0, overlap=0.5,
                               averaging=3, avgsamples=6, sample_rate=8000, data_block_size=1024),
}


def main():
    assert R.available("f64"), "run `make -C oracle` first"
    g1 = np.load(os.path.join(HERE, "glfer_ref_f64.npz"))
    x8 = synth.pcm16_to_float(g1["pcm"])                       # the 8 kHz stream of the round-1 fixtures
    pcm48 = synth.qrss_stream_int16(5 * 16384, fs=48000, seed=0xC4, dot_s=0.2)
    x48 = synth.pcm16_to_float(pcm48)
    out = {"pcm48": pcm48}
    # GUI with opt.autoscale == 0: sub_mean off AND first_buffer never cleared (g_main.c:1111-1120)
    out["zero_hist_rows"] = R.periodogram(x8[:20000], 1024, 0, 0.75, False, sticky_first_buffer=True)
    out["zero_hist_rows_odd"] = R.periodogram(x8[:12000], 512, 7, 0.9, False, sticky_first_buffer=True)
    # C3 at its own size: multitaper N=4096, mtm_k=7, NW=4, 50 %
    out["c3_rows_4096"], out["c3_ftest_4096"] = R.mtm_ftest(x8, 4096, 0.5, 4.0, 7, True)
    # F-test at N=1024 (the rows equal c3_rows of the round-1 file)
    rows, out["c3_ftest_1024"] = R.mtm_ftest(x8, 1024, 0.5, 4.0, 7, True)
    assert np.array_equal(rows, g1["c3_rows"])
    # C4 shape: N=16384 Hann 50 %
    out["c4_rows"] = R.periodogram(x48, 16384, 0, 0.5, True)
    # C5 shape: multitaper N=32768, mtm_k=15, at NW=8 (2NW-1 = 15) and at glfer's default NW=4
    out["c5_rows_nw8"] = R.mtm(x48, 32768, 0.5, 8.0, 15, True)
    out["c5_rows_nw4"] = R.mtm(x48, 32768, 0.5, 4.0, 15, True)
    # "all window types swept" at N=32768: the interior frame 2 of each periodogram
    out["win_rows_32768"] = np.stack([R.periodogram(x48, 32768, t, 0.5, True)[2] for t in range(8)])
    # display mapping: the reference's main_window_draw (g_main.c, compiled with GTK stubbed out) on the
    # C1 rows of the round-1 fixture; levels = palette index per pixel (pixel i = bin n-1-i), B/W palette
    from oracle import ref_gui as G
    rows = g1["c1_rows"]
    for name, kw in DISPLAY_CASES.items():
        r = G.draw_rows(rows, **kw)
        out[f"disp_{name}"] = r["levels"]
    out["disp_hot_rgb"] = G.draw_rows(rows[:8], scale_type=G.SCALE_LOG, autoscale=True, palette_id=G.HOT)["rgb"]
    out["palettes"] = np.stack([G.palette(p) for p in range(8)])
    path = os.path.join(HERE, "glfer_ref_f64_r2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
