"""Pins the numpy restatement (oracle/glfer_oracle.py) against the reference: the committed
fixtures generated from the unmodified reference sources, the known-answer values of
SURVEY.md section 8c, and (when oracle/_ref is present) the reference library itself."""
import os

import numpy as np
import pytest

from glfer_b200 import synth
from oracle import glfer_oracle as O
from oracle import ref_lib as R

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "glfer_ref_f64.npz"))
X = synth.pcm16_to_float(GOLD["pcm"])
have_ref = pytest.mark.skipif(not R.available("f64"), reason="oracle/_ref not built (needs /root/reference)")


def test_windows_bit_exact_vs_fixture():
    for t in range(8):
        assert np.array_equal(O.compute_window(1024, t), GOLD["windows_1024"][t]), O.WINDOW_NAMES[t]


def test_window_kat_values():
    # w[512] and sum(w) of the unit-energy windows at N=1024
    kat = {0: (5.105584487e-02, 26.11512763), 1: (5.664945394e-02, 24.34009806), 2: (4.042137787e-02, 30.89686565),
           3: (4.281169921e-02, 29.19757857), 4: (5.410012975e-02, 27.69926637), 5: (3.125e-02, 32.0),
           6: (4.959570244e-02, 27.40168354), 7: (5.162739381e-02, 26.40954478)}
    for t, (mid, total) in kat.items():
        w = O.compute_window(1024, t).astype(np.float64)
        assert abs(w[512] - mid) < 2e-9
        assert abs(w.sum() - total) < 2e-6


def test_periodogram_fixtures_bit_exact():
    assert np.array_equal(O.periodogram(X, 1024, 0, 0.5, True), GOLD["c1_rows"])
    assert np.array_equal(O.periodogram(X, 1024, 0, 0.5, False), GOLD["c1_rows_nomean"])
    assert np.array_equal(O.periodogram(X, 4096, 7, 0.75, True), GOLD["c2_rows"])
    for t in range(8):
        assert np.array_equal(O.periodogram(X[:8192], 512, t, 0.5, True), GOLD["win_rows"][t])
    assert np.array_equal(O.periodogram(X[:20000], 1024, 1, 0.9, True), GOLD["odd_rows"])
    assert np.array_equal(O.periodogram(X[:20000], 1024, 0, 0.5, True, a=0.01, limiter=1), GOLD["preop_rows"])


def test_sine_known_answers():
    i = np.arange(4096)
    xs = np.sin(2 * np.pi * i / 8).astype(np.float32)
    h = O.periodogram(xs, 1024, O.HANNING, 0.5)[:2, 128]
    r = O.periodogram(xs, 1024, O.RECTANGULAR, 0.5)[:2, 128]
    assert np.allclose(h, [4.146352410e-02, 1.665038764e-01], rtol=2e-7)
    assert np.allclose(r, [64.0, 256.0], rtol=1e-6)          # rectangular is not normalised
    assert np.array_equal(h, GOLD["kat_sine_hann"]) and np.array_equal(r, GOLD["kat_sine_rect"])


def test_round2_fixtures():
    """zeroed history on every frame (GUI, autoscale off), C4 shape, C5 at both NW, the F-test"""
    g2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "glfer_ref_f64_r2.npz"))
    x48 = synth.pcm16_to_float(g2["pcm48"])
    assert np.array_equal(O.periodogram(X[:20000], 1024, 0, 0.75, False, zero_history=True), g2["zero_hist_rows"])
    assert np.array_equal(O.periodogram(X[:12000], 512, 7, 0.9, False, zero_history=True), g2["zero_hist_rows_odd"])
    assert np.array_equal(O.periodogram(x48, 16384, 0, 0.5, True), g2["c4_rows"])
    p = O.multitaper(X, 4096, 0.5, 4.0, 7, True)
    assert (np.abs(p - g2["c3_rows_4096"]) / g2["c3_rows_4096"]).max() < 1e-6
    p = O.multitaper(x48[:3 * 16384], 32768, 0.5, 8.0, 15, True)
    assert (np.abs(p - g2["c5_rows_nw8"][:3]) / g2["c5_rows_nw8"][:3]).max() < 1e-5
    # F-test with the fixture's own tapers: bit exact, including the inf at Nyquist (mtm.c:229-233)
    ft = O.multitaper_ftest(X, 1024, 0.5, 4.0, 7, True, tapers=GOLD["c3_tapers"].astype(np.float64), lam=GOLD["c3_lambda"])
    ref = g2["c3_ftest_1024"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(ft), fin) and not fin[:, -1].any() and fin[:, :-1].all()
    assert np.allclose(ft[fin], ref[fin], rtol=2e-4)          # float32 copy of the tapers in the fixture


def _display_case(rows, kw, n=513):
    """the restatement driven with the arguments of one oracle.ref_gui.draw_rows case"""
    st = kw.get("scale_type", 2)
    log, max0 = st in (2, 3), st in (1, 3)
    shown = rows
    if kw.get("averaging", 0):
        fs, block = kw["sample_rate"], kw["data_block_size"]
        binsize = np.float32(fs) / np.float32(block)                       # g_main.c:1144-1146
        mn, mx = int(np.float32(400.0) / binsize), int(np.float32(1200.0) / binsize)
        shown = O.update_avg(kw["averaging"], rows, block, kw["avgsamples"], mn, mx, int(max0))[0][:, :n]
    return O.display_levels(rows, shown, kw.get("overlap", 0.5), log, kw.get("autoscale", True),
                            kw.get("max_level_db", -20.0), kw.get("min_level_db", -80.0), kw.get("thr_level", 0.0))[0]


def test_display_mapping_pinned():
    """main_window_draw (g_main.c:1072-1281): fixtures produced by the compiled reference GUI unit"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mg2", os.path.join(os.path.dirname(__file__), "golden", "make_golden_r2.py"))
    mg2 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg2)
    g2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "glfer_ref_f64_r2.npz"))
    rows = GOLD["c1_rows"]
    for name, kw in mg2.DISPLAY_CASES.items():
        assert np.array_equal(_display_case(rows, kw), g2[f"disp_{name}"]), name
    # palette look-up: the HOT palette applied to the log/autoscale levels
    lv = g2["disp_log_auto"][:8]
    assert np.array_equal(g2["palettes"][3][lv], g2["disp_hot_rgb"])
    assert np.array_equal(g2["palettes"][4][:, 0], np.arange(256))           # B/W: index = grey level


@have_ref
def test_display_mapping_matches_live_reference_gui_unit():
    from oracle import ref_gui as G
    if not G.available():
        pytest.skip("oracle/_ref/libglfer_ref_gui.so not built")
    rng = np.random.default_rng(11)
    rows = (rng.standard_normal((40, 257)) ** 2 * 1e-5).astype(np.float32)
    rows[:, 30] += 3e-3
    for kw in (dict(scale_type=2, autoscale=True, overlap=0.75, thr_level=5.0),
               dict(scale_type=0, autoscale=False, max_level_db=-25.0, min_level_db=-65.0, overlap=0.0)):
        assert np.array_equal(_display_case(rows, kw, n=257), G.draw_rows(rows, **kw)["levels"])


def test_hop_truncation():
    assert O.hop_size(4096, 0.9) == 409            # float overlap, double product, truncated
    assert O.hop_size(1024, 0.5) == 512 and O.hop_size(4096, 0.75) == 1024 and O.hop_size(1024, 0.0) == 1024


def test_dpss_and_multitaper_fixture():
    tap, lam = O.gl_dpss(1024, 4.0, 7)
    assert np.allclose(lam, GOLD["c3_lambda"], rtol=1e-12)
    sgn = np.sign(np.sum(tap * GOLD["c3_tapers"], axis=1))
    assert np.max(np.abs(tap * sgn[:, None] - GOLD["c3_tapers"])) < 1e-6
    # SURVEY 8c (probe of the unmodified reference): 1 - lambda_0..3 and lambda_4..7
    assert np.allclose(1.0 - lam[:4], [2.946e-10, 2.768e-08, 1.210e-06, 3.245e-05], rtol=2e-3)
    assert np.allclose(lam[4:], [0.999410076, 0.992504500, 0.936652233, 0.698835614], rtol=0, atol=2e-9)
    rows = O.multitaper(X, 1024, 0.5, 4.0, 7, True)
    rel = np.abs(rows.astype(np.float64) - GOLD["c3_rows"]) / GOLD["c3_rows"]
    assert rel.max() < 1e-6


def test_lmp_fixture_bit_exact():
    """lmp_do (lmp.c:101-192) restated: identical floats to the unmodified reference."""
    assert np.array_equal(O.lmp(X, 1024, 0.5, 4, True).view(np.uint32), GOLD["lmp_rows"].view(np.uint32))
    assert np.array_equal(O.lmp(X[:12000], 512, 0.75, 7, False).view(np.uint32), GOLD["lmp_rows_n512"].view(np.uint32))
    # bin 0 and the floor are pinned to 1e-3 (lmp.c:155-157)
    assert np.all(GOLD["lmp_rows"][:, 0] == np.float32(1e-3)) and GOLD["lmp_rows"].min() == np.float32(1e-3)
    # a shard that starts inside the run reproduces the ring from its nl - 1 frames of history
    full = O.lmp(X, 1024, 0.5, 4, True)
    assert np.array_equal(O.lmp(X, 1024, 0.5, 4, True, first_frame=9, nframes=20), full[9:29])


def test_avg_fixtures():
    mn, mx = GOLD["c2_band"]
    assert (mn, mx) == O.avg_bins(8000, 4096, 400.0, 1200.0)
    for mode, name in ((O.AVG_PLAIN, "plain"), (O.AVG_SUMAVG, "sumavg"), (O.AVG_SUMEXTREME, "sumextreme")):
        a, ret, pk, var = O.update_avg(mode, GOLD["c2_rows"], 4096, 4, int(mn), int(mx), 0)
        assert np.allclose(a[:, :2049], GOLD[f"c2_avg_{name}"], rtol=1e-12, atol=0)
        assert np.allclose(ret, GOLD[f"c2_ret_{name}"], rtol=1e-12)
        assert np.array_equal(pk, GOLD[f"c2_pk_{name}"])
        if mode == O.AVG_SUMAVG:
            assert np.allclose(var, GOLD[f"c2_var_{name}"], rtol=1e-12)


def test_avg_plain_known_answer():
    # depth 3, band [2,7), psd[b] = (f+1)(b+1)
    psd = np.array([[(f + 1) * (b + 1) for b in range(16)] for f in range(5)], dtype=np.float32)
    a, ret, pk, _ = O.update_avg(O.AVG_PLAIN, psd, 16, 3, 2, 7)
    assert np.allclose(a[:, 2], [1.5, 3.0, 4.5, 6.75, 9.0])
    assert np.allclose(ret, [2.25, 4.5, 6.75, 10.125, 13.5])
    assert (pk == 6).all() and a[0, 0] == 1e-15 and a[0, 7] == 1e-15


def test_wav_reader_stale_tail(tmp_path):
    pcm = GOLD["pcm"][:1000 + 300]
    p = tmp_path / "t.wav"
    synth.write_wav16(str(p), pcm, 8000)
    stream, rate, bits = O.read_wav_blocks(str(p), 500)
    assert rate == 8000 and bits == 16 and len(stream) == 1500
    assert np.array_equal(stream[:1300], synth.pcm16_to_float(pcm))
    assert np.array_equal(stream[1300:], stream[800:1000])        # stale tail of the previous block
    # with sub_mean the stale tail was already mean-subtracted in place (fft.c:93-95)
    s2, _, _ = O.read_wav_blocks(str(p), 500, sub_mean=True)
    prev = stream[500:1000]
    assert np.array_equal(s2[1300:], O.subtract_block_means(prev, 500)[300:])
    assert np.array_equal(s2[:1300], stream[:1300])


@have_ref
def test_headless_loop_on_reference_equals_restatement(tmp_path):
    """tools/glfer_headless.c (the per-block loop of source.c) linked against the unmodified
    reference (oracle/_ref/glfer_headless_ref) reproduces the restatement bit for bit,
    including the stale, mean-removed tail of a partial last block"""
    import subprocess
    exe = os.path.join(os.path.dirname(R.path()), "glfer_headless_ref")
    if not os.path.exists(exe):
        pytest.skip("glfer_headless_ref not built")
    pcm = GOLD["pcm"][:8000 * 2 + 77]
    wav = tmp_path / "t.wav"
    synth.write_wav16(str(wav), pcm, 8000)
    for args, fn in ((["-n", "1024", "-w", "0", "-o", "0.5", "-s", "1"],
                      lambda st: O.periodogram(st, 1024, 0, 0.5, True)),
                     (["-n", "512", "-w", "7", "-o", "0.75", "-s", "0"],
                      lambda st: O.periodogram(st, 512, 7, 0.75, False))):
        out = tmp_path / "rows.f32"
        subprocess.run([exe, "-f", str(wav), "-O", str(out)] + args, check=True, capture_output=True)
        n = int(args[1])
        hop = O.hop_size(n, float(args[5]))
        st, _, _ = O.read_wav_blocks(str(wav), hop, sub_mean=args[7] == "1")
        got = np.fromfile(str(out), dtype=np.float32).reshape(-1, n // 2 + 1)
        assert np.array_equal(got, fn(st))


@have_ref
def test_restatement_matches_reference_library():
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(30000) * 0.1).astype(np.float32)
    for n, wt, ov, sm, a, lim in ((256, 2, 0.25, True, 0.0, 0), (2048, 4, 0.6, False, 0.0, 0), (1024, 6, 0.5, True, 0.05, 1),
                                  (8192, 3, 0.5, True, 0.0, 0)):
        assert np.array_equal(O.periodogram(x, n, wt, ov, sm, a, lim), R.periodogram(x, n, wt, ov, sm, a, lim))
    for n in (64, 1024, 16384, 32768):
        for t in range(8):
            assert np.array_equal(O.compute_window(n, t), R.window(n, t))
    q = R.mtm(x, 2048, 0.5, 3.0, 5, True)
    p = O.multitaper(x, 2048, 0.5, 3.0, 5, True)
    assert (np.abs(p.astype(np.float64) - q) / q).max() < 1e-6
    rows = R.periodogram(x, 512, 0, 0.5, True)
    for mode in (1, 2, 3):
        for max0 in (0, 1):
            a1 = O.update_avg(mode, rows, 512, 5, 10, 200, max0)
            a2 = R.avg(mode, rows, 512, 5, 10, 200, max0)
            assert np.allclose(a1[0], a2[0], rtol=1e-12) and np.allclose(a1[1], a2[1], rtol=1e-12)
            assert np.array_equal(a1[2], a2[2])
    for n, ov, nl, sm in ((1024, 0.5, 4, True), (256, 0.75, 9, False), (2048, 0.0, 2, True)):
        assert np.array_equal(O.lmp(x, n, ov, nl, sm).view(np.uint32), R.lmp(x, n, ov, nl, sm).view(np.uint32))
    for row in rows[:40]:                        # compute_floor, bit for bit (fft.c:240-294)
        assert R.floor_stats(row) == tuple(np.float32(v) if i < 3 else v for i, v in enumerate(O.compute_floor(row)))
