"""Parity of the CUDA path against the oracle, through the C-ABI (ctypes -> libglfer_b200.so).
All tests here need a B200 (`-m gpu`); nothing reads /root/reference at run time.  When
oracle/_ref (the unmodified reference, built in the container) travelled with the
snapshot it is used as a second checker."""
import ctypes as C
import os

import numpy as np
import pytest

from glfer_b200 import synth
from oracle import glfer_oracle as O
from oracle import ref_lib as R
from parity import assert_psd_close, psd_stats

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "glfer_ref_f64.npz"))
XG = synth.pcm16_to_float(GOLD["pcm"])


def stream(n, fs=48000, seed=7):
    return synth.qrss_stream(n, fs=fs, seed=seed, dot_s=0.2)


# ------------------------------------------------------------------ golden fixtures
def test_c1_fixture(gpu_api):
    p = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=True)
    r = p.run(XG)
    assert r["psd"].shape == GOLD["c1_rows"].shape          # frame count and bin layout
    assert_psd_close(r["psd"], GOLD["c1_rows"], "C1")
    p2 = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=False)
    assert_psd_close(p2.run(XG)["psd"], GOLD["c1_rows_nomean"], "C1 no mean")


def test_c1_pcm16_ingest_matches_float(gpu_api):
    p = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=True)
    a = p.run(XG)["psd"]
    b = p.run(np.ascontiguousarray(GOLD["pcm"]))["psd"]
    assert np.array_equal(a, b)                              # (float) s / 32768 is exact on both sides


def test_all_windows_fixture(gpu_api):
    for t in range(8):
        p = gpu_api.GramPlan(n=512, window_type=t, overlap=0.5, sub_mean=True)
        assert_psd_close(p.run(XG[:8192])["psd"], GOLD["win_rows"][t], O.WINDOW_NAMES[t])
        assert np.array_equal(p.window(), O.compute_window(512, t))


def test_sine_known_answers(gpu_api):
    i = np.arange(4096)
    xs = np.sin(2 * np.pi * i / 8).astype(np.float32)
    h = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=False).run(xs)["psd"]
    r = gpu_api.GramPlan(n=1024, window_type=5, overlap=0.5, sub_mean=False).run(xs)["psd"]
    assert np.allclose(h[:2, 128], GOLD["kat_sine_hann"], rtol=1e-5)
    assert np.allclose(r[:2, 128], [64.0, 256.0], rtol=1e-5)   # rectangular quirk: not normalised
    assert np.argmax(h[1]) == 128


def test_c2_fixture_with_averaging(gpu_api):
    mn, mx = (int(v) for v in GOLD["c2_band"])
    for mode, name in ((2, "plain"), (1, "sumavg"), (3, "sumextreme")):
        p = gpu_api.GramPlan(n=4096, window_type=7, overlap=0.75, sub_mean=True, avg_mode=mode, avg_depth=4,
                             avg_minbin=mn, avg_maxbin=mx)
        r = p.run(XG)
        assert_psd_close(r["psd"], GOLD["c2_rows"], "C2 psd")
        ref = GOLD[f"c2_avg_{name}"]
        band = slice(mn, mx)
        assert np.allclose(r["avg"][:, band], ref[:, band], rtol=2e-4, atol=1e-15), name
        out = np.ones(2049, bool)
        out[band] = False
        assert np.all(r["avg"][:, out] == np.float32(1e-15))
        assert np.allclose(r["ret"], GOLD[f"c2_ret_{name}"], rtol=2e-4)
        assert np.array_equal(r["peakbin"], GOLD[f"c2_pk_{name}"])
        if mode == 1:
            assert np.allclose(r["variance"], GOLD[f"c2_var_{name}"], rtol=5e-4)


def test_c3_multitaper_fixture(gpu_api):
    p = gpu_api.GramPlan(n=1024, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=7)
    assert_psd_close(p.run(XG)["psd"], GOLD["c3_rows"], "C3")
    tap, lam = p.tapers()
    assert np.allclose(lam, GOLD["c3_lambda"], rtol=1e-9)


def test_odd_hop_and_preops_fixture(gpu_api):
    p = gpu_api.GramPlan(n=1024, window_type=1, overlap=0.9, sub_mean=True)
    assert p.hop == 102
    assert_psd_close(p.run(XG[:20000])["psd"], GOLD["odd_rows"], "hop 102")
    p = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=True, a=0.01, limiter=1)
    assert_psd_close(p.run(XG[:20000])["psd"], GOLD["preop_rows"], "RA9MB + limiter")


# ------------------------------------------------------------------ oracle on seeded inputs
@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768])
def test_every_fft_size(gpu_api, n):
    x = stream(max(8 * n, 20000), seed=n)
    for wt, ov, sm in ((0, 0.5, True), (7, 0.75, False)):
        p = gpu_api.GramPlan(n=n, window_type=wt, overlap=ov, sub_mean=sm)
        got = p.run(x)["psd"]
        ref = O.periodogram(x, n, wt, ov, sm)
        assert got.shape == ref.shape
        assert_psd_close(got, ref, f"N={n} w={wt} ov={ov}")


def test_every_kernel_family_on_regular_geometries(gpu_api):
    """Four kernel families serve the same launches: general (1), TMA ring (2), warp-per-frame
    (3; N <= 4096 periodograms, edge frames through the general kernel) and two-frames-per-thread
    (4; 50 % / 75 % overlap periodograms, N = 512..4096).  Each must match the oracle; the
    automatic choice is whatever is fastest."""
    x = stream(150000, seed=4)
    try:
        for kw in (dict(n=4096, window_type=0, overlap=0.5, sub_mean=True), dict(n=1024, window_type=7, overlap=0.75, sub_mean=True),
                   dict(n=4096, window_type=3, overlap=0.0, sub_mean=True), dict(n=2048, window_type=1, overlap=0.875, sub_mean=True),
                   dict(n=512, window_type=6, overlap=0.5, sub_mean=False), dict(n=4096, window_type=0, overlap=0.5, sub_mean=False, scale_db=True),
                   dict(n=16384, window_type=0, overlap=0.5, sub_mean=False),
                   dict(n=2048, mode=1, overlap=0.5, sub_mean=True, mtm_kmax=3, mtm_w=2.5)):
            if kw.get("mode"):
                ref = O.multitaper(x, kw["n"], kw["overlap"], kw["mtm_w"], kw["mtm_kmax"], True)
            else:
                ref = O.periodogram(x, kw["n"], kw["window_type"], kw["overlap"], kw["sub_mean"])
            for pref in (1, 2, 3, 4, 0):
                gpu_api.set_kernel_preference(pref)
                got = gpu_api.GramPlan(**kw).run(x)["psd"]
                if kw.get("scale_db"):
                    big = ref > ref.mean(axis=1, keepdims=True) * 1e-5
                    assert np.max(np.abs(got - 10 * np.log10(ref.astype(np.float64)))[big]) < 0.01, (pref, kw)
                else:
                    assert_psd_close(got, ref, f"family {pref} {kw}")
    finally:
        gpu_api.set_kernel_preference(0)


def test_pair_kernel_ragged_runs(gpu_api):
    """The two-frames-per-thread family on runs whose length is odd, shorter than one pair per
    group, or not a multiple of the group count: the unpaired last frame and the blocks past the
    end of the recording must not leak into the rows."""
    try:
        for n, ov in ((4096, 0.5), (2048, 0.75), (512, 0.5)):
            hop = O.hop_size(n, ov)
            for nframes in (2, 3, 5, 31, 298, 299):
                x = stream(nframes * hop - 7, seed=nframes)
                ref = O.periodogram(x, n, 0, ov, True)
                gpu_api.set_kernel_preference(4)
                got = gpu_api.GramPlan(n=n, window_type=0, overlap=ov, sub_mean=True).run(x)["psd"]
                assert got.shape == ref.shape
                assert_psd_close(got, ref, f"pair N={n} ov={ov} frames={nframes}")
    finally:
        gpu_api.set_kernel_preference(0)


@pytest.mark.parametrize("ov", [0.0, 0.25, 0.3, 0.5, 0.75, 0.9, 0.97])
def test_overlaps_including_odd_hops(gpu_api, ov):
    x = stream(60000, seed=3)
    for sm in (False, True):
        p = gpu_api.GramPlan(n=2048, window_type=6, overlap=ov, sub_mean=sm)
        assert p.hop == O.hop_size(2048, ov)
        assert_psd_close(p.run(x)["psd"], O.periodogram(x, 2048, 6, ov, sm), f"ov={ov} sm={sm}")


def test_multitaper_sizes(gpu_api):
    x = stream(50000, seed=11)
    for n, w, k, ov in ((256, 2.5, 3, 0.5), (4096, 4.0, 7, 0.5), (2048, 8.0, 15, 0.75), (16384, 4.0, 7, 0.5)):
        p = gpu_api.GramPlan(n=n, mode=1, overlap=ov, sub_mean=True, mtm_w=w, mtm_kmax=k)
        got = p.run(x)["psd"]
        ref = O.multitaper(x, n, ov, w, k, True)
        assert_psd_close(got, ref, f"MTM N={n} NW={w} k={k}")


def test_multitaper_n32768_k15(gpu_api):
    x = stream(32768 * 3, seed=12)
    p = gpu_api.GramPlan(n=32768, mode=1, overlap=0.5, sub_mean=True, mtm_w=8.0, mtm_kmax=15)
    assert_psd_close(p.run(x)["psd"], O.multitaper(x, 32768, 0.5, 8.0, 15, True), "C5 shape")


def test_db_output(gpu_api):
    x = stream(30000, seed=5)
    p = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=True, scale_db=True)
    got = p.run(x)["psd"]
    ref = 10 * np.log10(O.periodogram(x, 1024, 0, 0.5, True).astype(np.float64))
    big = ref > ref.mean(axis=1, keepdims=True) - 50
    assert np.max(np.abs(got - ref)[big]) < 0.01


def test_edge_cases(gpu_api):
    p = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=True)
    assert p.run(np.zeros(0, np.float32))["psd"].shape == (0, 513)           # empty stream
    assert p.run(np.zeros(511, np.float32))["psd"].shape == (0, 513)         # less than one block
    x = stream(512 + 100, seed=2)                                             # ragged: one block + tail
    got = p.run(x)["psd"]
    assert got.shape == (1, 513)
    assert_psd_close(got, O.periodogram(x, 1024, 0, 0.5, True), "single frame, zero history")
    z = p.run(np.zeros(4096, np.float32))["psd"]
    assert np.all(z == 0)                                                     # silence stays exactly zero
    with pytest.raises(gpu_api.GlferError):
        p.run(x, first_frame=1, nframes=5)                                    # frames beyond the samples


def test_frame_subranges_equal_full_run(gpu_api):
    x = stream(100000, seed=9)
    for kw in (dict(n=2048, window_type=7, overlap=0.75, sub_mean=True),
               dict(n=2048, window_type=0, overlap=0.6, sub_mean=True),
               dict(n=1024, mode=1, overlap=0.5, sub_mean=True, mtm_kmax=5, mtm_w=3.0)):
        p = gpu_api.GramPlan(**kw)
        full = p.run(x)["psd"]
        nf = full.shape[0]
        for first, cnt in ((0, 5), (7, 30), (nf - 3, 3)):
            lo, hi = p.required_span(first, cnt)
            lo = max(lo, 0)
            part = p.run(np.ascontiguousarray(x[lo:hi]), origin=lo, first_frame=first, nframes=cnt)["psd"]
            assert np.array_equal(part, full[first:first + cnt]), (kw, first, cnt)


def test_averaging_modes_vs_oracle(gpu_api):
    x = stream(200000, seed=21)
    n = 1024
    psd_ref = O.periodogram(x, n, 0, 0.5, True)
    for mode in (1, 2, 3):
        for max0 in (0, 1):
            for depth, mn, mx in ((4, 10, 40), (7, 0, 513), (1, 100, 140), (50, 200, 260)):
                p = gpu_api.GramPlan(n=n, window_type=0, overlap=0.5, sub_mean=True, avg_mode=mode, avg_depth=depth,
                                     avg_minbin=mn, avg_maxbin=mx, avg_max0=max0)
                r = p.run(x)
                # averaging parity is judged on the GPU's own float PSD rows (the oracle's avg.c
                # restatement is exact in double), so this isolates the averaging arithmetic
                a, ret, pk, var = O.update_avg(mode, r["psd"], n, depth, mn, mx, max0)
                assert np.allclose(r["avg"][:, mn:mx], a[:, mn:mx], rtol=1e-5, atol=1e-15), (mode, max0, depth)
                assert np.all(r["avg"][:, :mn] == np.float32(1e-15)) and np.all(r["avg"][:, mx:] == np.float32(1e-15))
                assert np.allclose(r["ret"], ret, rtol=1e-9)
                assert np.array_equal(r["peakbin"], pk)
                if mode == 1:
                    assert np.allclose(r["variance"], var, rtol=1e-9, equal_nan=True)
                # and end to end against the double-precision oracle rows
                a2 = O.update_avg(mode, psd_ref, n, depth, mn, mx, max0)[0]
                if mode == 2:
                    assert np.allclose(r["avg"][:, mn:mx], a2[:, mn:mx], rtol=2e-4)


def test_avg_plain_known_answer_via_per_call_api(gpu_api):
    lib = gpu_api.lib()
    ad = gpu_api.AvgData()
    lib.init_avg(C.byref(ad))
    lib.alloc_avg(C.byref(ad), 16, 3)
    pk = C.c_int(0)
    vals, rets = [], []
    for f in range(5):
        psd = np.array([(f + 1) * (b + 1) for b in range(16)], dtype=np.float32)
        rets.append(lib.update_avg_plain(C.byref(ad), 16, psd.ctypes.data, 2, 7, C.byref(pk)))
        avg = np.ctypeslib.as_array(ad.avg, shape=(16,)).copy()
        vals.append(avg[2])
        assert avg[0] == 1e-15 and avg[7] == 1e-15 and pk.value == 6
    assert np.allclose(vals, [1.5, 3.0, 4.5, 6.75, 9.0]) and np.allclose(rets, [2.25, 4.5, 6.75, 10.125, 13.5])
    assert ad.effdepth == 3
    lib.delete_avg(C.byref(ad))
    assert ad.avgwidth == 0


# ------------------------------------------------------------------ per-call drop-in interface
def _per_call_fft(api, x, n, wt, ov, sub_mean, a=0.0, limiter=0, want_phase=False):
    lib = api.lib()
    par = api.FftParams()
    par.n, par.window_type, par.overlap, par.a, par.limiter = n, wt, ov, a, limiter
    lib.glfer_b200_set_autoscale(int(sub_mean))
    lib.fft_init(C.byref(par))
    hop = api.host_hop(n, ov)
    rows, specs, phases = [], [], []
    lib.glfer_b200_set_first_buffer(1)
    for f in range(len(x) // hop):
        blk = np.ascontiguousarray(x[f * hop:(f + 1) * hop]).copy()
        lib.fft_do(blk.ctypes.data, C.byref(par))
        psd = np.empty(n // 2 + 1, np.float32)
        ph = np.empty(n // 2 + 1, np.float32)
        lib.fft_psd(psd.ctypes.data, ph.ctypes.data if want_phase else None, C.byref(par))
        rows.append(psd)
        phases.append(ph)
        specs.append(np.ctypeslib.as_array(par.outbuf, shape=(n,)).copy())
        lib.glfer_b200_set_first_buffer(0)
    lib.fft_close(C.byref(par))
    return np.array(rows), np.array(specs), np.array(phases)


def test_per_call_fft_equals_batch_and_oracle(gpu_api):
    x = stream(12000, seed=31)
    for n, wt, ov, sm, a, lim in ((1024, 0, 0.5, True, 0.0, 0), (512, 7, 0.75, True, 0.0, 0), (1024, 5, 0.3, False, 0.0, 0),
                                  (1024, 2, 0.5, True, 0.02, 1)):
        rows, specs, _ = _per_call_fft(gpu_api, x, n, wt, ov, sm, a, lim)
        ref, spec_ref = O.periodogram(x, n, wt, ov, sm, a, lim, return_spectrum=True)
        assert_psd_close(rows, ref, f"per-call N={n}")
        batch = gpu_api.GramPlan(n=n, window_type=wt, overlap=ov, sub_mean=sm, a=a, limiter=lim).run(x)["psd"]
        assert_psd_close(rows, batch.astype(np.float64), "per-call vs batch")
        # outbuf is the half-complex spectrum: out[k] = Re, out[n-k] = Im (fft.c:196-198)
        scale = np.sqrt(np.mean(np.abs(spec_ref) ** 2, axis=1, keepdims=True))
        assert np.max(np.abs(specs[:, :n // 2 + 1] - spec_ref.real) / scale) < 2e-6
        assert np.max(np.abs(specs[:, n // 2 + 1:] - spec_ref.imag[:, 1:n // 2][:, ::-1]) / scale) < 2e-6


def test_per_call_phase_and_foreign_spectrum(gpu_api):
    x = stream(4096, seed=33)
    rows, specs, phases = _per_call_fft(gpu_api, x, 1024, 0, 0.5, False, want_phase=True)
    ref, spec = O.periodogram(x, 1024, 0, 0.5, False, return_spectrum=True)
    ph_ref = O.phase_from_spectrum(spec, 1024)
    strong = ref > ref.mean(axis=1, keepdims=True)
    d = np.angle(np.exp(1j * (phases - ph_ref)))
    assert np.max(np.abs(d[strong])) < 1e-3
    # fft_psd on a spectrum the caller wrote into outbuf itself (lmp.c / hparma.c do that)
    lib = gpu_api.lib()
    par = gpu_api.FftParams()
    hc = np.random.default_rng(1).standard_normal(256).astype(np.float32)
    par.n = 256
    par.outbuf = hc.ctypes.data_as(C.POINTER(C.c_float))
    psd = np.empty(129, np.float32)
    lib.fft_psd(psd.ctypes.data, None, C.byref(par))
    want = np.empty(129)
    want[0] = hc[0] ** 2 / 256
    want[1:128] = (hc[1:128].astype(np.float64) ** 2 + hc[255:128:-1].astype(np.float64) ** 2) / 256
    want[128] = hc[128] ** 2 / 256
    assert np.allclose(psd, want, rtol=1e-6)


def test_per_call_mtm_equals_oracle(gpu_api):
    lib = gpu_api.lib()
    x = stream(9000, seed=35)
    par = gpu_api.MtmParams()
    par.fft.n, par.fft.window_type, par.fft.overlap = 1024, 5, 0.5
    par.w, par.kmax = 4.0, 7
    lib.glfer_b200_set_autoscale(1)
    lib.mtm_init(C.byref(par))
    lam = 1.0 + np.ctypeslib.as_array(par.sig, shape=(8,))
    assert np.allclose(lam, O.gl_dpss(1024, 4.0, 7)[1], rtol=1e-9)
    # NR-style 1-offset taper matrix window[i+1][k] (mtm.c:118)
    v0 = np.array([par.window[i + 1][0] for i in range(1024)])
    assert abs(np.sum(v0 * v0) - 1.0) < 1e-12
    rows = []
    lib.glfer_b200_set_first_buffer(1)
    for f in range(len(x) // 512):
        blk = np.ascontiguousarray(x[f * 512:(f + 1) * 512]).copy()
        psd = np.empty(513, np.float32)
        lib.mtm_do(blk.ctypes.data, psd.ctypes.data, None, C.byref(par))
        rows.append(psd)
        lib.glfer_b200_set_first_buffer(0)
    lib.mtm_close(C.byref(par))
    assert_psd_close(np.array(rows), O.multitaper(x, 1024, 0.5, 4.0, 7, True), "per-call MTM")


def assert_lmp_close(got, ref, what):
    """The LMP statistic divides by v_hat = (my - sqrt(my^2 - sy)) / 2, a difference of nearly
    equal numbers for a steady tone and a square root near 0 for noise: a 1e-6 relative error
    of the FP32 PSD rows becomes up to ~1e-3 of the statistic (measured with a float32 FFT on
    the CPU: max 9e-4, median 2e-7).  Bar: 2e-3 relative + 2e-3 absolute (the floor value is
    1e-3), median below 1e-5, NaN / inf in the same places."""
    assert got.shape == ref.shape, what
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), what
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))[fin]
    mag = np.abs(ref.astype(np.float64))[fin]
    assert np.all(err <= 2e-3 * mag + 2e-3), (what, float((err / (2e-3 * mag + 2e-3)).max()))
    assert np.median(err / mag) < 1e-5, what


def test_lmp_fixture_kernel_exactness_and_shards(gpu_api):
    """LMP mode (lmp.c): against the reference fixture; the statistic kernel alone against the
    oracle formula on the GPU's own PSD rows (IEEE double: identical floats); frame sub-ranges
    and time shards against the full run (the ring is rebuilt from lmp_av - 1 halo frames)."""
    for n, ov, nl, sm, key, x in ((1024, 0.5, 4, True, "lmp_rows", XG), (512, 0.75, 7, False, "lmp_rows_n512", XG[:12000])):
        p = gpu_api.GramPlan(n=n, mode=gpu_api.MODE_LMP, overlap=ov, sub_mean=sm, lmp_av=nl)
        got = p.run(x)["psd"]
        assert_lmp_close(got, GOLD[key], f"LMP fixture {key}")
        raw = gpu_api.GramPlan(n=n, window_type=5, overlap=ov, sub_mean=sm).run(x)["psd"]
        assert np.array_equal(got.view(np.uint32), O.lmp_statistic(raw, nl).view(np.uint32)), key
        assert np.all(got[:, 0] == np.float32(1e-3))
        sub = p.run(x, first_frame=11, nframes=17)["psd"]
        assert np.array_equal(sub, got[11:28])
    x = stream(200000, seed=43)
    kw = dict(n=2048, mode=gpu_api.MODE_LMP, overlap=0.5, sub_mean=True, lmp_av=5)
    one = gpu_api.run_sharded(x, 1, **kw)
    assert_lmp_close(one["psd"], O.lmp(x, 2048, 0.5, 5, True), "LMP N=2048")
    for shards in (2, 4):
        devs = [g % gpu_api.device_count() for g in range(shards)]
        assert np.array_equal(one["psd"], gpu_api.run_sharded(x, shards, devices=devs, **kw)["psd"]), shards
    # RA9MB / limiter never reach the LMP spectrum (lmp.c:112-114)
    q = gpu_api.GramPlan(n=1024, mode=gpu_api.MODE_LMP, overlap=0.5, sub_mean=True, lmp_av=4, a=0.05, limiter=1)
    assert np.array_equal(q.run(XG)["psd"], gpu_api.GramPlan(n=1024, mode=gpu_api.MODE_LMP, overlap=0.5, sub_mean=True, lmp_av=4).run(XG)["psd"])
    with pytest.raises(Exception):
        gpu_api.GramPlan(n=1024, mode=gpu_api.MODE_LMP, lmp_av=1)


def test_per_call_lmp_equals_batch(gpu_api):
    """lmp_init / lmp_do / lmp_close (lmp.h:49-51) one hop block per call, as source.c:155-156."""
    lib = gpu_api.lib()
    x = stream(12000, seed=36)
    par = gpu_api.LmpParams()
    par.fft.n, par.fft.window_type, par.fft.overlap, par.avg = 1024, 5, 0.5, 4
    lib.glfer_b200_set_autoscale(1)
    lib.lmp_init(C.byref(par))
    rows = []
    lib.glfer_b200_set_first_buffer(1)
    for f in range(len(x) // 512):
        blk = np.ascontiguousarray(x[f * 512:(f + 1) * 512]).copy()
        psd = np.empty(513, np.float32)
        lib.lmp_do(blk.ctypes.data, psd.ctypes.data, None, C.byref(par))
        rows.append(psd)
        lib.glfer_b200_set_first_buffer(0)
    lib.lmp_close(C.byref(par))
    rows = np.array(rows)
    assert_lmp_close(rows, O.lmp(x, 1024, 0.5, 4, True), "per-call LMP")
    batch = gpu_api.GramPlan(n=1024, mode=gpu_api.MODE_LMP, overlap=0.5, sub_mean=True, lmp_av=4).run(x)["psd"]
    assert_lmp_close(rows, batch, "per-call LMP vs batch")


def test_peak_carry_scan(gpu_api):
    """glb_launch_peak_carry (the carried *peakbin of avg.c:129-133) straight through the shim:
    random candidates, long runs of "not written" (-1) across chunk and tile borders, lengths that
    are not multiples of a tile, nothing written at all."""
    lib = gpu_api.lib()
    lib.glb_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    lib.glb_free.argtypes = [C.c_void_p]
    lib.glb_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.glb_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.glb_launch_peak_carry.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
    rng = np.random.default_rng(11)
    for n, p_valid in ((1, 1.0), (255, 0.5), (257, 0.01), (100003, 0.3), (100003, 0.0005), (300000, 0.0), (168750, 0.9)):
        cand = np.where(rng.random(n) < p_valid, rng.integers(0, 2049, n), -1).astype(np.int32)
        if n > 5000:
            cand[1000:4000] = -1                     # a run longer than several tiles
        ref = np.empty(n, np.int32)
        carry = 77
        for i in range(n):
            if cand[i] >= 0:
                carry = cand[i]
            ref[i] = carry
        d_c, d_o = C.c_void_p(), C.c_void_p()
        assert lib.glb_malloc(C.byref(d_c), cand.nbytes) == 0 and lib.glb_malloc(C.byref(d_o), cand.nbytes) == 0
        assert lib.glb_memcpy_h2d(d_c, cand.ctypes.data, cand.nbytes, None) == 0
        assert lib.glb_launch_peak_carry(d_c, d_o, n, 77, None) == 0
        got = np.empty(n, np.int32)
        assert lib.glb_memcpy_d2h(got.ctypes.data, d_o, got.nbytes, None) == 0
        lib.glb_free(d_c)
        lib.glb_free(d_o)
        assert np.array_equal(got, ref), (n, p_valid)


def test_compute_floor(gpu_api):
    lib = gpu_api.lib()
    x = stream(20000, seed=37)
    rows = O.periodogram(x, 4096, 0, 0.5, True)
    for row in rows[:3]:
        s, f, p, b = C.c_float(), C.c_float(), C.c_float(), C.c_uint()
        r = row.copy()
        lib.compute_floor(r.ctypes.data, len(r), C.byref(s), C.byref(f), C.byref(p), C.byref(b))
        s2, f2, p2, b2 = O.compute_floor(row)
        assert s.value == pytest.approx(s2) and p.value == pytest.approx(p2) and b.value == b2
        assert f.value == pytest.approx(f2, rel=1e-5)


# ------------------------------------------------------------------ display mapping
def test_display_levels_vs_oracle(gpu_api):
    """rows -> 8-bit levels as main_window_draw maps them (floor statistics, AGC recurrence,
    integer-dB level buffer, threshold, bin reversal).  Levels are compared to the restatement
    applied to the GPU's own float rows; a pixel may differ by one step when a dB value sits
    on an integer boundary of the `short` level buffer."""
    x = stream(120000, seed=51)
    n = 1024
    pal = (np.arange(768) % 251).astype(np.uint8)
    for kw, disp in ((dict(n=n, window_type=0, overlap=0.5, sub_mean=True), dict(log_scale=True, autoscale=True, thr_level=10.0)),
                     (dict(n=n, window_type=7, overlap=0.75, sub_mean=True), dict(log_scale=False, autoscale=True, thr_level=0.0)),
                     (dict(n=n, window_type=0, overlap=0.5, sub_mean=True), dict(log_scale=True, autoscale=False, max_level_db=-30.0, min_level_db=-90.0, thr_level=5.0)),
                     (dict(n=n, window_type=0, overlap=0.5, sub_mean=True, avg_mode=2, avg_depth=4, avg_minbin=20, avg_maxbin=200),
                      dict(log_scale=True, autoscale=True, thr_level=0.0))):
        p = gpu_api.GramPlan(**kw)
        r = p.run(x)
        shown = r["avg"] if kw.get("avg_mode") else r["psd"]
        d = p.run_display(x, colortab=pal, want_rgb=True, **disp)
        lev, rng, state = O.display_levels(r["psd"], shown, kw["overlap"], disp["log_scale"], disp["autoscale"],
                                           disp.get("max_level_db", -20.0), disp.get("min_level_db", -80.0), disp["thr_level"])
        assert d["levels"].shape == lev.shape
        diff = np.abs(d["levels"].astype(np.int32) - lev.astype(np.int32))
        assert np.mean(diff == 0) > 0.999, (kw, disp, np.mean(diff == 0))
        assert diff.max() <= (8 if disp["log_scale"] else 1)        # one integer-dB step of the level buffer at most
        if disp["autoscale"]:
            assert np.allclose(d["range"], rng, rtol=2e-5, atol=1e-3)
            assert np.allclose(d["agc_state"], state, rtol=2e-5)
        assert np.array_equal(d["rgb"], pal.reshape(256, 3)[d["levels"]])
    # chaining the AGC across two calls equals one call
    p = gpu_api.GramPlan(n=n, window_type=0, overlap=0.5, sub_mean=True)
    full = p.run_display(x, log_scale=True, autoscale=True)
    nf = full["levels"].shape[0]
    a = p.run_display(x, log_scale=True, autoscale=True, nframes=nf // 2)
    b = p.run_display(x, log_scale=True, autoscale=True, first_frame=nf // 2, nframes=nf - nf // 2, agc_state=a["agc_state"])
    assert np.array_equal(np.concatenate([a["levels"], b["levels"]]), full["levels"])


# ------------------------------------------------------------------ WAV source
def test_wav_source_with_stale_tail(gpu_api, tmp_path):
    pcm = GOLD["pcm"][:8000 * 3 + 123]
    path = str(tmp_path / "c1.wav")
    synth.write_wav16(path, pcm, 8000)
    p = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=True)
    r = p.run_wav(path)
    strm, rate, bits = O.read_wav_blocks(path, 512, sub_mean=True)
    assert r["sample_rate"] == 8000 and r["bits"] == 16 and rate == 8000
    ref = O.periodogram(strm, 1024, 0, 0.5, True)
    assert r["psd"].shape == ref.shape == (-(-len(pcm) // 512), 513)
    assert_psd_close(r["psd"], ref, "WAV")
    p0 = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=False)
    strm0, _, _ = O.read_wav_blocks(path, 512)
    assert_psd_close(p0.run_wav(path)["psd"], O.periodogram(strm0, 1024, 0, 0.5, False), "WAV no mean")
    # a stereo file: as the reference (interleaved samples taken as one stream), and one channel of it (extension)
    left = pcm[: len(pcm) // 2 * 2 // 2]
    right = np.roll(left, 137)
    st = np.stack([left, right], axis=1).reshape(-1)
    spath = str(tmp_path / "stereo.wav")
    synth.write_wav16(spath, st, 8000, channels=2)
    strm_i, _, _ = O.read_wav_blocks(spath, 512)
    assert_psd_close(p0.run_wav(spath)["psd"], O.periodogram(strm_i, 1024, 0, 0.5, False), "stereo WAV, interleaved as the reference")
    mono_r = str(tmp_path / "right.wav")
    synth.write_wav16(mono_r, right, 8000)
    assert np.array_equal(p0.run_wav(spath, channel=1)["psd"], p0.run_wav(mono_r)["psd"])


def test_headless_harness_per_call_and_batch(gpu_api, tmp_path):
    """tools/glfer_headless: the reference's per-block loop written against fft.h/mtm.h/avg.h,
    linked with libglfer_b200 (and, in oracle/_ref, with the unmodified reference)."""
    import subprocess
    from glfer_b200 import build
    exe = build.build_tools()
    pcm = GOLD["pcm"][:8000 * 2 + 77]
    wav = str(tmp_path / "t.wav")
    synth.write_wav16(wav, pcm, 8000)
    ref_exe = os.path.join(os.path.dirname(R.path()), "glfer_headless_ref")
    for args in (["-n", "1024", "-w", "0", "-o", "0.5", "-s", "1"], ["-n", "1024", "-m", "mtm", "-k", "7", "-W", "4", "-o", "0.5"],
                 ["-n", "2048", "-w", "7", "-o", "0.75", "-A", "2", "-d", "4"], ["-n", "1024", "-m", "lmp", "-L", "4", "-o", "0.5"]):
        outs = {}
        for tag, extra in (("percall", []), ("batch", ["-B"])):
            out = str(tmp_path / f"{tag}.f32")
            res = subprocess.run([exe, "-f", wav, "-O", out] + args + extra, capture_output=True, text=True)
            assert res.returncode == 0, res.stderr
            outs[tag] = np.fromfile(out, dtype=np.float32)
        assert outs["percall"].shape == outs["batch"].shape
        n = int(args[1])
        a = outs["percall"].reshape(-1, n // 2 + 1).astype(np.float64)
        b = outs["batch"].reshape(-1, n // 2 + 1).astype(np.float64)
        if "-A" in args:
            assert np.allclose(a, b, rtol=5e-4, atol=1e-15)
        elif "lmp" in args:
            assert_lmp_close(a.astype(np.float32), b.astype(np.float32), "LMP per-call vs batch")
        else:
            assert_psd_close(a, b, "per-call vs batch")
        if os.path.exists(ref_exe):
            out = str(tmp_path / "ref.f32")
            subprocess.run([ref_exe, "-f", wav, "-O", out] + args, check=True, capture_output=True)
            r = np.fromfile(out, dtype=np.float32).reshape(a.shape).astype(np.float64)
            if "-A" in args:
                assert np.allclose(a, r, rtol=5e-4, atol=1e-15)
            elif "lmp" in args:
                assert_lmp_close(a.astype(np.float32), r.astype(np.float32), "headless LMP product vs reference build")
            else:
                assert_psd_close(a, r, "headless product vs reference build")


# ------------------------------------------------------------------ time sharding
def test_sharded_run_is_bit_identical_to_one_shard(gpu_api):
    x = stream(300000, seed=41)
    ndev_avail = gpu_api.device_count()
    for kw in (dict(n=4096, window_type=0, overlap=0.5, sub_mean=True),
               dict(n=2048, window_type=7, overlap=0.75, sub_mean=True, avg_mode=2, avg_depth=4, avg_minbin=30, avg_maxbin=90)):
        one = gpu_api.run_sharded(x, 1, **kw)
        for shards in (2, 4, 8):
            devs = [g % ndev_avail for g in range(shards)]
            many = gpu_api.run_sharded(x, shards, devices=devs, **kw)
            assert np.array_equal(one["psd"], many["psd"]), (kw, shards)
            if kw.get("avg_mode"):
                assert np.array_equal(one["avg"], many["avg"])
                assert np.array_equal(one["peakbin"], many["peakbin"])
                assert np.allclose(one["ret"], many["ret"], rtol=1e-12)


# ------------------------------------------------------------------ full-size properties
def test_full_size_properties_parseval_and_linearity(gpu_api):
    """BASELINE metric shape at length the oracle cannot finish quickly: size-independent
    properties instead.  (1) Parseval: sum_k c_k psd[k] = sum_n (w x)^2 / ... per frame with
    the unit-energy Hann window; (2) spot parity on sampled frames; (3) scaling x by 2
    scales every PSD bin by exactly 4 (power of two: exact in float)."""
    n, hop = 4096, 2048
    x = synth.tiled_stream(48000 * 600, fs=48000, block_s=20.0)      # 10 minutes
    p = gpu_api.GramPlan(n=n, window_type=0, overlap=0.5, sub_mean=False)
    p.stage(x)
    nf = p.num_frames(len(x))
    p.exec(0, nf)
    full = p.fetch(nf)["psd"]
    assert full.shape == (nf, 2049) and np.isfinite(full).all()
    rng = np.random.default_rng(0)
    pick = np.unique(np.concatenate([np.arange(0, 64), np.arange(nf - 64, nf), rng.integers(0, nf, 500)]))
    w = O.compute_window(n, 0).astype(np.float64)
    for f in pick[::16]:
        lo = f * hop - (n - hop)
        fr = np.zeros(n)
        src = x[max(lo, 0): lo + n]
        fr[n - len(src):] = src
        energy = np.sum((fr * w) ** 2)
        c = np.full(2049, 2.0)
        c[0] = c[-1] = 1.0
        assert np.sum(c * full[f].astype(np.float64)) == pytest.approx(energy, rel=2e-5)
    ref = O.periodogram(x[: (pick[63] + 1) * hop], n, 0, 0.5, False)
    assert_psd_close(full[:64], ref[:64], "first 64 frames")
    p2 = gpu_api.GramPlan(n=n, window_type=0, overlap=0.5, sub_mean=False)
    twice = p2.run(2.0 * x[: hop * 2000])["psd"]
    assert np.array_equal(twice, 4.0 * full[:2000])
