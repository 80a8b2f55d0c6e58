"""Round-2 parity tests through the C-ABI: the BASELINE configurations at their own FFT sizes
against fixtures from the unmodified reference, frames beyond sample 2^32 of a 24-hour stream,
every averaging mode and the multitaper estimator under time sharding, and the GUI's
autoscale-off sequence (zeroed history on every frame).  All need a B200 (`-m gpu`)."""
import gc
import os
import weakref

import numpy as np
import pytest

from glfer_b200 import shard, synth
from oracle import glfer_oracle as O
from parity import assert_psd_close, assert_psd_strict

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(__file__)
G1 = np.load(os.path.join(HERE, "golden", "glfer_ref_f64.npz"))
G2 = np.load(os.path.join(HERE, "golden", "glfer_ref_f64_r2.npz"))
X8 = synth.pcm16_to_float(G1["pcm"])
X48 = synth.pcm16_to_float(G2["pcm48"])
FS = 48000


# ------------------------------------------------------------------ BASELINE configs at config size
def test_baseline_configs_strict_vs_reference_fixtures(gpu_api):
    """C1..C5 against rows produced by the compiled reference (double path), held to the strict bar."""
    st = {}
    st["c1"] = assert_psd_strict(gpu_api.GramPlan(n=1024, window_type=0, overlap=0.5, sub_mean=True).run(X8)["psd"],
                                 G1["c1_rows"], "C1")
    st["c2"] = assert_psd_strict(gpu_api.GramPlan(n=4096, window_type=7, overlap=0.75, sub_mean=True).run(X8)["psd"],
                                 G1["c2_rows"], "C2")
    st["c3"] = assert_psd_strict(gpu_api.GramPlan(n=4096, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=7).run(X8)["psd"],
                                 G2["c3_rows_4096"], "C3 N=4096")
    st["c4"] = assert_psd_strict(gpu_api.GramPlan(n=16384, window_type=0, overlap=0.5, sub_mean=True).run(X48)["psd"],
                                 G2["c4_rows"], "C4")
    st["c5_nw8"] = assert_psd_strict(gpu_api.GramPlan(n=32768, mode=1, overlap=0.5, sub_mean=True, mtm_w=8.0, mtm_kmax=15).run(X48)["psd"],
                                     G2["c5_rows_nw8"], "C5 NW=8")
    # NW=4 with 16 tapers weights the leaky tapers 10..15 by 1/lambda up to 1.2e8: all terms
    # are positive, so parity still holds
    st["c5_nw4"] = assert_psd_strict(gpu_api.GramPlan(n=32768, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=15).run(X48)["psd"],
                                     G2["c5_rows_nw4"], "C5 NW=4")
    print({k: (v["frac_1e4"], v["max_rel_above_floor"]) for k, v in st.items()})


def test_all_windows_at_n32768(gpu_api):
    for t in range(8):
        got = gpu_api.GramPlan(n=32768, window_type=t, overlap=0.5, sub_mean=True).run(X48)["psd"]
        # rectangular window + block means removed: the DC bin is the sum of two zero-mean blocks, i.e. pure
        # rounding noise in the reference too (1e-12 of the row mean); every other case is held to the strict bar
        check = assert_psd_close if t == O.RECTANGULAR else assert_psd_strict
        check(got[2:3], G2["win_rows_32768"][t:t + 1], f"N=32768 window {O.WINDOW_NAMES[t]}")


def test_zero_history_every_frame_fixture(gpu_api):
    """glfer.first_buffer never cleared: what the GUI does with opt.autoscale == 0 (g_main.c:1111-1120)."""
    got = gpu_api.GramPlan(n=1024, window_type=0, overlap=0.75, sub_mean=False, zero_history=True).run(X8[:20000])["psd"]
    assert got.shape == G2["zero_hist_rows"].shape
    assert_psd_close(got, G2["zero_hist_rows"], "zero history, 75 %")
    got = gpu_api.GramPlan(n=512, window_type=7, overlap=0.9, sub_mean=False, zero_history=True).run(X8[:12000])["psd"]
    assert_psd_close(got, G2["zero_hist_rows_odd"], "zero history, odd hop")
    # with block means (not a GUI combination, but the flag is orthogonal) and multitaper, vs the oracle
    x = X8[:30000]
    got = gpu_api.GramPlan(n=1024, window_type=2, overlap=0.5, sub_mean=True, zero_history=True).run(x)["psd"]
    assert_psd_close(got, O.periodogram(x, 1024, 2, 0.5, True, zero_history=True), "zero history + sub_mean")
    # and it differs from the default sequence on every frame after the first
    dflt = gpu_api.GramPlan(n=1024, window_type=2, overlap=0.5, sub_mean=True).run(x)["psd"]
    # (frame 0 has a zero history either way; the two plans run different kernel families, so equal to rounding)
    assert np.allclose(dflt[0], got[0], rtol=1e-4) and not np.allclose(dflt[1:], got[1:], rtol=1e-3)


# ------------------------------------------------------------------ 24-hour offsets
def _stream_span(lo, hi, seed=0x5EED):
    """samples [lo, hi) of the endless tiled synthetic stream (20 s block), by global index"""
    blk = synth.qrss_stream(20 * FS, FS, seed, dot_s=1.0)
    return blk[np.arange(lo, hi, dtype=np.int64) % len(blk)]


def _check_span(api, kw, first, cnt, what, oracle_fn, strict=True, min_frac=None):
    """frames [first, first + cnt) staged with their true stream origin vs the oracle run on the span"""
    n = kw["n"]
    p = api.GramPlan(**kw)
    hop = p.hop
    lo, hi = p.required_span(first, cnt)
    lo = max(lo, 0)
    x = np.ascontiguousarray(_stream_span(lo, hi))
    got = p.run(x, origin=lo, first_frame=first, nframes=cnt)["psd"]
    # the oracle sees the span as a stream of its own: its frame j is global frame lo / hop + j
    assert lo % hop == 0
    skip = first - lo // hop
    ref = oracle_fn(x, skip, cnt)
    if min_frac is None:
        (assert_psd_strict if strict else assert_psd_close)(got, ref, what)
    else:
        (assert_psd_strict if strict else assert_psd_close)(got, ref, what, min_frac)
    return got


def test_frames_beyond_sample_2_pow_32(gpu_api):
    """Every `long long` index path past 2^31 / 2^32: ring kernel (N=16384 and N=4096), general kernel
    with the block-mean table (odd hop), multitaper, averaging halo."""
    big = (1 << 32) + 12345
    f16 = big // 8192 + 7                       # N=16384 hop 8192: first sample ~4.29e9
    _check_span(gpu_api, dict(n=16384, window_type=0, overlap=0.5, sub_mean=True), f16, 80, "C4 shape beyond 2^32",
                lambda x, s, c: O.periodogram(x, 16384, 0, 0.5, True, first_frame=s, nframes=c))
    f4 = big // 2048 + 3
    _check_span(gpu_api, dict(n=4096, window_type=0, overlap=0.5, sub_mean=True), f4, 96, "metric shape beyond 2^32",
                lambda x, s, c: O.periodogram(x, 4096, 0, 0.5, True, first_frame=s, nframes=c))
    fo = big // 409 + 11                        # overlap 0.9 -> hop 409: general kernel, table means
    _check_span(gpu_api, dict(n=4096, window_type=7, overlap=0.9, sub_mean=True), fo, 64, "odd hop beyond 2^32",
                lambda x, s, c: O.periodogram(x, 4096, 7, 0.9, True, first_frame=s, nframes=c), strict=False)
    _check_span(gpu_api, dict(n=4096, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=7), f4, 64, "C3 beyond 2^32",
                lambda x, s, c: O.multitaper(x, 4096, 0.5, 4.0, 7, True, first_frame=s, nframes=c))
    # averaging: rows beyond 2^32 with the (depth-1)-frame halo, against avg.c restated on the GPU's own rows
    kw = dict(n=4096, window_type=7, overlap=0.75, sub_mean=True, avg_mode=2, avg_depth=4, avg_minbin=34, avg_maxbin=102)
    p = gpu_api.GramPlan(**kw)
    first, cnt = big // 1024 + 5, 64
    lo, hi = p.required_span(first, cnt)
    x = np.ascontiguousarray(_stream_span(lo, hi))
    r = p.run(x, origin=lo, first_frame=first, nframes=cnt)
    p0 = gpu_api.GramPlan(n=4096, window_type=7, overlap=0.75, sub_mean=True)
    rows = p0.run(x, origin=lo, first_frame=first - 3, nframes=cnt + 3)["psd"]
    assert np.array_equal(rows[3:], r["psd"])
    acc = rows[:, 34:102].astype(np.float64)
    want = (acc[0:cnt] + acc[1:cnt + 1] + acc[2:cnt + 2] + acc[3:cnt + 3]) / 5.0        # effdepth + 1 = 5 (avg.c:150)
    assert np.allclose(r["avg"][:, 34:102], want, rtol=1e-6)


def test_24h_recording_shards_first_last_random_frames(gpu_api):
    """the 24 h / 48 kHz recording of C4 (506 250 frames) time-sharded over 8 GPUs: the first
    and last 64 frames of every shard plus random frames, staged at their true 24-hour offsets."""
    n, hop = 16384, 8192
    total = 24 * 3600 * FS // hop
    assert total == 506250
    rng = np.random.default_rng(24)
    kw = dict(n=n, window_type=0, overlap=0.5, sub_mean=True)
    orc = lambda x, s, c: O.periodogram(x, n, 0, 0.5, True, first_frame=s, nframes=c)      # noqa: E731
    for g in range(8):
        first, cnt = shard.frame_range(total, 8, g)
        _check_span(gpu_api, kw, first, 64, f"shard {g} first 64", orc)
        _check_span(gpu_api, kw, first + cnt - 64, 64, f"shard {g} last 64", orc)
        for f in rng.integers(first, first + cnt, 4):
            # one frame = 8193 bins: a single bin beyond 1e-4 is already 1.2e-4 of the frame
            _check_span(gpu_api, kw, int(f), 1, f"shard {g} frame {f}", orc, min_frac=0.9995)


# ------------------------------------------------------------------ sharding: every averaging mode, multitaper
@pytest.mark.parametrize("mode", [1, 2, 3])
@pytest.mark.parametrize("max0", [0, 1])
def test_sharded_averaging_modes_bit_identical(gpu_api, mode, max0):
    x = synth.qrss_stream(300000, fs=FS, seed=43, dot_s=0.2)
    nd = gpu_api.device_count()
    for depth, mn, mx in ((4, 30, 90), (1, 50, 58), (9, 0, 1025)):
        kw = dict(n=2048, window_type=7, overlap=0.75, sub_mean=True, avg_mode=mode, avg_depth=depth, avg_minbin=mn,
                  avg_maxbin=mx, avg_max0=max0, avg_peakbin_init=5)
        one = gpu_api.run_sharded(x, 1, **kw)
        for shards in (2, 8):
            many = gpu_api.run_sharded(x, shards, devices=[g % nd for g in range(shards)], **kw)
            assert np.array_equal(one["psd"], many["psd"])
            assert np.array_equal(one["avg"], many["avg"]), (mode, max0, depth, shards)
            assert np.array_equal(one["peakbin"], many["peakbin"])
            assert np.array_equal(one["ret"], many["ret"])
            assert np.array_equal(one["variance"], many["variance"], equal_nan=True)


def test_sharded_sumavg_with_unresolved_leading_frames(gpu_api):
    """depth 1 and a stream whose strongest bin IS minbin: no frame ever writes *peakbin (avg.c:129-133
    needs cum > psd[minbin]), so every shard starts with the sentinel and the variance of its frames must
    exclude the CALLER's initial bin, as one shard does."""
    n, fs = 1024, 8000
    t = np.arange(120000) / fs
    rng = np.random.default_rng(3)
    x = (0.5 * np.sin(2 * np.pi * (fs / n * 40) * t) + 0.01 * rng.standard_normal(len(t))).astype(np.float32)
    kw = dict(n=n, window_type=0, overlap=0.5, sub_mean=True, avg_mode=1, avg_depth=1, avg_minbin=40, avg_maxbin=80,
              avg_peakbin_init=47)
    one = gpu_api.run_sharded(x, 1, **kw)
    assert (one["peakbin"] == 47).all()                    # never written: the carried initial value
    nd = gpu_api.device_count()
    many = gpu_api.run_sharded(x, 4, devices=[g % nd for g in range(4)], **kw)
    assert np.array_equal(one["peakbin"], many["peakbin"])
    assert np.array_equal(one["variance"], many["variance"], equal_nan=True)
    assert np.array_equal(one["avg"], many["avg"])
    a, ret, pk, var = O.update_avg(1, one["psd"], n, 1, 40, 80, 0, peakbin_init=47)
    assert np.array_equal(pk, one["peakbin"]) and np.allclose(var, one["variance"], rtol=1e-9, equal_nan=True)


def test_lmp_run_kernel_exact_for_every_ring_length(gpu_api):
    """The batch LMP kernel (one thread per bin walking a run of frames, ring in registers, Markstein division by
    nl and nl - 1) against the oracle formula on the GPU's own PSD rows: identical floats for every ring
    length it is instantiated for (2 .. 8) and for the general kernel behind it (9 .. 12), on runs that start
    and end off the run grid, and on a stream with a 200 dB step (sums that are not exact in double, so the
    slot order matters)."""
    x = synth.qrss_stream(512 * 150 + 77, fs=FS, seed=31, dot_s=0.2)
    x[512 * 60:] *= np.float32(1e-10)                       # PSD drops by 200 dB in the middle of the run
    x[512 * 100:] *= np.float32(1e10)
    raw = gpu_api.GramPlan(n=1024, window_type=5, overlap=0.5, sub_mean=True).run(x)["psd"]
    for nl in range(2, 13):
        p = gpu_api.GramPlan(n=1024, mode=gpu_api.MODE_LMP, overlap=0.5, sub_mean=True, lmp_av=nl)
        got = p.run(x)["psd"]
        want = O.lmp_statistic(raw, nl)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (nl, int((got.view(np.uint32) != want.view(np.uint32)).sum()))
        for first, cnt in ((1, 5), (nl - 1, 3 * nl + 1), (37, 70), (got.shape[0] - 3, 3)):
            lo, hi = p.required_span(first, cnt)
            lo = max(lo, 0)
            sub = p.run(np.ascontiguousarray(x[lo:hi]), origin=lo, first_frame=first, nframes=cnt)["psd"]
            assert np.array_equal(sub.view(np.uint32), want[first:first + cnt].view(np.uint32)), (nl, first, cnt)
    db = gpu_api.GramPlan(n=1024, mode=gpu_api.MODE_LMP, overlap=0.5, sub_mean=True, lmp_av=4, scale_db=True).run(x)["psd"]
    lin = O.lmp_statistic(raw, 4)
    fin = np.isfinite(lin)
    assert np.allclose(db[fin], 10.0 * np.log10(lin[fin]), atol=2e-4)


def test_sharded_multitaper_and_lmp_bit_identical(gpu_api):
    x = synth.qrss_stream(200000, fs=FS, seed=44, dot_s=0.2)
    nd = gpu_api.device_count()
    for kw in (dict(n=2048, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=7),
               dict(n=1024, mode=1, overlap=0.75, sub_mean=True, mtm_w=3.0, mtm_kmax=4),
               dict(n=1024, mode=3, overlap=0.5, sub_mean=True, lmp_av=5)):
        one = gpu_api.run_sharded(x, 1, **kw)
        for shards in (2, 8):
            many = gpu_api.run_sharded(x, shards, devices=[g % nd for g in range(shards)], **kw)
            assert np.array_equal(one["psd"], many["psd"]), (kw, shards)


# ------------------------------------------------------------------ housekeeping
def test_pinned_buffer_is_freed_with_its_last_view(gpu_api):
    a = gpu_api.pinned_empty((1000, 3), np.float32)
    owner = a.base
    while not isinstance(owner, gpu_api._PinnedBuffer):
        owner = owner.base
    ref = weakref.ref(owner)
    v = a[10:20]
    del a, owner
    gc.collect()
    assert ref() is not None                       # a view still holds the memory
    v[:] = 1.0
    del v
    gc.collect()
    assert ref() is None


def test_public_api_refuses_uncovered_span(gpu_api):
    """A span that does not cover the requested frames is refused by the host layer; underneath,
    glb_launch_gram only takes the unchecked TMA ring path when the span covers every block."""
    x = synth.qrss_stream(2048 * 40, fs=FS, seed=45)
    p = gpu_api.GramPlan(n=4096, window_type=0, overlap=0.5, sub_mean=False)
    full = p.run(x)["psd"]
    # the public API refuses outright
    with pytest.raises(gpu_api.GlferError):
        p.run(np.ascontiguousarray(x[: 2048 * 20]), origin=0, first_frame=0, nframes=30)
    assert full.shape[0] == 40


# ------------------------------------------------------------------ display mapping (pinned to g_main.c)
def _load_display_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mg2", os.path.join(HERE, "golden", "make_golden_r2.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.DISPLAY_CASES


def _product_display(api, x, kw, fs=8000):
    """the product's glfer_gram_run_display driven with one oracle.ref_gui.draw_rows case (C1 estimator)"""
    st = kw.get("scale_type", 2)
    log, max0 = st in (2, 3), st in (1, 3)
    plan_kw = dict(n=1024, window_type=0, overlap=0.5, sub_mean=True)
    if kw.get("averaging", 0):
        binsize = np.float32(fs) / np.float32(kw["data_block_size"])
        plan_kw.update(avg_mode=kw["averaging"], avg_depth=kw["avgsamples"], avg_minbin=int(np.float32(400.0) / binsize),
                       avg_maxbin=int(np.float32(1200.0) / binsize), avg_max0=int(max0))
    p = api.GramPlan(**plan_kw)
    d = p.run_display(x, log_scale=log, autoscale=kw.get("autoscale", True), max_level_db=kw.get("max_level_db", -20.0),
                      min_level_db=kw.get("min_level_db", -80.0), thr_level=kw.get("thr_level", 0.0))
    r = p.run(x)
    return d, r, plan_kw, log


def test_display_levels_exact_on_own_rows_and_close_to_reference_fixture(gpu_api):
    """Every display variant the compiled reference GUI unit produced a fixture for.  (1) The product's levels
    EQUAL main_window_draw's arithmetic (restatement pinned bit-exact to g_main.c) applied to the product's own
    float rows: floor statistics, AGC recurrence, integer-dB level buffer, threshold, clipping, bin reversal.
    (2) Against the fixture itself (the reference's rows in, so PSD differences of ~1e-6 relative can move a
    dB value across an integer) nearly every pixel is equal and none is more than one dB step away."""
    cases = _load_display_cases()
    for name, kw in cases.items():
        d, r, plan_kw, log = _product_display(gpu_api, X8, kw)
        shown = r["avg"] if plan_kw.get("avg_mode") else r["psd"]
        lev, rng, state = O.display_levels(r["psd"], shown, 0.5, log, kw.get("autoscale", True),
                                           kw.get("max_level_db", -20.0), kw.get("min_level_db", -80.0), kw.get("thr_level", 0.0))
        assert np.array_equal(d["levels"], lev), (name, int((d["levels"] != lev).sum()))
        if kw.get("autoscale", True):
            assert np.array_equal(d["range"], rng), name
        ref = G2[f"disp_{name}"]
        diff = np.abs(d["levels"].astype(np.int32) - ref.astype(np.int32))
        step = 255.0 / max(1.0, float(np.abs(rng[:, 0] - rng[:, 1]).min())) if log else 2.0
        assert np.mean(diff == 0) > 0.995, (name, np.mean(diff == 0))
        assert diff.max() <= np.ceil(step) + 1, (name, diff.max(), step)


def test_display_fused_levels_equal_two_pass(gpu_api):
    """Fixed display range, no averaging: the spectrogram kernel writes the 8-bit levels itself (no float row
    reaches HBM).  The result must equal the two-pass path (float rows -> levels_kernel) byte for byte, in every
    kernel family that can carry the fused epilogue."""
    x = synth.qrss_stream(400000, fs=FS, seed=61, dot_s=0.2)
    pcm = np.rint(x * 32768.0).astype(np.int16)
    for kw in (dict(n=4096, window_type=0, overlap=0.5, sub_mean=True),                    # ring, metric shape
               dict(n=4096, window_type=7, overlap=0.75, sub_mean=True),                   # ring, 75 %
               dict(n=512, window_type=1, overlap=0.875, sub_mean=False),                  # ring, several groups per CTA
               dict(n=16384, window_type=0, overlap=0.5, sub_mean=True),                   # ring, big frames
               dict(n=4096, window_type=0, overlap=0.9, sub_mean=True),                    # general kernel (odd hop)
               dict(n=2048, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=5),    # ring multitaper
               dict(n=32768, mode=1, overlap=0.5, sub_mean=True, mtm_w=8.0, mtm_kmax=3)):  # general multitaper
        p = gpu_api.GramPlan(**kw)
        for disp in (dict(log_scale=True, autoscale=False, max_level_db=-35.0, min_level_db=-75.0, thr_level=0.0),
                     dict(log_scale=True, autoscale=False, max_level_db=-30.0, min_level_db=-90.0, thr_level=12.5),
                     dict(log_scale=False, autoscale=False, max_level_db=-40.0, min_level_db=-70.0, thr_level=0.0)):
            gpu_api.set_fused_levels(True)
            n0 = gpu_api.kernel_launches()
            fused = p.run_display(x, **disp)["levels"]
            n_fused = gpu_api.kernel_launches() - n0
            gpu_api.set_fused_levels(False)
            n0 = gpu_api.kernel_launches()
            two = p.run_display(x, **disp)["levels"]
            n_two = gpu_api.kernel_launches() - n0
            gpu_api.set_fused_levels(True)
            assert np.array_equal(fused, two), (kw, disp, int((fused != two).sum()))
            assert n_fused < n_two                                   # one kernel per chunk instead of two
            assert fused.min() < fused.max()                         # the range actually exercises the mapping
        # 16-bit PCM in, 8-bit levels out: the same bytes
        lv = p.run_display(pcm, log_scale=True, autoscale=False, max_level_db=-35.0, min_level_db=-75.0)["levels"]
        lf = p.run_display(x, log_scale=True, autoscale=False, max_level_db=-35.0, min_level_db=-75.0)["levels"]
        assert np.array_equal(lv, lf)


def test_integer_db_levels_match_host_libm_on_boundaries(gpu_api):
    """(short) (10 log10 x) on the device equals the host's libm even ON the integer-dB boundaries: rows whose
    bins are the floats right at, below and above every threshold 10^(j/10), plus zero, subnormals, the
    float extremes, NaN and inf (x86 conversion overflow -> level of 0 dB)."""
    vals = []
    for j in range(-449, 386):
        c = np.float32(10.0 ** (j / 10.0))
        for k in range(-3, 4):
            v = c
            for _ in range(abs(k)):
                v = np.nextafter(v, np.float32(np.inf if k > 0 else 0), dtype=np.float32)
            vals.append(v)
    vals += [0.0, -1.0, 1e-45, 3e-45, 1.1754942e-38, 1.1754944e-38, 3.4028235e38, 1.0, np.inf, np.nan]
    rng = np.random.default_rng(8)
    vals += list(np.exp(rng.uniform(-100, 80, 20000)).astype(np.float32))
    nb = 1000
    row = np.resize(np.array(vals, dtype=np.float32), (-(-len(vals) // nb), nb))
    with np.errstate(divide="ignore", invalid="ignore"):
        d = 10.0 * np.log10(row.astype(np.float64))
    s_ref = np.where(np.abs(d) < 2147483648.0, np.trunc(d), 0.0).astype(np.float32)
    for mx, mn, thr in ((-20.0, -80.0, 0.0), (30.0, -300.0, 7.0), (380.0, -440.0, 0.0)):
        got = gpu_api.map_levels(row, True, mx, mn, thr)
        dmax, dmin = np.float32(10.0 * np.log10(float(np.float32(10.0 ** (mx / 10.0))))), np.float32(10.0 * np.log10(float(np.float32(10.0 ** (mn / 10.0)))))
        t = np.float32(np.float32(thr) / 100.0)
        fl = (np.float32(255) * ((s_ref - dmin) / (dmax - dmin))).astype(np.float32)
        sc = (fl.astype(np.float64) - 255.0 * float(t)) / (1.0 - float(t))
        want = np.where(fl.astype(np.float64) < 255.0 * float(t), 0, np.where(fl > 255, 255, sc)).astype(np.int64).astype(np.uint8)
        assert np.array_equal(got, want[:, ::-1]), (mx, mn, thr, int((got != want[:, ::-1]).sum()))
    # explicit per-row ranges (the autoscale form: no look-up table, arithmetic per pixel)
    pr = np.tile(np.array([[-15.5, -95.25]], dtype=np.float32), (row.shape[0], 1))
    got = gpu_api.map_levels(row, True, display_range=pr)
    fl = (np.float32(255) * ((s_ref - pr[0, 1]) / (pr[0, 0] - pr[0, 1]))).astype(np.float32)
    want = np.where(fl < 0, 0, np.where(fl > 255, 255, fl)).astype(np.int64).astype(np.uint8)
    assert np.array_equal(got, want[:, ::-1])


# ------------------------------------------------------------------ band-only averaged rows
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_band_only_averaged_rows_equal_the_band_of_full_rows(gpu_api, mode):
    x = synth.qrss_stream(400000, fs=FS, seed=71, dot_s=0.2)
    for depth, mn, mx, db in ((4, 34, 102, False), (40, 0, 2049, False), (3, 100, 101, True)):
        kw = dict(n=4096, window_type=7, overlap=0.75, sub_mean=True, avg_mode=mode, avg_depth=depth, avg_minbin=mn,
                  avg_maxbin=mx, avg_max0=1, scale_db=db)
        full = gpu_api.GramPlan(**kw).run(x)
        band = gpu_api.GramPlan(avg_band_only=True, **kw).run(x)
        assert band["avg"].shape == (full["avg"].shape[0], mx - mn)
        assert np.array_equal(band["avg"], full["avg"][:, mn:mx], equal_nan=True)
        for k in ("psd", "ret", "peakbin", "variance"):
            assert np.array_equal(band[k], full[k], equal_nan=True), k
    nd = gpu_api.device_count()
    kw = dict(n=2048, window_type=0, overlap=0.5, sub_mean=True, avg_mode=mode, avg_depth=5, avg_minbin=20, avg_maxbin=90,
              avg_band_only=True)
    one = gpu_api.run_sharded(x, 1, **kw)
    many = gpu_api.run_sharded(x, 4, devices=[g % nd for g in range(4)], **kw)
    assert one["avg"].shape[1] == 70 and np.array_equal(one["avg"], many["avg"])


# ------------------------------------------------------------------ the 32-points-per-thread kernel family
def test_big_frame_kernel_vs_oracle_and_16_point_families(gpu_api):
    """N = 16384 / 32768 periodograms at 0 / 50 / 75 % overlap run on gram_big_kernel (family 5) by default:
    against the oracle, against the 16-point ring and general kernels, on ragged frame counts, with and
    without block means, dB output, sub-ranges staged at an offset, and zero history at the stream start."""
    x = synth.qrss_stream(16384 * 41 + 777, fs=FS, seed=81, dot_s=0.2)
    for n in (16384, 32768):
        for ov, wt, sm in ((0.5, 0, True), (0.75, 7, True), (0.0, 1, True), (0.5, 6, False)):
            kw = dict(n=n, window_type=wt, overlap=ov, sub_mean=sm)
            gpu_api.set_kernel_preference(0)
            p = gpu_api.GramPlan(**kw)
            got = p.run(x)["psd"]
            assert gpu_api.last_kernel_family().startswith("gram_big_kernel"), gpu_api.last_kernel_family()
            ref = O.periodogram(x, n, wt, ov, sm)
            assert got.shape == ref.shape
            assert_psd_close(got, ref, f"big kernel N={n} ov={ov}")
            for pref in (1, 2):
                gpu_api.set_kernel_preference(pref)
                other = gpu_api.GramPlan(**kw).run(x)["psd"]
                assert not gpu_api.last_kernel_family().startswith("gram_big_kernel")
                st = assert_psd_close(got, other, f"big vs family {pref}")
            gpu_api.set_kernel_preference(0)
            # a sub-range staged at its own origin equals the rows of the full run (zero history only at frame 0)
            nf = got.shape[0]
            for first, cnt in ((0, 3), (5, nf - 7), (nf - 2, 2)):
                lo, hi = p.required_span(first, cnt)
                lo = max(lo, 0)
                part = p.run(np.ascontiguousarray(x[lo:hi]), origin=lo, first_frame=first, nframes=cnt)["psd"]
                assert np.array_equal(part, got[first:first + cnt]), (kw, first, cnt)
    p = gpu_api.GramPlan(n=16384, window_type=0, overlap=0.5, sub_mean=True, scale_db=True)
    lin = gpu_api.GramPlan(n=16384, window_type=0, overlap=0.5, sub_mean=True).run(x)["psd"]
    assert np.allclose(p.run(x)["psd"], 10.0 * np.log10(lin), atol=2e-4)
    # time shards on the big kernel are bit-identical to one shard
    nd = gpu_api.device_count()
    kw = dict(n=16384, window_type=0, overlap=0.5, sub_mean=True)
    one = gpu_api.run_sharded(x, 1, **kw)
    many = gpu_api.run_sharded(x, 8, devices=[g % nd for g in range(8)], **kw)
    assert np.array_equal(one["psd"], many["psd"])


def test_big_frame_pair_variant_is_bit_identical(gpu_api):
    """N = 16384 periodograms take the two-groups-per-CTA variant of gram_big_kernel (half taper in shared
    memory, second half mirrored) by default; kernel preference 6 keeps one group per CTA with the full taper
    from global memory.  Same arithmetic on the same bits: rows must be identical for every window type (the
    tables are mirror-symmetric bit for bit), every overlap, frame counts below / not a multiple of the
    resident group count, dB rows and staged sub-ranges."""
    x = synth.qrss_stream(16384 * 23 + 4321, fs=FS, seed=83, dot_s=0.2)
    long = synth.qrss_stream(16384 * 340 + 99, fs=FS, seed=84, dot_s=0.2)      # > 2 x 148 groups, ragged
    try:
        for wt in range(8):
            for ov, sm in ((0.5, True), (0.75, False), (0.0, True)):
                kw = dict(n=16384, window_type=wt, overlap=ov, sub_mean=sm)
                gpu_api.set_kernel_preference(0)
                a = gpu_api.GramPlan(**kw).run(x)["psd"]
                assert gpu_api.last_kernel_family().startswith("gram_big_kernel")
                gpu_api.set_kernel_preference(6)
                b = gpu_api.GramPlan(**kw).run(x)["psd"]
                assert gpu_api.last_kernel_family().startswith("gram_big_kernel")
                assert np.array_equal(a, b), (wt, ov, sm, int((a != b).sum()))
        for nfr in (1, 2, 3, 297, None):
            for db in (False, True):
                kw = dict(n=16384, window_type=0, overlap=0.5, sub_mean=True, scale_db=db)
                gpu_api.set_kernel_preference(0)
                p0 = gpu_api.GramPlan(**kw)
                nf = p0.num_frames(len(long)) if nfr is None else nfr
                lo, hi = p0.required_span(7, nf - 7 if nfr is None else nf)
                cnt = nf - 7 if nfr is None else nf
                a = p0.run(np.ascontiguousarray(long[lo:hi]), origin=lo, first_frame=7, nframes=cnt)["psd"]
                gpu_api.set_kernel_preference(6)
                b = gpu_api.GramPlan(**kw).run(np.ascontiguousarray(long[lo:hi]), origin=lo, first_frame=7, nframes=cnt)["psd"]
                assert a.shape[0] == cnt and np.array_equal(a, b), (nfr, db)
    finally:
        gpu_api.set_kernel_preference(0)


# ------------------------------------------------------------------ per-call interface: k blocks per call
def test_fft_do_batch_equals_the_sequence_of_single_calls(gpu_api):
    import ctypes as C
    lib = gpu_api.lib()
    x = synth.qrss_stream(60000, fs=8000, seed=91, dot_s=0.2)
    for n, wt, ov, sm, a, lim in ((1024, 0, 0.5, 1, 0.0, 0), (512, 7, 0.75, 1, 0.0, 0), (1024, 2, 0.9, 0, 0.0, 0),
                                  (2048, 0, 0.5, 1, 0.02, 1)):
        hop = gpu_api.host_hop(n, ov)
        nb = len(x) // hop
        bins = n // 2 + 1

        def fresh():
            p = gpu_api.FftParams()
            p.n, p.window_type, p.overlap, p.a, p.limiter = n, wt, ov, a, lim
            lib.glfer_b200_set_autoscale(sm)
            lib.glfer_b200_set_first_buffer(1)
            lib.fft_init(C.byref(p))
            return p
        # one block per call
        p1 = fresh()
        single = np.empty((nb, bins), np.float32)
        blocks1 = x[: nb * hop].copy().reshape(nb, hop)
        for b in range(nb):
            lib.fft_do(blocks1[b].ctypes.data, C.byref(p1))
            lib.fft_psd(single[b].ctypes.data, None, C.byref(p1))
            lib.glfer_b200_set_first_buffer(0)
        # the same blocks in three batched calls (13, then the bulk, then 1)
        p2 = fresh()
        batched = np.empty((nb, bins), np.float32)
        blocks2 = x[: nb * hop].copy().reshape(nb, hop)
        done = 0
        for k in (13, nb - 14, 1):
            lib.fft_do_batch(blocks2[done].ctypes.data, k, batched[done].ctypes.data, C.byref(p2))
            done += k
        assert done == nb
        assert_psd_close(batched, single, f"fft_do_batch n={n} ov={ov}")
        assert np.array_equal(blocks1, blocks2)                       # block means removed in place, identically
        h1 = np.ctypeslib.as_array(p1.inbuf_audio, shape=(n,)).copy()
        h2 = np.ctypeslib.as_array(p2.inbuf_audio, shape=(n,)).copy()
        assert np.array_equal(h1, h2)                                 # the overlap history the next call starts from
        o1 = np.ctypeslib.as_array(p1.outbuf, shape=(n,)).copy()
        o2 = np.ctypeslib.as_array(p2.outbuf, shape=(n,)).copy()
        assert np.array_equal(o1, o2)                                 # half-complex spectrum of the last frame
        assert_psd_close(single, O.periodogram(x, n, wt, ov, bool(sm), a, lim), "per call vs oracle")
        lib.fft_close(C.byref(p1))
        lib.fft_close(C.byref(p2))


def test_compute_floor_bit_exact_and_autoscale_display_on_wide_rows(gpu_api):
    """compute_floor through the radix-select kernel: sig / floor / peak EQUAL the restatement (pinned bit for bit
    to the compiled reference) on rows of every width, with ties, zeros and a row of equal values."""
    import ctypes as C
    lib = gpu_api.lib()
    rng = np.random.default_rng(12)
    for nb in (17, 33, 257, 513, 544, 545, 1056, 1057, 2049, 2080, 2081, 8193, 16385):
        rows = [(rng.standard_normal(nb) ** 2 * 1e-6).astype(np.float32), np.full(nb, 3e-7, np.float32),
                np.zeros(nb, np.float32), (rng.integers(0, 4, nb) * 1e-5).astype(np.float32),
                np.full(nb, 3e-7, np.float32), (rng.standard_normal(nb) ** 2 * 10.0 ** rng.uniform(-30, 5, nb)).astype(np.float32)]
        rows[0][nb // 3] = 2e-3
        rows[4][: nb // 50] = (rng.random(nb // 50) * 1e-7).astype(np.float32)    # a few bins below a wide tie at the threshold
        rows[4] = rng.permutation(rows[4])
        for row in rows:
            s, f, p, b = C.c_float(), C.c_float(), C.c_float(), C.c_uint()
            r = row.copy()
            lib.compute_floor(r.ctypes.data, len(r), C.byref(s), C.byref(f), C.byref(p), C.byref(b))
            s2, f2, p2, b2 = O.compute_floor(row)
            assert (np.float32(s.value), np.float32(f.value), np.float32(p.value), b.value) == \
                (np.float32(s2), np.float32(f2), np.float32(p2), b2), (nb, s.value, s2, f.value, f2)


# ------------------------------------------------------------------ harmonic F-test
def _ftest_stats(got, ref):
    fin = np.isfinite(ref)
    rel = np.abs(got[fin].astype(np.float64) - ref[fin]) / np.maximum(ref[fin], 1e-30)
    return fin, rel


def test_mtm_harmonic_ftest_vs_reference_fixture(gpu_api):
    """Thomson's harmonic F-test, computed by mtm_do into a file-static nothing reads (mtm.c:204-233), made live:
    against the values the compiled reference left in that static (oracle/ref_mtm_unit.c), double path.
    The statistic is a ratio whose denominator is what is LEFT of |y_k|^2 after removing the line component,
    so float32 spectra reproduce it less closely on the strongest lines (measured: median 2e-7, worst bin 1.5e-4)."""
    for n, key in ((1024, "c3_ftest_1024"), (4096, "c3_ftest_4096")):
        p = gpu_api.GramPlan(n=n, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=7, mtm_ftest=True)
        r = p.run_mtm_ftest(X8)
        ref = G2[key]
        assert r["ftest"].shape == ref.shape
        fin, rel = _ftest_stats(r["ftest"], ref)
        assert np.array_equal(np.isfinite(r["ftest"]), fin)            # inf at Nyquist (never accumulated denominator)
        assert not fin[:, -1].any() and fin[:, :-1].all()
        print(n, "ftest rel err: median %.2e  p99 %.2e  max %.2e  (F max %.1f)" % (np.median(rel), np.quantile(rel, 0.99), rel.max(), ref[fin].max()))
        # measured on B200: median 2e-7, 99th percentile 2e-6, worst bin 1.5e-4 (F = 493)
        assert np.median(rel) < 2e-6 and np.quantile(rel, 0.99) < 2e-5 and rel.max() < 1e-3
        # the multitaper rows of the same call are the plan's usual rows
        assert np.array_equal(r["psd"], p.run(X8)["psd"])
    # odd hop (block-mean table path) and a sub-range, against the oracle restatement
    x = X8[:30000]
    p = gpu_api.GramPlan(n=1024, mode=1, overlap=0.9, sub_mean=True, mtm_w=3.0, mtm_kmax=4, mtm_ftest=True)
    tap, lam = p.tapers()
    ref = O.multitaper_ftest(x, 1024, 0.9, 3.0, 4, True, tapers=tap, lam=lam)
    got = p.run_mtm_ftest(x)["ftest"]
    fin, rel = _ftest_stats(got, ref)
    assert np.median(rel) < 2e-6 and np.quantile(rel, 0.99) < 1e-4
    lo, hi = p.required_span(40, 25)
    part = p.run_mtm_ftest(np.ascontiguousarray(x[max(lo, 0):hi]), origin=max(lo, 0), first_frame=40, nframes=25)["ftest"]
    assert np.array_equal(part, got[40:65], equal_nan=True)
    with pytest.raises(gpu_api.GlferError):
        gpu_api.GramPlan(n=1024, mode=1, overlap=0.5, mtm_kmax=4).run_mtm_ftest(x)      # plan without mtm_ftest


# ------------------------------------------------------------------ the FFTW (double buffers) struct layout
def test_fftw_layout_library_per_call_interface(gpu_api):
    """libglfer_b200_fftw.so: fft_params_t as the reference built WITH FFTW lays it out (fft.h:36-48: plan first,
    fftw_real = double buffers, outbuf its own buffer).  The same calls, the same rows as the float-layout
    library; the buffers a caller reads (g_scope.c:189-197) are doubles."""
    import ctypes as C
    path = os.path.join(os.path.dirname(gpu_api.LIB_PATH), "libglfer_b200_fftw.so")
    assert os.path.exists(path)
    lf = C.CDLL(path)

    class FftParamsFFTW(C.Structure):
        _fields_ = [("plan", C.c_void_p), ("inbuf_audio", C.POINTER(C.c_double)), ("inbuf_fft", C.POINTER(C.c_double)),
                    ("outbuf", C.POINTER(C.c_double)), ("n", C.c_int), ("window", C.POINTER(C.c_float)),
                    ("window_type", C.c_int), ("overlap", C.c_float), ("a", C.c_float), ("limiter", C.c_int),
                    ("sub_mean", C.c_int)]

    class MtmParamsFFTW(C.Structure):
        _fields_ = [("fft", FftParamsFFTW), ("window", C.POINTER(C.POINTER(C.c_double))), ("sig", C.POINTER(C.c_double)),
                    ("w", C.c_float), ("kmax", C.c_int)]
    x = X8[:30000]
    for n, wt, ov, a, lim in ((1024, 0, 0.5, 0.0, 0), (512, 7, 0.75, 0.02, 1)):
        hop = gpu_api.host_hop(n, ov)
        nb, bins = len(x) // hop, n // 2 + 1
        p = FftParamsFFTW()
        p.n, p.window_type, p.overlap, p.a, p.limiter = n, wt, ov, a, lim
        lf.glfer_b200_set_autoscale(1)
        lf.glfer_b200_set_first_buffer(1)
        lf.fft_init(C.byref(p))
        rows = np.empty((nb, bins), np.float32)
        blocks = x[: nb * hop].copy().reshape(nb, hop)
        for b in range(nb):
            lf.fft_do(blocks[b].ctypes.data_as(C.c_void_p), C.byref(p))
            lf.fft_psd(rows[b].ctypes.data_as(C.c_void_p), None, C.byref(p))
            lf.glfer_b200_set_first_buffer(0)
        ref, spec = O.periodogram(x, n, wt, ov, True, a, lim, return_spectrum=True)
        assert_psd_close(rows, ref, f"FFTW layout n={n}")
        hist = np.ctypeslib.as_array(p.inbuf_audio, shape=(n,))
        assert hist.dtype == np.float64
        want = O.gather_frames(x, n, ov, True)[-1]
        assert np.array_equal(hist, want.astype(np.float64))          # doubles holding the float samples
        assert C.addressof(p.outbuf.contents) != C.addressof(p.inbuf_fft.contents)      # not aliased (fft.c:173)
        out = np.ctypeslib.as_array(p.outbuf, shape=(n,)).copy()
        scale = np.sqrt(np.mean(np.abs(spec[-1]) ** 2))
        assert np.abs(out[1:n // 2] - spec[-1].real[1:n // 2]).max() < 3e-6 * scale * np.sqrt(n)
        assert np.abs(out[n - 1:n // 2:-1] - spec[-1].imag[1:n // 2]).max() < 3e-6 * scale * np.sqrt(n)
        lf.fft_close(C.byref(p))
    # multitaper through the same layout
    m = MtmParamsFFTW()
    m.fft.n, m.fft.window_type, m.fft.overlap, m.w, m.kmax = 1024, 5, 0.5, 4.0, 7
    lf.glfer_b200_set_first_buffer(1)
    lf.mtm_init(C.byref(m))
    nb = len(x) // 512
    rows = np.empty((nb, 513), np.float32)
    blocks = x[: nb * 512].copy().reshape(nb, 512)
    for b in range(nb):
        lf.mtm_do(blocks[b].ctypes.data_as(C.c_void_p), rows[b].ctypes.data_as(C.c_void_p), None, C.byref(m))
        lf.glfer_b200_set_first_buffer(0)
    assert_psd_close(rows, O.multitaper(x, 1024, 0.5, 4.0, 7, True), "FFTW layout mtm_do")
    lf.mtm_close(C.byref(m))


def test_32_point_kernel_on_small_frames(gpu_api):
    """kernel preference 5 runs the 32-point family at N = 4096 / 8192 too (64 / 128 threads per frame, up to eight
    frames in flight per SM): an experiment kept selectable; the 16-point ring kernel stays the default there."""
    x = synth.qrss_stream(4096 * 300 + 17, fs=FS, seed=85, dot_s=0.2)
    try:
        for n, ov, wt in ((4096, 0.5, 0), (8192, 0.75, 7), (4096, 0.0, 2)):
            gpu_api.set_kernel_preference(5)
            got = gpu_api.GramPlan(n=n, window_type=wt, overlap=ov, sub_mean=True).run(x)["psd"]
            assert gpu_api.last_kernel_family().startswith("gram_big_kernel")
            assert_psd_close(got, O.periodogram(x, n, wt, ov, True), f"32-point kernel N={n}")
            gpu_api.set_kernel_preference(0)
            ring = gpu_api.GramPlan(n=n, window_type=wt, overlap=ov, sub_mean=True).run(x)["psd"]
            assert gpu_api.last_kernel_family().startswith("gram_ring_kernel")
            assert_psd_close(got, ring, "32-point vs ring")
        gpu_api.set_kernel_preference(5)
        p = gpu_api.GramPlan(n=4096, mode=1, overlap=0.5, sub_mean=True, mtm_w=4.0, mtm_kmax=7)
        got = p.run(x[:200000])["psd"]
        tap, lam = p.tapers()
        assert_psd_close(got, O.multitaper(x[:200000], 4096, 0.5, 4.0, 7, True, tapers=tap, lam=lam), "32-point multitaper N=4096")
    finally:
        gpu_api.set_kernel_preference(0)


def test_big_frame_multitaper_kernel(gpu_api):
    """multitaper at N = 16384 / 32768 on the 32-point kernel (eigenspectra summed in shared memory, the frame
    re-landed by TMA per taper): against the oracle and the 16-point general kernel, 0 / 50 / 75 % overlap."""
    x = synth.qrss_stream(32768 * 9 + 333, fs=FS, seed=83, dot_s=0.2)
    for n, ov, kmax, nw in ((16384, 0.5, 7, 4.0), (32768, 0.5, 15, 8.0), (32768, 0.75, 3, 3.0), (16384, 0.0, 2, 2.5)):
        kw = dict(n=n, mode=1, overlap=ov, sub_mean=True, mtm_w=nw, mtm_kmax=kmax)
        gpu_api.set_kernel_preference(0)
        p = gpu_api.GramPlan(**kw)
        got = p.run(x)["psd"]
        assert gpu_api.last_kernel_family().startswith("gram_big_kernel"), gpu_api.last_kernel_family()
        tap, lam = p.tapers()
        ref = O.multitaper(x, n, ov, nw, kmax, True, tapers=tap, lam=lam)
        assert_psd_close(got, ref, f"big multitaper N={n} ov={ov} K'={kmax + 1}")
        gpu_api.set_kernel_preference(1)
        other = gpu_api.GramPlan(**kw).run(x)["psd"]
        gpu_api.set_kernel_preference(0)
        assert_psd_close(got, other, "big multitaper vs general kernel")
        nf = got.shape[0]
        lo, hi = p.required_span(2, nf - 3)
        part = p.run(np.ascontiguousarray(x[max(lo, 0):hi]), origin=max(lo, 0), first_frame=2, nframes=nf - 3)["psd"]
        assert np.array_equal(part, got[2:nf - 1])


# ------------------------------------------------------------------ averaging fused into the ring kernel
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_fused_averaging_equals_two_pass(gpu_api, mode):
    """avg_band_only plans at N = 4096 / 8192: the ring kernel averages the frames it computes (band history in
    shared memory, one warp per frame).  Rows, return values, peak bins and variances must EQUAL the two-pass
    path (float PSD rows -> averaging kernel): both run avg_frame_warp()."""
    x = synth.qrss_stream(2_000_000, fs=FS, seed=93, dot_s=0.2)
    for n, wt, ov, depth, mn, mx in ((4096, 7, 0.75, 4, 34, 102), (4096, 0, 0.5, 7, 300, 360), (8192, 0, 0.5, 3, 68, 204),
                                     (4096, 1, 0.875, 2, 0, 40)):
        kw = dict(n=n, window_type=wt, overlap=ov, sub_mean=True, avg_mode=mode, avg_depth=depth, avg_minbin=mn, avg_maxbin=mx,
                  avg_max0=1, avg_band_only=True, avg_peakbin_init=mn + 1)
        gpu_api.set_fused_avg(True)
        n0 = gpu_api.kernel_launches()
        fused = gpu_api.GramPlan(**kw).run(x)
        n_fused = gpu_api.kernel_launches() - n0
        gpu_api.set_fused_avg(False)
        n0 = gpu_api.kernel_launches()
        two = gpu_api.GramPlan(**kw).run(x)
        n_two = gpu_api.kernel_launches() - n0
        assert n_fused < n_two, (n_fused, n_two)
        gpu_api.set_fused_avg(True)
        for k in ("psd", "avg", "ret", "peakbin", "variance"):
            assert np.array_equal(fused[k], two[k], equal_nan=True), (kw, k)
        # sub-range at an offset: the groups' pre-roll reads before the first requested frame
        p = gpu_api.GramPlan(**kw)
        nf = fused["psd"].shape[0]
        first, cnt = nf // 3, nf // 2
        lo, hi = p.required_span(first, cnt)
        part = p.run(np.ascontiguousarray(x[max(lo, 0):hi]), origin=max(lo, 0), first_frame=first, nframes=cnt)
        assert np.array_equal(part["avg"], fused["avg"][first:first + cnt])
        assert np.array_equal(part["ret"], fused["ret"][first:first + cnt])
    gpu_api.set_fused_avg(False)
    # against avg.c restated on the GPU's own rows
    a, ret, pk, var = O.update_avg(mode, fused["psd"], 4096, 2, 0, 40, 1, peakbin_init=1)
    assert np.allclose(fused["avg"], a[:, 0:40], rtol=1e-5, atol=1e-15) and np.array_equal(fused["peakbin"], pk)
