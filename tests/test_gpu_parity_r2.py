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
        assert_psd_strict(got[2:3], G2["win_rows_32768"][t:t + 1], f"N=32768 window {O.WINDOW_NAMES[t]}")


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
    assert np.array_equal(dflt[0], got[0]) and not np.allclose(dflt[1:], got[1:], rtol=1e-3)


# ------------------------------------------------------------------ 24-hour offsets
def _stream_span(lo, hi, seed=0x5EED):
    """samples [lo, hi) of the endless tiled synthetic stream (20 s block), by global index"""
    blk = synth.qrss_stream(20 * FS, FS, seed, dot_s=1.0)
    return blk[np.arange(lo, hi, dtype=np.int64) % len(blk)]


def _check_span(api, kw, first, cnt, what, oracle_fn, strict=True):
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
    (assert_psd_strict if strict else assert_psd_close)(got, ref, what)
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
            _check_span(gpu_api, kw, int(f), 1, f"shard {g} frame {f}", orc)


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
