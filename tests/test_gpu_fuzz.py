"""Seeded random sweep of the batched path against the oracle: FFT size, window, overlap
(regular and odd hops), block-mean removal, estimator mode, frame sub-ranges.  Complements the
targeted parity tests: every kernel family / plan type is reached through the automatic choice."""
import numpy as np
import pytest

from glfer_b200 import synth
from oracle import glfer_oracle as O
from parity import assert_psd_close

pytestmark = pytest.mark.gpu

OVERLAPS = [0.0, 0.5, 0.75, 0.875, 0.9375, 0.25, 0.3, 0.6, 0.9]


def _cases(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        n = int(2 ** rng.integers(5, 16))                  # 32 .. 32768
        ov = float(rng.choice(OVERLAPS))
        mode = int(rng.choice([0, 0, 0, 1, 3]))
        out.append(dict(n=n, overlap=ov, mode=mode, window_type=int(rng.integers(0, 8)), sub_mean=bool(rng.integers(0, 2)),
                        kmax=int(rng.integers(0, 6)), nw=float(rng.choice([2.0, 2.5, 4.0])), nl=int(rng.integers(2, 7)),
                        frames=int(rng.integers(3, 40)), seed=int(rng.integers(1, 1 << 30))))
    return out


@pytest.mark.parametrize("case", _cases(20261018, 36), ids=lambda c: f"n{c['n']}-m{c['mode']}-ov{c['overlap']}")
def test_random_configuration(gpu_api, case):
    n, ov = case["n"], case["overlap"]
    hop = O.hop_size(n, ov)
    if hop < 1:
        pytest.skip("overlap leaves no new samples")
    frames = case["frames"] if n <= 8192 else min(case["frames"], 12)
    x = synth.qrss_stream(frames * hop + int(case["seed"] % hop), fs=48000, seed=case["seed"] % 1000, dot_s=0.05)
    sm = case["sub_mean"]
    # the DC bin of a mean-removed frame is rounding noise: 1 bin in n/2+1
    min_frac = 0.995 - (2.0 / (n // 2 + 1) if sm else 0.0)
    if case["mode"] == 1:
        if n < 64:
            pytest.skip("DPSS tapers need a few points per taper")
        p = gpu_api.GramPlan(n=n, mode=gpu_api.MODE_MTM, overlap=ov, sub_mean=sm, mtm_w=case["nw"], mtm_kmax=case["kmax"])
        ref = O.multitaper(x, n, ov, case["nw"], case["kmax"], sm)
    elif case["mode"] == 3:
        p = gpu_api.GramPlan(n=n, mode=gpu_api.MODE_LMP, overlap=ov, sub_mean=sm, lmp_av=case["nl"])
        raw = gpu_api.GramPlan(n=n, window_type=5, overlap=ov, sub_mean=sm).run(x)["psd"]
        got = p.run(x)["psd"]
        assert_psd_close(raw, O.periodogram(x, n, 5, ov, sm), f"LMP raw rows {case}", min_frac)
        assert np.array_equal(got.view(np.uint32), O.lmp_statistic(raw, case["nl"]).view(np.uint32)), case
        return
    else:
        p = gpu_api.GramPlan(n=n, window_type=case["window_type"], overlap=ov, sub_mean=sm)
        ref = O.periodogram(x, n, case["window_type"], ov, sm)
    got = p.run(x)["psd"]
    assert got.shape == ref.shape
    assert_psd_close(got, ref, str(case), min_frac)
    # a sub-range of the run gives the same rows as the full run
    if got.shape[0] > 4:
        a, b = 1, got.shape[0] - 1
        assert np.array_equal(p.run(x, first_frame=a, nframes=b - a)["psd"], got[a:b]), case
