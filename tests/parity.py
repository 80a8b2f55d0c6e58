"""Parity criterion shared by the tests (BASELINE.json: 'PSD values within 1e-4 relative
per bin (or 0.01 dB) against the reference's double-precision path').

The GPU computes the FFT in float32 with exact (double-derived) twiddles; its error per
bin is absolute, ~3e-7 of the frame's RMS spectrum amplitude, so bins far below the
frame's mean power cannot meet a *relative* bound ('Precision').  The test
therefore checks, per row:
  * every bin at or above FLOOR x (row mean power): within 0.01 dB of the oracle;
  * bins below that floor: absolute error below 0.01 dB of the floor level;
  * at least MIN_FRAC of all bins within 1e-4 relative.
and returns the statistics so callers can print them."""
import numpy as np

RTOL = 1e-4
DB_TOL = 0.01
FLOOR = 1e-6
MIN_FRAC = 0.995
# BASELINE configurations (C1..C5 and the metric shape) are held to the stated bar itself: EVERY bin
# is judged relatively (0.01 dB), with no exemption for bins below the floor (their count is reported:
# the reference's own rows hold one such bin in C2, a Kaiser side-lobe null), and the share of bins
# within 1e-4 relative must be what was measured on B200 in round 1/2 (>= 99.99 %)
STRICT_MIN_FRAC = 0.9999


def psd_stats(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    mean = np.maximum(ref.mean(axis=-1, keepdims=True), 1e-300)
    floor = FLOOR * mean
    big = ref >= floor
    err = np.abs(got - ref)
    rel = err / np.maximum(ref, 1e-300)
    db_lin = 10.0 ** (DB_TOL / 10.0) - 1.0
    ok_big = rel <= db_lin
    ok_small = err <= db_lin * floor
    ok = np.where(big, ok_big, ok_small)
    frac = float(np.mean(rel <= RTOL)) if rel.size else 1.0
    return dict(ok=bool(ok.all()), nbad=int((~ok).sum()), frac_1e4=frac, nbelow=int((~big).sum()),
                nbad_relative=int((rel > db_lin).sum()), max_rel=float(rel.max()) if rel.size else 0.0,
                max_rel_above_floor=float(rel[big].max()) if big.any() else 0.0,
                median_rel=float(np.median(rel)) if rel.size else 0.0)


def assert_psd_close(got, ref, what="", min_frac=MIN_FRAC):
    """min_frac: with block means removed and no overlap the DC bin of every row is pure rounding
    noise; at N = 32 that is 1 bin in 17, so callers sweeping tiny sizes lower the share."""
    assert np.isfinite(np.asarray(got)).all(), f"{what}: non-finite PSD"
    st = psd_stats(got, ref)
    assert st["ok"], f"{what}: {st}"
    assert st["frac_1e4"] >= min_frac, f"{what}: {st}"
    return st


def assert_psd_strict(got, ref, what="", min_frac=STRICT_MIN_FRAC):
    """The bar for the BASELINE configurations: every bin within 0.01 dB *relative* -- bins below the
    floor get no absolute-error exemption -- and >= min_frac of the bins within 1e-4 relative."""
    st = assert_psd_close(got, ref, what, min_frac)
    assert st["nbad_relative"] == 0, f"{what}: {st['nbad_relative']} bins beyond 0.01 dB relative: {st}"
    return st
