"""CPU restatement (numpy) of glfer's spectrum-estimator hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product library
(glfer_b200/) never does and has no CPU fallback.

Every function cites the reference lines (file:line under /root/reference) it
restates.  The arithmetic follows the reference's *double-precision path*
(-DHAVE_LIBRFFTW: fftw_real = double buffers, float window table, float PSD
output), which is what BASELINE.json names as the parity target.

Pinning: this restatement is checked in tests/test_oracle.py against
  (a) outputs of the unmodified reference sources built by oracle/Makefile
      (oracle/_ref/libglfer_ref_f64.so) when that library is present, and
  (b) the committed fixtures tests/golden/*.npz generated from that same library by
      tests/golden/make_golden.py, and the known-answer values of SURVEY.md section 8c.
The reference itself ships no tests or golden vectors; its FFT
in the double path is the third-party FFTW 2.x rfftw (version not pinned,
configure.in:17-18), restated here as the mathematical DFT (numpy.fft.rfft).
"""
from __future__ import annotations

import math
import struct

import numpy as np

# window enum, fft.h:67
HANNING, BLACKMAN, GAUSSIAN, WELCH, BARTLETT, RECTANGULAR, HAMMING, KAISER = range(8)
WINDOW_NAMES = ["Hanning", "Blackman", "Gaussian", "Welch", "Bartlett", "Rectangular", "Hamming", "Kaiser"]
# avgmode_t, glfer.h:53-55
NO_AVG, AVG_SUMAVG, AVG_PLAIN, AVG_SUMEXTREME = range(4)


def hop_size(n: int, overlap: float) -> int:
    """n_eff of prepare_audio, fft.c:70: `int n_eff = N * (1.0 - params->overlap)` with
    overlap a C float promoted to double; truncation toward zero."""
    return int(n * (1.0 - float(np.float32(overlap))))


def _seq_sum_f32(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """Sequential (left-to-right) float32 accumulation, as a C `float acc; acc += x[i]`
    loop does (np.cumsum accumulates sequentially in the requested dtype)."""
    x = np.asarray(x, dtype=np.float32)
    if x.shape[axis] == 0:
        return np.zeros(np.delete(x.shape, axis), dtype=np.float32)
    return np.take(np.cumsum(x, axis=axis, dtype=np.float32), -1, axis=axis)


def _seq_sum_f64(x: np.ndarray) -> float:
    """Sequential double accumulation (`double acc; acc += x[i]`)."""
    return float(np.cumsum(np.asarray(x, dtype=np.float64))[-1]) if len(x) else 0.0


def bessel_i0(x: np.ndarray) -> np.ndarray:
    """util.c:222-237: polynomial approximations of I0 (Abramowitz & Stegun 9.8.1 / 9.8.2)."""
    x = np.asarray(x, dtype=np.float64)
    ax = np.abs(x)
    small = ax < 3.75
    y = (x / 3.75) ** 2
    ans_s = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-01 + y * 0.45813e-02)))))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        yl = 3.75 / ax
        ans_l = (np.exp(ax) / np.sqrt(ax)) * (0.39894228 + yl * (0.1328592e-01 + yl * (0.225319e-02 + yl * (-0.157565e-02 + yl * (0.916281e-02 + yl * (-0.2057706e-01 + yl * (0.2635537e-01 + yl * (-0.1647633e-01 + yl * 0.392377e-02))))))))
    return np.where(small, ans_s, ans_l)


def compute_window(n: int, window_type: int) -> np.ndarray:
    """compute_window, fft.c:309-360.  Shape in double, stored to float[N]; the power
    sum is accumulated in a C float (fft.c:314,353-356) and each tap is divided by
    sqrt(w_pwr) in double and stored back to float (fft.c:357-359).  Unknown types
    fall to the `default:` arm = all ones (fft.c:348-349)."""
    i = np.arange(n, dtype=np.float64)
    if window_type == HANNING:          # fft.c:322
        w = 0.5 - 0.5 * np.cos(2.0 * math.pi * i / (n - 1.0))
    elif window_type == BLACKMAN:       # fft.c:325
        w = 0.42 - 0.5 * np.cos(2.0 * math.pi * i / (n - 1.0)) + 0.08 * np.cos(4.0 * math.pi * i / (n - 1.0))
    elif window_type == GAUSSIAN:       # fft.c:328-329 (alpha = 1)
        w = np.exp(-1.0 * (2.0 * i - n + 1.0) * (2.0 * i - n + 1.0) / ((n - 1.0) * (n - 1.0)))
    elif window_type == WELCH:          # fft.c:332
        w = 1.0 - ((2.0 * i - n + 1.0) / (n - 1.0)) * ((2.0 * i - n + 1.0) / (n - 1.0))
    elif window_type == BARTLETT:       # fft.c:335
        w = 1.0 - np.abs((2.0 * i - n + 1.0) / (n - 1.0))
    elif window_type == HAMMING:        # fft.c:341
        w = 0.54 - 0.46 * np.cos(2.0 * math.pi * i / (n - 1.0))
    elif window_type == KAISER:         # fft.c:344-346: t and alpha are C floats
        # `t*t - (i-t)*(i-t)` is float arithmetic in C (t float, i int -> float); sqrt()
        # promotes to double; alpha * sqrt() is float * double; alpha * t is float * float
        t = np.float32((n - 1.0) / 2.0)
        alpha = np.float32(6.0 / float(t))
        it = np.arange(n).astype(np.float32) - t
        rad = (t * t - it * it).astype(np.float32)
        with np.errstate(invalid="ignore"):
            arg = float(alpha) * np.sqrt(rad.astype(np.float64))
        w = bessel_i0(arg) / bessel_i0(np.array(float(np.float32(alpha * t))))
    else:                               # RECTANGULAR fft.c:338 and default fft.c:349
        w = np.ones(n, dtype=np.float64)
    w32 = w.astype(np.float32)
    w_pwr = _seq_sum_f32(w32 * w32)     # float * float -> float, float accumulator
    return (w32.astype(np.float64) / math.sqrt(float(w_pwr))).astype(np.float32)


def num_frames(nsamples: int, n: int, overlap: float) -> int:
    """One frame per complete hop block delivered by the source (source.c:130:
    `for i < n` blocks; wav_fmt.c:119 delivers one block per non-empty read)."""
    hop = hop_size(n, overlap)
    return nsamples // hop if hop > 0 else 0


def subtract_block_means(samples: np.ndarray, hop: int) -> np.ndarray:
    """prepare_audio, fft.c:86-96: per hop block, mean of the hop NEW samples only,
    float accumulator, `sig_mean /= n_eff` in float, subtracted in float."""
    nblk = len(samples) // hop
    blk = np.asarray(samples[: nblk * hop], dtype=np.float32).reshape(nblk, hop)
    mean = (_seq_sum_f32(blk, axis=1) / np.float32(hop)).astype(np.float32)
    return (blk - mean[:, None]).astype(np.float32).reshape(-1)


def block_means(samples: np.ndarray, hop: int) -> np.ndarray:
    nblk = len(samples) // hop
    blk = np.asarray(samples[: nblk * hop], dtype=np.float32).reshape(nblk, hop)
    return (_seq_sum_f32(blk, axis=1) / np.float32(hop)).astype(np.float32)


def gather_frames(samples: np.ndarray, n: int, overlap: float, sub_mean: bool,
                  first_frame: int = 0, nframes: int | None = None, zero_history: bool = False) -> np.ndarray:
    """inbuf_audio for frames [first_frame, first_frame+nframes), fft.c:98-113: frame f
    holds stream samples [f*hop - n_ov, f*hop + hop); the history is zero before the
    first block (glfer.first_buffer, fft.c:103-108).  Returns float32 [nframes][n]."""
    hop = hop_size(n, overlap)
    n_ov = n - hop
    total = len(samples) // hop
    if nframes is None:
        nframes = total - first_frame
    x = np.asarray(samples[: total * hop], dtype=np.float32)
    if sub_mean:
        x = subtract_block_means(x, hop)
    # With first_buffer TRUE only on frame 0 the history of frame f is the last n_ov
    # samples of frame f-1's buffer; unrolled, that is the stream itself zero-prefixed
    # by n_ov samples, *also when n_ov > hop* (older history shifts through, fft.c:100-102).
    xp = np.concatenate([np.zeros(n_ov, dtype=np.float32), x])
    idx = (np.arange(first_frame, first_frame + nframes)[:, None] * hop) + np.arange(n)[None, :]
    frames = xp[idx]
    if zero_history:
        # glfer.first_buffer never cleared (the GUI with opt.autoscale == 0, g_main.c:1111-1120):
        # fft.c:103-108 zeroes the first n_ov samples of inbuf_audio on every block
        frames[:, :n_ov] = 0.0
    return frames


def _preops(frames: np.ndarray, window: np.ndarray, window_type: int, a: float, limiter: int) -> np.ndarray:
    """inbuf_fft (double) from inbuf_audio, fft.c:127-156."""
    a32 = np.float32(a)
    if a32 > 0.0:
        # fft.c:130-131: inp_val is a C float; inp_val / (a + inp_val*inp_val) in float
        v = (frames / (a32 + frames * frames)).astype(np.float32).astype(np.float64)
        if window_type != RECTANGULAR:
            v = v * window.astype(np.float64)[None, :]          # fft.c:134 (double *= float)
    else:
        if window_type != RECTANGULAR:
            v = window.astype(np.float64)[None, :] * frames.astype(np.float64)   # fft.c:143
        else:
            v = frames.astype(np.float64)                       # fft.c:148 (no normalisation)
    if limiter == 1:
        # fft.c:153-154: ftmp is a C float holding log(|x|); exp(ftmp * 0.1) in double
        with np.errstate(divide="ignore"):
            ftmp = np.log(np.abs(v)).astype(np.float32).astype(np.float64)
        e = np.exp(ftmp * 0.1)
        v = np.where(v > 0, e, -e)
    return v


def psd_from_spectrum(spec: np.ndarray, n: int) -> np.ndarray:
    """fft_psd, fft.c:203-217: psd[0] = Re0^2/N, psd[i] = (Re^2+Im^2)/N, psd[N/2] =
    ReNyq^2/N; computed in double (outbuf is fftw_real), stored to float.  No
    one-sided doubling, no sample-rate scaling.  spec is the rfft result (N/2+1 bins)."""
    p = (spec.real * spec.real + spec.imag * spec.imag) / n
    p[..., 0] = spec[..., 0].real * spec[..., 0].real / n
    if n % 2 == 0:
        p[..., n // 2] = spec[..., n // 2].real * spec[..., n // 2].real / n
    return p.astype(np.float32)


def periodogram(samples: np.ndarray, n: int, window_type: int, overlap: float, sub_mean: bool = False,
                a: float = 0.0, limiter: int = 0, first_frame: int = 0, nframes: int | None = None,
                return_spectrum: bool = False, zero_history: bool = False):
    """fft_do + fft_psd per hop block (source.c:143-144 -> fft.c:190-217).
    Returns float32 rows [nframes][n/2+1] (bin i <-> i*fs/N, DC first)."""
    window = compute_window(n, window_type)
    frames = gather_frames(samples, n, overlap, sub_mean, first_frame, nframes, zero_history)
    v = _preops(frames, window, window_type, a, limiter)
    spec = np.fft.rfft(v, axis=1)          # forward e^{-2 pi i jk/N}, un-normalised (fft.c:196)
    rows = psd_from_spectrum(spec, n)
    if return_spectrum:
        return rows, spec
    return rows


def phase_from_spectrum(spec: np.ndarray, n: int) -> np.ndarray:
    """fft_psd phase branch, fft.c:218-225: atan2(Re, Im) (argument order as in the
    reference), 0 at DC and Nyquist."""
    ph = np.arctan2(spec.real, spec.imag).astype(np.float32)
    ph[..., 0] = 0
    if n % 2 == 0:
        ph[..., n // 2] = 0
    return ph


# ----------------------------------------------------------------------------- LMP
def lmp_statistic(psd_rows: np.ndarray, nl: int, first_frame: int = 0, history: np.ndarray | None = None) -> np.ndarray:
    """The per-bin statistic of lmp_do, lmp.c:131-160, over a sequence of float PSD rows.
    psdbufl is a ring of nl rows, zero at init (lmp.c:86-93), written at slot j_l = frame mod nl
    (lmp.c:124,188-190).  Per bin, in double: my = sum_j psdbufl[j] / nl in SLOT order
    (lmp.c:131-137), sy = sum_j (psdbufl[j] - my)^2 / (nl - 1) (lmp.c:140-146),
    v_hat = 0.5 (my - sqrt(max(my^2 - sy, 0))) (lmp.c:150-152),
    out = -sqrt(nl/2) + nl my / (2 sqrt(2 nl) v_hat) stored to float, then 1e-3 where
    out <= 1e-3 (NaN and inf pass through), out[0] = 1e-3 (lmp.c:154-157).
    `first_frame` / `history` ([first_frame rows or at least nl-1][bins], the rows before
    psd_rows[0]) let a shard reproduce the ring of a longer run."""
    rows = np.asarray(psd_rows, dtype=np.float32)
    nf, bins = rows.shape
    ring = np.zeros((nl, bins), dtype=np.float32)
    if history is not None and first_frame > 0:
        h = np.asarray(history, dtype=np.float32)
        for i in range(max(0, len(h) - nl), len(h)):
            g = first_frame - len(h) + i
            if g >= 0:
                ring[g % nl] = h[i]
    out = np.empty((nf, bins), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        for f in range(nf):
            ring[(first_frame + f) % nl] = rows[f]
            my = np.zeros(bins, dtype=np.float64)
            for j in range(nl):
                my += ring[j].astype(np.float64)
            my /= nl
            sy = np.zeros(bins, dtype=np.float64)
            for j in range(nl):
                d = ring[j].astype(np.float64) - my
                sy += d * d
            sy /= (nl - 1)
            v = my * my - sy
            v = np.where(v < 0.0, 0.0, v)
            v = 0.5 * (my - np.sqrt(v))
            o = (-math.sqrt(nl / 2.0) + (nl * my) / (2.0 * math.sqrt(2.0 * nl) * v)).astype(np.float32)
            o = np.where(o.astype(np.float64) <= 1.0e-3, np.float32(1e-3), o)
            o[0] = np.float32(1e-3)
            out[f] = o
    return out


def lmp(samples: np.ndarray, n: int, overlap: float, nl: int, sub_mean: bool = False,
        first_frame: int = 0, nframes: int | None = None) -> np.ndarray:
    """lmp_do per hop block (source.c:155-156 -> lmp.c:101-192).  prepare_audio runs with a
    rectangular window (source.c:395) and its RA9MB / limiter results are overwritten: the FFT
    input is rebuilt from inbuf_audio (lmp.c:112-114), so only the block-mean removal reaches
    the spectrum.  Returns float32 rows [nframes][n/2+1]."""
    total = num_frames(len(samples), n, overlap)
    if nframes is None:
        nframes = total - first_frame
    raw = periodogram(samples, n, RECTANGULAR, overlap, sub_mean, first_frame=0, nframes=first_frame + nframes)
    return lmp_statistic(raw[first_frame:], nl, first_frame, raw[:first_frame])


# ----------------------------------------------------------------------------- DPSS
def gl_dpss(n: int, w: float, kmax: int):
    """gl_dpss, g-l_dpss.c:288-347: c = pi*w with w = N*W given directly (:295-297);
    32-point Gauss-Legendre discretisation of the sinc kernel (:303-313); symmetric
    eigen-decomposition (reference: cyclic Jacobi :315; same eigen-pairs up to sign),
    sorted by |lambda| descending (:316, eigen_symmv_sort :35-72); sinc interpolation to
    n points (:319-328); unit-energy normalisation (:331-339).
    Returns (tapers float64 [kmax+1][n], lam float64 [kmax+1]) with lam = 1 + sig."""
    w = float(np.float32(w))            # mtm_params_t.w is a C float (mtm.h:42)
    c = math.pi * w
    gx, gw = np.polynomial.legendre.leggauss(32)       # tables g-l_dpss.c:213-282
    d = gx[:, None] - gx[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.where(d == 0.0, c / math.pi, np.sin(c * d) / (math.pi * d))
    k = k * np.sqrt(gw[:, None] * gw[None, :])
    ev, evec = np.linalg.eigh(k)
    order = np.argsort(-np.abs(ev), kind="stable")
    ev, evec = ev[order], evec[:, order]
    i = np.arange(n, dtype=np.float64)
    argm = (2.0 * (i[:, None] + 0.5) / n) - 1.0 - gx[None, :]          # [n][32]
    s = np.sqrt(gw)[None, :] * np.sin(c * argm) / (math.pi * argm)      # [n][32]
    v = s @ evec[:, : kmax + 1]                                         # [n][kmax+1]
    v = v / np.sqrt(np.sum(v * v, axis=0))[None, :]
    return np.ascontiguousarray(v.T), ev[: kmax + 1].copy()


def multitaper(samples: np.ndarray, n: int, overlap: float, w: float, kmax: int, sub_mean: bool = False,
               first_frame: int = 0, nframes: int | None = None, tapers=None, lam=None) -> np.ndarray:
    """mtm_do, mtm.c:154-220 (the F-test branch :165-174,204-210,222-233 writes only the
    file-static ftest[] that nothing reads: dead output, not restated).  K' = kmax+1
    tapers (loops `j <= k`, :189); per taper: inbuf_fft = taper * inbuf_audio (:190-192),
    FFT, fft_psd into a float temporary (:212), then psd += tmp / (1 + sig[j]) with the
    division in double and the running sum held in the caller's float buffer (:214-219).
    RA9MB / limiter do not reach the spectrum here: mtm_do overwrites inbuf_fft from
    inbuf_audio."""
    if tapers is None:
        tapers, lam = gl_dpss(n, w, kmax)
    frames = gather_frames(samples, n, overlap, sub_mean, first_frame, nframes).astype(np.float64)
    acc = np.zeros((frames.shape[0], n // 2 + 1), dtype=np.float32)
    for j in range(kmax + 1):
        spec = np.fft.rfft(frames * tapers[j][None, :], axis=1)
        tmp = psd_from_spectrum(spec, n)
        acc = (acc.astype(np.float64) + tmp.astype(np.float64) / lam[j]).astype(np.float32)
    return acc


# ------------------------------------------------------------------------- averaging
def multitaper_ftest(samples: np.ndarray, n: int, overlap: float, w: float, kmax: int, sub_mean: bool = False,
                     first_frame: int = 0, nframes: int | None = None, tapers=None, lam=None) -> np.ndarray:
    """Thomson's harmonic F-test as mtm_do computes it into its file-static `ftest` (double /
    FFTW-layout build; in the float build `mu` is never written).
      U0[j] = sum_i v[i][j]                                   mtm.c:78-84 (double)
      sum_U0_sqr = sum_j U0[j]^2                              mtm.c:125-128 (float accumulator)
      hn[i] = sum_j U0[j] v[i][j] / sum_U0_sqr                mtm.c:130-136 (float accumulator)
      mu = FFT(inbuf_audio * hn)                              mtm.c:165-171
      den[i] = sum_j |y_j[i] - mu[i] U0[j]|^2                 mtm.c:204-210 (float accumulator `ftest[i] +=`)
      ftest[i] = kmax |mu[i]|^2 sum_U0_sqr / den[i]           mtm.c:222-233 (note: kmax = K' - 1)
    DC uses the real part only (:205-206, :223-224); for even n the Nyquist denominator is never
    accumulated (the loops stop at (n+1)/2), so ftest[n/2] = num / 0 = inf (nan when num is 0)."""
    if tapers is None:
        tapers, lam = gl_dpss(n, w, kmax)
    frames = gather_frames(samples, n, overlap, sub_mean, first_frame, nframes).astype(np.float64)
    u0 = np.array([_seq_sum_f64(tapers[j]) for j in range(kmax + 1)])
    s2 = np.float32(0.0)
    for j in range(kmax + 1):
        s2 = np.float32(np.float64(s2) + u0[j] * u0[j])
    hn = np.zeros(n, dtype=np.float32)
    for j in range(kmax + 1):
        hn = (hn.astype(np.float64) + u0[j] * tapers[j]).astype(np.float32)
    hn = (hn / s2).astype(np.float32)
    mu = np.fft.rfft(frames * hn.astype(np.float64)[None, :], axis=1)
    half = (n + 1) // 2
    den = np.zeros((frames.shape[0], n // 2 + 1), dtype=np.float32)
    for j in range(kmax + 1):
        y = np.fft.rfft(frames * tapers[j][None, :], axis=1)
        d = y - mu * u0[j]
        t = d.real * d.real + d.imag * d.imag
        t[:, 0] = d[:, 0].real * d[:, 0].real
        den[:, :half] = (den[:, :half].astype(np.float64) + t[:, :half]).astype(np.float32)
    m2 = mu.real * mu.real + mu.imag * mu.imag
    m2[:, 0] = mu[:, 0].real * mu[:, 0].real
    if n % 2 == 0:
        m2[:, n // 2] = 2.0 * mu[:, n // 2].real * mu[:, n // 2].real     # mu[i]^2 + mu[n-i]^2 with i = n/2
    num = kmax * m2 * np.float64(s2)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (num / den.astype(np.float64)).astype(np.float32)


def avg_bins(sample_rate: int, n: int, min_band_hz: float, max_band_hz: float):
    """g_main.c:1144-1146: binsize is a float quotient, bins are truncated float quotients."""
    binsize = np.float32(sample_rate) / np.float32(n)
    return int(np.float32(min_band_hz) / binsize), int(np.float32(max_band_hz) / binsize)


def update_avg(mode: int, psd_rows: np.ndarray, width: int, depth: int, minbin: int, maxbin: int,
               max0: int = 0, peakbin_init: int = 0):
    """update_avg_plain / _sumextreme / _sumavg applied frame by frame from a fresh
    alloc_avg (avg.c:38-60, 108-298).  Returns (avg float64 [F][width], ret float64 [F],
    peakbin int [F], variance float64 [F]).  State: cum[] running sums in double with a
    depth-long shift register per bin (avg.c:116-127), effdepth incremented until depth
    (:138-139), divisor effdepth+1 for PLAIN (:147,155), 1e-15 outside [minbin,maxbin)."""
    nf = psd_rows.shape[0]
    band = np.arange(minbin, maxbin)
    nb = len(band)
    cum = np.zeros(width, dtype=np.float64)
    ring = np.zeros((width, depth), dtype=np.float64)
    eff = 0
    out = np.empty((nf, width), dtype=np.float64)
    ret = np.empty(nf, dtype=np.float64)
    pk = np.empty(nf, dtype=np.int64)
    var = np.zeros(nf, dtype=np.float64)
    peakbin = peakbin_init
    with np.errstate(divide="ignore", invalid="ignore"):
        for f in range(nf):
            psd = psd_rows[f].astype(np.float64)
            p = psd[band] if nb else psd[:0]
            if eff < depth:
                ring[band, eff] = p
                cum[band] += p
            else:
                cum[band] += p - ring[band, 0]
                ring[band, :-1] = ring[band, 1:]
                ring[band, depth - 1] = p
            mx = float(psd_rows[f][minbin])                  # `double max = psd[minbin]` (raw row)
            cb = cum[band]
            # sequential `if (cum > max) { max = cum; *peakbin = index; }` == first strict maximum
            if nb and cb.max() > mx:
                peakbin = int(band[int(np.argmax(cb))])
                mx = float(cb.max())
            avgspec = float(np.sum(cb))                       # summation order differs only at 1e-16
            if eff < depth:
                eff += 1
            out[f, :] = 1e-15
            if mode == AVG_PLAIN:                             # avg.c:147-156
                ret[f] = (avgspec - mx) / (float(maxbin - minbin - 1) * float(eff + 1))
                out[f, band] = cb / float(eff + 1)
            elif mode == AVG_SUMEXTREME:                      # avg.c:163-218
                mn = min(1.0, float(cb.min())) if nb else 1.0
                avgspec = (avgspec - mx) / float(maxbin - minbin - 1)
                out[f, band] = (cb - mn) / (mx - mn) if max0 else cb / avgspec
                ret[f] = mx / avgspec
            elif mode == AVG_SUMAVG:                          # avg.c:224-297
                avgspec = (avgspec - mx) / float(maxbin - minbin - 1)
                pos = (cb - avgspec) > 0
                val = (cb - avgspec) / (mx - avgspec) if max0 else cb / avgspec
                out[f, band] = np.where(pos, val, 1e-15)
                sel = pos & (band != peakbin)
                r = cb[sel] / avgspec
                var[f] = np.sum(r * r) / float(np.count_nonzero(sel))
                ret[f] = mx / avgspec
            else:
                raise ValueError("mode")
            pk[f] = peakbin
    return out, ret, pk, var


def compute_floor(psd: np.ndarray):
    """compute_floor, fft.c:240-294: sig = largest bin, floor = sum of the lowest 5 %
    (indices >= (int)(n*0.95) of the descending sort) / 0.05 / n, peak = first bin
    strictly greater than 0 that is the maximum."""
    n = len(psd)
    srt = np.sort(psd.astype(np.float32))[::-1]
    floor = _seq_sum_f32(srt[int(n * 0.95):])
    floor = np.float32(float(floor) / 0.05)          # `floor_pwr /= 0.05`: float / double -> double, stored to float
    floor = np.float32(floor / np.float32(n))        # `floor_pwr /= N2`: float / int -> float
    peak_bin = int(np.argmax(psd)) if psd.max() > 0 else 0
    return float(srt[0]), float(floor), float(max(psd.max(), 0.0)), peak_bin


# ------------------------------------------------------------------------ display mapping
def display_levels(psd_rows: np.ndarray, shown_rows: np.ndarray, overlap: float, log_scale: bool, autoscale: bool,
                   max_level_db: float = -20.0, min_level_db: float = -80.0, thr_level: float = 0.0,
                   first_frame: int = 0, agc_state=(0.0, 0.0)):
    """main_window_draw, g_main.c:1109-1229, without the GTK drawing.  PINNED (round 2): bit-identical
    to the reference's own main_window_draw -- g_main.c compiled unmodified with GTK stubbed out,
    oracle/ref_gui_unit.c -- in every scale / autoscale / threshold / averaging variant
    (tests/test_oracle.py::test_display_mapping_pinned, fixtures disp_* of glfer_ref_f64_r2.npz).
    Per frame: compute_floor on the PSD row (:1109); AGC with float state and double
    arithmetic (:1111-1124) or fixed levels (:1126-1128); dB range (:1132-1135); per pixel i:
    bin n-1-i, level through the `short` level buffer in the log scales (:68,1193-1195),
    f = 255 (level - min) / (max - min) in float (:1206), threshold / clip / scale (:1221-1229).
    Returns (levels uint8 [F][n], range float32 [F][2], final agc state)."""
    nf, n = shown_rows.shape
    thr = np.float32(np.float32(thr_level) / 100.0)
    mx_lvl, mn_lvl = np.float32(agc_state[0]), np.float32(agc_state[1])
    levels = np.empty((nf, n), dtype=np.uint8)
    rng = np.empty((nf, 2), dtype=np.float32)
    ov = np.float32(overlap)
    for f in range(nf):
        if autoscale:
            sig, flo, _, _ = compute_floor(psd_rows[f])
            sig, flo = np.float32(sig), np.float32(flo)
            if first_frame + f == 0:
                if ov > 0.0:
                    sig = np.float32(sig / ov)
                    flo = np.float32(flo / ov)
                mx_lvl, mn_lvl = sig, flo
            else:
                mx_lvl = np.float32((1.0 - 0.99) * float(sig) + 0.99 * float(mx_lvl))
                mn_lvl = np.float32((1.0 - 0.99) * float(flo) + 0.99 * float(mn_lvl))
        else:
            mx_lvl = np.float32(10.0 ** (float(np.float32(max_level_db)) / 10.0))
            mn_lvl = np.float32(10.0 ** (float(np.float32(min_level_db)) / 10.0))
            mn_lvl = mn_lvl if mx_lvl > mn_lvl else np.float32(float(mx_lvl) / 10.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            if log_scale:
                dmax = np.float32(10.0 * np.log10(float(mx_lvl)))
                dmin = np.float32(10.0 * np.log10(float(mn_lvl)))
                d = 10.0 * np.log10(shown_rows[f][::-1].astype(np.float64))
                sig_level = np.where(np.abs(d) < 2147483648.0, np.trunc(d), 0.0).astype(np.float32)
            else:
                dmax, dmin = mx_lvl, mn_lvl
                sig_level = shown_rows[f][::-1].astype(np.float32)
            fl = (np.float32(255) * ((sig_level - dmin) / (dmax - dmin))).astype(np.float32)
            scaled = (fl.astype(np.float64) - 255.0 * float(thr)) / (1.0 - float(thr))
            v = np.where(fl.astype(np.float64) < 255.0 * float(thr), 0, np.where(fl > 255, 255, np.nan_to_num(scaled)))
        levels[f] = v.astype(np.int64).astype(np.uint8)
        rng[f] = (dmax, dmin)
    return levels, rng, (float(mx_lvl), float(mn_lvl))


# ------------------------------------------------------------------------------ WAV
def read_wav_blocks(path: str, hop: int, sub_mean: bool = False):
    """wav_fmt.c:45-121 restated for LP64 (the reference's header struct uses u_long and
    mis-parses on x86-64): skip the canonical 44-byte header
    (:58-66), rate / bits from it (:69-70), then one block of hop samples per read
    (:87,102); u8 -> (x-128)/128, s16 -> x/32768 (:105-116), channels not
    de-interleaved; a short final read leaves the stale tail of the previous block in
    place (:102-119, buffer is calloc'ed once :91-100).  With sub_mean the estimator has
    already subtracted the previous block's mean IN PLACE in that very buffer
    (fft.c:93-95 mutates audio_buf = wav_fmt.c's buff), so the stale tail carries the
    mean-removed values.  Returns (float32 stream of whole blocks, sample_rate, bits)."""
    with open(path, "rb") as fh:
        hd = fh.read(44)
        rate = struct.unpack_from("<I", hd, 24)[0]
        bits = struct.unpack_from("<H", hd, 34)[0]
        data = fh.read()
    bps = bits // 8
    blocks = []
    buff = np.zeros(hop, dtype=np.float32)
    for off in range(0, len(data), hop * bps):
        chunk = data[off: off + hop * bps]
        if bits == 8:
            v = (np.frombuffer(chunk, dtype=np.uint8).astype(np.float32) - 128) / np.float32(128)
        else:
            v = np.frombuffer(chunk[: len(chunk) // 2 * 2], dtype="<i2").astype(np.float32) / np.float32(32768)
        buff = buff.copy()
        buff[: len(v)] = v
        blocks.append(buff)
        if sub_mean:
            # what prepare_audio leaves in the reader's buffer for the next read
            buff = subtract_block_means(buff, hop)
    stream = np.concatenate(blocks) if blocks else np.zeros(0, dtype=np.float32)
    return stream, rate, bits
