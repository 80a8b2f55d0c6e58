"""ctypes access to oracle/_ref/libglfer_ref_{f32,f64}.so (the UNMODIFIED reference
sources + oracle/ref_harness.c, built by oracle/Makefile).  TEST INFRASTRUCTURE ONLY.

f64 = the reference's double-precision path (golden); f32 = float radix-2 build
(what glfer ships without FFTW; CPU timing baseline)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, C.CDLL] = {}

_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def path(kind: str = "f64") -> str:
    return os.path.join(_HERE, "_ref", f"libglfer_ref_{kind}.so")


def available(kind: str = "f64") -> bool:
    return os.path.exists(path(kind))


def lib(kind: str = "f64") -> C.CDLL:
    if kind not in _LIBS:
        l = C.CDLL(path(kind))
        l.refh_hop.argtypes = [C.c_int, C.c_float]
        l.refh_hop.restype = C.c_int
        l.refh_window.argtypes = [C.c_int, C.c_int, _fp]
        l.refh_periodogram.argtypes = [_fp, C.c_long, C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int,
                                       C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        l.refh_periodogram.restype = C.c_long
        l.refh_dpss.argtypes = [C.c_int, C.c_float, C.c_int, _dp, _dp]
        l.refh_mtm.argtypes = [_fp, C.c_long, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int,
                               C.c_long, _fp]
        l.refh_mtm.restype = C.c_long
        l.refh_mtm_ftest.argtypes = [_fp, C.c_long, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_long, _fp, _fp]
        l.refh_mtm_ftest.restype = C.c_long
        l.refh_lmp.argtypes = [_fp, C.c_long, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_long, _fp]
        l.refh_lmp.restype = C.c_long
        l.refh_avg.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, C.c_long, C.c_long, C.c_int,
                               _dp, _dp, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"), _dp, C.c_int]
        l.refh_floor.argtypes = [_fp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                 C.POINTER(C.c_uint)]
        l.refh_time_periodogram.argtypes = [_fp, C.c_long, C.c_int, C.c_int, C.c_float, C.c_int,
                                            C.POINTER(C.c_long), C.POINTER(C.c_double)]
        l.refh_time_periodogram.restype = C.c_double
        l.refh_time_mtm.argtypes = [_fp, C.c_long, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int,
                                    C.POINTER(C.c_long), C.POINTER(C.c_double)]
        l.refh_time_mtm.restype = C.c_double
        l.refh_set_sticky_first_buffer.argtypes = [C.c_int]
        l.refh_set_sticky_first_buffer.restype = None
        _LIBS[kind] = l
    return _LIBS[kind]


def hop(n: int, overlap: float, kind="f64") -> int:
    return lib(kind).refh_hop(n, overlap)


def window(n: int, wtype: int, kind="f64") -> np.ndarray:
    out = np.empty(n, dtype=np.float32)
    lib(kind).refh_window(n, wtype, out)
    return out


def periodogram(samples, n, wtype, overlap, sub_mean=False, a=0.0, limiter=0, max_frames=1 << 40, kind="f64",
                want_spec=False, want_phase=False, sticky_first_buffer=False):
    """sticky_first_buffer: glfer.first_buffer is never cleared (GUI with opt.autoscale == 0)."""
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    h = hop(n, overlap, kind)
    nf = min(len(samples) // h, max_frames)
    rows = np.empty((nf, n // 2 + 1), dtype=np.float32)
    spec = np.empty((nf, n), dtype=np.float64) if want_spec else None
    phase = np.empty((nf, n // 2 + 1), dtype=np.float32) if want_phase else None
    lib(kind).refh_set_sticky_first_buffer(int(sticky_first_buffer))
    try:
        got = lib(kind).refh_periodogram(samples, len(samples), n, wtype, overlap, int(sub_mean), a, limiter, nf,
                                         rows.ctypes.data, spec.ctypes.data if want_spec else None,
                                         phase.ctypes.data if want_phase else None)
    finally:
        lib(kind).refh_set_sticky_first_buffer(0)
    assert got == nf
    res = [rows]
    if want_spec:
        res.append(spec)
    if want_phase:
        res.append(phase)
    return res[0] if len(res) == 1 else tuple(res)


def dpss(n, w, kmax, kind="f64"):
    tap = np.empty((kmax + 1, n), dtype=np.float64)
    lam = np.empty(kmax + 1, dtype=np.float64)
    lib(kind).refh_dpss(n, w, kmax, tap, lam)
    return tap, lam


def mtm(samples, n, overlap, w, kmax, sub_mean=False, a=0.0, limiter=0, max_frames=1 << 40, kind="f64"):
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    h = hop(n, overlap, kind)
    nf = min(len(samples) // h, max_frames)
    rows = np.empty((nf, n // 2 + 1), dtype=np.float32)
    got = lib(kind).refh_mtm(samples, len(samples), n, overlap, int(sub_mean), a, limiter, w, kmax, nf, rows)
    assert got == nf
    return rows


def mtm_ftest(samples, n, overlap, w, kmax, sub_mean=False, max_frames=1 << 40, kind="f64"):
    """mtm_do rows plus the harmonic F-test it leaves in its file-static buffer (double build)."""
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    h = hop(n, overlap, kind)
    nf = min(len(samples) // h, max_frames)
    rows = np.empty((nf, n // 2 + 1), dtype=np.float32)
    ft = np.empty((nf, n // 2 + 1), dtype=np.float32)
    got = lib(kind).refh_mtm_ftest(samples, len(samples), n, overlap, int(sub_mean), w, kmax, nf, rows, ft)
    assert got == nf
    return rows, ft


def lmp(samples, n, overlap, nl, sub_mean=False, a=0.0, limiter=0, max_frames=1 << 40, kind="f64"):
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    h = hop(n, overlap, kind)
    nf = min(len(samples) // h, max_frames)
    rows = np.empty((nf, n // 2 + 1), dtype=np.float32)
    got = lib(kind).refh_lmp(samples, len(samples), n, overlap, int(sub_mean), a, limiter, nl, nf, rows)
    assert got == nf
    return rows


def avg(mode, psd_rows, width, depth, minbin, maxbin, max0=0, nbins_out=None, peakbin_init=0, kind="f64"):
    psd_rows = np.ascontiguousarray(psd_rows, dtype=np.float32)
    nf, stride = psd_rows.shape
    nbins_out = width if nbins_out is None else nbins_out
    out = np.empty((nf, nbins_out), dtype=np.float64)
    ret = np.empty(nf, dtype=np.float64)
    pk = np.empty(nf, dtype=np.int32)
    var = np.empty(nf, dtype=np.float64)
    rc = lib(kind).refh_avg(mode, width, depth, minbin, maxbin, max0, psd_rows, nf, stride, nbins_out, out, ret, pk,
                            var, peakbin_init)
    assert rc == 0
    return out, ret, pk, var


def floor_stats(psd_row, kind="f64"):
    psd_row = np.ascontiguousarray(psd_row, dtype=np.float32).copy()
    s, f, p, b = C.c_float(), C.c_float(), C.c_float(), C.c_uint()
    lib(kind).refh_floor(psd_row, len(psd_row), C.byref(s), C.byref(f), C.byref(p), C.byref(b))
    return s.value, f.value, p.value, b.value


def time_periodogram(samples, n, wtype, overlap, sub_mean=False, kind="f32"):
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    nf, cs = C.c_long(), C.c_double()
    t = lib(kind).refh_time_periodogram(samples, len(samples), n, wtype, overlap, int(sub_mean), C.byref(nf), C.byref(cs))
    return t, nf.value


def time_mtm(samples, n, overlap, w, kmax, sub_mean=False, kind="f32"):
    samples = np.ascontiguousarray(samples, dtype=np.float32)
    nf, cs = C.c_long(), C.c_double()
    t = lib(kind).refh_time_mtm(samples, len(samples), n, overlap, int(sub_mean), w, kmax, C.byref(nf), C.byref(cs))
    return t, nf.value
