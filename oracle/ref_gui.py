"""ctypes access to oracle/_ref/libglfer_ref_gui.so: the reference's main_window_draw() and
set_palette() (g_main.c, unmodified, GTK stubbed out; see ref_gui_unit.c).  TEST INFRASTRUCTURE
ONLY: pins the display mapping.  One GUI state per process (the reference keeps it in statics)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libglfer_ref_gui.so")
_lib = None
SCALE_LIN, SCALE_LIN_MAX0, SCALE_LOG, SCALE_LOG_MAX0 = range(4)          # glfer.h:43
HSV, THRESH, COOL, HOT, BW, BONE, COPPER, OTD = range(8)                 # glfer.h:47


def available() -> bool:
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(_PATH)
        l.refh_gui_setup.argtypes = [C.c_int, C.c_int]
        l.refh_gui_palette.argtypes = [C.c_int, C.c_void_p]
        l.refh_gui_draw.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        l.refh_gui_options.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int,
                                       C.c_float, C.c_float, C.c_int, C.c_int]
        l.refh_gui_first_buffer.argtypes = [C.c_int]
        l.refh_gui_alloc_avg.argtypes = [C.c_int, C.c_int]
        for f in (l.refh_gui_setup, l.refh_gui_palette, l.refh_gui_draw, l.refh_gui_options, l.refh_gui_first_buffer,
                  l.refh_gui_alloc_avg):
            f.restype = None
        _lib = l
    return _lib


def palette(p: int) -> np.ndarray:
    tab = np.empty(768, dtype=np.uint8)
    lib().refh_gui_palette(p, tab.ctypes.data)
    return tab.reshape(256, 3)


def draw_rows(psd_rows: np.ndarray, scale_type: int = SCALE_LOG, autoscale: bool = True, max_level_db: float = -20.0,
              min_level_db: float = -80.0, thr_level: float = 0.0, overlap: float = 0.5, palette_id: int = BW,
              averaging: int = 0, avgsamples: int = 4, min_avgband: float = 400.0, max_avgband: float = 1200.0,
              sample_rate: int = 48000, data_block_size: int | None = None):
    """main_window_draw on each row in turn from a fresh display state (first_buffer TRUE).
    Returns dict(rgb [F][n][3], levels [F][n] (= the red channel under the B/W palette), lev [F][n] int16,
    scal [F][6])."""
    rows = np.ascontiguousarray(psd_rows, dtype=np.float32)
    nf, n = rows.shape
    block = data_block_size if data_block_size is not None else 2 * (n - 1)
    L = lib()
    L.refh_gui_options(scale_type, int(autoscale), max_level_db, min_level_db, thr_level, overlap, averaging, avgsamples,
                       min_avgband, max_avgband, sample_rate, block)
    L.refh_gui_setup(n, palette_id)
    if averaging:
        L.refh_gui_alloc_avg(block, avgsamples)              # source.c:312: width = block size
    rgb = np.empty((nf, n, 3), dtype=np.uint8)
    lev = np.empty((nf, n), dtype=np.int16)
    scal = np.empty((nf, 6), dtype=np.float32)
    for f in range(nf):
        row = rows[f].copy()
        L.refh_gui_draw(row.ctypes.data, rgb[f].ctypes.data, lev[f].ctypes.data, scal[f].ctypes.data)
    return dict(rgb=rgb, levels=rgb[:, :, 0].copy(), lev=lev, scal=scal)
