"""Host-side logic of the product library, runnable without a GPU: the library loads and
exports every symbol include/*.h declares, the host table generators match the oracle, the
FFT core's index arithmetic is emulated on the CPU, and the product fails loudly (no
fallback) when no CUDA device is present."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import glfer_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("fft.h", "mtm.h", "avg.h", "lmp.h", "glfer_b200.h", "glb_shim.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"^[ \t]*(?:extern\s+)?(?:const\s+)?(?:unsigned\s+)?[A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+\**"
                             r"([a-z_][A-Za-z0-9_]*)\s*\(", txt, flags=re.M):
            names.add(m.group(1))
    names |= {"fft_windows", "num_fft_windows"}
    return names - {"defined", "sizeof"}


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    decl = _declared_symbols()
    assert len(decl) > 60
    missing = sorted(decl - exported)
    assert not missing, missing


def test_fftw_layout_library_exports_the_same_interface(built_lib):
    """libglfer_b200_fftw.so (fft_params_t in the reference's HAVE_LIBRFFTW layout) carries every declared symbol too"""
    path = os.path.join(os.path.dirname(built_lib), "libglfer_b200_fftw.so")
    assert os.path.exists(path)
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    missing = sorted(_declared_symbols() - exported)
    assert not missing, missing


def test_library_is_built_for_sm100a(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_host_windows_match_oracle_bit_exact(api):
    for n in (32, 1024, 4096, 32768):
        for t in range(8):
            assert np.array_equal(api.host_window(n, t), O.compute_window(n, t)), (n, t)
    assert np.array_equal(api.host_window(256, 99), O.compute_window(256, 99))     # default arm: all ones


def test_host_hop_matches_reference_truncation(api):
    for n in (64, 1024, 4096, 32768):
        for ov in (0.0, 0.25, 0.5, 0.75, 0.9, 0.95, 0.3, 0.999):
            assert api.host_hop(n, ov) == O.hop_size(n, ov)
    assert api.host_hop(4096, 0.9) == 409


def test_host_dpss_matches_oracle(api):
    for n, w, k in ((1024, 4.0, 7), (4096, 4.0, 7), (512, 2.5, 3), (256, 3.0, 5)):
        ta, la = api.host_dpss(n, w, k)
        tb, lb = O.gl_dpss(n, w, k)
        assert np.allclose(la, lb, rtol=1e-9, atol=1e-15)
        sgn = np.sign(np.sum(ta * tb, axis=1))
        assert np.max(np.abs(ta * sgn[:, None] - tb)) < 1e-7
        assert np.allclose(np.sum(ta * ta, axis=1), 1.0, rtol=1e-12)


def test_host_dpss_degenerate_cluster_spans_same_subspace(api):
    # NW=8: lambda_0..7 equal 1 to machine precision, eigenvectors are defined up to a
    # rotation inside the cluster; the projector (and hence the multitaper PSD) is invariant
    ta, la = api.host_dpss(512, 8.0, 15)
    tb, lb = O.gl_dpss(512, 8.0, 15)
    assert np.allclose(la, lb, rtol=1e-7, atol=1e-15)
    pa = (ta / np.sqrt(la)[:, None]).T @ (ta / np.sqrt(la)[:, None])
    pb = (tb / np.sqrt(lb)[:, None]).T @ (tb / np.sqrt(lb)[:, None])
    assert np.max(np.abs(pa - pb)) < 1e-6 * np.max(np.abs(pb))


def test_shard_ranges_partition_the_frames(api):
    for nframes in (0, 1, 7, 84375, 506250):
        for ndev in (1, 2, 3, 4, 8):
            pos = 0
            for g in range(ndev):
                first, count = api.shard_range(nframes, ndev, g)
                assert first == pos and count >= 0
                pos += count
            assert pos == nframes


def test_fft_core_emulation(tmp_path):
    """the kernels' FFT index/twiddle/swizzle arithmetic, run on the CPU thread by thread"""
    exe = str(tmp_path / "emu_fft")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_fft.cpp")], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    rows = [l.split() for l in res.stdout.strip().splitlines()]
    # table-twiddle variant for every plan up to M=2048, then the register-twiddle variant
    assert [int(r[0]) for r in rows if r[4] == "table"] == [16, 32, 64, 128, 256, 512, 1024, 2048]
    assert [int(r[0]) for r in rows if r[4] == "rt"] == [16, 64, 128, 256, 512, 1024, 2048]
    assert all(int(r[3]) == 0 and float(r[2]) < 1e-6 for r in rows)


def test_big_frame_core_emulation(tmp_path):
    """fft_big.cuh (three passes, 32 points per thread: the N = 16384 / 32768 kernel) on the CPU, thread by thread"""
    exe = str(tmp_path / "emu_big")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_big.cpp")], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    rows = [l.split() for l in res.stdout.strip().splitlines()]
    assert [int(r[0]) for r in rows] == [2048, 4096, 8192, 16384, 8192, 16384]      # the last two: table last pass
    assert all(int(r[3]) == 0 and float(r[1]) < 1e-4 and float(r[2]) < 1e-6 for r in rows)


def test_warp_per_frame_core_emulation(tmp_path):
    """fft_wpf.cuh (two-pass, 64 points per lane) on the CPU, lane by lane"""
    exe = str(tmp_path / "emu_wpf")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_wpf.cpp")], check=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    rows = [l.split() for l in res.stdout.strip().splitlines()]
    assert [int(r[0]) for r in rows] == [256, 512, 1024, 2048]
    assert all(int(r[3]) == 0 and float(r[2]) < 1e-6 for r in rows)


def test_no_cpu_fallback(api):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.GlferError, match="no CUDA device"):
        api.GramPlan(n=1024)


def test_invalid_configs_are_rejected(api):
    if api.device_count() < 1:
        pytest.skip("argument validation past the device check needs a device")
    for kw in (dict(n=1000), dict(n=65536), dict(n=1024, overlap=1.0), dict(n=1024, mode=1, mtm_kmax=32)):
        with pytest.raises(api.GlferError):
            api.GramPlan(**kw)


def test_palettes_match_reference_fixture(api):
    """glfer_palette vs set_palette of the compiled reference GUI unit (g_main.c:649-762)"""
    g2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "glfer_ref_f64_r2.npz"))
    for pn in range(8):
        assert np.array_equal(api.palette(pn), g2["palettes"][pn]), pn


def test_db_threshold_tables_reproduce_host_libm(api):
    """host/levels.c: the number of thresholds <= x IS (short) (10 log10 x), for floats and doubles"""
    import ctypes as C
    lib = api.lib()
    n = 390 + 450 + 1
    tf = np.empty(n, dtype=np.float32)
    td = np.empty(n, dtype=np.float64)
    lib.glb_db_thresholds_f(tf.ctypes.data_as(C.c_void_p))
    lib.glb_db_thresholds_d(td.ctypes.data_as(C.c_void_p))
    assert np.all(np.diff(tf[np.isfinite(tf)]) >= 0) and np.all(np.diff(td) >= 0)
    rng = np.random.default_rng(2)
    x = np.exp(rng.uniform(-103, 88, 20000)).astype(np.float32)
    x = np.concatenate([x, tf[np.isfinite(tf)], np.nextafter(tf[np.isfinite(tf)], np.float32(0))])
    x = x[x > 0]
    d = 10.0 * np.log10(x.astype(np.float64))
    want = np.trunc(d).astype(np.int64)
    got = np.searchsorted(tf, x, side="right") - 1 - 450
    assert np.array_equal(got, want)
    xd = np.exp(rng.uniform(-103, 88, 20000))
    got = np.searchsorted(td, xd, side="right") - 1 - 450
    assert np.array_equal(got, np.trunc(10.0 * np.log10(xd)).astype(np.int64))


def test_wav_channel_selection(api, tmp_path):
    """glfer_wav_select_channel (an extension: the reference feeds a stereo file's interleaved samples to the
    estimator as one channel): the kept channel, the sample count, a partial trailing frame dropped, errors."""
    import ctypes as C
    from glfer_b200 import synth
    lib = api.lib()
    rng = np.random.default_rng(5)
    for channels, nfr, extra in ((2, 1000, 0), (3, 777, 2), (1, 500, 0)):
        pcm = rng.integers(-30000, 30000, nfr * channels + extra).astype(np.int16)
        path = str(tmp_path / f"c{channels}.wav")
        synth.write_wav16(path, pcm, 8000, channels=channels)
        for ch in range(channels):
            wav = api.Wav()
            assert lib.glfer_wav_load(path.encode(), C.byref(wav)) == 0
            assert wav.channels == channels and wav.nsamples == len(pcm)
            assert lib.glfer_wav_select_channel(C.byref(wav), ch) == 0
            assert wav.channels == 1 and wav.nsamples == nfr
            got = np.ctypeslib.as_array(C.cast(wav.data, C.POINTER(C.c_short)), shape=(nfr,)).copy()
            assert np.array_equal(got, pcm[: nfr * channels].reshape(nfr, channels)[:, ch])
            lib.glfer_wav_free(C.byref(wav))
        wav = api.Wav()
        assert lib.glfer_wav_load(path.encode(), C.byref(wav)) == 0
        assert lib.glfer_wav_select_channel(C.byref(wav), channels) != 0
        assert lib.glfer_wav_select_channel(C.byref(wav), -1) != 0
        lib.glfer_wav_free(C.byref(wav))
