"""glfer_b200: B200-native spectrogram engine behind glfer's fft.h / mtm.h / avg.h API.

The product is the C-ABI shared library glfer_b200/libglfer_b200.so (C host layer +
CUDA kernels for sm_100a).  This Python package only holds the build recipe, a thin
ctypes binding used by tests/bench, the time-sharding arithmetic and the synthetic
signal generator; it never computes spectra itself and has no CPU fallback."""
