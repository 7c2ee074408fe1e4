"""pixie_b200 — B200-native raster hot path behind Pixie's public procs.

Layout: ``csrc/host`` (C++ mirror of the Nim host producers -> libpixie_host.so),
``csrc/cuda`` (sm_100a kernels + the C ABI of include/pixie_cuda.h -> pixie_cuda.so),
``host.py`` / ``device.py`` (ctypes bindings) and ``api.py`` (Pixie's public procs:
fillPath / strokePath / draw / blur / shadow on Image).  There is no CPU fallback: anything
that renders needs pixie_cuda.so and a GPU and fails loudly otherwise.
"""
from .common import *  # noqa: F401,F403
from .common import PixieError  # noqa: F401
