"""B200-native path-tracing core: drop-in for the PathTracer render path of Khrylx/DSGPURayTracing.

The product is the CUDA library `libdsrt.so` (hand-written sm_100a kernels behind the C ABI of include/dsrt.h)
plus the C++ host side (`pathtracer` CLI, COLLADA loader, SAH builder).  This Python package is only the thin
ctypes binding used by tests/, bench.py and __graft_entry__.py; it fails loudly when the CUDA library is missing.
"""
from ._lib import (Core, Stats, DsrtError, lib_path, load_library, build_bvh2, EXPORTED_SYMBOLS,  # noqa: F401
                   HOST_EXPORTED_SYMBOLS, load_host_library, load_dae, load_envmap, render_file, set_loader_option)
