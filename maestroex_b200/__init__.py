"""maestroex_b200 — B200-native (sm_100a, fp64) scalar-advection hot path of MAESTROeX.

The product is the C-ABI library ``csrc/libmxadv.so`` (include/mxadv.h) plus the C++ overlay
in ``overlay/`` that keeps the ``Maestro::`` member-function signatures.  This Python package
is only the test / benchmark harness around that ABI: ``abi`` (ctypes structs), ``calls``
(call layer), ``lib`` (loader, fails loudly if the CUDA library is missing), ``synth``
(seeded synthetic inputs) and ``decomp`` (brick decomposition + halo exchange for N GPUs).
"""
__all__ = ["abi", "calls", "lib", "synth"]
