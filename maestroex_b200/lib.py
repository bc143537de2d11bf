"""Loader for the product library maestroex_b200/csrc/libmxadv.so (C ABI: include/mxadv.h).

Fails loudly when the library has not been built: there is no Python / CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from functools import lru_cache

from .calls import Calls

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# MXADV_LIB: A/B builds of the same ABI (kernel experiments); default is the in-tree build
LIB_PATH = os.environ.get("MXADV_LIB") or os.path.join(CSRC, "libmxadv.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "mxadv.h")


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("building libmxadv.so failed")
    return LIB_PATH


def declared_symbols() -> list:
    """Every `mx_*` function include/mxadv.h declares."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mx_[a-zA-Z0-9_]+)\s*\(", src)))


@lru_cache(maxsize=None)
def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C maestroex_b200/csrc`).  There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    lib.mx_last_error.restype = C.c_char_p
    lib.mx_launch_count.restype = C.c_longlong
    lib.mx_tma_launch_count.restype = C.c_longlong
    lib.mx_halo_count.restype = C.c_longlong
    return lib


@lru_cache(maxsize=None)
def calls() -> Calls:
    return Calls(load(), "mx_", with_stream=True)
