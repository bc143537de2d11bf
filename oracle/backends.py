"""TEST INFRASTRUCTURE ONLY: loaders for the two CPU checkers.

* ``oracle()``     -> oracle/libmxadv_oracle.so, our plain-C restatement (``orc_*``)
* ``reference(dim)`` -> oracle/_ref/libmaestro_ref{2,3}d.so, the reference's own sources
  compiled against oracle/shim (``ref_*``); ``None`` if it has not been built.
Both are wrapped in the same ``Calls`` layer the product uses, so a test can run the three
back ends on identical inputs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

from maestroex_b200.calls import Calls

HERE = os.path.dirname(os.path.abspath(__file__))


def build(ref: bool = True) -> None:
    """Compile the C oracle (always) and, when /root/reference exists, oracle/_ref."""
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


@lru_cache(maxsize=None)
def oracle() -> Calls:
    path = os.path.join(HERE, "libmxadv_oracle.so")
    if not os.path.exists(path):
        build(ref=False)
    return Calls(C.CDLL(path), "orc_", with_stream=False)


@lru_cache(maxsize=None)
def reference(dim: int = 3):
    path = os.path.join(HERE, "_ref", f"libmaestro_ref{dim}d.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/Source"):
            build(ref=True)
        if not os.path.exists(path):
            return None
    return Calls(C.CDLL(path), "ref_", with_stream=False)


@lru_cache(maxsize=None)
def overlay(dim: int = 3):
    """The PRODUCT's C++ overlay (maestroex_b200/overlay: `Maestro::` members on libmxadv.so)
    behind the same driver as ``reference`` (exports ``ovl_*``); ``None`` if not built."""
    path = os.path.join(HERE, "_ref", f"libmxadv_overlay{dim}d.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/Source"):
            build(ref=True)
        if not os.path.exists(path):
            return None
    return Calls(C.CDLL(path), "ovl_", with_stream=False)
