"""CPU-side checks of the drop-in layers (no GPU, no reference tree needed): every C-ABI
function the overlay translation units call is declared in include/mxadv.h and exported by
libmxadv.so; the e2e recipes parse; fcompare's table parser reads AMReX's format."""
import glob
import os
import re
import subprocess

import pytest

from maestroex_b200 import lib as mxlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OVERLAYS = sorted(glob.glob(os.path.join(ROOT, "maestroex_b200", "overlay*", "*.cpp")) +
                  glob.glob(os.path.join(ROOT, "maestroex_b200", "overlay*", "*.H")))


def test_overlay_calls_only_declared_exported_symbols():
    mxlib.build()
    lib = mxlib.load()
    declared = set(mxlib.declared_symbols())
    used = set()
    for f in OVERLAYS:
        used |= set(re.findall(r"\b(mx_[a-z0-9_]+)\s*\(", open(f).read()))
    assert {"mx_make_edge_scal", "mx_density_edge_flux", "mx_enthalpy_edge_flux",
            "mx_update_scal"} <= used
    for name in sorted(used):
        assert name in declared, f"{name} is not declared in include/mxadv.h"
        assert hasattr(lib, name), f"libmxadv.so does not export {name}"


def test_overlay_members_shadow_reference_files():
    """same file names as the reference translation units they replace (make vpath shadowing)"""
    names = {os.path.basename(f) for f in OVERLAYS}
    for n in ("MaestroMakeEdgeScalars.cpp", "MaestroMakeFlux.cpp", "MaestroCelltoEdge.cpp",
              "MaestroMakew0.cpp", "MaestroDensityAdvance.cpp", "MaestroEnthalpyAdvance.cpp"):
        assert n in names
    # the fused layer has no fallback inside: unsupported options must abort
    for n in ("MaestroDensityAdvance.cpp", "MaestroEnthalpyAdvance.cpp"):
        src = open(os.path.join(ROOT, "maestroex_b200", "overlay_fused", n)).read()
        assert "Abort(" in src and "mxadv::check(" in src


@pytest.mark.parametrize("script", ["build_e2e.sh", "build_e2e_cuda.sh", "run_cuda_compare.sh"])
def test_e2e_recipes_parse(script):
    p = os.path.join(ROOT, "oracle", "e2e", script)
    assert os.access(p, os.X_OK)
    subprocess.run(["bash", "-n", p], check=True)


def test_fcompare_parser():
    from test_gpu_fcompare import parse_fcompare
    txt = """
            variable name            absolute error            relative error
                                        (||A - B||)         (||A - B||/||A||)
 ----------------------------------------------------------------------------
 level = 0
 velx                                             0                         0
 rhoX(C12)                          8.344650269e-07           5.056392296e-16
 omegadot(O16)                      1.590974308e-12               1.333333333
 PLOTFILE AGREE
"""
    e = parse_fcompare(txt)
    assert e["velx"] == (0.0, 0.0)
    assert e["rhoX(C12)"][1] == pytest.approx(5.056392296e-16)
    assert e["omegadot(O16)"][1] == pytest.approx(1.333333333)
    assert len(e) == 3
