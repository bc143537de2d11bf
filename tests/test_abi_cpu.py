"""CPU-side checks of the product library: it builds, loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C

import numpy as np
import pytest

from maestroex_b200 import lib as mxlib


@pytest.fixture(scope="module")
def lib():
    mxlib.build()
    return mxlib.load()


def test_exports_every_declared_symbol(lib):
    names = mxlib.declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"libmxadv.so does not export {n}"


def test_abi_version(lib):
    assert lib.mx_abi_version() == 1


def test_no_cpu_fallback(lib):
    """Without a CUDA device the compute entry points must fail loudly."""
    if lib.mx_device_count() > 0:
        pytest.skip("a GPU is present")
    from common import make_problem
    from maestroex_b200.synth import RHO
    from maestroex_b200.calls import MxError
    prob = make_problem(n=(8, 8, 8))
    sedge, _, _ = prob.clone_outputs()
    with pytest.raises(MxError) as ei:
        mxlib.calls().make_edge_scal(prob.geom, prob.state, sedge, prob.umac, prob.force,
                                     prob.edge_params(RHO, 1))
    assert "no CPU fallback" in str(ei.value)
    for f in sedge:
        assert np.all(f.data == -7.25e33)  # nothing was computed


def test_argument_validation(lib):
    from common import make_problem
    from maestroex_b200.synth import RHO
    from maestroex_b200.calls import MxError
    prob = make_problem(n=(8, 8, 8), ng=3)
    sedge, _, _ = prob.clone_outputs()
    bad = prob.edge_params(RHO, 1, ppm_type=7)
    with pytest.raises(MxError) as ei:
        mxlib.calls().make_edge_scal(prob.geom, prob.state, sedge, prob.umac, prob.force, bad)
    assert "ppm_type" in str(ei.value)
