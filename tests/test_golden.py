"""Golden vectors produced by the reference itself (tests/golden/make_golden.py, from
oracle/_ref = the reference's own hot-path sources compiled unmodified): the C oracle (CPU)
and libmxadv.so (GPU, device and host residency) must reproduce them bit for bit.  The
fixtures carry their own inputs, so nothing here depends on the generator's RNG."""
import glob
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from make_golden import CASES, build, run_case  # noqa: E402
from maestroex_b200.abi import Base  # noqa: E402


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    prob, ekw, fkw = build(name)
    # overwrite every generated input with the stored one
    prob.state.data[...] = z["state"]
    prob.force.data[...] = z["force"]
    prob.bc[...] = z["bc"]
    prob.dt, prob.rel_eps = float(z["dt"]), float(z["rel_eps"])
    prob.etarhoflux.data[...] = z["eta_in"]
    for i, u in enumerate(prob.umac):
        u.data[...] = z[f"umac{i}"]
    for k in list(prob.base):
        prob.base[k] = Base(np.ascontiguousarray(z[f"base_{k}"]))
    return prob, ekw, fkw, z


def check(z, se, sf, eta, nscal):
    for d in range(len(se)):
        assert np.array_equal(se[d].data, z[f"sedge{d}"]), f"sedge dir {d}"
        assert np.array_equal(sf[d].data, z[f"sflux{d}"]), f"sflux dir {d}"
    assert np.array_equal(eta.data, z["eta_out"])


def test_fixtures_present():
    have = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))}
    assert set(CASES) | {"cell_to_edge"} <= have


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(oracle_calls, name):
    prob, ekw, fkw, z = load(name)
    se, sf, eta = run_case(oracle_calls, prob, ekw, fkw)
    check(z, se, sf, eta, prob.nscal)


def test_oracle_cell_to_edge_golden(oracle_calls):
    z = np.load(os.path.join(HERE, "golden", "cell_to_edge.npz"))
    nr = z["cell"].size
    for lo, hi in ((0, nr - 1), (8, 39)):
        edge = np.full(nr + 1, -3.0)
        oracle_calls.cell_to_edge(z["cell"], edge, nr, lo, hi)
        assert np.array_equal(edge, z[f"edge_{lo}_{hi}"])


@pytest.mark.gpu
@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_matches_golden(name, residency):
    from gpu_util import to_device, to_host
    from maestroex_b200 import lib as mxlib
    prob, ekw, fkw, z = load(name)
    calls = mxlib.calls()
    if residency == "device":
        dprob = to_device(prob)
        se, sf, eta = run_case(calls, dprob, ekw, fkw)
        se, sf = to_host(se), to_host(sf)
        eta = to_host([eta])[0]
    else:
        se, sf, eta = run_case(calls, prob, ekw, fkw)
    check(z, se, sf, eta, prob.nscal)


@pytest.mark.gpu
def test_gpu_cell_to_edge_golden():
    from maestroex_b200 import lib as mxlib
    z = np.load(os.path.join(HERE, "golden", "cell_to_edge.npz"))
    nr = z["cell"].size
    for lo, hi in ((0, nr - 1), (8, 39)):
        edge = np.full(nr + 1, -3.0)
        mxlib.calls().cell_to_edge(z["cell"], edge, nr, lo, hi)
        assert np.array_equal(edge, z[f"edge_{lo}_{hi}"])
