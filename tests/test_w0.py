"""Maestro::Makew0 (Source/MaestroMakew0.cpp:5-272,537-740,942-969 + MakeGravEdge,
Source/MaestroMakeGrav.cpp:303-322): oracle vs the compiled reference, oracle vs the golden
vectors generated from the reference (tests/golden/make_golden.py), and — on the GPU —
libmxadv.so (host and device residency) and the C++ overlay vs both.  Bit-exact everywhere."""
import os

import numpy as np
import pytest

from maestroex_b200.synth import W0_CASES, make_w0_problem, run_w0

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = sorted(W0_CASES)


def golden(name):
    z = np.load(os.path.join(HERE, "golden", "w0_" + name + ".npz"))
    sph, par, d = make_w0_problem(name)
    for k in d:  # the fixture carries its own inputs
        d[k] = np.ascontiguousarray(z["in_" + k])
    return sph, par, d, z


def f0(name):
    """first w0_force entry that is compared: Makew0SphrIrreg reads r_edge_loc(0,-1) for cell 0
    (an out-of-bounds read in the reference, Source/MaestroMakew0.cpp:926), so its w0_force(0)
    is not defined by the reference"""
    return 1 if name.startswith("irreg") else 0


def check(z, w0, force, pb, name=""):
    assert np.array_equal(w0, z["w0"])
    assert np.array_equal(force[f0(name):], z["w0_force"][f0(name):])
    assert pb == float(z["p0bdot"])


@pytest.mark.parametrize("name", NAMES)
def test_oracle_vs_reference(oracle_calls, ref3d, name):
    sph, par, d = make_w0_problem(name)
    rw, rf, rpb = run_w0(ref3d, sph, par, d)
    ow, of, opb = run_w0(oracle_calls, sph, par, d)
    assert np.isfinite(rw).all() and np.abs(rw).max() > 0.0
    assert np.array_equal(rw, ow)
    assert np.array_equal(rf[f0(name):], of[f0(name):])
    assert rpb == opb


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_golden(oracle_calls, name):
    sph, par, d, z = golden(name)
    check(z, *run_w0(oracle_calls, sph, par, d), name=name)


def test_planar_closed_box_property(oracle_calls):
    """add_pb makes the top-of-domain w0 vanish (MaestroMakew0.cpp:175-184): a
    size-independent property of the closed-box correction."""
    sph, par, d = make_w0_problem("planar_odd_pb")
    w0, _, pb = run_w0(oracle_calls, sph, par, d)
    assert pb != 0.0
    assert abs(w0[-1]) <= 1e-12 * np.abs(w0).max()


@pytest.mark.gpu
@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_golden_and_oracle(oracle_calls, name, residency):
    import torch
    from maestroex_b200 import lib as mxlib
    sph, par, d, z = golden(name)
    put = (lambda a: a) if residency == "host" else (lambda a: torch.from_numpy(a).cuda())
    mxlib.load().mx_reset_launch_count()
    w0, force, pb = run_w0(mxlib.calls(), sph, par, d, put=put)
    assert mxlib.load().mx_launch_count() == 1
    if residency == "device":
        w0, force = w0.cpu().numpy(), force.cpu().numpy()
    check(z, w0, force, pb, name=name)
    ow, of, opb = run_w0(oracle_calls, sph, par, d)
    assert np.array_equal(w0, ow) and np.array_equal(force, of) and pb == opb


@pytest.mark.gpu
def test_gpu_tridiag_failure_is_reported():
    """B(0) is fixed to 1 by the boundary row, so provoke `bet == 0` instead: with p0 == 0
    everywhere every interior row is all-zero and the elimination breaks down at j = 1."""
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.calls import MxError
    sph, par, d = make_w0_problem("sphr_small")
    d["p0_old"] = np.zeros_like(d["p0_old"])
    d["p0_new"] = np.zeros_like(d["p0_new"])
    with pytest.raises(MxError) as ei:
        run_w0(mxlib.calls(), sph, par, d)
    assert "TRIDIAG FAILED" in str(ei.value)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["planar_corr_pb", "planar_wd", "sphr_odd", "sphr_wd", "irreg_odd"])
def test_overlay_makew0(oracle_calls, name):
    """maestroex_b200/overlay/MaestroMakew0.cpp (Maestro::Makew0 on libmxadv.so) driven like
    the compiled reference."""
    from oracle import backends
    ovl = backends.overlay(3)
    if ovl is None:
        pytest.skip("oracle/_ref/libmxadv_overlay3d.so not built")
    sph, par, d, z = golden(name)
    check(z, *run_w0(ovl, sph, par, d), name=name)
