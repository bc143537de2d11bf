"""Shared helpers for the parity tests."""
import itertools

import numpy as np

from maestroex_b200.abi import Fab
from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH, make_problem  # noqa: F401

# north_star: "edge states and fluxes must agree to a maximum relative error of 1e-12 per zone"
REL_TOL = 1.0e-12
# Tolerance the GPU parity tests apply.  0 = bit-exact, which the shipped (strict) kernels meet;
# MXADV_TEST_TOL=1e-12 is for A/B runs of tolerance-mode builds (FMA contraction on).
import os as _os
STRICT_TOL = float(_os.environ.get("MXADV_TEST_TOL", "0.0"))
assert STRICT_TOL <= REL_TOL


def max_rel_err(a: np.ndarray, ref: np.ndarray) -> float:
    """max over zones of |a-ref|/|ref| (0 where both are exactly equal, incl. +-0)."""
    a = np.asarray(a)
    ref = np.asarray(ref)
    diff = np.abs(a - ref)
    den = np.abs(ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(diff == 0.0, 0.0, diff / den)
    return float(rel.max()) if rel.size else 0.0


def assert_same(a, ref, what="", tol=0.0):
    err = max_rel_err(a, ref)
    assert err <= tol, f"{what}: max rel err {err:.3e} > {tol:.1e}"


def run_edge(calls, prob, par, sedge=None):
    sedge = sedge if sedge is not None else prob.clone_outputs()[0]
    calls.make_edge_scal(prob.geom, prob.state, sedge, prob.umac, prob.force, par)
    return sedge


def run_rhoX(calls, prob, sedge, fpar):
    _, sflux, eta = prob.clone_outputs()
    b = prob.base
    calls.make_rhoX_flux(prob.geom, sflux, eta, sedge, prob.umac, b["rho0_old"],
                         b["rho0_edge_old"], prob.mac.get("r0mac_old"), b["rho0_new"],
                         b["rho0_edge_new"], prob.mac.get("r0mac_new"),
                         b["rho0_predicted_edge"], b["w0"], fpar)
    return sflux, eta


def run_rhoh(calls, prob, sedge, fpar, sflux=None):
    sflux = sflux if sflux is not None else prob.clone_outputs()[1]
    b = prob.base
    m = prob.mac
    calls.make_rhoh_flux(prob.geom, sflux, sedge, prob.umac, b["rho0_old"], b["rho0_edge_old"],
                         m.get("r0mac_old"), b["rho0_new"], b["rho0_edge_new"],
                         m.get("r0mac_new"), b["rhoh0_old"], b["rhoh0_edge_old"],
                         m.get("rh0mac_old"), b["rhoh0_new"], b["rhoh0_edge_new"],
                         m.get("rh0mac_new"), m.get("h0mac_old"), m.get("h0mac_new"), fpar)
    return sflux


def compare_faces(got, want, comps, what, tol=0.0):
    for d, (g, w) in enumerate(zip(got, want)):
        for c in comps:
            assert_same(g.data[c], w.data[c], f"{what} dir {d} comp {c}", tol)
        # nothing outside `comps` may have been written
        others = [c for c in range(g.data.shape[0]) if c not in comps]
        for c in others:
            assert np.array_equal(g.data[c], w.data[c]), f"{what}: comp {c} dir {d} was touched"
