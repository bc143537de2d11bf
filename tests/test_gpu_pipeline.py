"""Host-fab callers (CPU AMReX builds, bench.py's `e2e`): mx_density_edge_flux /
mx_enthalpy_edge_flux cut the box into z slabs and overlap the upload of slab s+1, the kernel
of slab s and the download of slab s-1 (Session::run_slabs, maestroex_b200/csrc/mx_api.cu).
The results must be what the CPU oracle computes, bit for bit, for every slab thickness —
including the shared z faces between slabs, the accumulated etarhoflux (read-modify-write:
uploaded per slab in kernel order) and untouched components."""
import numpy as np
import pytest

from common import (FIRSTSPEC, RHO, RHOH, assert_same, compare_faces, make_problem, run_edge,
                    run_rhoh, run_rhoX)

pytestmark = pytest.mark.gpu
from common import STRICT_TOL as TOL  # 0 unless MXADV_TEST_TOL is set


@pytest.fixture(scope="module")
def mx():
    from maestroex_b200 import lib as mxlib
    lib = mxlib.load()
    assert lib.mx_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return mxlib.calls()


@pytest.mark.parametrize("slab", [None, "16", "7", "33", "off"])
@pytest.mark.parametrize("ppm,bck,nspec,n", [
    (1, "reacting_bubble", 3, (20, 18, 70)),
    (2, "reacting_bubble", 4, (33, 17, 45)),
    (0, "periodic", 5, (18, 31, 64)),
])
def test_density_pass_host_slabs(mx, oracle_calls, monkeypatch, ppm, bck, nspec, n, slab):
    from maestroex_b200.fusedcalls import density_edge_flux
    if slab == "off":
        monkeypatch.setenv("MXADV_NO_PIPELINE", "1")
    elif slab:
        monkeypatch.setenv("MXADV_SLAB", slab)
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=n, nspec=nspec, ng=ng, bc_kind=bck, seed=91)
    epar = prob.edge_params(FIRSTSPEC, nspec, ppm_type=ppm)
    fpar = prob.flux_params(species_pred_type=1)
    want_e = run_edge(oracle_calls, prob, epar)
    run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm), sedge=want_e)
    want_f, want_eta = run_rhoX(oracle_calls, prob, want_e, fpar)
    ge, gf, geta = density_edge_flux(mx, prob, epar, fpar)  # host fabs
    comps = list(range(prob.nscal))  # swept AND untouched components
    compare_faces(ge, want_e, comps, f"sedge slab={slab}", TOL)
    compare_faces(gf, want_f, comps, f"sflux slab={slab}", TOL)
    assert_same(geta.data, want_eta.data, f"etarhoflux slab={slab}", TOL)


@pytest.mark.parametrize("slab", [None, "9", "off"])
@pytest.mark.parametrize("ppm,ept", [(1, 1), (2, 1), (0, 2)])
def test_enthalpy_pass_host_slabs(mx, oracle_calls, monkeypatch, ppm, ept, slab):
    from maestroex_b200.abi import Fab
    from maestroex_b200.fusedcalls import enthalpy_edge_flux
    if slab == "off":
        monkeypatch.setenv("MXADV_NO_PIPELINE", "1")
    elif slab:
        monkeypatch.setenv("MXADV_SLAB", slab)
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=(21, 19, 50), nspec=3, ng=ng, bc_kind="reacting_bubble", seed=92)
    base_e = run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm))
    epar = prob.edge_params(RHOH, 1, ppm_type=ppm)
    fpar = prob.flux_params(enthalpy_pred_type=ept)
    want_e = [Fab(f.data.copy(), f.lo, f.hi) for f in base_e]
    run_edge(oracle_calls, prob, epar, sedge=want_e)
    want_f = run_rhoh(oracle_calls, prob, want_e, fpar)
    got_e = [Fab(f.data.copy(), f.lo, f.hi) for f in base_e]
    _, got_f, _ = prob.clone_outputs()
    enthalpy_edge_flux(mx, prob, epar, fpar, outputs=(got_e, got_f, None))
    compare_faces(got_e, want_e, list(range(prob.nscal)), f"sedge slab={slab}", TOL)
    compare_faces(got_f, want_f, [RHOH], f"rhoh flux slab={slab}", TOL)
