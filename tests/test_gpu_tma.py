"""k_stream4 with its shared-memory ring filled by TMA (cp.async.bulk.tensor + mbarrier) and
with the cp.async fallback, against the oracle on the same inputs, bit for bit, with a check of
WHICH staging ran.

TMA needs fab rows of an even number of doubles and a 16-byte aligned base; boxes that do not
qualify take the cp.async instantiation.  Every case here has even rows, so the default
dispatch must launch the TMA instantiation (mx_tma_launch_count() > 0) and MXADV_NO_TMA=1 must
not.  (tests/test_gpu_parity.py holds the odd-row boxes.)"""
import numpy as np
import pytest

from common import (FIRSTSPEC, RHO, RHOH, assert_same, compare_faces, make_problem, run_edge,
                    run_rhoh, run_rhoX)
from common import STRICT_TOL as TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mxl():
    from maestroex_b200 import lib as mxlib
    lib = mxlib.load()
    assert lib.mx_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return lib, mxlib.calls()


def _expect_kernel(lib, kernel):
    n_all, n_tma = lib.mx_launch_count(), lib.mx_tma_launch_count()
    assert n_all > 0
    if kernel == "cp.async":
        assert n_tma == 0, "MXADV_NO_TMA=1 still launched the TMA instantiation"
    else:
        assert n_tma > 0, "even-row box did not take the TMA instantiation"


KERNELS = [None, "cp.async"]
# (ppm, slope_order, bc kind, is_vel)
EDGE_CASES = [
    (0, 4, "periodic", 0), (0, 2, "mixed", 0), (0, 4, "outflow", 0), (0, 0, "periodic", 0),
    (1, 4, "periodic", 0), (1, 4, "reacting_bubble", 0), (1, 4, "mixed", 0), (1, 4, "outflow", 1),
    (2, 4, "periodic", 0), (2, 4, "mixed", 0), (2, 4, "reacting_bubble", 1),
]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("case", EDGE_CASES)
@pytest.mark.parametrize("n,zchunk", [((64, 30, 40), None), ((34, 44, 70), "9"), ((12, 10, 14), None)])
def test_edge_states(mxl, oracle_calls, case, n, zchunk, kernel, monkeypatch):
    """MakeEdgeScal: several tiles in x and y, several z chunks (incl. the short boundary
    chunks), partial tiles, all ppm types, physical boundaries on every face."""
    from gpu_util import to_device, to_host
    lib, mx = mxl
    ppm, so, bck, is_vel = case
    if kernel:
        monkeypatch.setenv("MXADV_NO_TMA", "1")
    if zchunk:
        monkeypatch.setenv("MXADV_ZCHUNK", zchunk)
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=n, nspec=2, ng=ng, bc_kind=bck, seed=61)
    if is_vel:
        par = prob.edge_params(0, 3, ppm_type=ppm, slope_order=so, is_vel=1)
        comps = [0, 1, 2]
    else:
        par = prob.edge_params(RHO, prob.nscal, ppm_type=ppm, slope_order=so)
        comps = list(range(prob.nscal))
    want = run_edge(oracle_calls, prob, par)
    lib.mx_reset_launch_count()
    got = to_host(run_edge(mx, to_device(prob), par))
    if bck != "mixed":  # "mixed" has ext_dir faces: those sweeps take the staged kernels
        _expect_kernel(lib, kernel)
    compare_faces(got, want, comps, f"sedge {case} n={n} kernel={kernel}", TOL)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("ppm,bck,spt,nspec,n,sph", [
    (1, "reacting_bubble", 1, 3, (62, 30, 33), False),
    (1, "periodic", 1, 13, (40, 24, 20), False),
    (2, "reacting_bubble", 1, 3, (32, 18, 21), False),
    (0, "outflow", 3, 5, (18, 34, 12), False),
    (1, "outflow", 1, 4, (36, 20, 24), True),
    (2, "outflow", 3, 3, (20, 18, 40), True),
    (0, "outflow", 1, 13, (34, 32, 12), True),
])
def test_density_pass(mxl, oracle_calls, ppm, bck, spt, nspec, n, sph, kernel, monkeypatch):
    from gpu_util import to_device, to_host
    from maestroex_b200.fusedcalls import density_edge_flux
    lib, mx = mxl
    if kernel:
        monkeypatch.setenv("MXADV_NO_TMA", "1")
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=n, nspec=nspec, ng=ng, bc_kind=bck, seed=62, spherical=sph)
    epar = prob.edge_params(FIRSTSPEC, nspec, ppm_type=ppm)
    fpar = prob.flux_params(species_pred_type=spt, spherical=1 if sph else 0)
    want_e = run_edge(oracle_calls, prob, epar)
    run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm), sedge=want_e)
    want_f, want_eta = run_rhoX(oracle_calls, prob, want_e, fpar)
    lib.mx_reset_launch_count()
    ge, gf, geta = density_edge_flux(mx, to_device(prob), epar, fpar)
    ge, gf, geta = to_host(ge), to_host(gf), to_host([geta])[0]
    _expect_kernel(lib, kernel)
    comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + nspec))
    compare_faces(ge, want_e, comps, "fused sedge", TOL)
    compare_faces(gf, want_f, comps, "fused sflux", TOL)
    assert_same(geta.data, want_eta.data, "fused etarhoflux", TOL)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("ppm,ept,spt", [(1, 1, 1), (1, 0, 1), (2, 5, 3), (0, 2, 1), (1, 2, 3)])
def test_enthalpy_pass(mxl, oracle_calls, ppm, ept, spt, kernel, monkeypatch):
    from gpu_util import to_device, to_host
    from maestroex_b200.fusedcalls import enthalpy_edge_flux
    lib, mx = mxl
    if kernel:
        monkeypatch.setenv("MXADV_NO_TMA", "1")
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=(34, 30, 26), nspec=3, ng=ng, bc_kind="reacting_bubble", seed=63)
    # sedge(Rho) from an earlier density pass is an input of the h-type flux forms
    import torch
    from maestroex_b200.abi import Fab
    base_e = run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm))
    epar = prob.edge_params(RHOH, 1, ppm_type=ppm)
    fpar = prob.flux_params(species_pred_type=spt, enthalpy_pred_type=ept)
    want_e = [Fab(f.data.copy(), f.lo, f.hi) for f in base_e]
    run_edge(oracle_calls, prob, epar, sedge=want_e)
    want_f = run_rhoh(oracle_calls, prob, want_e, fpar)
    dprob = to_device(prob)
    d_e = [Fab(torch.from_numpy(np.ascontiguousarray(f.data)).cuda(), f.lo, f.hi) for f in base_e]
    _, d_f, _ = dprob.clone_outputs()
    lib.mx_reset_launch_count()
    ge, gf = enthalpy_edge_flux(mx, dprob, epar, fpar, outputs=(d_e, d_f, None))
    ge, gf = to_host(ge), to_host(gf)
    _expect_kernel(lib, kernel)
    compare_faces(ge, want_e, [RHO, RHOH], "fused sedge", TOL)
    compare_faces(gf, want_f, [RHOH], "fused rhoh flux", TOL)


@pytest.mark.parametrize("residency", ["device", "host"])
@pytest.mark.parametrize("ppm,ept,spt,exact", [(1, 1, 1, 0), (1, 2, 1, 0), (0, 5, 1, 0), (2, 0, 3, 0),
                                              (1, 2, 3, 0), (1, 1, 1, 1)])
def test_enthalpy_pass_spherical(mxl, oracle_calls, ppm, ept, spt, exact, residency):
    """mx_enthalpy_edge_flux_sph == MakeEdgeScal(RhoH) + MakeRhoHFlux with the spherical flux
    forms (MaestroMakeFlux.cpp:788-1021): fused in k_stream4 unless use_exact_base_state."""
    import torch
    from gpu_util import to_device, to_host
    from maestroex_b200.abi import Fab
    from maestroex_b200.fusedcalls import enthalpy_edge_flux_sph
    lib, mx = mxl
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=(34, 30, 26), nspec=3, ng=ng, bc_kind="outflow", seed=64,
                        spherical=True)
    base_e = run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm))
    epar = prob.edge_params(RHOH, 1, ppm_type=ppm)
    fpar = prob.flux_params(species_pred_type=spt, enthalpy_pred_type=ept, spherical=1,
                            use_exact_base_state=exact)
    want_e = [Fab(f.data.copy(), f.lo, f.hi) for f in base_e]
    run_edge(oracle_calls, prob, epar, sedge=want_e)
    want_f = run_rhoh(oracle_calls, prob, want_e, fpar)
    lib.mx_reset_launch_count()
    if residency == "device":
        dprob = to_device(prob)
        d_e = [Fab(torch.from_numpy(np.ascontiguousarray(f.data)).cuda(), f.lo, f.hi) for f in base_e]
        _, d_f, _ = dprob.clone_outputs()
        ge, gf = enthalpy_edge_flux_sph(mx, dprob, epar, fpar, outputs=(d_e, d_f, None))
        ge, gf = to_host(ge), to_host(gf)
    else:
        h_e = [Fab(f.data.copy(), f.lo, f.hi) for f in base_e]
        _, h_f, _ = prob.clone_outputs()
        ge, gf = enthalpy_edge_flux_sph(mx, prob, epar, fpar, outputs=(h_e, h_f, None))
    if not exact:
        assert lib.mx_tma_launch_count() == 1, "spherical enthalpy pass did not run fused"
    compare_faces(ge, want_e, [RHO, RHOH], "fused spherical sedge", TOL)
    compare_faces(gf, want_f, [RHOH], "fused spherical rhoh flux", TOL)
