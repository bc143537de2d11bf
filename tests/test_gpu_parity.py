"""GPU parity: the CUDA path (through the C ABI of libmxadv.so) against the CPU oracle on the
same seeded inputs, bit-for-bit (tolerance written below: 0, i.e. tighter than the 1e-12
relative error per zone the north star asks for).  Every case runs twice: with HOST buffers
(the library stages them) and with DEVICE buffers (torch tensors)."""
import numpy as np
import pytest

from common import (FIRSTSPEC, REL_TOL, RHO, RHOH, assert_same, compare_faces, make_problem,
                    run_edge, run_rhoh, run_rhoX)
from test_oracle_vs_reference import CASES_3D

pytestmark = pytest.mark.gpu

from common import STRICT_TOL as TOL  # 0 unless MXADV_TEST_TOL is set


@pytest.fixture(scope="module")
def mx():
    from maestroex_b200 import lib as mxlib
    lib = mxlib.load()
    assert lib.mx_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return mxlib.calls()


def _edge_case(mx, oracle_calls, dim, n, case, seed, residency, **kw):
    from gpu_util import to_device, to_host
    ppm, so, bck, cons, trace, is_vel = case
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=dim, n=n, nspec=3, ng=ng, bc_kind=bck, seed=seed,
                        force_ng=ng if trace else 1, **kw)
    if is_vel:
        par = prob.edge_params(0, dim, ppm_type=ppm, slope_order=so, ppm_trace_forces=trace,
                               is_vel=1, is_conservative=cons)
        comps = list(range(dim))
    else:
        par = prob.edge_params(RHO, prob.nscal, ppm_type=ppm, slope_order=so,
                               ppm_trace_forces=trace, is_vel=0, is_conservative=cons)
        comps = list(range(prob.nscal))
    want = run_edge(oracle_calls, prob, par)
    if residency == "host":
        got = run_edge(mx, prob, par)
    else:
        dprob = to_device(prob)
        got = to_host(run_edge(mx, dprob, par))
    compare_faces(got, want, comps, f"sedge {case} [{residency}]", TOL)


@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("case", CASES_3D)
def test_edge_3d_boundary_box(mx, oracle_calls, case, residency):
    _edge_case(mx, oracle_calls, 3, (12, 10, 14), case, 11, residency)


@pytest.mark.parametrize("case", CASES_3D[:9])
def test_edge_3d_interior_box(mx, oracle_calls, case):
    _edge_case(mx, oracle_calls, 3, (10, 9, 8), case, 5, "device", valid_lo=(9, 8, 10),
               dom_n=(32, 32, 32))


@pytest.mark.parametrize("case", CASES_3D[:9])
def test_edge_3d_partial_boundary_box(mx, oracle_calls, case):
    _edge_case(mx, oracle_calls, 3, (10, 9, 8), case, 6, "device", valid_lo=(0, 8, 24),
               dom_n=(32, 32, 32))


@pytest.mark.parametrize("case", [(1, 4, "reacting_bubble", 0, 0, 0), (2, 4, "mixed", 0, 0, 0),
                                  (0, 4, "mixed", 0, 0, 0)])
def test_edge_3d_larger_box(mx, oracle_calls, case):
    # several thread-block tiles in every direction, ragged sizes
    _edge_case(mx, oracle_calls, 3, (70, 37, 45), case, 8, "device")


@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("case", CASES_3D)
def test_edge_2d(mx, oracle_calls, case, residency):
    _edge_case(mx, oracle_calls, 2, (20, 17), case, 3, residency)


@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("spt", [1, 2, 3])
@pytest.mark.parametrize("evolve,exact", [(1, 0), (0, 0), (1, 1)])
def test_rhoX_flux(mx, oracle_calls, dim, spt, evolve, exact, residency):
    from gpu_util import to_device, to_host
    prob = make_problem(dim=dim, n=(9, 8, 11), nspec=5, bc_kind="reacting_bubble", seed=21)
    sedge = run_edge(oracle_calls, prob, prob.edge_params(RHO, prob.nscal))
    fpar = prob.flux_params(species_pred_type=spt, evolve_base_state=evolve,
                            use_exact_base_state=exact)
    fw, ew = run_rhoX(oracle_calls, prob, sedge, fpar)
    if residency == "host":
        fg, eg = run_rhoX(mx, prob, sedge, fpar)
    else:
        dprob = to_device(prob)
        dsedge = to_device_fabs(sedge)
        fg, eg = run_rhoX(mx, dprob, dsedge, fpar)
        fg = to_host(fg)
        eg = to_host([eg])[0]
    compare_faces(fg, fw, [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + 5)), "sflux", TOL)
    assert_same(eg.data, ew.data, "etarhoflux", TOL)


def to_device_fabs(fabs):
    import torch
    from maestroex_b200.abi import Fab
    return [Fab(torch.from_numpy(np.ascontiguousarray(f.data)).cuda(), f.lo, f.hi) for f in fabs]


@pytest.mark.parametrize("spt", [1, 2, 3])
def test_rhoX_flux_spherical(mx, oracle_calls, spt):
    prob = make_problem(dim=3, n=(9, 8, 11), nspec=4, bc_kind="outflow", seed=22, spherical=True)
    sedge = run_edge(oracle_calls, prob, prob.edge_params(RHO, prob.nscal))
    fpar = prob.flux_params(species_pred_type=spt, spherical=1)
    fw, ew = run_rhoX(oracle_calls, prob, sedge, fpar)
    fg, eg = run_rhoX(mx, prob, sedge, fpar)
    compare_faces(fg, fw, [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + 4)), "sflux", TOL)
    assert np.array_equal(eg.data, ew.data)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("ept", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("spt", [1, 2, 3])
def test_rhoh_flux_planar(mx, oracle_calls, dim, ept, spt):
    prob = make_problem(dim=dim, n=(9, 8, 11), nspec=3, bc_kind="reacting_bubble", seed=23)
    sedge = run_edge(oracle_calls, prob, prob.edge_params(RHO, prob.nscal))
    fpar = prob.flux_params(species_pred_type=spt, enthalpy_pred_type=ept)
    fw = run_rhoh(oracle_calls, prob, sedge, fpar)
    fg = run_rhoh(mx, prob, sedge, fpar)
    compare_faces(fg, fw, [RHOH], "rhoh flux", TOL)


@pytest.mark.parametrize("ept", [0, 1, 2, 5])
@pytest.mark.parametrize("spt", [1, 3])
@pytest.mark.parametrize("exact", [0, 1])
def test_rhoh_flux_spherical(mx, oracle_calls, ept, spt, exact):
    prob = make_problem(dim=3, n=(9, 8, 11), nspec=3, bc_kind="outflow", seed=24, spherical=True)
    sedge = run_edge(oracle_calls, prob, prob.edge_params(RHO, prob.nscal))
    fpar = prob.flux_params(species_pred_type=spt, enthalpy_pred_type=ept, spherical=1,
                            use_exact_base_state=exact)
    fw = run_rhoh(oracle_calls, prob, sedge, fpar)
    fg = run_rhoh(mx, prob, sedge, fpar)
    compare_faces(fg, fw, [RHOH], "rhoh flux", TOL)


@pytest.mark.parametrize("nr,lo,hi", [(64, 0, 63), (33, 0, 32), (64, 16, 47), (5, 0, 4),
                                      (2218, 0, 2217)])
def test_cell_to_edge(mx, oracle_calls, nr, lo, hi):
    rng = np.random.default_rng(nr)
    cell = np.exp(-5 * np.linspace(0, 1, nr)) * (1 + 0.2 * rng.random(nr))
    e_want = np.full(nr + 1, -1.0)
    e_got = np.full(nr + 1, -1.0)
    oracle_calls.cell_to_edge(cell, e_want, nr, lo, hi)
    mx.cell_to_edge(cell, e_got, nr, lo, hi)
    assert np.array_equal(e_want, e_got)


@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("ppm,bck,spt,nspec,n", [
    (1, "reacting_bubble", 1, 3, (33, 18, 21)),
    (1, "periodic", 1, 13, (40, 24, 20)),
    (2, "reacting_bubble", 1, 3, (33, 18, 21)),
    (0, "outflow", 3, 5, (17, 33, 12)),
    (1, "mixed", 1, 4, (16, 16, 16)),
])
def test_fused_density_pass(mx, oracle_calls, ppm, bck, spt, nspec, n, residency):
    """mx_density_edge_flux == MakeEdgeScal(species) + MakeEdgeScal(Rho) + MakeRhoXFlux."""
    from gpu_util import to_device, to_host
    from maestroex_b200.fusedcalls import density_edge_flux
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=n, nspec=nspec, ng=ng, bc_kind=bck, seed=31)
    epar = prob.edge_params(FIRSTSPEC, nspec, ppm_type=ppm)
    fpar = prob.flux_params(species_pred_type=spt)
    want_e = run_edge(oracle_calls, prob, epar)
    run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm), sedge=want_e)
    want_f, want_eta = run_rhoX(oracle_calls, prob, want_e, fpar)
    if residency == "host":
        ge, gf, geta = density_edge_flux(mx, prob, epar, fpar)
    else:
        ge, gf, geta = density_edge_flux(mx, to_device(prob), epar, fpar)
        ge, gf, geta = to_host(ge), to_host(gf), to_host([geta])[0]
    comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + nspec))
    compare_faces(ge, want_e, comps, "fused sedge", TOL)
    compare_faces(gf, want_f, comps, "fused sflux", TOL)
    assert_same(geta.data, want_eta.data, "fused etarhoflux", TOL)


@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("ppm,spt,nspec,n", [
    (1, 1, 3, (37, 19, 23)),
    (0, 1, 13, (33, 33, 12)),
    (2, 3, 4, (20, 18, 40)),
    (1, 1, 5, (31, 29, 35)),
])
def test_fused_density_pass_spherical(mx, oracle_calls, ppm, spt, nspec, n, residency):
    """wdconvect-like: spherical flux form (r0mac face arrays, no etarhoflux), outflow BCs;
    the time-centred rho0 on faces is pre-averaged by the library (mac_average)."""
    from gpu_util import to_device, to_host
    from maestroex_b200.fusedcalls import density_edge_flux
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=n, nspec=nspec, ng=ng, bc_kind="outflow", seed=77, spherical=True)
    epar = prob.edge_params(FIRSTSPEC, nspec, ppm_type=ppm)
    fpar = prob.flux_params(species_pred_type=spt, spherical=1)
    want_e = run_edge(oracle_calls, prob, epar)
    run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm), sedge=want_e)
    want_f, want_eta = run_rhoX(oracle_calls, prob, want_e, fpar)
    if residency == "host":
        ge, gf, geta = density_edge_flux(mx, prob, epar, fpar)
    else:
        ge, gf, geta = density_edge_flux(mx, to_device(prob), epar, fpar)
        ge, gf, geta = to_host(ge), to_host(gf), to_host([geta])[0]
    comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + nspec))
    compare_faces(ge, want_e, comps, "fused spherical sedge", TOL)
    compare_faces(gf, want_f, comps, "fused spherical sflux", TOL)
    assert np.array_equal(geta.data, want_eta.data)  # spherical: etarhoflux is not touched


@pytest.mark.parametrize("ppm,ept,cons", [(1, 1, 0), (2, 1, 0), (0, 1, 0), (1, 0, 1), (1, 2, 0)])
def test_fused_enthalpy_pass(mx, oracle_calls, ppm, ept, cons):
    from gpu_util import to_device, to_host
    from maestroex_b200.fusedcalls import enthalpy_edge_flux
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=(33, 18, 21), nspec=3, ng=ng, bc_kind="reacting_bubble", seed=32)
    # sedge(Rho) from an earlier density pass is an input of the h-type flux forms
    base_e = run_edge(oracle_calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm))
    epar = prob.edge_params(RHOH, 1, ppm_type=ppm, is_conservative=cons)
    fpar = prob.flux_params(enthalpy_pred_type=ept)
    from maestroex_b200.abi import Fab
    want_e = [Fab(f.data.copy(), f.lo, f.hi) for f in base_e]
    run_edge(oracle_calls, prob, epar, sedge=want_e)
    want_f = run_rhoh(oracle_calls, prob, want_e, fpar)
    dprob = to_device(prob)
    d_e = to_device_fabs(base_e)
    _, d_f, _ = dprob.clone_outputs()
    ge, gf = enthalpy_edge_flux(mx, dprob, epar, fpar, outputs=(d_e, d_f, None))
    ge, gf = to_host(ge), to_host(gf)
    compare_faces(ge, want_e, [RHO, RHOH], "fused sedge", TOL)
    compare_faces(gf, want_f, [RHOH], "fused rhoh flux", TOL)


@pytest.mark.parametrize("zchunk", ["1", "5", "16", "1000"])
@pytest.mark.parametrize("ppm", [0, 1, 2])
def test_fused_zchunk_independence(mx, oracle_calls, ppm, zchunk, monkeypatch):
    """The z-chunk size of the streaming kernel must not change a single bit."""
    from gpu_util import to_device, to_host
    monkeypatch.setenv("MXADV_ZCHUNK", zchunk)
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=(35, 29, 23), nspec=2, ng=ng, bc_kind="reacting_bubble", seed=41)
    par = prob.edge_params(RHO, prob.nscal, ppm_type=ppm)
    want = run_edge(oracle_calls, prob, par)
    got = to_host(run_edge(mx, to_device(prob), par))
    compare_faces(got, want, list(range(prob.nscal)), f"sedge zchunk={zchunk}", TOL)


def test_staged_and_fused_agree(mx, oracle_calls, monkeypatch):
    """MXADV_FORCE_STAGED routes the same call through the general kernels."""
    from gpu_util import to_device, to_host
    prob = make_problem(dim=3, n=(40, 20, 18), nspec=3, ng=3, bc_kind="reacting_bubble", seed=42)
    par = prob.edge_params(RHO, prob.nscal, ppm_type=1)
    fused = to_host(run_edge(mx, to_device(prob), par))
    monkeypatch.setenv("MXADV_FORCE_STAGED", "1")
    staged = to_host(run_edge(mx, to_device(prob), par))
    compare_faces(fused, staged, list(range(prob.nscal)), "fused vs staged", TOL)
