"""Pin the C oracle (oracle/mxadv_oracle.c) bit-for-bit against the REFERENCE's own sources
compiled into oracle/_ref (see oracle/Makefile), over the option space the reference
supports: ppm_type 0/1/2, slope_order 0/2/4, every BC code on every face, conservative and
non-conservative form, ppm_trace_forces, is_vel, 2-D and 3-D, boundary and interior boxes,
and all species / enthalpy flux forms (planar and spherical)."""
import numpy as np
import pytest

from common import (FIRSTSPEC, RHO, RHOH, compare_faces, assert_same, make_problem, run_edge,
                    run_rhoh, run_rhoX)

CASES_3D = [
    # (ppm_type, slope_order, bc_kind, conservative, trace_forces, is_vel)
    (1, 4, "periodic", 0, 0, 0),
    (1, 4, "reacting_bubble", 0, 0, 0),
    (1, 4, "mixed", 0, 0, 0),
    (1, 4, "mixed", 1, 0, 0),
    (1, 4, "mixed", 0, 1, 0),
    (1, 4, "mixed", 0, 0, 1),
    (2, 4, "periodic", 0, 0, 0),
    (2, 4, "mixed", 0, 0, 0),
    (2, 4, "mixed", 1, 1, 0),
    (2, 4, "reacting_bubble", 0, 0, 1),
    (0, 4, "periodic", 0, 0, 0),
    (0, 4, "mixed", 0, 0, 0),
    (0, 4, "mixed", 1, 0, 1),
    (0, 2, "mixed", 0, 0, 0),
    (0, 2, "reacting_bubble", 1, 0, 0),
    (0, 0, "mixed", 0, 0, 0),
    (0, 4, "outflow", 0, 0, 0),
]


def _edge_case(oracle_calls, ref, dim, n, case, seed, **kw):
    ppm, so, bck, cons, trace, is_vel = case
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=dim, n=n, nspec=3, ng=ng, bc_kind=bck, seed=seed,
                        force_ng=ng if trace else 1, **kw)
    if is_vel:
        # velocity sweep: components 0..dim-1 with their own BCs (comp == dir clamps)
        par = prob.edge_params(0, dim, ppm_type=ppm, slope_order=so, ppm_trace_forces=trace,
                               is_vel=1, is_conservative=cons)
        comps = list(range(dim))
    else:
        par = prob.edge_params(RHO, prob.nscal, ppm_type=ppm, slope_order=so,
                               ppm_trace_forces=trace, is_vel=0, is_conservative=cons)
        comps = list(range(prob.nscal))
    want = run_edge(ref, prob, par)
    got = run_edge(oracle_calls, prob, par)
    compare_faces(got, want, comps, f"sedge {case}")
    for f in want:  # every swept entry was really written by the reference
        assert not np.any(f.data[comps] == -7.25e33)


@pytest.mark.parametrize("case", CASES_3D)
def test_edge_3d_boundary_box(oracle_calls, ref3d, case):
    _edge_case(oracle_calls, ref3d, 3, (12, 10, 14), case, seed=11)


@pytest.mark.parametrize("case", CASES_3D[:9])
def test_edge_3d_interior_box(oracle_calls, ref3d, case):
    # a box strictly inside a larger domain: no BC row may fire
    _edge_case(oracle_calls, ref3d, 3, (10, 9, 8), case, seed=5, valid_lo=(9, 8, 10),
               dom_n=(32, 32, 32))


@pytest.mark.parametrize("case", CASES_3D[:9])
def test_edge_3d_partial_boundary_box(oracle_calls, ref3d, case):
    # touches the lo-x and hi-z domain faces only
    _edge_case(oracle_calls, ref3d, 3, (10, 9, 8), case, seed=6, valid_lo=(0, 8, 24),
               dom_n=(32, 32, 32))


@pytest.mark.parametrize("case", CASES_3D)
def test_edge_2d(oracle_calls, ref2d, case):
    _edge_case(oracle_calls, ref2d, 2, (20, 17), case, seed=3)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("spt", [1, 2, 3])
@pytest.mark.parametrize("evolve,exact", [(1, 0), (0, 0), (1, 1)])
def test_rhoX_flux(oracle_calls, ref2d, ref3d, dim, spt, evolve, exact):
    ref = ref3d if dim == 3 else ref2d
    prob = make_problem(dim=dim, n=(9, 8, 11), nspec=5, bc_kind="reacting_bubble", seed=21)
    par = prob.edge_params(RHO, prob.nscal)
    sedge = run_edge(ref, prob, par)
    fpar = prob.flux_params(species_pred_type=spt, evolve_base_state=evolve,
                            use_exact_base_state=exact)
    fw, ew = run_rhoX(ref, prob, sedge, fpar)
    fg, eg = run_rhoX(oracle_calls, prob, sedge, fpar)
    compare_faces(fg, fw, [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + 5)), "sflux")
    assert_same(eg.data, ew.data, "etarhoflux")


@pytest.mark.parametrize("spt", [1, 2, 3])
def test_rhoX_flux_spherical(oracle_calls, ref3d, spt):
    prob = make_problem(dim=3, n=(9, 8, 11), nspec=4, bc_kind="outflow", seed=22,
                        spherical=True)
    sedge = run_edge(ref3d, prob, prob.edge_params(RHO, prob.nscal))
    fpar = prob.flux_params(species_pred_type=spt, spherical=1)
    fw, ew = run_rhoX(ref3d, prob, sedge, fpar)
    fg, eg = run_rhoX(oracle_calls, prob, sedge, fpar)
    compare_faces(fg, fw, [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + 4)), "sflux")
    assert np.array_equal(eg.data, ew.data)  # untouched in the spherical branch


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("ept", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("spt", [1, 2, 3])
def test_rhoh_flux_planar(oracle_calls, ref2d, ref3d, dim, ept, spt):
    ref = ref3d if dim == 3 else ref2d
    prob = make_problem(dim=dim, n=(9, 8, 11), nspec=3, bc_kind="reacting_bubble", seed=23)
    sedge = run_edge(ref, prob, prob.edge_params(RHO, prob.nscal))
    fpar = prob.flux_params(species_pred_type=spt, enthalpy_pred_type=ept)
    fw = run_rhoh(ref, prob, sedge, fpar)
    fg = run_rhoh(oracle_calls, prob, sedge, fpar)
    for d in range(dim):
        assert_same(fg[d].data, fw[d].data, f"rhoh flux dir {d}")


@pytest.mark.parametrize("ept", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("spt", [1, 3])
@pytest.mark.parametrize("exact", [0, 1])
def test_rhoh_flux_spherical(oracle_calls, ref3d, ept, spt, exact):
    prob = make_problem(dim=3, n=(9, 8, 11), nspec=3, bc_kind="outflow", seed=24,
                        spherical=True)
    sedge = run_edge(ref3d, prob, prob.edge_params(RHO, prob.nscal))
    fpar = prob.flux_params(species_pred_type=spt, enthalpy_pred_type=ept, spherical=1,
                            use_exact_base_state=exact)
    fw = run_rhoh(ref3d, prob, sedge, fpar)
    fg = run_rhoh(oracle_calls, prob, sedge, fpar)
    for d in range(3):
        assert_same(fg[d].data, fw[d].data, f"rhoh flux dir {d}")


@pytest.mark.parametrize("nr,lo,hi", [(64, 0, 63), (33, 0, 32), (64, 16, 47), (5, 0, 4)])
def test_cell_to_edge(oracle_calls, ref3d, nr, lo, hi):
    rng = np.random.default_rng(nr)
    cell = np.exp(-5 * np.linspace(0, 1, nr)) * (1 + 0.2 * rng.random(nr))
    e_ref = np.full(nr + 1, -1.0)
    e_orc = np.full(nr + 1, -1.0)
    ref3d.cell_to_edge(cell, e_ref, nr, lo, hi)
    oracle_calls.cell_to_edge(cell, e_orc, nr, lo, hi)
    assert np.array_equal(e_ref, e_orc)
