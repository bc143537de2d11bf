"""The drop-in boundary end to end: the overlay translation units in maestroex_b200/overlay
(same `Maestro::MakeEdgeScal / MakeRhoXFlux / MakeRhoHFlux / CelltoEdge` signatures as
Source/Maestro.H:230-295,401-402,913-919) are compiled against the single-box AMReX stand-in
and driven exactly like the compiled reference (oracle/ref_driver.cpp); inputs are host
MultiFabs, so the library stages them to the GPU.  Results must equal the C oracle bit for bit
(and hence the reference, tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest

from common import REL_TOL, compare_faces, make_problem, run_edge, run_rhoh, run_rhoX
from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH

pytestmark = pytest.mark.gpu


def _overlay(dim):
    from oracle import backends
    o = backends.overlay(dim)
    if o is None:
        pytest.skip("oracle/_ref/libmxadv_overlay*.so not built")
    return o


@pytest.mark.parametrize("ppm", [0, 1, 2])
@pytest.mark.parametrize("bc_kind", ["periodic", "reacting_bubble", "mixed"])
def test_overlay_density_enthalpy_3d(oracle_calls, ppm, bc_kind):
    ovl = _overlay(3)
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=(20, 18, 16), nspec=3, ng=ng, bc_kind=bc_kind, seed=11 + ppm)
    comps = [RHO, RHOH] + list(range(FIRSTSPEC, FIRSTSPEC + prob.nspec))
    res = []
    for calls in (oracle_calls, ovl):
        se = run_edge(calls, prob, prob.edge_params(FIRSTSPEC, prob.nspec, ppm_type=ppm))
        run_edge(calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm), sedge=se)
        run_edge(calls, prob, prob.edge_params(RHOH, 1, ppm_type=ppm), sedge=se)
        sf, eta = run_rhoX(calls, prob, se, prob.flux_params())
        run_rhoh(calls, prob, se, prob.flux_params(), sflux=sf)
        res.append((se, sf, eta))
    (we, wf, weta), (ge, gf, geta) = res
    compare_faces(ge, we, comps, "overlay sedge", REL_TOL)
    compare_faces(gf, wf, comps, "overlay sflux", REL_TOL)
    assert np.array_equal(geta.data, weta.data)


def test_overlay_2d(oracle_calls):
    ovl = _overlay(2)
    prob = make_problem(dim=2, n=(24, 20), nspec=3, ng=3, bc_kind="reacting_bubble", seed=4)
    comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + prob.nspec))
    res = []
    for calls in (oracle_calls, ovl):
        se = run_edge(calls, prob, prob.edge_params(FIRSTSPEC, prob.nspec, ppm_type=1))
        run_edge(calls, prob, prob.edge_params(RHO, 1, ppm_type=1), sedge=se)
        sf, eta = run_rhoX(calls, prob, se, prob.flux_params())
        res.append((se, sf, eta))
    (we, wf, weta), (ge, gf, geta) = res
    compare_faces(ge, we, comps, "overlay 2-D sedge", REL_TOL)
    compare_faces(gf, wf, comps, "overlay 2-D sflux", REL_TOL)
    assert np.array_equal(geta.data, weta.data)


def test_overlay_cell_to_edge(oracle_calls):
    ovl = _overlay(3)
    rng = np.random.default_rng(0)
    nr = 96
    cell = np.exp(-5.0 * np.arange(nr) / nr) * (1.0 + 0.1 * rng.random(nr))
    want = np.full(nr + 1, -7.0)
    got = want.copy()
    oracle_calls.cell_to_edge(cell, want, nr, 0, nr - 1)
    ovl.cell_to_edge(cell, got, nr, 0, nr - 1)
    assert np.array_equal(got, want)
