"""mx_density_advance / mx_enthalpy_advance (the update fused
behind the flux): one library call against the ORACLE's sequence of the separate reference steps
  MakeEdgeScal(X_k), MakeEdgeScal(rho'), MakeRhoXFlux, PutInPertForm(add), ConvertRhoXToX(mult),
  scal_force = 0, UpdateScal        (Source/MaestroDensityAdvance.cpp:115-258)
  MakeEdgeScal((rho h)'), MakeRhoHFlux, PutInPertForm(add), UpdateScal(RhoH)
                                    (Source/MaestroEnthalpyAdvance.cpp:141-340)
bit for bit, with device- and host-resident fabs."""
import numpy as np
import pytest

from common import FIRSTSPEC, RHO, RHOH, make_problem, run_edge, run_rhoh, run_rhoX
from maestroex_b200.abi import Fab, UpdateParams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mx():
    from maestroex_b200 import lib as mxlib
    lib = mxlib.load()
    assert lib.mx_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return mxlib.calls()


def _valid(prob, fab):
    sl = tuple(slice(prob.geom.valid_lo[d] - fab.lo[d], prob.geom.valid_hi[d] - fab.lo[d] + 1)
               for d in (2, 1, 0))
    return (slice(None),) + sl


def _prep(ppm, nspec, n, bck, seed):
    ng = 4 if ppm == 2 else 3
    prob = make_problem(dim=3, n=n, nspec=nspec, ng=ng, bc_kind=bck, seed=seed)
    # species force is zero in DensityAdvance (MaestroDensityAdvance.cpp:36-39)
    prob.force.data[FIRSTSPEC:FIRSTSPEC + nspec] = 0.0
    return prob


@pytest.mark.parametrize("residency", ["device", "host", "host_pinned"])
@pytest.mark.parametrize("ppm,nspec,n,bck,spt,zero_flag", [
    (1, 3, (34, 30, 33), "reacting_bubble", 1, True),
    (1, 13, (40, 24, 20), "periodic", 1, False),
    (2, 3, (32, 18, 21), "reacting_bubble", 3, True),
    (0, 5, (18, 34, 40), "outflow", 1, True),
])
def test_density_advance(mx, oracle_calls, ppm, nspec, n, bck, spt, zero_flag, residency):
    from gpu_util import to_device, to_host
    from maestroex_b200.fusedcalls import density_advance
    prob = _prep(ppm, nspec, n, bck, 71)
    epar = prob.edge_params(FIRSTSPEC, nspec, ppm_type=ppm)
    fpar = prob.flux_params(species_pred_type=spt)
    upar = UpdateParams(dt=prob.dt, start_comp=FIRSTSPEC, num_comp=nspec, num_spec=nspec,
                        base_cutoff_density=0.9 * float(np.abs(prob.state.data[RHO]).max()))
    # ---- oracle: the reference's sequence of separate steps
    orc = oracle_calls
    want_e = run_edge(orc, prob, epar)
    run_edge(orc, prob, prob.edge_params(RHO, 1, ppm_type=ppm), sedge=want_e)
    want_f, want_eta = run_rhoX(orc, prob, want_e, fpar)
    sold = Fab(prob.state.data.copy(), prob.state.lo, prob.state.hi)
    if spt == 1:
        orc.put_in_pert_form(prob.geom, sold, prob.base["rho0_old"], 0, RHO, False)
    orc.convert_by_rho(prob.geom, sold, RHO, FIRSTSPEC, nspec, False)
    zero = Fab(np.zeros_like(prob.force.data), prob.force.lo, prob.force.hi)
    want_new = Fab(np.full_like(prob.state.data, -7.0), prob.state.lo, prob.state.hi)
    orc.update_scal(prob.geom, sold, want_new, want_f, zero, upar)
    # ---- product: one call
    snew = Fab(np.full_like(prob.state.data, -7.0), prob.state.lo, prob.state.hi)
    if residency == "device":
        dprob = to_device(prob)
        dnew = to_device_fab(snew)
        got_new, got_eta = density_advance(mx, dprob, dnew, epar, fpar, upar, species_force_is_zero=zero_flag)
        got_new, got_eta = to_host([got_new])[0], to_host([got_eta])[0]
    else:
        _, _, eta = prob.clone_outputs()
        hprob = prob
        if residency == "host_pinned":
            hprob, snew, eta = pinned_problem(prob), pinned_fab(snew), pinned_fab(eta)
        got_new, got_eta = density_advance(mx, hprob, snew, epar, fpar, upar, species_force_is_zero=zero_flag,
                                           outputs=(None, None, eta))
        got_new, got_eta = as_numpy(got_new), as_numpy(got_eta)
    v = _valid(prob, got_new)
    comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + nspec))
    for c in comps:
        assert np.array_equal(got_new.data[v][c], want_new.data[v][c]), f"snew comp {c}"
    assert np.array_equal(got_eta.data, want_eta.data), "etarhoflux"
    # untouched: every other component of snew, and its ghost zones
    assert (got_new.data[1:FIRSTSPEC] == -7.0).all()


def pinned_fab(f):
    """the same fab in page-locked host memory (what a CUDA-aware host application registers):
    the library then returns valid zones with zero-copy stores instead of strided copies"""
    import torch
    return Fab(torch.from_numpy(np.ascontiguousarray(f.data)).pin_memory(), f.lo, f.hi)


def pinned_problem(prob):
    import copy
    p = copy.copy(prob)
    p.state, p.force = pinned_fab(prob.state), pinned_fab(prob.force)
    p.umac = [pinned_fab(f) for f in prob.umac]
    return p


def as_numpy(f):
    return Fab(f.data.numpy() if hasattr(f.data, "numpy") else f.data, f.lo, f.hi)


def to_device_fab(f):
    import torch
    return Fab(torch.from_numpy(np.ascontiguousarray(f.data)).cuda(), f.lo, f.hi)


@pytest.mark.parametrize("residency", ["device", "host", "host_pinned"])
@pytest.mark.parametrize("ppm,ept,n,cut", [(1, 1, (34, 30, 26), 0.0), (1, 0, (34, 30, 26), 0.6),
                                           (2, 1, (32, 18, 40), 0.6), (0, 1, (20, 34, 33), 0.0)])
def test_enthalpy_advance(mx, oracle_calls, ppm, ept, n, cut, residency):
    from gpu_util import to_device, to_host
    from maestroex_b200.fusedcalls import enthalpy_advance
    rng = np.random.default_rng(5)
    prob = _prep(ppm, 3, n, "reacting_bubble", 72)
    epar = prob.edge_params(RHOH, 1, ppm_type=ppm)
    fpar = prob.flux_params(enthalpy_pred_type=ept)
    rho_new = np.abs(prob.state.data[RHO]) * (1.0 + 0.1 * rng.random(prob.state.data[RHO].shape)) + 1.0
    bcd = cut * float(rho_new.max())
    upar = UpdateParams(dt=prob.dt, start_comp=RHOH, num_comp=1, num_spec=3, base_cutoff_density=bcd,
                        do_eos_h_above_cutoff=1)
    fupd = Fab(rng.standard_normal(prob.force.data.shape), prob.force.lo, prob.force.hi)
    orc = oracle_calls
    want_e = run_edge(orc, prob, epar)
    want_f = run_rhoh(orc, prob, want_e, fpar)
    sold = Fab(prob.state.data.copy(), prob.state.lo, prob.state.hi)
    if ept == 1:
        orc.put_in_pert_form(prob.geom, sold, prob.base["rhoh0_old"], 0, RHOH, False)
    nz = int(np.prod([prob.geom.valid_hi[d] - prob.geom.valid_lo[d] + 1 for d in range(3)]))

    def fresh_new():
        a = np.full_like(prob.state.data, -7.0)
        a[RHO] = rho_new
        return Fab(a, prob.state.lo, prob.state.hi)

    want_new, want_mask = fresh_new(), np.zeros(nz, dtype=np.uint8)
    want_n = orc.update_scal(prob.geom, sold, want_new, want_f, fupd, upar, eos_mask=want_mask)
    snew, mask = fresh_new(), np.zeros(nz, dtype=np.uint8)
    if residency == "device":
        import torch
        dprob = to_device(prob)
        dnew, dfu = to_device_fab(snew), to_device_fab(fupd)
        dmask = torch.zeros(nz, dtype=torch.uint8, device="cuda")
        got_n = enthalpy_advance(mx, dprob, dnew, dfu, epar, fpar, upar, eos_mask=dmask)
        got_new, mask = to_host([dnew])[0], dmask.cpu().numpy()
    else:
        hprob = prob
        if residency == "host_pinned":
            hprob, snew, fupd = pinned_problem(prob), pinned_fab(snew), pinned_fab(fupd)
        got_n = enthalpy_advance(mx, hprob, snew, fupd, epar, fpar, upar, eos_mask=mask)
        got_new = as_numpy(snew)
    v = _valid(prob, got_new)
    assert got_n == want_n
    assert np.array_equal(mask, want_mask)
    assert np.array_equal(got_new.data[v][RHOH], want_new.data[v][RHOH])
    assert (got_new.data[2:] == -7.0).all()
