"""Maestro::UpdateScal (Source/MaestroAdvection.cpp:86-251):
oracle vs the compiled reference, and -- on the GPU -- libmxadv.so (host and device
residency) and the C++ overlay vs the oracle / reference.  Bit-exact.

The reference calls the Microphysics EOS for zones with rho_new <= base_cutoff_density
(:137-159).  That EOS is third-party and stays with the caller: oracle and product flag those
zones instead (eos_mask).  The compiled reference here runs with the stand-in EOS of
oracle/shim/Maestro.H; `apply_standin_eos` is the same function in numpy, applied to the
flagged zones before comparing."""
import numpy as np
import pytest

from maestroex_b200.abi import Fab, UpdateParams
from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH, TEMP, make_problem

CASES = [
    # dim, n, nspec, seed, base_cutoff_density scale (of max rho), negative species?
    (3, (17, 12, 9), 3, 1, 0.0, False),
    (3, (16, 16, 16), 13, 2, 0.5, True),
    (3, (9, 21, 14), 5, 3, 3.0, True),
    (2, (24, 20), 3, 4, 0.5, True),
    (2, (33, 7), 4, 5, 0.0, False),
]


def build(dim, n, nspec, seed, cut_scale, negatives):
    """stateold = conserved (rho, rho h, T, pi, rho X_k); fluxes and forces sized so that a
    fraction of the zones goes through the density floor and the negative-species branch."""
    rng = np.random.default_rng(100 + seed)
    prob = make_problem(dim=dim, n=n, nspec=nspec, ng=3, bc_kind="periodic", seed=seed)
    so = prob.state.data
    shp = so.shape[1:]
    rho = 1.0e5 * (1.0 + rng.random(shp))
    X = rng.random((nspec,) + shp) + 0.05
    X /= X.sum(axis=0, keepdims=True)
    so[RHO] = rho
    so[RHOH] = rho * 1.0e8 * (1.0 + 0.2 * rng.random(shp))
    so[TEMP] = 1.0e8 * (1.0 + 0.1 * rng.random(shp))
    so[FIRSTSPEC:] = rho * X
    if negatives:
        so[FIRSTSPEC + nspec - 1] *= 1.0e-6  # a trace species the fluxes can drive negative
    dx = prob.geom.dx
    amp = 0.3 * min(dx) / prob.dt  # |dt div F| up to ~ the zone content
    for d, f in enumerate(prob.sflux):
        f.data[...] = amp * 1.0e5 * rng.standard_normal(f.data.shape) * 0.02
        f.data[RHOH] *= 1.0e8
        if negatives:
            f.data[FIRSTSPEC + nspec - 1] *= 1.0e-3
    force = prob.force
    force.data[...] = 0.0
    force.data[RHOH] = 1.0e13 * rng.standard_normal(force.data.shape[1:])
    force.data[FIRSTSPEC:] = 1.0e3 * rng.standard_normal((nspec,) + force.data.shape[1:])
    snew = Fab(np.full_like(so, -4.5e33), prob.state.lo, prob.state.hi)
    glo, ghi = prob.geom.cell_box(0)
    p0 = Fab(1.0e22 * (1.0 + rng.random((1,) + tuple(ghi[q] - glo[q] + 1 for q in (2, 1, 0)))), glo, ghi)
    bcd = cut_scale * 1.0e5
    return prob, snew, p0, bcd


def params(prob, start, bcd, do_eos=1):
    return UpdateParams(dt=prob.dt, start_comp=start, num_comp=prob.nspec if start == FIRSTSPEC else 1,
                        num_spec=prob.nspec, base_cutoff_density=bcd, do_eos_h_above_cutoff=do_eos)


def valid(prob, fab):
    g = prob.geom
    sl = tuple(slice(g.valid_lo[d] - fab.lo[d], g.valid_hi[d] - fab.lo[d] + 1) if d < g.dim
               else slice(None) for d in (2, 1, 0))
    return (slice(None),) + sl


def apply_standin_eos(prob, snew, sold, p0, mask):
    """oracle/shim/Maestro.H eos(): h = 2.5 p / rho (1 + 0.1 sum_n X_n (1 + 0.25 n))."""
    v = valid(prob, snew)
    sn = snew.data[v]
    m = mask.reshape(sn.shape[1:]).astype(bool)
    rho = sn[RHO]
    abar = np.zeros_like(rho)
    for n in range(prob.nspec):
        abar = abar + sn[FIRSTSPEC + n] / rho * (1.0 + 0.25 * n)
    h = 2.5 * p0.data[0] / rho * (1.0 + 0.1 * abar)
    sn[RHOH] = np.where(m, rho * h, sn[RHOH])
    snew.data[v] = sn


def run_both_passes(calls, prob, snew, p0, bcd, reference_like, put=lambda f: f, get=lambda f: f):
    """species + density update, then the enthalpy update (DensityAdvance / EnthalpyAdvance
    order, so that the EOS flag sees the new density)."""
    so, fo, fl = put(prob.state), put(prob.force), [put(f) for f in prob.sflux]
    sn = put(snew)
    nz = int(np.prod([prob.geom.valid_hi[d] - prob.geom.valid_lo[d] + 1 for d in range(prob.geom.dim)]))
    if reference_like:
        calls.update_scal(prob.geom, so, sn, fl, fo, params(prob, FIRSTSPEC, bcd), p0_cart=p0)
        calls.update_scal(prob.geom, so, sn, fl, fo, params(prob, RHOH, bcd), p0_cart=p0)
        return get(sn), None
    calls.update_scal(prob.geom, so, sn, fl, fo, params(prob, FIRSTSPEC, bcd))
    mask = np.zeros(nz, dtype=np.uint8)
    n = calls.update_scal(prob.geom, so, sn, fl, fo, params(prob, RHOH, bcd), eos_mask=mask)
    assert n == int(mask.sum())
    out = get(sn)
    apply_standin_eos(prob, out, prob.state, p0, mask)
    return out, mask


@pytest.mark.parametrize("case", CASES)
def test_oracle_vs_reference(oracle_calls, ref3d, ref2d, case):
    prob, snew, p0, bcd = build(*case)
    ref = ref3d if case[0] == 3 else ref2d
    want, _ = run_both_passes(ref, prob, Fab(snew.data.copy(), snew.lo, snew.hi), p0, bcd, True)
    got, mask = run_both_passes(oracle_calls, prob, Fab(snew.data.copy(), snew.lo, snew.hi), p0, bcd, False)
    assert np.array_equal(got.data, want.data, equal_nan=True)
    v = valid(prob, got)
    assert np.isfinite(got.data[v][[RHO, RHOH] + list(range(FIRSTSPEC, FIRSTSPEC + prob.nspec))]).all()
    if case[5]:
        assert (got.data[v][FIRSTSPEC:] == 0.0).any(), "negative-species branch not exercised"
    if case[4] > 1.0:
        assert mask.all() and (got.data[v][RHO] == 0.5 * bcd).any(), "density floor not exercised"
    assert (got.data[0] == -4.5e33).any(), "ghost zones must stay untouched"


def test_invalid_start_comp(oracle_calls):
    prob, snew, p0, bcd = build(*CASES[0])
    from maestroex_b200.calls import MxError
    with pytest.raises(MxError):
        oracle_calls.update_scal(prob.geom, prob.state, snew, prob.sflux, prob.force,
                                 params(prob, TEMP, bcd))


@pytest.mark.gpu
@pytest.mark.parametrize("residency", ["host", "device"])
@pytest.mark.parametrize("case", CASES)
def test_gpu_vs_oracle(oracle_calls, case, residency):
    import torch
    from maestroex_b200 import lib as mxlib
    prob, snew, p0, bcd = build(*case)
    want, wmask = run_both_passes(oracle_calls, prob, Fab(snew.data.copy(), snew.lo, snew.hi), p0, bcd, False)
    if residency == "host":
        put, get = (lambda f: f), (lambda f: f)
    else:
        put = lambda f: Fab(torch.from_numpy(f.data).cuda(), f.lo, f.hi)  # noqa: E731
        get = lambda f: Fab(f.data.cpu().numpy(), f.lo, f.hi)  # noqa: E731
    got, gmask = run_both_passes(mxlib.calls(), prob, Fab(snew.data.copy(), snew.lo, snew.hi), p0, bcd,
                                 False, put=put, get=get)
    assert np.array_equal(gmask, wmask)
    assert np.array_equal(got.data, want.data, equal_nan=True)


@pytest.mark.gpu
def test_gpu_eos_mask_is_required():
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.calls import MxError
    prob, snew, p0, bcd = build(*CASES[1])
    with pytest.raises(MxError) as ei:
        mxlib.calls().update_scal(prob.geom, prob.state, snew, prob.sflux, prob.force,
                                  params(prob, RHOH, bcd))
    assert "eos_mask" in str(ei.value)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [CASES[1], CASES[2], CASES[3]])
def test_overlay_update_scal(ref3d, ref2d, case):
    """maestroex_b200/overlay/MaestroUpdateScal.cpp driven like the compiled reference."""
    from oracle import backends
    ovl = backends.overlay(case[0])
    if ovl is None:
        pytest.skip("overlay library not built")
    prob, snew, p0, bcd = build(*case)
    ref = ref3d if case[0] == 3 else ref2d
    want, _ = run_both_passes(ref, prob, Fab(snew.data.copy(), snew.lo, snew.hi), p0, bcd, True)
    got, _ = run_both_passes(ovl, prob, Fab(snew.data.copy(), snew.lo, snew.hi), p0, bcd, True)
    assert np.array_equal(got.data, want.data, equal_nan=True)
