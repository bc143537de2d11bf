"""Parity at BASELINE.json's full sizes through a size-independent property of the path:
BOX INVARIANCE.  MakeEdgeScal / MakeRhoXFlux / MakeRhoHFlux on a box depend only on that box
plus its ghost cells (SURVEY.md section 8a "index-space extents"; this is what lets the
reference tile and distribute boxes freely), so the edge states and fluxes the GPU produces
for the whole 256^3 / 512^3 brick must equal, bit for bit, what the CPU oracle produces for
any small sub-box cut out of the same arrays (sub-box valid cells + ghost cells copied to the
host, same domain, same BCs).  Sub-boxes are placed at domain corners (physical BCs), in the
interior at offsets incommensurate with the kernel's 30 x 14 tiles and z chunks, and across
the |u| == 0 / |u| < rel_eps velocity slabs.  Also checked on the full arrays: sflux(Rho) is
the species sum in the reference's order, and nothing outside the swept components is touched.
"""
import numpy as np
import pytest

from common import run_edge, run_rhoh, run_rhoX
from maestroex_b200.abi import Base, Fab, Geom
from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH, Problem

pytestmark = pytest.mark.gpu


def cut(fab, lo, hi):
    """Host numpy copy of index range [lo, hi] (inclusive, i/j/k order) of a device fab."""
    sl = tuple(slice(lo[d] - fab.lo[d], hi[d] - fab.lo[d] + 1) for d in (2, 1, 0))
    return Fab(np.ascontiguousarray(fab.data[(slice(None),) + sl].cpu().numpy()), tuple(lo), tuple(hi))


def sub_problem(big, lo, hi):
    """The inputs of sub-box [lo, hi] of the device problem `big`, on the host."""
    g = big.geom
    geom = Geom(3, tuple(lo), tuple(hi), g.dom_lo, g.dom_hi, g.dx)
    ng = big.ng
    state = cut(big.state, *geom.cell_box(ng))
    force = cut(big.force, *geom.cell_box(1))
    umac = [cut(big.umac[d], *geom.face_box(d, 1)) for d in range(3)]
    nscal = big.nscal
    sedge, sflux = [], []
    for d in range(3):
        flo, fhi = geom.face_box(d, 0)
        shp = tuple(fhi[q] - flo[q] + 1 for q in (2, 1, 0))
        sedge.append(Fab(np.full((nscal,) + shp, -7.25e33), flo, fhi))
        sflux.append(Fab(np.full((nscal,) + shp, 3.5e33), flo, fhi))
    eta = cut(big.extras["eta_in"], *geom.face_box(2, 0))
    base = {k: Base(np.ascontiguousarray(v.data.cpu().numpy())) for k, v in big.base.items()}
    return Problem(geom=geom, nspec=big.nspec, ng=ng, state=state, force=force, umac=umac,
                   sedge=sedge, sflux=sflux, etarhoflux=eta, base=base, mac={}, dt=big.dt,
                   rel_eps=big.rel_eps, bc=big.bc)


def boxes_for(n):
    """Sub-boxes (lo, size): corners, tile/chunk-straddling interior, velocity slabs."""
    m = n // 4  # the |u| == 0 slab sits at index ~n/4 of a transverse direction, n/2 the tiny one
    return [
        ((0, 0, 0), (21, 17, 19)),                        # low corner: lo-z BC
        ((n - 23, n - 19, n - 18), (23, 19, 18)),         # high corner: hi-z BC
        ((101 % (n - 40), 77 % (n - 40), 59 % (n - 40)), (33, 29, 31)),  # interior, unaligned
        ((27, 11, 29), (8, 9, 40)),                       # thin column across z chunks 32 / 64
        ((m - 6, m - 5, m - 7), (14, 13, 15)),            # exactly-zero velocity slabs
        ((n // 2 - 7, n // 2 - 6, n // 2 - 5), (15, 14, 13)),  # |u| < rel_eps slabs
    ]


@pytest.mark.parametrize("n,nspec,ppm", [(256, 13, 1), (256, 3, 2), (256, 3, 0), (512, 3, 1),
                                         (512, 13, 1)])
def test_box_invariance_full_size(oracle_calls, n, nspec, ppm):
    import torch
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.devsynth import make_device_problem
    from maestroex_b200.fusedcalls import density_edge_flux, enthalpy_edge_flux

    # what the problem really allocates (bench.py runs the same 512^3 x 13-species brick in
    # ~150 GB): state + force + 3 sedge + 3 sflux of Nscal components, three MAC velocities,
    # etarhoflux twice (the test keeps its input), plus 5 GB for the generator's and the
    # whole-array checks' temporaries.  Memory cached by earlier tests is returned first.
    import gc
    ng_ = 4 if ppm == 2 else 3
    nscal = 4 + nspec
    doubles = (nscal * ((n + 2 * ng_) ** 3 + (n + 2) ** 3 + 6 * n * n * (n + 1))
               + 3 * (n + 2) ** 2 * (n + 3) + 2 * n * n * (n + 1))
    need = 8 * doubles + 5e9
    gc.collect()
    mxlib.load().mx_free_workspace()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < need:
        pytest.skip(f"needs {need / 1e9:.0f} GB of HBM, {free / 1e9:.0f} GB free")
    ng = 4 if ppm == 2 else 3
    big = make_device_problem(n=(n, n, n), nspec=nspec, ng=ng, bc_kind="reacting_bubble", seed=4242)
    big.etarhoflux.data.normal_(generator=torch.Generator(device="cuda").manual_seed(7))
    big.extras["eta_in"] = Fab(big.etarhoflux.data.clone(), big.etarhoflux.lo, big.etarhoflux.hi)
    for f in big.sedge:
        f.data.fill_(-7.25e33)
    for f in big.sflux:
        f.data.fill_(3.5e33)
    calls = mxlib.calls()
    outs = (big.sedge, big.sflux, big.etarhoflux)
    epd = big.edge_params(FIRSTSPEC, nspec, ppm_type=ppm)
    eph = big.edge_params(RHOH, 1, ppm_type=ppm)
    fpar = big.flux_params()
    mxlib.load().mx_reset_launch_count()
    density_edge_flux(calls, big, epd, fpar, outputs=outs)
    enthalpy_edge_flux(calls, big, eph, fpar, outputs=outs)
    torch.cuda.synchronize()
    assert mxlib.load().mx_launch_count() == 2  # one fused launch per pass

    swept = [RHO, RHOH] + list(range(FIRSTSPEC, FIRSTSPEC + nspec))
    for lo, size in boxes_for(n):
        hi = tuple(lo[d] + size[d] - 1 for d in range(3))
        sp = sub_problem(big, lo, hi)
        se = run_edge(oracle_calls, sp, sp.edge_params(FIRSTSPEC, nspec, ppm_type=ppm), sedge=sp.sedge)
        run_edge(oracle_calls, sp, sp.edge_params(RHO, 1, ppm_type=ppm), sedge=se)
        # run_rhoX clones the (poisoned) outputs of sp, accumulating into a copy of eta_in
        sf, eta = run_rhoX(oracle_calls, sp, se, sp.flux_params())
        run_edge(oracle_calls, sp, sp.edge_params(RHOH, 1, ppm_type=ppm), sedge=se)
        run_rhoh(oracle_calls, sp, se, sp.flux_params(), sflux=sf)
        for d in range(3):
            got_e = cut(big.sedge[d], se[d].lo, se[d].hi).data
            got_f = cut(big.sflux[d], sf[d].lo, sf[d].hi).data
            for c in swept:
                assert np.array_equal(got_e[c], se[d].data[c]), f"sedge dir {d} comp {c} box {lo}"
                assert np.array_equal(got_f[c], sf[d].data[c]), f"sflux dir {d} comp {c} box {lo}"
        got_eta = cut(big.etarhoflux, eta.lo, eta.hi).data
        assert np.array_equal(got_eta, eta.data), f"etarhoflux box {lo}"

    # whole-array properties
    for d in range(3):
        acc = torch.zeros_like(big.sflux[d].data[RHO])
        for c in range(FIRSTSPEC, FIRSTSPEC + nspec):  # reference order, :223
            acc = acc + big.sflux[d].data[c]
        assert torch.equal(acc, big.sflux[d].data[RHO]), f"sflux(Rho) != species sum, dir {d}"
        for c in (2, 3):  # Temp, Pi are not swept by either pass
            assert bool((big.sedge[d].data[c] == -7.25e33).all())
            assert bool((big.sflux[d].data[c] == 3.5e33).all())
        for c in swept:
            assert bool(torch.isfinite(big.sedge[d].data[c]).all())
    del big
    torch.cuda.empty_cache()
