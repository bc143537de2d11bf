"""Seeded synthetic single-level inputs for the scalar-advection path.

`make_problem` builds, with numpy on the host, everything MakeEdgeScal / MakeRhoXFlux /
MakeRhoHFlux read for one box: state (X_k, rho', (rho h)'), MAC velocities (already
containing w0, as at Source/MaestroAdvance.cpp:448), forcing, the 1-D base-state arrays and
their edge-centred images, plus zero-initialised outputs.  Fields mix a smooth part with
noise so limiters, extrema logic and both upwind directions are exercised; the velocity
has an exactly-zero slab and a |u| < rel_eps slab (averaging branches,
Source/MaestroMakeEdgeScalars.cpp:745-748,968-971).  Ghost cells are filled by the generator
(the halo fill is not part of the path).

State component indices follow Source/param/_variables: Rho=0, RhoH=1, Temp=2, Pi=3,
FirstSpec=4, Nscal = 4 + nspec.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import numpy as np

from .abi import (BC_FOEXTRAP, BC_HOEXTRAP, BC_INT_DIR, Base, EdgeParams, Fab, FluxParams,
                  Geom, box_shape)

RHO, RHOH, TEMP, PI, FIRSTSPEC = 0, 1, 2, 3, 4


@dataclass
class Problem:
    geom: Geom
    nspec: int
    ng: int
    state: Fab
    force: Fab
    umac: List[Fab]
    sedge: List[Fab]
    sflux: List[Fab]
    etarhoflux: Fab
    base: Dict[str, Base]
    mac: Dict[str, List[Fab]]  # spherical face arrays: r0mac_old/new, rh0mac_*, h0mac_*
    dt: float
    rel_eps: float
    bc: np.ndarray  # (Nscal, 6)
    extras: dict = field(default_factory=dict)

    @property
    def nscal(self) -> int:
        return 4 + self.nspec

    def edge_params(self, start_scomp, num_comp, **kw) -> EdgeParams:
        return EdgeParams(dt=self.dt, rel_eps=self.rel_eps, start_scomp=start_scomp,
                          start_bccomp=start_scomp, num_comp=num_comp, bc=self.bc, **kw)

    def flux_params(self, **kw) -> FluxParams:
        d = dict(rho_comp=RHO, rhoh_comp=RHOH, first_spec=FIRSTSPEC, num_spec=self.nspec,
                 start_comp=FIRSTSPEC, num_comp=self.nspec)
        d.update(kw)
        return FluxParams(**d)

    def clone_outputs(self):
        """Fresh copies of the output fabs (so several back ends can run on one problem)."""
        cp = lambda f: Fab(f.data.copy(), f.lo, f.hi)  # noqa: E731
        return [cp(f) for f in self.sedge], [cp(f) for f in self.sflux], cp(self.etarhoflux)


def bc_set(kind: str, nscal: int, dim: int) -> np.ndarray:
    """BCRec table for all Nscal components.

    'periodic'         : int_dir everywhere
    'reacting_bubble'  : periodic laterally, lo-vertical slip wall (hoextrap for scalars),
                         hi-vertical outflow (foextrap) -- Source/MaestroSetup.cpp:366-421
    'outflow'          : foextrap on all faces (wdconvect)
    'mixed'            : a different type on every face (stress test, incl. reflect/ext_dir)
    """
    bc = np.zeros((nscal, 6), dtype=np.int32)
    vert = dim - 1
    if kind == "periodic":
        bc[:] = BC_INT_DIR
    elif kind == "reacting_bubble":
        bc[:, vert] = BC_HOEXTRAP
        bc[:, 3 + vert] = BC_FOEXTRAP
    elif kind == "outflow":
        for d in range(dim):
            bc[:, d] = BC_FOEXTRAP
            bc[:, 3 + d] = BC_FOEXTRAP
    elif kind == "mixed":
        codes = [4, 2, 1, 3, -1, 2, 4, 3, 1, 0, 2, 4]
        for n in range(nscal):
            for d in range(dim):
                bc[n, d] = codes[(n + 2 * d) % len(codes)]
                bc[n, 3 + d] = codes[(n + 2 * d + 5) % len(codes)]
    else:
        raise ValueError(kind)
    return bc


def make_problem(dim=3, n=(16, 16, 16), nspec=3, ng=3, bc_kind="periodic", seed=20240807,
                 valid_lo=None, dom_n=None, spherical=False, force_ng=1) -> Problem:
    """Build one box of size `n` (valid cells) inside a domain `dom_n` (defaults to the box).

    `valid_lo` places the box inside the domain so that interior boxes (no face on the
    domain boundary) and boundary boxes can both be generated.
    """
    rng = np.random.default_rng(seed)
    n = tuple(n)[:dim]
    dom_n = tuple(dom_n)[:dim] if dom_n is not None else n
    valid_lo = tuple(valid_lo)[:dim] if valid_lo is not None else (0,) * dim
    valid_hi = tuple(valid_lo[d] + n[d] - 1 for d in range(dim))
    dx = tuple(1.0 / dom_n[d] for d in range(dim))
    geom = Geom(dim, valid_lo, valid_hi, (0,) * dim, tuple(m - 1 for m in dom_n), dx)
    nscal = 4 + nspec
    vert = dim - 1

    def coords(lo, hi, stagger=-1):
        """cell-centre (or face, in direction `stagger`) coordinates as broadcastable arrays"""
        out = []
        for d in range(3):
            idx = np.arange(lo[d], hi[d] + 1, dtype=np.float64)
            if d < dim:
                x = (idx + (0.0 if d == stagger else 0.5)) * dx[d]
            else:
                x = idx * 0.0
            shape = [1, 1, 1]
            shape[2 - d] = x.size
            out.append(x.reshape(shape))
        return out  # x, y, z with shapes (1,1,nx), (1,ny,1), (nz,1,1)

    # --- 1-D base state: exponential stratification over ~3 decades along the vertical
    nr = dom_n[vert]
    r_cc = (np.arange(nr) + 0.5) / nr
    r_ed = np.arange(nr + 1) / nr
    rho0_old = 1.0e6 * np.exp(-6.9 * r_cc)
    rho0_new = rho0_old * (1.0 + 1.0e-3 * np.sin(2 * np.pi * r_cc))
    rhoh0_old = rho0_old * 1.0e8 * (1.0 + 0.3 * r_cc)
    rhoh0_new = rhoh0_old * (1.0 - 2.0e-3 * np.cos(2 * np.pi * r_cc))
    w0 = 0.05 * np.sin(np.pi * r_ed) ** 2

    def cell_to_edge_np(c):  # plain 2-point average is enough for inputs
        e = np.empty(nr + 1)
        e[1:-1] = 0.5 * (c[1:] + c[:-1])
        e[0], e[-1] = c[0], c[-1]
        return e

    base = {
        "rho0_old": Base(np.ascontiguousarray(rho0_old)),
        "rho0_new": Base(np.ascontiguousarray(rho0_new)),
        "rho0_edge_old": Base(cell_to_edge_np(rho0_old)),
        "rho0_edge_new": Base(cell_to_edge_np(rho0_new)),
        "rhoh0_old": Base(np.ascontiguousarray(rhoh0_old)),
        "rhoh0_new": Base(np.ascontiguousarray(rhoh0_new)),
        "rhoh0_edge_old": Base(cell_to_edge_np(rhoh0_old)),
        "rhoh0_edge_new": Base(cell_to_edge_np(rhoh0_new)),
        "rho0_predicted_edge": Base(cell_to_edge_np(0.5 * (rho0_old + rho0_new))),
        "w0": Base(np.ascontiguousarray(w0)),
    }

    # --- state
    slo, shi = geom.cell_box(ng)
    shp = box_shape(slo, shi)
    x, y, z = coords(slo, shi)
    state = np.zeros((nscal,) + shp)
    vcoord = (z if dim == 3 else y)
    rho0_3d = 1.0e6 * np.exp(-6.9 * np.clip(vcoord, 0.0, 1.0))
    # rho' = rho - rho0: smooth + noise, passes through zero
    state[RHO] = rho0_3d * (0.05 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
                            + 0.02 * (rng.random(shp) - 0.5))
    state[RHOH] = rho0_3d * 1.0e8 * (0.04 * np.cos(2 * np.pi * x) * np.sin(4 * np.pi * y)
                                      + 0.02 * (rng.random(shp) - 0.5))
    state[TEMP] = 1.0e8 * (1.0 + 0.1 * rng.random(shp))
    state[PI] = rng.standard_normal(shp)
    # mass fractions: positive, normalised, last species is a 1e-30 trace (X(Mg24) floor)
    X = np.empty((nspec,) + shp)
    for s_ in range(nspec):
        X[s_] = 1.0 + 0.5 * np.sin(2 * np.pi * (s_ + 1) * x + 0.3 * s_) * np.cos(
            2 * np.pi * y + 0.7 * s_) + 0.4 * rng.random(shp)
        # add sharp features (top-hat) so limiters switch
        X[s_] += (np.abs(x - 0.5) < 0.13 + 0.02 * s_) * 0.8
    if nspec > 1:
        X[-1] = 1.0e-30
    X[: max(nspec - 1, 1)] /= X[: max(nspec - 1, 1)].sum(axis=0, keepdims=True)
    state[FIRSTSPEC:] = X

    # --- MAC velocity (face-centred, 1 ghost), includes w0 on the vertical faces
    umac = []
    for d in range(dim):
        lo, hi = geom.face_box(d, 1)
        fs = box_shape(lo, hi)
        xf, yf, zf = coords(lo, hi, stagger=d)
        ph = 0.37 * d
        u = (np.sin(2 * np.pi * (xf + ph)) * np.cos(2 * np.pi * (yf - ph))
             * (np.cos(2 * np.pi * zf + ph) if dim == 3 else 1.0))
        u = u + 0.1 * (rng.random(fs) - 0.5) + np.zeros(fs)
        if d == vert:
            r_face = np.clip(zf if dim == 3 else yf, 0.0, 1.0)
            u = u + 0.05 * np.sin(np.pi * r_face) ** 2
        umac.append(np.ascontiguousarray(u))
    umax = max(float(np.abs(u).max()) for u in umac)
    rel_eps = 1.0e-8 * umax
    dt = 0.7 * min(dx) / umax
    for d in range(dim):
        u = umac[d]
        # an exactly-zero slab and a |u| < rel_eps slab (both signs)
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        ax = 2 - ((d + 1) % dim)  # a direction transverse to d
        m = u.shape[ax]
        sl0[ax] = slice(m // 4, m // 4 + 2)
        sl1[ax] = slice(m // 2, m // 2 + 2)
        u[tuple(sl0)] = 0.0
        tiny = (rng.random(u[tuple(sl1)].shape) - 0.5) * rel_eps
        u[tuple(sl1)] = tiny
    umac_f = [Fab(np.ascontiguousarray(umac[d][None]), *geom.face_box(d, 1)) for d in range(dim)]

    # --- forcing: non-zero for rho', (rho h)'; zero for species (MaestroDensityAdvance.cpp:36-39)
    flo, fhi = geom.cell_box(force_ng)
    fshp = box_shape(flo, fhi)
    force = np.zeros((nscal,) + fshp)
    xf, yf, zf = coords(flo, fhi)
    force[RHO] = 1.0e5 * (np.sin(2 * np.pi * xf) + 0.3 * rng.standard_normal(fshp))
    force[RHOH] = 1.0e13 * (np.cos(2 * np.pi * yf) + 0.3 * rng.standard_normal(fshp))
    force[TEMP] = rng.standard_normal(fshp)
    force[PI] = rng.standard_normal(fshp)

    # --- outputs, poisoned so that untouched entries are detectable
    sedge, sflux = [], []
    for d in range(dim):
        lo, hi = geom.face_box(d, 0)
        sedge.append(Fab(np.full((nscal,) + box_shape(lo, hi), -7.25e33), lo, hi))
        sflux.append(Fab(np.full((nscal,) + box_shape(lo, hi), 3.5e33), lo, hi))
    elo, ehi = geom.face_box(vert, 0)
    etarho = Fab(rng.standard_normal((1,) + box_shape(elo, ehi)), elo, ehi)

    # --- spherical face arrays (r0mac etc. are MultiFabs in the reference): 1 ghost
    mac: Dict[str, List[Fab]] = {}
    if spherical:
        for name, scale in (("r0mac_old", 1.0e6), ("r0mac_new", 1.0e6), ("rh0mac_old", 1.0e14),
                            ("rh0mac_new", 1.0e14), ("h0mac_old", 1.0e8), ("h0mac_new", 1.0e8)):
            lst = []
            for d in range(dim):
                lo, hi = geom.face_box(d, 1)
                fs = box_shape(lo, hi)
                xf, yf, zf = coords(lo, hi, stagger=d)
                rad = np.sqrt((xf - 0.5) ** 2 + (yf - 0.5) ** 2 + (zf - 0.5) ** 2)
                lst.append(Fab(np.ascontiguousarray(
                    (scale * np.exp(-3.0 * rad) * (1.0 + 1e-3 * rng.random(fs)))[None]), lo, hi))
            mac[name] = lst

    return Problem(geom=geom, nspec=nspec, ng=ng, state=Fab(state, slo, shi),
                   force=Fab(force, flo, fhi), umac=umac_f, sedge=sedge, sflux=sflux,
                   etarhoflux=etarho, base=base, mac=mac, dt=dt, rel_eps=rel_eps,
                   bc=bc_set(bc_kind, nscal, dim))


# ---------------------------------------------------------------------------------------
# Makew0 inputs (Source/MaestroMakew0.cpp:5-14): a stratified base state with noise
# ---------------------------------------------------------------------------------------
W0_CASES = {
    # name -> (spherical, nr_fine, base_cutoff_density_coord, is_predictor, add_pb,
    #          dpdt_factor, base_cutoff_density, seed)
    "planar_pred": (0, 64, 50, 1, 0, 0.0, -1.0, 1),
    "planar_corr_pb": (0, 64, 64, 0, 1, 0.3, -1.0, 2),
    "planar_odd_pb": (0, 257, 200, 1, 1, 0.5, -1.0, 3),
    "planar_wd": (0, 2218, 1800, 0, 0, 0.2, -1.0, 4),
    "planar_tiny": (0, 5, 3, 1, 0, 0.1, -1.0, 5),
    "sphr_small": (1, 64, 50, 1, 0, 0.0, 1.0e4, 6),
    "sphr_cut_above": (1, 64, 80, 1, 0, 0.3, 1.0e5, 7),
    "sphr_odd": (1, 257, 200, 1, 0, 0.5, 3.0e4, 8),
    "sphr_wd": (1, 2218, 1800, 1, 0, 0.2, 1.0e4, 9),
    # spherical == 2: use_exact_base_state (Makew0SphrIrreg), irregular radial bins
    "irreg_small": (2, 64, 50, 1, 0, 0.0, 1.0e4, 10),
    "irreg_odd": (2, 257, 200, 1, 0, 0.5, 3.0e4, 11),
    "irreg_wd": (2, 2218, 1800, 1, 0, 0.2, 1.0e4, 12),
}


def make_w0_problem(name: str):
    """(W0Params, dict of input arrays) of one Makew0 case."""
    from .abi import W0Params
    sph, nr, cut, pred, pb, dpdt, bcd, seed = W0_CASES[name]
    rng = np.random.default_rng(1000 + seed)
    dr = (5.0e8 if sph else 1.0e8) / nr
    z = (np.arange(nr) + 0.5) / nr
    rho = 1.0e7 * np.exp(-8.0 * z) * (1.0 + 0.01 * rng.random(nr))
    p0 = 1.0e24 * np.exp(-9.0 * z) * (1.0 + 0.01 * rng.random(nr))
    g1 = 1.4 + 0.2 * rng.random(nr)
    d = {
        "w0_old": 1.0e4 * np.cumsum(rng.standard_normal(nr + 1)) / nr,
        "Sbar": 1.0e-3 * rng.standard_normal(nr),
        "rho0_old": rho,
        "rho0_new": rho * (1.0 + 1.0e-3 * rng.standard_normal(nr)),
        "p0_old": p0,
        "p0_new": p0 * (1.0 + 1.0e-3 * rng.standard_normal(nr)),
        "gamma1bar_old": g1,
        "gamma1bar_new": g1 + 1.0e-3 * rng.standard_normal(nr),
        "p0_minus_peosbar": 1.0e18 * rng.standard_normal(nr),
        "etarho_cc": 1.0e2 * rng.standard_normal(nr),
        "etarho_ec": 1.0e2 * rng.standard_normal(nr + 1),
        "r_cc_loc": (np.arange(nr) + 0.5) * dr,
        "r_edge_loc": np.arange(nr + 1) * dr,
    }
    if sph == 2:
        # the bins of the exact base state (Source/MaestroBaseStateGeometry... init: one bin per
        # distinct cell-centre radius): r_cc = sqrt(0.75 + 2 r) dx, edges half way in r^2
        dxf = 5.0e8 / np.sqrt(0.75 + 2.0 * nr)
        d["r_cc_loc"] = np.sqrt(0.75 + 2.0 * np.arange(nr)) * dxf
        e = np.sqrt(0.75 + 2.0 * (np.arange(nr + 1) - 0.5).clip(min=0.0)) * dxf
        e[0] = 0.0
        d["r_edge_loc"] = e
    d = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in d.items()}
    par = W0Params(nr_fine=nr, base_cutoff_density_coord=cut, dr=dr, dt=0.37, dt_in=0.4,
                   dtold_in=0.33, is_predictor=pred, add_pb=pb, dpdt_factor=dpdt,
                   base_cutoff_density=bcd)
    return int(sph), par, d


def run_w0(calls, sph: bool, par, d, xp=np, put=lambda a: a):
    """Makew0 through `calls`; returns (w0, w0_force, p0bdot).  `put` moves an input array to
    where the back end wants it (identity for host arrays)."""
    nr = par.nr_fine
    w0 = put(np.full(nr + 1, 7.0))
    force = put(np.full(nr, 7.0))
    a = {k: put(v) for k, v in d.items()}
    if sph:
        calls.make_w0_sphr(par, a["w0_old"], a["Sbar"], a["rho0_old"], a["rho0_new"],
                           a["p0_old"], a["p0_new"], a["gamma1bar_old"], a["gamma1bar_new"],
                           a["p0_minus_peosbar"], a["etarho_cc"], a["etarho_ec"],
                           a["r_cc_loc"], a["r_edge_loc"], w0, force, irreg=(sph == 2))
        pb = 0.0
    else:
        pb = calls.make_w0_planar(par, a["w0_old"], a["Sbar"], a["p0_old"], a["p0_new"],
                                  a["gamma1bar_old"], a["gamma1bar_new"],
                                  a["p0_minus_peosbar"], a["etarho_cc"], w0, force)
    return w0, force, pb
