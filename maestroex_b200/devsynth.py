"""Device-resident synthetic inputs for bench.py, generated with torch
directly in HBM so that 512^3 x 13 species (~150 GB of fabs) never touches the host.

Same field recipe as synth.make_problem (smooth part + noise, exponential base-state
stratification, a 1e-30 trace species, an exactly-zero and a |u| < rel_eps velocity slab,
non-zero forcing only for rho' and (rho h)'), but seeded with torch's device generator; the
parity tests use synth.py (numpy, bit-reproducible on the host), bench.py uses this.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np

from .abi import Base, Fab, Geom, box_shape
from .synth import FIRSTSPEC, PI, RHO, RHOH, TEMP, Problem, bc_set


def make_device_problem(n=(64, 64, 64), nspec=3, ng=3, bc_kind="reacting_bubble", seed=20240807,
                        device="cuda:0", dom_n=None, valid_lo=None, spherical=False,
                        alloc_outputs=True) -> Problem:
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    f64 = torch.float64
    dim = 3
    n = tuple(n)
    dom_n = tuple(dom_n) if dom_n is not None else n
    valid_lo = tuple(valid_lo) if valid_lo is not None else (0, 0, 0)
    valid_hi = tuple(valid_lo[d] + n[d] - 1 for d in range(3))
    dx = tuple(1.0 / dom_n[d] for d in range(3))
    geom = Geom(dim, valid_lo, valid_hi, (0, 0, 0), tuple(m - 1 for m in dom_n), dx)
    nscal = 4 + nspec
    two_pi = 2.0 * math.pi

    def coords(lo, hi, stagger=-1):
        out = []
        for d in range(3):
            idx = torch.arange(lo[d], hi[d] + 1, dtype=f64, device=dev)
            x = (idx + (0.0 if d == stagger else 0.5)) * dx[d]
            shape = [1, 1, 1]
            shape[2 - d] = x.numel()
            out.append(x.reshape(shape))
        return out

    def noise(shape):
        return torch.rand(shape, dtype=f64, device=dev, generator=gen)

    # ---- 1-D base state (host numpy, tiny), as in synth.py
    nr = dom_n[2]
    r_cc = (np.arange(nr) + 0.5) / nr
    r_ed = np.arange(nr + 1) / nr
    rho0_old = 1.0e6 * np.exp(-6.9 * r_cc)
    rho0_new = rho0_old * (1.0 + 1.0e-3 * np.sin(2 * np.pi * r_cc))
    rhoh0_old = rho0_old * 1.0e8 * (1.0 + 0.3 * r_cc)
    rhoh0_new = rhoh0_old * (1.0 - 2.0e-3 * np.cos(2 * np.pi * r_cc))
    w0 = 0.05 * np.sin(np.pi * r_ed) ** 2

    def c2e(c):
        e = np.empty(nr + 1)
        e[1:-1] = 0.5 * (c[1:] + c[:-1])
        e[0], e[-1] = c[0], c[-1]
        return e

    host_base = {
        "rho0_old": rho0_old, "rho0_new": rho0_new,
        "rho0_edge_old": c2e(rho0_old), "rho0_edge_new": c2e(rho0_new),
        "rhoh0_old": rhoh0_old, "rhoh0_new": rhoh0_new,
        "rhoh0_edge_old": c2e(rhoh0_old), "rhoh0_edge_new": c2e(rhoh0_new),
        "rho0_predicted_edge": c2e(0.5 * (rho0_old + rho0_new)), "w0": w0,
    }
    base = {k: Base(torch.from_numpy(np.ascontiguousarray(v)).to(dev)) for k, v in host_base.items()}

    # ---- state
    slo, shi = geom.cell_box(ng)
    shp = box_shape(slo, shi)
    x, y, z = coords(slo, shi)
    state = torch.empty((nscal,) + shp, dtype=f64, device=dev)
    rho0_3d = 1.0e6 * torch.exp(-6.9 * z.clamp(0.0, 1.0))
    smooth_xy = torch.sin(two_pi * x) * torch.cos(two_pi * y)
    state[RHO] = rho0_3d * (0.05 * smooth_xy + 0.02 * (noise(shp) - 0.5))
    state[RHOH] = rho0_3d * 1.0e8 * (0.04 * torch.cos(two_pi * x) * torch.sin(2 * two_pi * y)
                                      + 0.02 * (noise(shp) - 0.5))
    state[TEMP] = 1.0e8 * (1.0 + 0.1 * noise(shp))
    state[PI] = noise(shp) - 0.5
    nlive = max(nspec - 1, 1)
    tot = torch.zeros(shp, dtype=f64, device=dev)
    for s_ in range(nlive):
        X = (1.0 + 0.5 * torch.sin(two_pi * (s_ + 1) * x + 0.3 * s_) * torch.cos(two_pi * y + 0.7 * s_)
             + 0.4 * noise(shp) + 0.2 * torch.cos(two_pi * z * (1 + s_ % 3)))
        X += ((x - 0.5).abs() < 0.13 + 0.02 * (s_ % 5)).to(f64) * 0.8
        state[FIRSTSPEC + s_] = X
        tot += X
        del X
    for s_ in range(nlive):
        state[FIRSTSPEC + s_] /= tot
    del tot
    if nspec > 1:
        state[FIRSTSPEC + nspec - 1] = 1.0e-30

    # ---- MAC velocities (1 ghost), w0 included on the vertical faces
    umac: List[Fab] = []
    umax = 0.0
    for d in range(3):
        lo, hi = geom.face_box(d, 1)
        fs = box_shape(lo, hi)
        xf, yf, zf = coords(lo, hi, stagger=d)
        ph = 0.37 * d
        u = (torch.sin(two_pi * (xf + ph)) * torch.cos(two_pi * (yf - ph))
             * torch.cos(two_pi * zf + ph)) + 0.1 * (noise(fs) - 0.5)
        if d == 2:
            u = u + 0.05 * torch.sin(math.pi * zf.clamp(0.0, 1.0)) ** 2
        umax = max(umax, float(u.abs().max()))
        umac.append(Fab(u.reshape((1,) + fs).contiguous(), lo, hi))
    rel_eps = 1.0e-8 * umax
    dt = 0.7 * min(dx) / umax
    for d in range(3):
        u = umac[d].data[0]
        ax = 2 - ((d + 1) % 3)
        m = u.shape[ax]
        sl0 = [slice(None)] * 3
        sl1 = [slice(None)] * 3
        sl0[ax] = slice(m // 4, m // 4 + 2)
        sl1[ax] = slice(m // 2, m // 2 + 2)
        u[tuple(sl0)] = 0.0
        u[tuple(sl1)] = (noise(u[tuple(sl1)].shape) - 0.5) * rel_eps

    # ---- forcing (1 ghost): non-zero for rho' and (rho h)' only
    flo, fhi = geom.cell_box(1)
    fshp = box_shape(flo, fhi)
    xf, yf, zf = coords(flo, fhi)
    force = torch.zeros((nscal,) + fshp, dtype=f64, device=dev)
    force[RHO] = 1.0e5 * (torch.sin(two_pi * xf) + 0.3 * (noise(fshp) - 0.5))
    force[RHOH] = 1.0e13 * (torch.cos(two_pi * yf) + 0.3 * (noise(fshp) - 0.5))

    # ---- outputs
    sedge, sflux = [], []
    for d in range(3):
        lo, hi = geom.face_box(d, 0)
        if alloc_outputs:
            sedge.append(Fab(torch.zeros((nscal,) + box_shape(lo, hi), dtype=f64, device=dev), lo, hi))
            sflux.append(Fab(torch.zeros((nscal,) + box_shape(lo, hi), dtype=f64, device=dev), lo, hi))
    elo, ehi = geom.face_box(2, 0)
    etarho = Fab(torch.zeros((1,) + box_shape(elo, ehi), dtype=f64, device=dev), elo, ehi)

    mac: Dict[str, List[Fab]] = {}
    if spherical:
        # what the spherical flux forms read (MaestroMakeFlux.cpp:311-407, :871-1021)
        for name, scale in (("r0mac_old", 1.0e6), ("r0mac_new", 1.0e6), ("h0mac_old", 1.0e8),
                            ("h0mac_new", 1.0e8)):
            lst = []
            for d in range(3):
                lo, hi = geom.face_box(d, 1)
                fs = box_shape(lo, hi)
                xf, yf, zf = coords(lo, hi, stagger=d)
                rad = torch.sqrt((xf - 0.5) ** 2 + (yf - 0.5) ** 2 + (zf - 0.5) ** 2)
                lst.append(Fab((scale * torch.exp(-3.0 * rad)).reshape((1,) + fs).contiguous(), lo, hi))
            mac[name] = lst

    prob = Problem(geom=geom, nspec=nspec, ng=ng, state=Fab(state, slo, shi),
                   force=Fab(force, flo, fhi), umac=umac, sedge=sedge, sflux=sflux,
                   etarhoflux=etarho, base=base, mac=mac, dt=dt, rel_eps=rel_eps,
                   bc=bc_set(bc_kind, nscal, 3))
    torch.cuda.synchronize(dev)
    return prob


def to_pinned_host(prob: Problem) -> Problem:
    """Copy of a device problem in pinned host memory (the e2e leg of bench.py)."""
    import copy
    import torch

    def fab(f):
        return Fab(f.data.cpu().pin_memory(), f.lo, f.hi)

    p = copy.copy(prob)
    p.state = fab(prob.state)
    p.force = fab(prob.force)
    p.umac = [fab(f) for f in prob.umac]
    p.sedge = [fab(f) for f in prob.sedge]
    p.sflux = [fab(f) for f in prob.sflux]
    p.etarhoflux = fab(prob.etarhoflux)
    p.base = {k: Base(v.data.cpu().pin_memory()) for k, v in prob.base.items()}
    p.mac = {k: [fab(f) for f in v] for k, v in prob.mac.items()}
    torch.cuda.synchronize()
    return p


def to_numpy_host(prob: Problem) -> Problem:
    """numpy copy (for the CPU oracle / reference library)."""
    import copy

    def fab(f):
        return Fab(np.ascontiguousarray(f.data.cpu().numpy()), f.lo, f.hi)

    p = copy.copy(prob)
    p.state = fab(prob.state)
    p.force = fab(prob.force)
    p.umac = [fab(f) for f in prob.umac]
    p.sedge = [fab(f) for f in prob.sedge]
    p.sflux = [fab(f) for f in prob.sflux]
    p.etarhoflux = fab(prob.etarhoflux)
    p.base = {k: Base(np.ascontiguousarray(v.data.cpu().numpy())) for k, v in prob.base.items()}
    p.mac = {k: [fab(f) for f in v] for k, v in prob.mac.items()}
    return p
