"""Python call wrappers for the fused entry points mx_density_edge_flux /
mx_enthalpy_edge_flux (include/mxadv.h), used by the tests, smoke() and bench.py."""
from __future__ import annotations

import ctypes as C

from .abi import fab_array
from .synth import RHO, RHOH


def _b(x):
    return C.byref(x.c()) if x is not None else None


def density_edge_flux(calls, prob, epar, fpar, outputs=None, stream: int = 0):
    """MakeEdgeScal(species) + MakeEdgeScal(Rho) + MakeRhoXFlux in one library call
    (Source/MaestroDensityAdvance.cpp:115,128,182)."""
    sedge, sflux, eta = outputs if outputs is not None else prob.clone_outputs()
    g, st, fo, ep, fp, et = (prob.geom.c(), prob.state.c(), prob.force.c(), epar.c(), fpar.c(),
                             eta.c())
    b = prob.base
    fn = calls.lib.mx_density_edge_flux
    fn.restype = C.c_int
    rc = fn(C.byref(g), C.byref(st), fab_array(sedge), fab_array(sflux), C.byref(et),
            fab_array(prob.umac), C.byref(fo), C.byref(ep), C.c_int(RHO), _b(b["rho0_old"]),
            _b(b["rho0_edge_old"]),
            fab_array(prob.mac["r0mac_old"]) if "r0mac_old" in prob.mac else None,
            _b(b["rho0_new"]), _b(b["rho0_edge_new"]),
            fab_array(prob.mac["r0mac_new"]) if "r0mac_new" in prob.mac else None,
            _b(b["rho0_predicted_edge"]), _b(b["w0"]), C.byref(fp), C.c_void_p(stream))
    calls._check(rc, "density_edge_flux")
    return sedge, sflux, eta


def enthalpy_edge_flux(calls, prob, epar, fpar, outputs=None, stream: int = 0):
    """MakeEdgeScal(RhoH) + MakeRhoHFlux in one library call
    (Source/MaestroEnthalpyAdvance.cpp:141-146,229)."""
    sedge, sflux, _ = outputs if outputs is not None else prob.clone_outputs()
    g, st, fo, ep, fp = prob.geom.c(), prob.state.c(), prob.force.c(), epar.c(), fpar.c()
    b = prob.base
    fn = calls.lib.mx_enthalpy_edge_flux
    fn.restype = C.c_int
    rc = fn(C.byref(g), C.byref(st), fab_array(sedge), fab_array(sflux), fab_array(prob.umac),
            C.byref(fo), C.byref(ep), _b(b["rho0_old"]), _b(b["rho0_edge_old"]),
            _b(b["rho0_new"]), _b(b["rho0_edge_new"]), _b(b["rhoh0_old"]),
            _b(b["rhoh0_edge_old"]), _b(b["rhoh0_new"]), _b(b["rhoh0_edge_new"]), C.byref(fp),
            C.c_void_p(stream))
    calls._check(rc, "enthalpy_edge_flux")
    return sedge, sflux


def enthalpy_edge_flux_sph(calls, prob, epar, fpar, outputs=None, stream: int = 0):
    """The spherical form (wdconvect): base state on faces (r0mac / h0mac / rh0mac)."""
    sedge, sflux, _ = outputs if outputs is not None else prob.clone_outputs()
    g, st, fo, ep, fp = prob.geom.c(), prob.state.c(), prob.force.c(), epar.c(), fpar.c()
    m = prob.mac
    fa = lambda k: fab_array(m[k]) if k in m else None  # noqa: E731
    fn = calls.lib.mx_enthalpy_edge_flux_sph
    fn.restype = C.c_int
    rc = fn(C.byref(g), C.byref(st), fab_array(sedge), fab_array(sflux), fab_array(prob.umac),
            C.byref(fo), C.byref(ep), fa("r0mac_old"), fa("r0mac_new"), fa("rh0mac_old"),
            fa("rh0mac_new"), fa("h0mac_old"), fa("h0mac_new"), C.byref(fp), C.c_void_p(stream))
    calls._check(rc, "enthalpy_edge_flux_sph")
    return sedge, sflux


def density_advance(calls, prob, snew, epar, fpar, upar, species_force_is_zero=False,
                    outputs=None, with_edges=False, sphr=None, stream: int = 0):
    """mx_density_advance: mx_density_edge_flux + PutInPertForm / ConvertRhoXToX back +
    UpdateScal(species, density), intermediates device resident
    (Source/MaestroDensityAdvance.cpp:115-258).  Returns (snew, etarhoflux[, sedge, sflux])."""
    sedge, sflux, eta = outputs if outputs is not None else prob.clone_outputs()
    g, st, sn, fo, ep, fp, et, up = (prob.geom.c(), prob.state.c(), snew.c(), prob.force.c(),
                                     epar.c(), fpar.c(), eta.c(), upar.c())
    b = prob.base
    fn = calls.lib.mx_density_advance
    fn.restype = C.c_int
    rc = fn(C.byref(g), C.byref(st), C.byref(sn), C.byref(et), fab_array(prob.umac), C.byref(fo),
            C.c_int(1 if species_force_is_zero else 0), C.byref(ep), C.c_int(RHO),
            _b(b["rho0_old"]), _b(b["rho0_edge_old"]),
            fab_array(prob.mac["r0mac_old"]) if "r0mac_old" in prob.mac else None,
            _b(b["rho0_new"]), _b(b["rho0_edge_new"]),
            fab_array(prob.mac["r0mac_new"]) if "r0mac_new" in prob.mac else None,
            _b(b["rho0_predicted_edge"]), _b(b["w0"]), C.byref(fp), _b(b["rho0_old"]),
            C.byref(sphr.c()) if sphr is not None else None, C.byref(up),
            fab_array(sedge) if with_edges else None, fab_array(sflux) if with_edges else None,
            C.c_void_p(stream))
    calls._check(rc, "density_advance")
    return (snew, eta, sedge, sflux) if with_edges else (snew, eta)


def enthalpy_advance(calls, prob, snew, force_update, epar, fpar, upar, eos_mask=None,
                     stream: int = 0):
    """mx_enthalpy_advance: mx_enthalpy_edge_flux + PutInPertForm back + UpdateScal(RhoH)
    (Source/MaestroEnthalpyAdvance.cpp:141-340); returns the number of zones left to the
    caller's EOS."""
    from .abi import _ptr_of
    g, st, sn, fo, fu, ep, fp, up = (prob.geom.c(), prob.state.c(), snew.c(), prob.force.c(),
                                     force_update.c(), epar.c(), fpar.c(), upar.c())
    b = prob.base
    n = C.c_longlong(0)
    fn = calls.lib.mx_enthalpy_advance
    fn.restype = C.c_int
    rc = fn(C.byref(g), C.byref(st), C.byref(sn), fab_array(prob.umac), C.byref(fo), C.byref(fu),
            C.byref(ep), _b(b["rho0_old"]), _b(b["rho0_edge_old"]), _b(b["rho0_new"]),
            _b(b["rho0_edge_new"]), _b(b["rhoh0_old"]), _b(b["rhoh0_edge_old"]),
            _b(b["rhoh0_new"]), _b(b["rhoh0_edge_new"]), C.byref(fp), _b(b["rhoh0_old"]),
            C.byref(up), C.c_void_p(_ptr_of(eos_mask)) if eos_mask is not None else None,
            C.byref(n), C.c_void_p(stream))
    calls._check(rc, "enthalpy_advance")
    return n.value
