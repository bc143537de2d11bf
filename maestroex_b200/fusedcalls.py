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
