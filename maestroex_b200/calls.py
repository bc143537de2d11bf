"""Uniform Python call layer over the C ABI declared in include/mxadv.h.

``Calls(lib, prefix, with_stream)`` wraps a loaded shared library whose entry points are
``<prefix>make_edge_scal`` etc.  The product library uses prefix ``mx_`` and takes a trailing
``void* stream``; the test-only oracle / reference libraries use ``orc_`` / ``ref_`` and the
same argument lists without the stream (see oracle/backends.py).
Argument order and meaning follow the reference member functions
(Source/MaestroMakeEdgeScalars.cpp:6, Source/MaestroMakeFlux.cpp:18,467,
Source/MaestroCelltoEdge.cpp:5, Source/MaestroMakew0.cpp:5).
"""
from __future__ import annotations

import ctypes as C

from .abi import Base, EdgeParams, Fab, FluxParams, Geom, Sphr, fab_array  # noqa: F401


class MxError(RuntimeError):
    pass


class Calls:
    def __init__(self, lib: C.CDLL, prefix: str, with_stream: bool):
        self.lib = lib
        self.prefix = prefix
        self.with_stream = with_stream

    # -- helpers ----------------------------------------------------------------------
    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def _check(self, rc, name):
        if rc != 0:
            msg = ""
            if self.prefix == "mx_":
                self.lib.mx_last_error.restype = C.c_char_p
                msg = self.lib.mx_last_error().decode()
            raise MxError(f"{self.prefix}{name} failed with code {rc}: {msg}")

    def _tail(self, stream):
        return (C.c_void_p(stream),) if self.with_stream else ()

    @staticmethod
    def _faces(fabs):
        return fab_array(fabs) if fabs is not None else None

    @staticmethod
    def _base(b):
        return C.byref(b.c()) if b is not None else None

    # -- Maestro::MakeEdgeScal --------------------------------------------------------
    def make_edge_scal(self, geom: Geom, state: Fab, sedge, umac, force: Fab,
                       par: EdgeParams, stream: int = 0):
        g, st, fo, p = geom.c(), state.c(), force.c(), par.c()
        rc = self._fn("make_edge_scal")(C.byref(g), C.byref(st), self._faces(sedge),
                                        self._faces(umac), C.byref(fo), C.byref(p),
                                        *self._tail(stream))
        self._check(rc, "make_edge_scal")

    # -- Maestro::MakeRhoXFlux --------------------------------------------------------
    def make_rhoX_flux(self, geom: Geom, sflux, etarhoflux: Fab, sedge, umac,
                       rho0_old: Base, rho0_edge_old: Base, r0mac_old, rho0_new: Base,
                       rho0_edge_new: Base, r0mac_new, rho0_predicted_edge: Base, w0: Base,
                       par: FluxParams, stream: int = 0):
        g, eta, p = geom.c(), etarhoflux.c(), par.c()
        rc = self._fn("make_rhoX_flux")(
            C.byref(g), self._faces(sflux), C.byref(eta), self._faces(sedge),
            self._faces(umac), self._base(rho0_old), self._base(rho0_edge_old),
            self._faces(r0mac_old), self._base(rho0_new), self._base(rho0_edge_new),
            self._faces(r0mac_new), self._base(rho0_predicted_edge), self._base(w0),
            C.byref(p), *self._tail(stream))
        self._check(rc, "make_rhoX_flux")

    # -- Maestro::MakeRhoHFlux --------------------------------------------------------
    def make_rhoh_flux(self, geom: Geom, sflux, sedge, umac, rho0_old, rho0_edge_old,
                       r0mac_old, rho0_new, rho0_edge_new, r0mac_new, rhoh0_old,
                       rhoh0_edge_old, rh0mac_old, rhoh0_new, rhoh0_edge_new, rh0mac_new,
                       h0mac_old, h0mac_new, par: FluxParams, stream: int = 0):
        g, p = geom.c(), par.c()
        rc = self._fn("make_rhoh_flux")(
            C.byref(g), self._faces(sflux), self._faces(sedge), self._faces(umac),
            self._base(rho0_old), self._base(rho0_edge_old), self._faces(r0mac_old),
            self._base(rho0_new), self._base(rho0_edge_new), self._faces(r0mac_new),
            self._base(rhoh0_old), self._base(rhoh0_edge_old), self._faces(rh0mac_old),
            self._base(rhoh0_new), self._base(rhoh0_edge_new), self._faces(rh0mac_new),
            self._faces(h0mac_old), self._faces(h0mac_new), C.byref(p), *self._tail(stream))
        self._check(rc, "make_rhoh_flux")

    # -- Maestro::CelltoEdge ----------------------------------------------------------
    def cell_to_edge(self, cell, edge, nr: int, r_lo: int, r_hi: int, stream: int = 0):
        from .abi import _ptr_of
        rc = self._fn("cell_to_edge")(C.c_void_p(_ptr_of(cell)), C.c_void_p(_ptr_of(edge)),
                                      C.c_int(nr), C.c_int(r_lo), C.c_int(r_hi),
                                      *self._tail(stream))
        self._check(rc, "cell_to_edge")

    # -- Maestro::Makew0 --------------------------------------------------------------
    @staticmethod
    def _vp(a):
        from .abi import _ptr_of
        return C.c_void_p(_ptr_of(a)) if a is not None else None

    def make_w0_planar(self, par, w0_old, Sbar, p0_old, p0_new, gamma1bar_old, gamma1bar_new,
                       p0_minus_peosbar, etarho_cc, w0, w0_force, stream: int = 0):
        """Returns p0bdot (0.0 unless par.add_pb)."""
        p = par.c()
        pb = C.c_double(0.0)
        v = self._vp
        rc = self._fn("make_w0_planar")(
            C.byref(p), v(w0_old), v(Sbar), v(p0_old), v(p0_new), v(gamma1bar_old),
            v(gamma1bar_new), v(p0_minus_peosbar), v(etarho_cc), v(w0), v(w0_force),
            C.byref(pb), *self._tail(stream))
        self._check(rc, "make_w0_planar")
        return pb.value

    def make_w0_sphr(self, par, w0_old, Sbar, rho0_old, rho0_new, p0_old, p0_new,
                     gamma1bar_old, gamma1bar_new, p0_minus_peosbar, etarho_cc, etarho_ec,
                     r_cc_loc, r_edge_loc, w0, w0_force, stream: int = 0, irreg: bool = False):
        """irreg: Makew0SphrIrreg (use_exact_base_state, Source/MaestroMakew0.cpp:742-940)"""
        p = par.c()
        v = self._vp
        rc = self._fn("make_w0_sphr_irreg" if irreg else "make_w0_sphr")(
            C.byref(p), v(w0_old), v(Sbar), v(rho0_old), v(rho0_new), v(p0_old), v(p0_new),
            v(gamma1bar_old), v(gamma1bar_new), v(p0_minus_peosbar), v(etarho_cc),
            v(etarho_ec), v(r_cc_loc), v(r_edge_loc), v(w0), v(w0_force), *self._tail(stream))
        self._check(rc, "make_w0_sphr")

    def make_w0_planar_ml(self, par, rg, w0_old, Sbar, p0_old, p0_new, gamma1bar_old, gamma1bar_new,
                          p0_minus_peosbar, etarho_cc, w0, w0_force, stream: int = 0):
        """Makew0Planar for a multilevel base state (host numpy arrays, shape (nlev, len));
        returns p0bdot."""
        p, g = par.c(), rg.c()
        pb = C.c_double(0.0)
        v = self._vp
        rc = self._fn("make_w0_planar_ml")(
            C.byref(p), C.byref(g), v(w0_old), v(Sbar), v(p0_old), v(p0_new), v(gamma1bar_old),
            v(gamma1bar_new), v(p0_minus_peosbar), v(etarho_cc), v(w0), v(w0_force), C.byref(pb),
            *self._tail(stream))
        self._check(rc, "make_w0_planar_ml")
        return pb.value

    # -- Maestro::UpdateScal ----------------------------------------------------------
    def update_scal(self, geom: Geom, stateold: Fab, statenew: Fab, sflux, force: Fab, par,
                    eos_mask=None, p0_cart: Fab = None, stream: int = 0):
        """Zone loop of UpdateScal.  Product / oracle (``mx_`` / ``orc_``): the zones that need
        the caller's EOS are flagged in ``eos_mask`` (uint8, one per valid zone) and their
        number is returned.  Compiled reference / overlay (``ref_`` / ``ovl_``): the whole
        member function incl. its EOS call, ``p0_cart`` required, returns None."""
        g, so, sn, fo, p = geom.c(), stateold.c(), statenew.c(), force.c(), par.c()
        if self.prefix in ("ref_", "ovl_"):
            p0 = p0_cart.c() if p0_cart is not None else None
            rc = self._fn("update_scal")(C.byref(g), C.byref(so), C.byref(sn),
                                         self._faces(sflux), C.byref(fo),
                                         C.byref(p0) if p0 is not None else None, C.byref(p))
            self._check(rc, "update_scal")
            return None
        n = C.c_longlong(0)
        rc = self._fn("update_scal")(C.byref(g), C.byref(so), C.byref(sn), self._faces(sflux),
                                     C.byref(fo), C.byref(p), self._vp(eos_mask), C.byref(n),
                                     *self._tail(stream))
        self._check(rc, "update_scal")
        return n.value

    # ==== the passes either side of the path ==============
    @staticmethod
    def _sphr(s):
        return C.byref(s.c()) if s is not None else None

    @staticmethod
    def _fab(f):
        return C.byref(f.c()) if f is not None else None

    # -- Maestro::ConvertRhoXToX / ConvertRhoHToH (Source/MaestroConvert.cpp:64-114) ------
    def convert_by_rho(self, geom: Geom, scal: Fab, rho_comp: int, start_comp: int, num_comp: int,
                       flag: bool, stream: int = 0):
        g, s = geom.c(), scal.c()
        rc = self._fn("convert_by_rho")(C.byref(g), C.byref(s), rho_comp, start_comp, num_comp,
                                        int(bool(flag)), *self._tail(stream))
        self._check(rc, "convert_by_rho")

    # -- Maestro::Put1dArrayOnCart (Source/MaestroFill3dData.cpp:48-337) ------------------
    def put_1d_array_on_cart(self, geom: Geom, s0: Base, lev: int, s0_cart: Fab, edge: bool,
                             vec: bool, sphr: Sphr = None, stream: int = 0):
        g, f = geom.c(), s0_cart.c()
        rc = self._fn("put_1d_array_on_cart")(C.byref(g), self._base(s0), lev, C.byref(f),
                                              int(bool(edge)), int(bool(vec)), self._sphr(sphr),
                                              *self._tail(stream))
        self._check(rc, "put_1d_array_on_cart")

    # -- Maestro::PutInPertForm (Source/MaestroConvert.cpp:6-62) --------------------------
    def put_in_pert_form(self, geom: Geom, scal: Fab, s0: Base, lev: int, comp: int, flag: bool,
                         sphr: Sphr = None, stream: int = 0):
        g, f = geom.c(), scal.c()
        rc = self._fn("put_in_pert_form")(C.byref(g), C.byref(f), self._base(s0), lev, comp,
                                          int(bool(flag)), self._sphr(sphr), *self._tail(stream))
        self._check(rc, "put_in_pert_form")

    # -- Maestro::ModifyScalForce (Source/MaestroForce.cpp:261-467) -----------------------
    def modify_scal_force(self, geom: Geom, scal_force: Fab, state: Fab, umac, s0_edge: Base,
                          lev: int, s0_cart: Fab, w0_cart: Fab, w0, comp: int, fullform: int,
                          bcf_hi: int, sphr: Sphr = None, stream: int = 0, s0_edge_cart: Fab = None):
        g, f, st, s0c = geom.c(), scal_force.c(), state.c(), s0_cart.c()
        rc = self._fn("modify_scal_force")(
            C.byref(g), C.byref(f), C.byref(st), self._faces(umac), self._base(s0_edge), lev,
            C.byref(s0c), self._fab(w0_cart), self._vp(w0), comp, fullform, bcf_hi,
            self._fab(s0_edge_cart), self._sphr(sphr), *self._tail(stream))
        self._check(rc, "modify_scal_force")

    # -- Maestro::Addw0 (Source/MaestroFill3dData.cpp:357-461) ----------------------------
    def addw0(self, geom: Geom, uedge, w0mac, w0: Base, lev: int, mult: float, spherical: int,
              stream: int = 0):
        g = geom.c()
        rc = self._fn("addw0")(C.byref(g), self._faces(uedge), self._faces(w0mac), self._base(w0),
                               lev, C.c_double(mult), spherical, *self._tail(stream))
        self._check(rc, "addw0")

    # -- Maestro::MakeW0mac / MakeS0mac (Source/MaestroFill3dData.cpp:464-1188) -----------
    def make_w0mac(self, geom: Geom, w0mac, w0_cart: Fab, w0, sphr: Sphr, stream: int = 0):
        g = geom.c()
        rc = self._fn("make_w0mac")(C.byref(g), self._faces(w0mac), self._fab(w0_cart),
                                    self._vp(w0), self._sphr(sphr), *self._tail(stream))
        self._check(rc, "make_w0mac")

    def make_s0mac(self, geom: Geom, s0: Base, s0mac, sphr: Sphr, stream: int = 0, s0_cart: Fab = None):
        g = geom.c()
        rc = self._fn("make_s0mac")(C.byref(g), self._base(s0), self._faces(s0mac), self._fab(s0_cart),
                                    self._sphr(sphr), *self._tail(stream))
        self._check(rc, "make_s0mac")

    # -- Maestro::UpdateVel (Source/MaestroAdvection.cpp:299-526) -------------------------
    def update_vel(self, geom: Geom, uold: Fab, unew: Fab, umac, uedge, force: Fab, sponge: Fab,
                   w0_cart: Fab, w0mac, dt: float, do_sponge: int, spherical: int,
                   stream: int = 0):
        g, uo, un, fo = geom.c(), uold.c(), unew.c(), force.c()
        rc = self._fn("update_vel")(C.byref(g), C.byref(uo), C.byref(un), self._faces(umac),
                                    self._faces(uedge), C.byref(fo), self._fab(sponge),
                                    self._fab(w0_cart), self._faces(w0mac), C.c_double(dt),
                                    do_sponge, spherical, *self._tail(stream))
        self._check(rc, "update_vel")

    # -- Maestro::MakeUtrans (Source/MaestroMakeUtrans.cpp:6-648) --------------------------
    def make_utrans(self, geom: Geom, utilde: Fab, ufull: Fab, utrans, w0_cart: Fab, w0mac, par,
                    stream: int = 0):
        g, ut, uf, p = geom.c(), utilde.c(), ufull.c(), par.c()
        rc = self._fn("make_utrans")(C.byref(g), C.byref(ut), C.byref(uf), self._faces(utrans),
                                     self._fab(w0_cart), self._faces(w0mac), C.byref(p),
                                     *self._tail(stream))
        self._check(rc, "make_utrans")

    # -- Maestro::VelPred (Source/MaestroVelPred.cpp:5-1917) --------------------------------
    def vel_pred(self, geom: Geom, utilde: Fab, ufull: Fab, utrans, umac, w0mac, w0_cart: Fab,
                 force: Fab, par, stream: int = 0):
        g, ut, uf, fo, p = geom.c(), utilde.c(), ufull.c(), force.c(), par.c()
        rc = self._fn("vel_pred")(C.byref(g), C.byref(ut), C.byref(uf), self._faces(utrans),
                                  self._faces(umac), self._faces(w0mac), self._fab(w0_cart),
                                  C.byref(fo), C.byref(p), *self._tail(stream))
        self._check(rc, "vel_pred")
