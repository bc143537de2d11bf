"""ctypes mirror of include/mxadv.h (the C ABI of libmxadv.so).

The structs here are shared by three consumers with identical signatures:
  * the product library ``maestroex_b200/csrc/libmxadv.so``  (``mx_*`` entry points),
  * the C oracle ``oracle/libmxadv_oracle.so``              (``orc_*``, tests only),
  * the compiled reference ``oracle/_ref/libmaestro_ref{2,3}d.so`` (``ref_*``, tests only).

Data layout follows AMReX ``Array4`` (external/amrex/Src/Base/AMReX_Array4.H:86-91):
i fastest, component slowest, over the ghost-grown box.  A numpy / torch array of shape
``(ncomp, nz, ny, nx)`` in C order therefore *is* a fab.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class MxFab(C.Structure):
    _fields_ = [("p", C.c_void_p), ("lo", C.c_int * 3), ("hi", C.c_int * 3), ("ncomp", C.c_int)]


class MxGeom(C.Structure):
    _fields_ = [
        ("dim", C.c_int),
        ("valid_lo", C.c_int * 3),
        ("valid_hi", C.c_int * 3),
        ("dom_lo", C.c_int * 3),
        ("dom_hi", C.c_int * 3),
        ("dx", C.c_double * 3),
    ]


class MxBase(C.Structure):
    _fields_ = [("p", C.c_void_p), ("nlev", C.c_int), ("len", C.c_int), ("nvar", C.c_int)]


class MxEdgeParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double),
        ("rel_eps", C.c_double),
        ("ppm_type", C.c_int),
        ("slope_order", C.c_int),
        ("ppm_trace_forces", C.c_int),
        ("is_vel", C.c_int),
        ("is_conservative", C.c_int),
        ("start_scomp", C.c_int),
        ("start_bccomp", C.c_int),
        ("num_comp", C.c_int),
        ("nbccomp", C.c_int),
        ("bc", c_int_p),
    ]


class MxVelPredParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double),
        ("rel_eps", C.c_double),
        ("ppm_type", C.c_int),
        ("slope_order", C.c_int),
        ("ppm_trace_forces", C.c_int),
        ("spherical", C.c_int),
        ("phys_bc_lo", C.c_int * 3),
        ("phys_bc_hi", C.c_int * 3),
        ("bc", c_int_p),
    ]


class MxFluxParams(C.Structure):
    _fields_ = [
        ("species_pred_type", C.c_int),
        ("enthalpy_pred_type", C.c_int),
        ("spherical", C.c_int),
        ("use_exact_base_state", C.c_int),
        ("evolve_base_state", C.c_int),
        ("rho_comp", C.c_int),
        ("rhoh_comp", C.c_int),
        ("first_spec", C.c_int),
        ("num_spec", C.c_int),
        ("start_comp", C.c_int),
        ("num_comp", C.c_int),
        ("lev", C.c_int),
    ]


class MxW0Params(C.Structure):
    _fields_ = [
        ("nr_fine", C.c_int),
        ("base_cutoff_density_coord", C.c_int),
        ("is_predictor", C.c_int),
        ("add_pb", C.c_int),
        ("dr", C.c_double),
        ("dt", C.c_double),
        ("dt_in", C.c_double),
        ("dtold_in", C.c_double),
        ("grav_const", C.c_double),
        ("dpdt_factor", C.c_double),
        ("base_cutoff_density", C.c_double),
    ]


class MxUpdateParams(C.Structure):
    _fields_ = [
        ("dt", C.c_double),
        ("base_cutoff_density", C.c_double),
        ("do_eos_h_above_cutoff", C.c_int),
        ("rho_comp", C.c_int),
        ("rhoh_comp", C.c_int),
        ("temp_comp", C.c_int),
        ("first_spec", C.c_int),
        ("num_spec", C.c_int),
        ("start_comp", C.c_int),
        ("num_comp", C.c_int),
    ]


# AMReX BCType codes (external/amrex/Src/Base/AMReX_BC_TYPES.H:68-76)
BC_REFLECT_ODD, BC_INT_DIR, BC_REFLECT_EVEN, BC_FOEXTRAP, BC_EXT_DIR, BC_HOEXTRAP = -1, 0, 1, 2, 3, 4


def _ptr_of(arr) -> int:
    """Raw address of a numpy array or a torch tensor (host or device)."""
    if isinstance(arr, np.ndarray):
        return arr.ctypes.data
    return int(arr.data_ptr())  # torch.Tensor


@dataclass
class Fab:
    """One box of a MultiFab: ``data`` has shape (ncomp, nz, ny, nx), C-contiguous fp64."""

    data: object  # np.ndarray or torch.Tensor
    lo: tuple
    hi: tuple

    @property
    def ncomp(self) -> int:
        return int(self.data.shape[0])

    def c(self) -> MxFab:
        shp = tuple(int(x) for x in self.data.shape)
        ext = tuple(self.hi[d] - self.lo[d] + 1 for d in range(3))
        assert shp[1:] == (ext[2], ext[1], ext[0]), (shp, ext)
        if isinstance(self.data, np.ndarray):
            assert self.data.dtype == np.float64 and self.data.flags["C_CONTIGUOUS"]
        else:
            assert self.data.is_contiguous()
        f = MxFab()
        f.p = _ptr_of(self.data)
        f.lo[:] = self.lo
        f.hi[:] = self.hi
        f.ncomp = shp[0]
        return f


def fab_array(fabs: Sequence[Fab]):
    arr = (MxFab * len(fabs))()
    for i, f in enumerate(fabs):
        arr[i] = f.c()
    return arr


@dataclass
class Geom:
    dim: int
    valid_lo: tuple
    valid_hi: tuple
    dom_lo: tuple
    dom_hi: tuple
    dx: tuple

    def c(self) -> MxGeom:
        g = MxGeom()
        g.dim = self.dim
        pad = lambda t, z: tuple(t) + (z,) * (3 - len(t))  # noqa: E731
        g.valid_lo[:] = pad(self.valid_lo, 0)
        g.valid_hi[:] = pad(self.valid_hi, 0)
        g.dom_lo[:] = pad(self.dom_lo, 0)
        g.dom_hi[:] = pad(self.dom_hi, 0)
        g.dx[:] = pad(self.dx, 0.0)
        return g

    def cell_box(self, ng: int):
        """(lo, hi) of the valid box grown by ng in the first `dim` directions."""
        lo = [0, 0, 0]
        hi = [0, 0, 0]
        for d in range(self.dim):
            lo[d] = self.valid_lo[d] - ng
            hi[d] = self.valid_hi[d] + ng
        return tuple(lo), tuple(hi)

    def face_box(self, d: int, ng: int):
        lo, hi = self.cell_box(ng)
        hi = list(hi)
        hi[d] += 1
        return lo, tuple(hi)


def box_shape(lo, hi):
    return (hi[2] - lo[2] + 1, hi[1] - lo[1] + 1, hi[0] - lo[0] + 1)


@dataclass
class Base:
    """A BaseState array: data shape (nlev, len, nvar) C-order (Source/BaseState.H:80)."""

    data: object

    def c(self) -> MxBase:
        b = MxBase()
        shp = tuple(int(x) for x in self.data.shape)
        if len(shp) == 1:
            shp = (1, shp[0], 1)
        b.p = _ptr_of(self.data)
        b.nlev, b.len, b.nvar = shp
        return b


@dataclass
class EdgeParams:
    dt: float
    rel_eps: float
    ppm_type: int = 1
    slope_order: int = 4
    ppm_trace_forces: int = 0
    is_vel: int = 0
    is_conservative: int = 0
    start_scomp: int = 0
    start_bccomp: int = 0
    num_comp: int = 1
    bc: Optional[np.ndarray] = None  # (nbccomp, 6) int32: lo_x,lo_y,lo_z,hi_x,hi_y,hi_z
    _keep: list = field(default_factory=list, repr=False)

    def c(self) -> MxEdgeParams:
        p = MxEdgeParams()
        p.dt = self.dt
        p.rel_eps = self.rel_eps
        p.ppm_type = self.ppm_type
        p.slope_order = self.slope_order
        p.ppm_trace_forces = self.ppm_trace_forces
        p.is_vel = self.is_vel
        p.is_conservative = self.is_conservative
        p.start_scomp = self.start_scomp
        p.start_bccomp = self.start_bccomp
        p.num_comp = self.num_comp
        bc = np.ascontiguousarray(self.bc, dtype=np.int32)
        assert bc.ndim == 2 and bc.shape[1] == 6
        self._keep.append(bc)
        p.nbccomp = bc.shape[0]
        p.bc = bc.ctypes.data_as(c_int_p)
        return p


# amrex::PhysBCType (AMReX_BC_TYPES.H:61-64)
PHYS_INTERIOR, PHYS_INFLOW, PHYS_OUTFLOW, PHYS_SYMMETRY, PHYS_SLIPWALL, PHYS_NOSLIPWALL = range(6)


@dataclass
class VelPredParams:
    """Runtime parameters of Maestro::MakeUtrans / VelPred (MxVelPredParams)."""
    dt: float
    rel_eps: float
    ppm_type: int = 1
    slope_order: int = 4
    ppm_trace_forces: int = 0
    spherical: int = 0
    phys_bc_lo: tuple = (0, 0, 0)
    phys_bc_hi: tuple = (0, 0, 0)
    bc: Optional[np.ndarray] = None  # (dim, 6) int32: bcs_u, lo_x,lo_y,lo_z,hi_x,hi_y,hi_z
    _keep: list = field(default_factory=list, repr=False)

    def c(self) -> MxVelPredParams:
        p = MxVelPredParams()
        p.dt = self.dt
        p.rel_eps = self.rel_eps
        p.ppm_type = self.ppm_type
        p.slope_order = self.slope_order
        p.ppm_trace_forces = self.ppm_trace_forces
        p.spherical = self.spherical
        for d in range(3):
            p.phys_bc_lo[d] = self.phys_bc_lo[d] if d < len(self.phys_bc_lo) else 0
            p.phys_bc_hi[d] = self.phys_bc_hi[d] if d < len(self.phys_bc_hi) else 0
        bc = np.ascontiguousarray(self.bc, dtype=np.int32)
        assert bc.ndim == 2 and bc.shape[1] == 6
        self._keep.append(bc)
        p.bc = bc.ctypes.data_as(c_int_p)
        return p


@dataclass
class FluxParams:
    species_pred_type: int = 1
    enthalpy_pred_type: int = 1
    spherical: int = 0
    use_exact_base_state: int = 0
    evolve_base_state: int = 1
    rho_comp: int = 0
    rhoh_comp: int = 1
    first_spec: int = 4
    num_spec: int = 3
    start_comp: int = 4
    num_comp: int = 3
    lev: int = 0

    def c(self) -> MxFluxParams:
        p = MxFluxParams()
        for name, _ in MxFluxParams._fields_:
            setattr(p, name, int(getattr(self, name)))
        return p


@dataclass
class W0Params:
    """Arguments / parameters of Maestro::Makew0 (Source/MaestroMakew0.cpp:5-14,60-69)."""
    nr_fine: int
    base_cutoff_density_coord: int
    dr: float
    dt: float
    dt_in: float
    dtold_in: float
    is_predictor: int = 1
    add_pb: int = 0
    grav_const: float = -1.5e10
    dpdt_factor: float = 0.0
    base_cutoff_density: float = -1.0

    def c(self) -> MxW0Params:
        p = MxW0Params()
        for name, typ in MxW0Params._fields_:
            v = getattr(self, name)
            setattr(p, name, int(v) if typ is C.c_int else float(v))
        return p


@dataclass
class UpdateParams:
    """Parameters of Maestro::UpdateScal (Source/MaestroAdvection.cpp:86-91)."""
    dt: float
    start_comp: int
    num_comp: int
    num_spec: int
    base_cutoff_density: float = -1.0
    do_eos_h_above_cutoff: int = 1
    rho_comp: int = 0
    rhoh_comp: int = 1
    temp_comp: int = 2
    first_spec: int = 4

    def c(self) -> MxUpdateParams:
        p = MxUpdateParams()
        for name, typ in MxUpdateParams._fields_:
            v = getattr(self, name)
            setattr(p, name, int(v) if typ is C.c_int else float(v))
        return p


class MxSphr(C.Structure):
    _fields_ = [
        ("spherical", C.c_int),
        ("nr_fine", C.c_int),
        ("dr_fine", C.c_double),
        ("prob_lo", C.c_double * 3),
        ("center", C.c_double * 3),
        ("r_cc_loc", C.c_void_p),
        ("r_edge_loc", C.c_void_p),
        ("s0_interp_type", C.c_int),
        ("w0_interp_type", C.c_int),
        ("s0mac_interp_type", C.c_int),
        ("w0mac_interp_type", C.c_int),
    ]


@dataclass
class Sphr:
    """base_geom / center / interp types read by the 1-D -> 3-D mappings
    (Source/MaestroFill3dData.cpp); r_cc_loc / r_edge_loc are HOST numpy arrays."""
    nr_fine: int
    dr_fine: float
    prob_lo: tuple
    center: tuple
    r_cc_loc: object
    r_edge_loc: object
    s0_interp_type: int = 3
    w0_interp_type: int = 2
    s0mac_interp_type: int = 1
    w0mac_interp_type: int = 1
    spherical: int = 1

    def c(self) -> MxSphr:
        p = MxSphr()
        p.spherical = self.spherical
        p.nr_fine = self.nr_fine
        p.dr_fine = self.dr_fine
        p.prob_lo[:] = self.prob_lo
        p.center[:] = self.center
        p.r_cc_loc = _ptr_of(self.r_cc_loc)
        p.r_edge_loc = _ptr_of(self.r_edge_loc)
        for n in ("s0_interp_type", "w0_interp_type", "s0mac_interp_type", "w0mac_interp_type"):
            setattr(p, n, int(getattr(self, n)))
        return p


class MxRadialGeom(C.Structure):
    _fields_ = [
        ("max_radial_level", C.c_int), ("finest_radial_level", C.c_int), ("nr_fine", C.c_int),
        ("dr", C.c_void_p), ("nr", C.c_void_p), ("base_cutoff_density_coord", C.c_void_p),
        ("numdisjointchunks", C.c_void_p), ("r_start_coord", C.c_void_p), ("r_end_coord", C.c_void_p),
        ("chunk_len", C.c_int),
    ]


@dataclass
class RadialGeom:
    """base_geom of a multilevel planar base state (Source/BaseStateGeometry.H); numpy arrays:
    dr float64 [nlev], nr / base_cutoff_density_coord int32 [nlev], numdisjointchunks int32 [nfine],
    r_start_coord / r_end_coord int32 [nfine, chunk_len] (entry (n, j), j = 1 .. chunks)."""
    nr_fine: int
    dr: object
    nr: object
    base_cutoff_density_coord: object
    numdisjointchunks: object
    r_start_coord: object
    r_end_coord: object

    def c(self) -> MxRadialGeom:
        g = MxRadialGeom()
        g.max_radial_level = len(self.dr) - 1
        g.finest_radial_level = len(self.numdisjointchunks) - 1
        g.nr_fine = self.nr_fine
        for k in ("dr", "nr", "base_cutoff_density_coord", "numdisjointchunks", "r_start_coord", "r_end_coord"):
            a = getattr(self, k)
            assert a.flags["C_CONTIGUOUS"] and a.dtype == (np.float64 if k == "dr" else np.int32), k
            setattr(g, k, a.ctypes.data)
        g.chunk_len = int(self.r_start_coord.shape[1])
        return g


def byref_or_null(x):
    return C.byref(x) if x is not None else None
