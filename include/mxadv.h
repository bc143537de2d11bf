/*
 * mxadv.h — C ABI of libmxadv.so: B200-native (sm_100a, fp64) scalar-advection hot path of
 * MAESTROeX: Godunov edge-state prediction (slopes / PPM / corner-coupled transverse terms /
 * upwinding) and (rho X_k), (rho h) face-flux assembly, plus the 1-D base-state helpers.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * MAESTROeX tree, commit 543ea58f).  The reference has no FFI: these are the functions the
 * source-shadowing overlay in maestroex_b200/overlay/ (same `Maestro::` member signatures)
 * binds, see INTEGRATION.md.
 *
 * Conventions
 *  - Plain pointers + POD structs only.  All floating point data is IEEE binary64.
 *  - A fab (`MxFab`) is one box of an AMReX MultiFab in Array4 order
 *    (external/amrex/Src/Base/AMReX_Array4.H:86-91): i fastest, then j, k, component slowest,
 *    over the GHOST-GROWN box [lo,hi] (inclusive).  2-D data has lo[2] = hi[2] = 0.
 *    value(i,j,k,n) = p[(i-lo[0]) + nx*((j-lo[1]) + ny*((k-lo[2]) + nz*n))].
 *  - Face-centred fabs in direction d are indexed by the low-face convention
 *    (face i sits between cells i-1 and i) and their valid box is [vlo, vhi + e_d].
 *  - Pointers may be device pointers (AMReX built USE_CUDA=TRUE: The_Arena memory, or managed
 *    memory as used by BaseState, Source/BaseState.H:300) or plain host pointers (CPU AMReX
 *    build).  The library detects which (cudaPointerGetAttributes); host fabs are staged
 *    through a library-owned device workspace (pinned double-buffered z-slabs).  There is no
 *    CPU compute path: without a CUDA device every compute entry point fails with
 *    MX_ERR_NO_DEVICE.
 *  - Caller owns every buffer.  Outputs: only components [start_scomp, start_scomp+num_comp)
 *    of `sedge`, the listed components of `sflux`, and `etarhoflux` (read-modify-write) are
 *    written, and only on their valid (nodal) boxes; everything else is left untouched.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls on
 *    device data are asynchronous on that stream; calls on host data return after the results
 *    are back in host memory.
 *  - Return value: 0 on success, negative MX_ERR_* otherwise; mx_last_error() returns a
 *    thread-local message.  (The reference aborts via amrex::Abort; the overlay turns a
 *    non-zero return into amrex::Abort(mx_last_error()).)
 */
#ifndef MXADV_H
#define MXADV_H

#ifdef __cplusplus
extern "C" {
#endif

#define MXADV_ABI_VERSION 1

enum {
    MX_OK = 0,
    MX_ERR_INVALID = -1,     /* bad argument / inconsistent boxes */
    MX_ERR_UNSUPPORTED = -2, /* option combination the reference itself rejects or that is
                                not built (never a silent fallback) */
    MX_ERR_CUDA = -3,        /* CUDA runtime error (message has the cudaError string) */
    MX_ERR_NO_DEVICE = -4    /* no CUDA device: there is no CPU fallback */
};

/* AMReX mathematical BC codes, external/amrex/Src/Base/AMReX_BC_TYPES.H:68-76 */
enum {
    MX_BC_REFLECT_ODD = -1,
    MX_BC_INT_DIR = 0,
    MX_BC_REFLECT_EVEN = 1,
    MX_BC_FOEXTRAP = 2,
    MX_BC_EXT_DIR = 3,
    MX_BC_HOEXTRAP = 4
};

/* species_pred_type / enthalpy_pred_type, Source/MaestroMakeFlux.cpp:6-16 */
enum { MX_PRED_RHOPRIME_AND_X = 1, MX_PRED_RHOX = 2, MX_PRED_RHO_AND_X = 3 };
enum {
    MX_PRED_RHOH = 0,
    MX_PRED_RHOHPRIME = 1,
    MX_PRED_H = 2,
    MX_PRED_T_THEN_RHOHPRIME = 3,
    MX_PRED_T_THEN_H = 4,
    MX_PRED_HPRIME = 5,
    MX_PRED_TPRIME_THEN_H = 6
};

/* One box of a MultiFab: mf[mfi].dataPtr(), mf[mfi].box(), mf.nComp()
 * (external/amrex/Src/Base/AMReX_BaseFab.H:277,291,352). */
typedef struct MxFab {
    double* p;
    int lo[3];
    int hi[3];
    int ncomp;
} MxFab;

/* Index space: mfi.validbox(), geom[lev].Domain(), geom[lev].CellSizeArray(). */
typedef struct MxGeom {
    int dim;         /* AMREX_SPACEDIM: 2 or 3 */
    int valid_lo[3]; /* cell-centred valid box of this fab */
    int valid_hi[3];
    int dom_lo[3]; /* cell-centred problem domain */
    int dom_hi[3];
    double dx[3];
} MxGeom;

/* A BaseState<Real> viewed as BaseStateArray (Source/BaseState.H:14-81):
 * value(lev, r, n) = p[(lev*len + r)*nvar + n]. */
typedef struct MxBase {
    const double* p;
    int nlev;
    int len;
    int nvar;
} MxBase;

/* Arguments + runtime parameters of Maestro::MakeEdgeScal
 * (Source/MaestroMakeEdgeScalars.cpp:6-12; maestro.* parameters from
 * Source/param/_cpp_parameters; dt and rel_eps are Maestro members, Source/Maestro.H:1873,1974). */
typedef struct MxEdgeParams {
    double dt;
    double rel_eps;
    int ppm_type;         /* 0 (slopes), 1 (CW84), 2 (Colella-Sekora 2008) */
    int slope_order;      /* 0, 2, 4 (used when ppm_type == 0) */
    int ppm_trace_forces; /* 0 / 1 */
    int is_vel;
    int is_conservative;
    int start_scomp;  /* first component of state/force/sedge to sweep */
    int start_bccomp; /* first entry of bc[] to use */
    int num_comp;
    int nbccomp;   /* number of BCRec entries in bc */
    const int* bc; /* HOST pointer, nbccomp x 6: {lo_x,lo_y,lo_z,hi_x,hi_y,hi_z} per BCRec */
} MxEdgeParams;

/* Runtime parameters + state indices used by MakeRhoXFlux / MakeRhoHFlux
 * (Source/MaestroMakeFlux.cpp:35-37,495-504; state_indices.H from Source/param/_variables). */
typedef struct MxFluxParams {
    int species_pred_type;
    int enthalpy_pred_type;
    int spherical;
    int use_exact_base_state;
    int evolve_base_state;
    int rho_comp;   /* Rho */
    int rhoh_comp;  /* RhoH */
    int first_spec; /* FirstSpec */
    int num_spec;   /* NumSpec */
    int start_comp; /* MakeRhoXFlux: first species component to flux */
    int num_comp;   /* MakeRhoXFlux: number of components */
    int lev;        /* level index into the BaseState arrays */
} MxFluxParams;

/* ---- library / device management ------------------------------------------------------ */

int mx_abi_version(void);
const char* mx_last_error(void);
/* Number of CUDA devices visible (0 if none / no driver); never fails. */
int mx_device_count(void);
/* Number of kernels this library has launched since load (or since mx_reset_launch_count). */
long long mx_launch_count(void);
void mx_reset_launch_count(void);
/* How many of those were the TMA-fed instantiation of the streaming kernel k_stream4 (scalar
 * planes staged by cp.async.bulk.tensor); the remainder of the fused launches ran its cp.async
 * instantiation (fabs whose rows are an odd number of doubles or whose base is not 16-byte
 * aligned cannot be described to TMA). */
long long mx_tma_launch_count(void);
/* Release the cached device workspaces. */
int mx_free_workspace(void);

/* ---- Maestro::MakeEdgeScal ---------------------------------------------------------------
 * Replaces Maestro::MakeEdgeScal (Source/MaestroMakeEdgeScalars.cpp:6-302, decl
 * Source/Maestro.H:913-919) for ONE box of ONE level, including everything it calls:
 * Slopex/y/z (MaestroSlopes.cpp), PPM (MaestroPPM.cpp), MakeDivU, MakeEdgeScalPredictor,
 * MakeEdgeScalTransverse, MakeEdgeScalEdges.  state: >= 3 ghost cells (4 if ppm_type == 2);
 * umac[d]: face-centred in d, 1 ghost; force: 1 ghost (state's ghosts if ppm_trace_forces);
 * sedge[d]: face-centred in d, any ghosts.  AverageDownFaces (is_vel only) is the caller's. */
int mx_make_edge_scal(const MxGeom* geom, const MxFab* state, const MxFab* sedge /*[dim]*/,
                      const MxFab* umac /*[dim]*/, const MxFab* force,
                      const MxEdgeParams* par, void* stream);

/* ---- Maestro::MakeRhoXFlux ---------------------------------------------------------------
 * Replaces Maestro::MakeRhoXFlux (Source/MaestroMakeFlux.cpp:18-465, decl Maestro.H:230-245)
 * for one box: zeroes sflux(Rho), computes the species fluxes, sums them into sflux(Rho) and
 * accumulates etarhoflux on the vertical faces.  r0mac_old/new (3 fabs each) are only read
 * when par->spherical; pass NULL otherwise.  Flux registers / AverageDownFaces (AMR) stay
 * with the caller. */
int mx_make_rhoX_flux(const MxGeom* geom, const MxFab* sflux /*[dim]*/,
                      const MxFab* etarhoflux, const MxFab* sedge /*[dim]*/,
                      const MxFab* umac /*[dim]*/, const MxBase* rho0_old,
                      const MxBase* rho0_edge_old, const MxFab* r0mac_old /*[3] or NULL*/,
                      const MxBase* rho0_new, const MxBase* rho0_edge_new,
                      const MxFab* r0mac_new /*[3] or NULL*/,
                      const MxBase* rho0_predicted_edge, const MxBase* w0,
                      const MxFluxParams* par, void* stream);

/* ---- Maestro::MakeRhoHFlux ---------------------------------------------------------------
 * Replaces Maestro::MakeRhoHFlux (Source/MaestroMakeFlux.cpp:467-1067, decl
 * Maestro.H:270-295) for one box.  The *mac fabs are only read when par->spherical. */
int mx_make_rhoh_flux(const MxGeom* geom, const MxFab* sflux /*[dim]*/,
                      const MxFab* sedge /*[dim]*/, const MxFab* umac /*[dim]*/,
                      const MxBase* rho0_old, const MxBase* rho0_edge_old,
                      const MxFab* r0mac_old, const MxBase* rho0_new,
                      const MxBase* rho0_edge_new, const MxFab* r0mac_new,
                      const MxBase* rhoh0_old, const MxBase* rhoh0_edge_old,
                      const MxFab* rh0mac_old, const MxBase* rhoh0_new,
                      const MxBase* rhoh0_edge_new, const MxFab* rh0mac_new,
                      const MxFab* h0mac_old, const MxFab* h0mac_new,
                      const MxFluxParams* par, void* stream);

/* ---- fused density / enthalpy passes -----------------------------------------------------
 * What DensityAdvance does between Source/MaestroDensityAdvance.cpp:115 and :217 in one
 * pass over the data: MakeEdgeScal(species X_k) + MakeEdgeScal(rho' or rho) + MakeRhoXFlux.
 * `state` holds X_k in the species components and rho' (or rho) in Rho, i.e. the state as it
 * is after ConvertRhoXToX / PutInPertForm (:98-104).  `epar` describes the species sweep
 * (start_scomp = FirstSpec, num_comp = NumSpec, start_bccomp = FirstSpec); Rho is swept with
 * BCRec index `rho_bccomp`.  Requires 3-D, is_vel = 0, is_conservative = 0,
 * species_pred_type 1 or 3.  Results are bit-identical to the three separate calls. */
int mx_density_edge_flux(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                         const MxFab* sflux, const MxFab* etarhoflux, const MxFab* umac,
                         const MxFab* force, const MxEdgeParams* epar, int rho_bccomp,
                         const MxBase* rho0_old, const MxBase* rho0_edge_old,
                         const MxFab* r0mac_old, const MxBase* rho0_new,
                         const MxBase* rho0_edge_new, const MxFab* r0mac_new,
                         const MxBase* rho0_predicted_edge, const MxBase* w0,
                         const MxFluxParams* fpar, void* stream);

/* EnthalpyAdvance's MakeEdgeScal(pred_comp) + MakeRhoHFlux
 * (Source/MaestroEnthalpyAdvance.cpp:141-146,229,312) in one pass; epar sweeps the single
 * component RhoH (enthalpy_pred_type 0/1/2/5) and reads sedge(Rho) where the flux form needs
 * it.  3-D, plane-parallel base state (spherical: mx_enthalpy_edge_flux_sph). */
int mx_enthalpy_edge_flux(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                          const MxFab* sflux, const MxFab* umac, const MxFab* force,
                          const MxEdgeParams* epar, const MxBase* rho0_old,
                          const MxBase* rho0_edge_old, const MxBase* rho0_new,
                          const MxBase* rho0_edge_new, const MxBase* rhoh0_old,
                          const MxBase* rhoh0_edge_old, const MxBase* rhoh0_new,
                          const MxBase* rhoh0_edge_new, const MxFluxParams* fpar,
                          void* stream);

/* The same for a spherical base state (wdconvect): the flux forms of
 * Source/MaestroMakeFlux.cpp:788-1021 read the base state from face arrays (3 fabs each, as
 * MakeRhoHFlux receives them): r0mac / h0mac when !use_exact_base_state, rh0mac otherwise; the
 * 1-D BaseState arrays are not used.  use_exact_base_state composes the two separate kernels. */
int mx_enthalpy_edge_flux_sph(const MxGeom* geom, const MxFab* state, const MxFab* sedge,
                              const MxFab* sflux, const MxFab* umac, const MxFab* force,
                              const MxEdgeParams* epar, const MxFab* r0mac_old,
                              const MxFab* r0mac_new, const MxFab* rh0mac_old,
                              const MxFab* rh0mac_new, const MxFab* h0mac_old,
                              const MxFab* h0mac_new, const MxFluxParams* fpar, void* stream);

/* ---- Maestro::CelltoEdge -----------------------------------------------------------------
 * Replaces Maestro::CelltoEdge (Source/MaestroCelltoEdge.cpp:5-55) for one radial level with
 * one disjoint chunk covering [r_lo, r_hi] (base_geom.r_start_coord / r_end_coord) of nr
 * cells: cell has nr entries, edge nr+1.  Planar only (the reference aborts if spherical). */
int mx_cell_to_edge(const double* cell, double* edge, int nr, int r_lo, int r_hi,
                    void* stream);

/* ---- Maestro::Makew0 ---------------------------------------------------------------------
 * Runtime parameters and base-state geometry read by Maestro::Makew0Planar / Makew0Sphr
 * (Source/MaestroMakew0.cpp:60-272,537-740; base_geom = Source/BaseStateGeometry.H;
 * maestro.* parameters Source/param/_cpp_parameters:309,331,385,391). */
typedef struct MxW0Params {
    int nr_fine;                   /* base_geom.nr_fine */
    int base_cutoff_density_coord; /* base_geom.base_cutoff_density_coord(0) */
    int is_predictor;              /* planar only (:151-163) */
    int add_pb;                    /* planar closed-box pressure correction (:175-184) */
    double dr;                     /* base_geom.dr(0) */
    double dt;                     /* the Maestro member `dt` (planar delta_chi_w0, :103) */
    double dt_in, dtold_in;        /* the dt_in / dtold_in arguments */
    double grav_const, dpdt_factor, base_cutoff_density;
} MxW0Params;

/* Replaces Maestro::Makew0 -> Makew0Planar (Source/MaestroMakew0.cpp:5-58,60-272) for ONE
 * radial level (max_radial_level == 0, one disjoint chunk [0, nr_fine-1]; multilevel
 * restriction / offset :185-213 is the caller's).  Cell-centred inputs have
 * nr_fine entries, w0_old and w0 nr_fine+1.  Writes the Maestro member w0 (all nr_fine+1
 * edges), w0_force (nr_fine cells) and, if add_pb, *p0bdot (HOST pointer, may be NULL; the
 * caller sets p0b = p0bdot * dt, :178-180).  Arrays may be host or device resident. */
int mx_make_w0_planar(const MxW0Params* par, const double* w0_old, const double* Sbar,
                      const double* p0_old, const double* p0_new, const double* gamma1bar_old,
                      const double* gamma1bar_new, const double* p0_minus_peosbar,
                      const double* etarho_cc, double* w0, double* w0_force, double* p0bdot,
                      void* stream);

/* base_geom of a MULTILEVEL planar base state (Source/BaseStateGeometry.H:22-44): radial level n
 * has nr[n] cells of width dr[n] (nr[n] = nr_fine / 2^(max_radial_level-n)) and is defined on
 * numdisjointchunks[n] chunks [r_start_coord(n,j), r_end_coord(n,j)], j = 1 .. chunks; entry (n,j)
 * of the two coordinate tables is at [n*chunk_len + j] (the reference's BaseStateArray layout). */
typedef struct MxRadialGeom {
    int max_radial_level, finest_radial_level, nr_fine;
    const double* dr;                     /* [max_radial_level+1] */
    const int* nr;                        /* [max_radial_level+1] */
    const int* base_cutoff_density_coord; /* [max_radial_level+1] */
    const int* numdisjointchunks;         /* [finest_radial_level+1] */
    const int* r_start_coord;             /* [(finest_radial_level+1) * chunk_len] */
    const int* r_end_coord;
    int chunk_len;
} MxRadialGeom;

/* Maestro::Makew0 -> Makew0Planar for a multilevel base state (Source/MaestroMakew0.cpp:60-272
 * with the level loop, the restriction / offset of :185-213, the zeroing of :216-236 and the
 * RestrictBase / FillGhostBase calls, Source/MaestroBaseStateGeometry.cpp:236-355).  Every 1-D
 * array is the reference's BaseState: (max_radial_level+1) levels x nr_fine cells (w0_old / w0:
 * nr_fine+1 edges), level n at offset n*len.  par->dr / base_cutoff_density_coord are not read
 * (they come per level from rg).  HOST arrays (they are a few kB; staged by the library). */
int mx_make_w0_planar_ml(const MxW0Params* par, const MxRadialGeom* rg, const double* w0_old,
                         const double* Sbar, const double* p0_old, const double* p0_new,
                         const double* gamma1bar_old, const double* gamma1bar_new,
                         const double* p0_minus_peosbar, const double* etarho_cc, double* w0,
                         double* w0_force, double* p0bdot, void* stream);

/* Replaces Maestro::Makew0 -> Makew0Sphr (Source/MaestroMakew0.cpp:537-740) including the
 * calls it makes: MakeGravEdge, spherical branch (Source/MaestroMakeGrav.cpp:303-322) and
 * Tridiag (:942-969).  r_cc_loc (nr_fine) and r_edge_loc (nr_fine+1) are level 0 of
 * base_geom.r_cc_loc / r_edge_loc; etarho_ec has nr_fine+1 entries. */
int mx_make_w0_sphr(const MxW0Params* par, const double* w0_old, const double* Sbar,
                    const double* rho0_old, const double* rho0_new, const double* p0_old,
                    const double* p0_new, const double* gamma1bar_old,
                    const double* gamma1bar_new, const double* p0_minus_peosbar,
                    const double* etarho_cc, const double* etarho_ec, const double* r_cc_loc,
                    const double* r_edge_loc, double* w0, double* w0_force, void* stream);

/* Replaces Maestro::Makew0 -> Makew0SphrIrreg (Source/MaestroMakew0.cpp:742-940; spherical with
 * use_exact_base_state: irregularly spaced radial bins).  Same arguments; par->dr is not read.
 * w0_force(0): the reference evaluates r_edge_loc(0,-1) (an out-of-bounds read, :926); this
 * library uses the first cell's own width there. */
int mx_make_w0_sphr_irreg(const MxW0Params* par, const double* w0_old, const double* Sbar,
                          const double* rho0_old, const double* rho0_new, const double* p0_old,
                          const double* p0_new, const double* gamma1bar_old,
                          const double* gamma1bar_new, const double* p0_minus_peosbar,
                          const double* etarho_cc, const double* etarho_ec,
                          const double* r_cc_loc, const double* r_edge_loc, double* w0,
                          double* w0_force, void* stream);

/* ---- Maestro::UpdateScal ------------------------------------
 * Parameters read by the per-box body of Maestro::UpdateScal
 * (Source/MaestroAdvection.cpp:86-251; dt = the Maestro member, base_cutoff_density and
 * do_eos_h_above_cutoff = Source/param/_cpp_parameters). */
typedef struct MxUpdateParams {
    double dt;
    double base_cutoff_density;
    int do_eos_h_above_cutoff;
    int rho_comp, rhoh_comp, temp_comp, first_spec, num_spec; /* Rho, RhoH, Temp, FirstSpec, NumSpec */
    int start_comp; /* RhoH (enthalpy update) or FirstSpec (species + density update) */
    int num_comp;   /* as passed by the caller; the reference ignores it for both branches */
} MxUpdateParams;

/* Replaces the zone loop of Maestro::UpdateScal for ONE box (2-D or 3-D, no aux variables):
 *   start_comp == FirstSpec: (rho X)_k^new = (rho X)_k^old + dt (-div F_k + force_k), then
 *       rho^new = rho^old + sum_k ((rho X)_k^new - (rho X)_k^old), the density floor at
 *       base_cutoff_density / 2 and the redistribution of negative species (:186-251);
 *   start_comp == RhoH: (rho h)^new = (rho h)^old + dt (-div F + force) (:117-135).
 * stateold / statenew: Nscal components, any (equal or different) ghost widths; sflux[d]:
 * face centred; force: Nscal components.  Only valid zones of statenew are written.
 * The EOS call the reference makes for zones with rho^new <= base_cutoff_density when
 * do_eos_h_above_cutoff (:137-159) is third-party Microphysics code and stays with the
 * caller: those zones are flagged in eos_mask (one byte per valid zone, i fastest, 1 = the
 * caller must overwrite statenew(RhoH) with rho * h_eos; required when the flag is set, may
 * be host or device memory) and counted in *n_eos (HOST pointer, may be NULL).  Reflux /
 * AverageDown / FillPatch at the end of UpdateScal (:257-296) remain the caller's. */
int mx_update_scal(const MxGeom* geom, const MxFab* stateold, const MxFab* statenew,
                   const MxFab* sflux /*[dim]*/, const MxFab* force,
                   const MxUpdateParams* par, unsigned char* eos_mask, long long* n_eos,
                   void* stream);

/* ---- ghost-cell exchange between the bricks of one box (multi-GPU) -----------------------
 * Packs / unpacks the `ng`-deep face slab `side` (0 = lo, 1 = hi) of direction `dir` of
 * components [comp0, comp0+ncomp) of a fab into / from a contiguous device buffer, i.e. the
 * FB_local_copy_gpu step of FabArray::FillBoundary
 * (external/amrex/Src/Base/AMReX_FabArrayCommI.H:49-150); the transfer itself is an NCCL
 * send/recv (or P2P copy) issued by the host layer.  `interior` = 1 packs the valid cells
 * adjacent to the face (what a neighbour needs), 0 unpacks into the ghost cells. */
int mx_halo_pack(const MxGeom* geom, const MxFab* fab, int dir, int side, int ng, int comp0,
                 int ncomp, double* buf, void* stream);
int mx_halo_unpack(const MxGeom* geom, const MxFab* fab, int dir, int side, int ng,
                   int comp0, int ncomp, const double* buf, void* stream);
/* Number of doubles mx_halo_pack writes for these arguments. */
long long mx_halo_count(const MxGeom* geom, const MxFab* fab, int dir, int ng, int ncomp);

/* ---- multi-GPU: the bricks of one level, one per GPU ---------------
 * mx_halo_exchange fills the `ng` ghost layers of components [comp0, comp0+ncomp) of this
 * rank's fab from its face neighbours in a px*py*pz brick grid (x fastest, AMReX box order),
 * direction by direction so that edge and corner ghosts are carried along: what
 * FabArray::FillBoundary does for this path (external/amrex/Src/Base/AMReX_FabArrayCommI.H:8-150;
 * reached from Source/MaestroDensityAdvance.cpp:98-104 / MaestroEnthalpyAdvance.cpp:120-136
 * through FillPatch).  Directions with a single brick are left to the caller (periodic image
 * of the brick itself or physical BC).  periodic[d] != 0: wrap around in direction d;
 * otherwise the bricks on the domain faces have no neighbour there.  Asynchronous on `stream`.
 *
 * One rank per GPU: processes (MPI, torchrun) or host threads of one process.  Rank 0 calls
 * mx_comm_unique_id and hands the 128 bytes to every rank by its own means (MPI_Bcast,
 * torch.distributed, a file); every rank then calls mx_comm_create with its device current.
 * Transports: MX_HALO_NCCL = packed slabs by ncclSend/ncclRecv; MX_HALO_P2P = the sender's pack
 * kernel stores directly into the receiver's slab buffer over NVLink (CUDA IPC / peer access),
 * completion by device-side epoch flags, no host synchronisation. */
typedef struct MxComm MxComm;
#define MX_COMM_ID_BYTES 128
#define MX_HALO_NCCL 0
#define MX_HALO_P2P 1
int mx_comm_unique_id(unsigned char id[MX_COMM_ID_BYTES]);
int mx_comm_create(MxComm** comm, const unsigned char id[MX_COMM_ID_BYTES], int rank, int nranks,
                   const int grid[3], int transport);
int mx_comm_destroy(MxComm* comm);
int mx_halo_exchange(MxComm* comm, const MxGeom* geom, const MxFab* fab, int ng, int comp0,
                     int ncomp, const int periodic[3], void* stream);
/* bytes this rank sent during its last mx_halo_exchange */
long long mx_comm_bytes_sent(const MxComm* comm);


/* ==== the passes either side of the path ==================
 * One box of one level each, like everything above; ghost fills (FillPatch), AverageDown and
 * the AMR tails of the reference members stay with the caller.  Fabs and 1-D arrays may be
 * host or device resident. */

/* Maestro::ConvertRhoXToX / ConvertRhoHToH (Source/MaestroConvert.cpp:64-114): components
 * [start_comp, start_comp+num_comp) of `scal` divided (flag != 0) or multiplied (flag == 0) by
 * component rho_comp on the valid box (MultiFab::Divide / Multiply with 0 ghost cells). */
int mx_convert_by_rho(const MxGeom* geom, const MxFab* scal, int rho_comp, int start_comp,
                      int num_comp, int flag, void* stream);

/* What the 1-D -> 3-D mappings of Source/MaestroFill3dData.cpp read besides their arguments:
 * base_geom (Source/BaseStateGeometry.H: nr_fine, dr_fine, r_cc_loc, r_edge_loc at level 0),
 * the Maestro members `center` and geom[lev].ProbLo, and the maestro.*_interp_type runtime
 * parameters (Source/param/_cpp_parameters:527-546).  Only read when `spherical`. */
typedef struct MxSphr {
    int spherical;
    int nr_fine;
    double dr_fine;
    double prob_lo[3];
    double center[3];
    const double* r_cc_loc;   /* nr_fine */
    const double* r_edge_loc; /* nr_fine + 1 */
    int s0_interp_type, w0_interp_type, s0mac_interp_type, w0mac_interp_type;
} MxSphr;

/* Maestro::Put1dArrayOnCart, per-level body (Source/MaestroFill3dData.cpp:48-337) on the valid
 * box of s0_cart: planar s0_cart(i,j,k,outcomp) = s0(lev,r) or 0.5 (s0(lev,r) + s0(lev,r+1))
 * with r the last index (outcomp = dim-1 if is_output_a_vector else 0); spherical (3-D,
 * !use_exact_base_state) the piecewise-constant / linear / quadratic interpolations in radius
 * selected by w0_interp_type (edge-centred input) or s0_interp_type (cell-centred input), times
 * the unit normal if is_output_a_vector (3 components). */
int mx_put_1d_array_on_cart(const MxGeom* geom, const MxBase* s0, int lev, const MxFab* s0_cart,
                            int is_input_edge_centered, int is_output_a_vector,
                            const MxSphr* sphr, void* stream);

/* Maestro::PutInPertForm (Source/MaestroConvert.cpp:6-62) without the s0_cart temporary:
 * scal(comp) -= s0 mapped onto the box (flag != 0) or += (flag == 0), valid box only; the
 * mapping is Put1dArrayOnCart(s0, ., false, false). */
int mx_put_in_pert_form(const MxGeom* geom, const MxFab* scal, const MxBase* s0, int lev, int comp,
                        int flag, const MxSphr* sphr, void* stream);

/* Maestro::ModifyScalForce (Source/MaestroForce.cpp:261-467), zone loop on the valid box:
 * scal_force(comp) -= scal div(U)                         (fullform == 1)
 * scal_force(comp) += -(scal - s0) div(U) - div(s0 U)     (fullform == 0)
 * umac does NOT contain w0 here.  Planar: s0_cart (1 comp, >= 0 ghosts: only the zone itself is
 * read), w0_cart (the Maestro member, dim comps, 1 ghost) and the 1-D s0_edge, which the
 * reference maps onto a 1-ghost MultiFab with Put1dArrayOnCart(s0_edge, ., false, false, bcs_f):
 * the face above the box's top zone is s0_edge(r+1) inside the domain and the bcs_f fill of it
 * at the top domain face (bcf_hi: MX_BC_FOEXTRAP / MX_BC_REFLECT_EVEN -> s0_edge(domhi),
 * MX_BC_INT_DIR (periodic) -> s0_edge(domlo)).  Spherical (3-D): s0_cart with 1 ghost,
 * div(w0) evaluated from the 1-D w0 (nr_fine+1) as the reference does (divw0 table, then
 * Put1dArrayOnCart(divw0, ., false, false)); w0_cart and s0_edge are not read.
 * s0_edge_cart (optional, planar): the 1-ghost MultiFab itself, as the reference builds it with
 * Put1dArrayOnCart + FillPatch; when given it is read instead of s0_edge / bcf_hi (what an AMR
 * caller passes: at coarse-fine boundaries FillPatch interpolates from the coarser level). */
int mx_modify_scal_force(const MxGeom* geom, const MxFab* scal_force, const MxFab* state,
                         const MxFab* umac /*[dim]*/, const MxBase* s0_edge, int lev,
                         const MxFab* s0_cart, const MxFab* w0_cart, const double* w0,
                         int comp, int fullform, int bcf_hi, const MxFab* s0_edge_cart,
                         const MxSphr* sphr, void* stream);

/* Maestro::Addw0 (Source/MaestroFill3dData.cpp:357-461), the add itself: planar
 * uedge[dim-1](i,j,r) += mult * w0(lev,r) on the nodal valid box grown by one in the transverse
 * directions; spherical uedge[d] += mult * w0mac[d] on the nodal valid boxes.  The ghost fill
 * that follows (FillBoundary / FillUmacGhost) is the caller's. */
int mx_addw0(const MxGeom* geom, const MxFab* uedge /*[dim]*/, const MxFab* w0mac /*[3] or NULL*/,
             const MxBase* w0, int lev, double mult, int spherical, void* stream);

/* Maestro::MakeW0mac (Source/MaestroFill3dData.cpp:464-710; spherical, 3-D): w0mac[d] on the
 * nodal valid box grown by one (w0mac has 1 ghost cell), from w0_cart (w0mac_interp_type 1:
 * the Maestro member, 3 comps, >= 2 ghost cells... the average reads one cell beyond the grown
 * face box) or from the 1-D w0 directly (types 2, 3; type 4 through nodal values). */
int mx_make_w0mac(const MxGeom* geom, const MxFab* w0mac /*[3]*/, const MxFab* w0_cart,
                  const double* w0, const MxSphr* sphr, void* stream);

/* Maestro::MakeS0mac (Source/MaestroFill3dData.cpp:712-1188; spherical, 3-D,
 * !use_exact_base_state): s0mac[d] on the nodal valid box grown by one.  s0mac_interp_type 1
 * averages the cell-centred mapping of s0 (Put1dArrayOnCart + FillPatch with bcs_f in the
 * reference: cells outside the domain take the first-order extrapolation, i.e. the value of the
 * nearest cell inside; evaluated on the fly here, no s0_cart temporary); types 2 and 3
 * interpolate to the face centres directly.  s0_cart (optional, type 1): the 2-ghost MultiFab
 * itself after the caller's FillPatch, read instead of the on-the-fly mapping (AMR callers). */
int mx_make_s0mac(const MxGeom* geom, const MxBase* s0, const MxFab* s0mac /*[3]*/,
                  const MxFab* s0_cart, const MxSphr* sphr, void* stream);

/* Maestro::UpdateVel (Source/MaestroAdvection.cpp:299-526), zone loop on the valid box:
 * unew = uold - dt (Utilde . grad) Utilde + dt force - dt (w0 . grad) Utilde, times the sponge.
 * umac, uedge[d] (dim comps, face centred), force (dim comps), sponge (1 comp, read if
 * do_sponge), w0_cart (planar; dim comps, 1 ghost in the last direction) or w0mac (spherical). */
int mx_update_vel(const MxGeom* geom, const MxFab* uold, const MxFab* unew, const MxFab* umac,
                  const MxFab* uedge, const MxFab* force, const MxFab* sponge,
                  const MxFab* w0_cart, const MxFab* w0mac, double dt, int do_sponge,
                  int spherical, void* stream);

/* ---- MAC-velocity predictor ----------------------------------
 * Runtime parameters of Maestro::MakeUtrans / Maestro::VelPred: dt and rel_eps are Maestro
 * members (Source/Maestro.H:1873,1974), ppm_type / slope_order / ppm_trace_forces / spherical
 * come from Source/param/_cpp_parameters, phys_bc is Maestro::phys_bc (Source/Maestro.H:1917,
 * amrex::PhysBCType codes: 0 interior, 1 inflow, 2 outflow, 3 symmetry, 4 slipwall,
 * 5 noslipwall), bc is Maestro::bcs_u. */
typedef struct MxVelPredParams {
    double dt;
    double rel_eps;
    int ppm_type;         /* 0 (slopes), 1, 2 */
    int slope_order;      /* 0, 2, 4 (used when ppm_type == 0) */
    int ppm_trace_forces; /* VelPred only; 3-D only (the reference's 2-D branch reads Ipfy / Imfy
                             without ever filling them, MaestroVelPred.cpp:134-139) */
    int spherical;
    int phys_bc_lo[3];
    int phys_bc_hi[3];
    const int* bc; /* HOST pointer, dim x 6: {lo_x,lo_y,lo_z,hi_x,hi_y,hi_z} per velocity component */
} MxVelPredParams;

/* Replaces Maestro::MakeUtrans (Source/MaestroMakeUtrans.cpp:6-648, decl Maestro.H:1172-1177)
 * for one box: the transverse velocities on the faces of the valid box from Slopex/y/z or
 * PPM(is_umac = false) of the normal component, phys_bc, and the Riemann problem with the full
 * velocity (w0mac when spherical, w0_cart's last component in the last direction otherwise).
 * utilde: dim components, >= 3 ghost cells (4 if ppm_type == 2); ufull: dim components, >= 1
 * ghost cell; utrans[d]: face-centred in d, written on the faces of the valid box; w0_cart
 * (planar; one ghost cell in the last direction) or w0mac[3] (spherical).  FillBoundary /
 * FillUmacGhost / AverageDownFaces / FillPatchUedge (:630-647) stay with the caller. */
int mx_make_utrans(const MxGeom* geom, const MxFab* utilde, const MxFab* ufull,
                   const MxFab* utrans /*[dim]*/, const MxFab* w0_cart, const MxFab* w0mac,
                   const MxVelPredParams* par, void* stream);

/* Replaces Maestro::VelPred (Source/MaestroVelPred.cpp:5-1917, decl Maestro.H:1711-1842) for one
 * box, including VelPredInterface, VelPredTransverse (3-D) and VelPredVelocities.  utrans[d]
 * (already containing w0, one ghost cell) is read; umac[d] is written on the faces of the valid
 * box; force: dim components, >= 1 ghost cell (utilde's ghost cells if ppm_trace_forces).
 * AverageDownFaces (:293) stays with the caller. */
int mx_vel_pred(const MxGeom* geom, const MxFab* utilde, const MxFab* ufull,
                const MxFab* utrans /*[dim]*/, const MxFab* umac /*[dim]*/, const MxFab* w0mac,
                const MxFab* w0_cart, const MxFab* force, const MxVelPredParams* par, void* stream);

/* ---- fused advances (the update that follows the fluxes) -------
 * mx_density_advance = mx_density_edge_flux + the part of DensityAdvance that follows it
 * (Source/MaestroDensityAdvance.cpp:144-258): rho' -> rho (PutInPertForm, flag false; only for
 * predict_rhoprime_and_X), X -> rho X (ConvertRhoXToX, flag false), scal_force = 0, UpdateScal
 * of the species and the density.  Edge states, fluxes, the zero force and the rebuilt old
 * state live in device workspace and never cross PCIe: a host caller uploads the converted
 * state (X_k, rho' with ghost cells), the prediction force and the MAC velocities and gets back
 * the species / density components of `snew` (valid zones) and `etarhoflux`.
 * species_force_is_zero != 0: only force(Rho) is read (the species components are zero in the
 * reference, MaestroDensityAdvance.cpp:36-39; saves their upload).  sedge / sflux may be NULL;
 * if given (device or host fabs, as in mx_density_edge_flux) they receive the edge states and
 * fluxes as well.  rho0_pert is the base state PutInPertForm adds back (rho0_old); upar as in
 * mx_update_scal with start_comp = FirstSpec.  3-D; results are bit-identical to the separate
 * calls. */
int mx_density_advance(const MxGeom* geom, const MxFab* state, const MxFab* snew,
                       const MxFab* etarhoflux, const MxFab* umac, const MxFab* force,
                       int species_force_is_zero, const MxEdgeParams* epar, int rho_bccomp,
                       const MxBase* rho0_old, const MxBase* rho0_edge_old, const MxFab* r0mac_old,
                       const MxBase* rho0_new, const MxBase* rho0_edge_new, const MxFab* r0mac_new,
                       const MxBase* rho0_predicted_edge, const MxBase* w0, const MxFluxParams* fpar,
                       const MxBase* rho0_pert, const MxSphr* sphr, const MxUpdateParams* upar,
                       const MxFab* sedge /*[3] or NULL*/, const MxFab* sflux /*[3] or NULL*/,
                       void* stream);

/* mx_enthalpy_advance = mx_enthalpy_edge_flux + (rho h)' -> (rho h) (PutInPertForm with
 * rhoh0_pert = rhoh0_old, enthalpy_pred_type predict_rhohprime; nothing for predict_rhoh) +
 * UpdateScal of RhoH with `force_update` (the time-centred force of
 * Source/MaestroEnthalpyAdvance.cpp:316-336, RhoH component).  `snew` must hold the NEW density
 * in its Rho component (it decides which zones need the caller's EOS, see mx_update_scal:
 * eos_mask / n_eos).  Planar base state, 3-D, enthalpy_pred_type 0 or 1.  sedge must be given
 * when the flux form reads sedge(Rho) (never for types 0 / 1) - pass NULL otherwise. */
int mx_enthalpy_advance(const MxGeom* geom, const MxFab* state, const MxFab* snew,
                        const MxFab* umac, const MxFab* force, const MxFab* force_update,
                        const MxEdgeParams* epar, const MxBase* rho0_old, const MxBase* rho0_edge_old,
                        const MxBase* rho0_new, const MxBase* rho0_edge_new, const MxBase* rhoh0_old,
                        const MxBase* rhoh0_edge_old, const MxBase* rhoh0_new,
                        const MxBase* rhoh0_edge_new, const MxFluxParams* fpar,
                        const MxBase* rhoh0_pert, const MxUpdateParams* upar,
                        unsigned char* eos_mask, long long* n_eos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MXADV_H */
