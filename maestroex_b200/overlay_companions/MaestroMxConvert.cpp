// Overlay for Maestro::ConvertRhoXToX / ConvertRhoHToH / PutInPertForm
// (Source/MaestroConvert.cpp:6-114) on libmxadv.so: the zone arithmetic runs on the GPU
// (mx_convert_by_rho, mx_put_in_pert_form - the latter without the s0_cart temporary), the
// AverageDown / FillPatch tails are the reference's own calls, unchanged.
// Build: compile the reference's MaestroConvert.cpp with
//   -DConvertRhoXToX=ConvertRhoXToX_reference -DConvertRhoHToH=ConvertRhoHToH_reference
//   -DPutInPertForm=PutInPertForm_reference
// and add this file to CEXE_sources.
#include <MxadvSphr.H>

using namespace amrex;

namespace {
void convert(const Vector<Geometry>& gms, int finest, Vector<MultiFab>& scal, int c0, int nc, bool flag,
             const char* who) {
    for (int lev = 0; lev <= finest; ++lev) {
        for (MFIter mfi(scal[lev], false); mfi.isValid(); ++mfi) {
            const MxGeom g = mxadv::geom_of(gms[lev], mfi.validbox());
            const MxFab f = mxadv::fab_of(scal[lev].array(mfi));
            mxadv::check(mx_convert_by_rho(&g, &f, Rho, c0, nc, flag ? 1 : 0, mxadv::stream()), who);
        }
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
}
template <class Center>
void pert(const Geometry& gm, const BaseStateGeometry& bg, const Center& center, int lev, MultiFab& scal,
          const BaseState<Real>& s0, int comp, bool flag) {
    mxadv::no_exact_base_state("PutInPertForm");
    const MxSphr sp = mxadv::sphr_of(gm, bg, center);
    const MxBase b = mxadv::base_of(s0);
    for (MFIter mfi(scal, false); mfi.isValid(); ++mfi) {
        const MxGeom g = mxadv::geom_of(gm, mfi.validbox());
        const MxFab f = mxadv::fab_of(scal.array(mfi));
        mxadv::check(mx_put_in_pert_form(&g, &f, &b, lev, comp, flag ? 1 : 0, &sp, mxadv::stream()),
                     "PutInPertForm");
    }
#ifdef AMREX_USE_CUDA
    Gpu::streamSynchronize();
#endif
}
}  // namespace

void Maestro::PutInPertForm(Vector<MultiFab>& scal, const BaseState<Real>& s0, int comp, int bccomp,
                            const Vector<BCRec>& bcs, bool flag) {
    BL_PROFILE_VAR("Maestro::PutInPertForm()", PutInPertForm);
    for (int lev = 0; lev <= finest_level; ++lev) pert(geom[lev], base_geom, center, lev, scal[lev], s0, comp, flag);
    AverageDown(scal, comp, 1);
    FillPatch(t_old, scal, scal, scal, comp, comp, 1, bccomp, bcs);  // :33
}

void Maestro::PutInPertForm(int level, Vector<MultiFab>& scal, const BaseState<Real>& s0, int comp,
                            int bccomp, const Vector<BCRec>& bcs, bool flag) {
    BL_PROFILE_VAR("Maestro::PutInPertForm_lev()", PutInPertForm);
    pert(geom[level], base_geom, center, level, scal[level], s0, comp, flag);
    if (level > 0) {
        average_down(scal[level], scal[level - 1], geom[level], geom[level - 1], comp, 1,
                     refRatio(level - 1));
    }
    FillPatch(level, t_old, scal[level], scal, scal, comp, comp, 1, bccomp, bcs);  // :61
}

void Maestro::ConvertRhoXToX(Vector<MultiFab>& scal, bool flag) {
    BL_PROFILE_VAR("Maestro::ConvertRhoXToX()", ConvertRhoXToX);
    convert(geom, finest_level, scal, FirstSpec, NumSpec, flag, "ConvertRhoXToX");
    AverageDown(scal, FirstSpec, NumSpec);
    if (flag) {
        FillPatch(t_old, scal, scal, scal, FirstSpec, FirstSpec, NumSpec, 0, bcs_f);
    } else {
        FillPatch(t_old, scal, scal, scal, FirstSpec, FirstSpec, NumSpec, FirstSpec, bcs_s);
    }
}

void Maestro::ConvertRhoHToH(Vector<MultiFab>& scal, bool flag) {
    BL_PROFILE_VAR("Maestro::ConvertRhoHToH()", ConvertRhoHToH);
    convert(geom, finest_level, scal, RhoH, 1, flag, "ConvertRhoHToH");
    AverageDown(scal, RhoH, 1);
    if (flag) {
        FillPatch(t_old, scal, scal, scal, RhoH, RhoH, 1, 0, bcs_f);
    } else {
        FillPatch(t_old, scal, scal, scal, RhoH, RhoH, 1, RhoH, bcs_s);
    }
}
