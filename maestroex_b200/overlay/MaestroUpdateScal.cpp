// Overlay for Maestro::UpdateScal (Source/MaestroAdvection.cpp:86-297, decl Source/Maestro.H)
// on libmxadv.so (mx_update_scal).  UpdateScal shares its translation unit with AdvancePremac
// and UpdateVel, which stay the reference's: build the reference's MaestroAdvection.cpp with
// `-DUpdateScal=UpdateScal_reference` and add this file to
// CEXE_sources.  The zone loop runs on the GPU; the EOS call for zones at or below
// base_cutoff_density (:137-159, Microphysics) is made here, by the reference's own eos(), on
// the zones the library flags; the AMR / ghost-fill tail (:257-296) is unchanged.
#include <MxadvAmrex.H>

// mx_update_scal carries no auxiliary composition variables: the reference updates them in
// the species branch (Source/MaestroAdvection.cpp:203-211) and fills eos_state.aux before
// eos() (:148-153).  Refuse to build against a network that has them instead of silently
// producing a wrong statenew.
#if defined(NAUX_NET) && NAUX_NET > 0
#error "maestroex_b200 overlay: UpdateScal does not handle NAUX_NET > 0 (auxiliary variables); build this problem with Source/MaestroAdvection.cpp"
#endif

using namespace amrex;

void Maestro::UpdateScal(const Vector<MultiFab>& stateold, Vector<MultiFab>& statenew,
                         const Vector<std::array<MultiFab, AMREX_SPACEDIM> >& sflux,
                         const Vector<MultiFab>& force, int start_comp, int num_comp,
                         const Vector<MultiFab>& p0_cart) {
    BL_PROFILE_VAR("Maestro::UpdateScal()", UpdateScal);

    if (start_comp != RhoH && start_comp != FirstSpec) {
        Abort("Invalid scalar in UpdateScal().");
    }
    MxUpdateParams par;
    par.dt = dt;
    par.base_cutoff_density = base_cutoff_density;
    par.do_eos_h_above_cutoff = do_eos_h_above_cutoff ? 1 : 0;
    par.rho_comp = Rho;
    par.rhoh_comp = RhoH;
    par.temp_comp = Temp;
    par.first_spec = FirstSpec;
    par.num_spec = NumSpec;
    par.start_comp = start_comp;
    par.num_comp = num_comp;

    for (int lev = 0; lev <= finest_level; ++lev) {
#ifdef AMREX_USE_GPU
        // zone flags of all local boxes of this level in one device allocation (1 byte per
        // zone): consecutive MFIter iterations run on different AMReX streams, so the boxes
        // must not share a buffer, and a per-box allocation would cost a synchronisation each
        const bool want_mask = (start_comp == RhoH) && do_eos_h_above_cutoff;
        std::vector<Long> mask_off;
        Long mask_total = 0;
        if (want_mask) {
            for (MFIter mfi(stateold[lev], false); mfi.isValid(); ++mfi) {
                mask_off.push_back(mask_total);
                mask_total += mfi.validbox().numPts();
            }
        }
        Gpu::DeviceVector<unsigned char> dmask(mask_total);
        int ibox = 0;
#endif
        for (MFIter mfi(stateold[lev], false); mfi.isValid(); ++mfi) {
            const Box& vbx = mfi.validbox();
            const MxGeom g = mxadv::geom_of(geom[lev], vbx);
            const MxFab so = mxadv::fab_of(stateold[lev].const_array(mfi));
            const MxFab sn = mxadv::fab_of(statenew[lev].array(mfi));
            const MxFab fo = mxadv::fab_of(force[lev].const_array(mfi));
            const mxadv::Faces sf = mxadv::faces_of(sflux[lev], mfi);
            const bool enth = start_comp == RhoH;
#ifdef AMREX_USE_GPU
            // GPU AMReX: fabs are device memory.  The zone flags stay on the device and the
            // reference's own eos() runs there, on the flagged zones, in AMReX's stream
            // (same statements as :137-159)
            unsigned char* mk_box = want_mask ? dmask.data() + mask_off[ibox++] : nullptr;
            mxadv::check(mx_update_scal(&g, &so, &sn, sf.f, &fo, &par, mk_box, nullptr,
                                        mxadv::stream()),
                         "UpdateScal");
            if (mk_box != nullptr) {
                const Array4<const Real> sold_arr = stateold[lev].const_array(mfi);
                const Array4<Real> snew_arr = statenew[lev].array(mfi);
                const Array4<const Real> p0_arr = p0_cart[lev].const_array(mfi);
                const unsigned char* AMREX_RESTRICT mk = mk_box;
                const Dim3 vlo = amrex::lbound(vbx);
                const Dim3 vln = amrex::length(vbx);
                ParallelFor(vbx, [=] AMREX_GPU_DEVICE(int i, int j, int k) {
                    const long q = (long)(i - vlo.x) +
                                   (long)vln.x * ((long)(j - vlo.y) + (long)vln.y * (k - vlo.z));
                    if (!mk[q]) return;
                    eos_rh_t eos_state;
                    eos_state.rho = snew_arr(i, j, k, Rho);
                    eos_state.T = sold_arr(i, j, k, Temp);
                    eos_state.p = p0_arr(i, j, k);
                    for (auto n = 0; n < NumSpec; ++n) {
                        eos_state.xn[n] = snew_arr(i, j, k, FirstSpec + n) / eos_state.rho;
                    }
                    eos(eos_input_rp, eos_state);
                    snew_arr(i, j, k, RhoH) = snew_arr(i, j, k, Rho) * eos_state.h;
                });
            }
#else
            std::vector<unsigned char> mask;
            long long n_eos = 0;
            if (enth && do_eos_h_above_cutoff) mask.assign(vbx.numPts(), 0);
            mxadv::check(mx_update_scal(&g, &so, &sn, sf.f, &fo, &par,
                                        mask.empty() ? nullptr : mask.data(), &n_eos,
                                        mxadv::stream()),
                         "UpdateScal");
            if (n_eos > 0) {
                // :137-159 on the flagged zones, with the reference's own EOS
                const Array4<const Real> sold_arr = stateold[lev].const_array(mfi);
                const Array4<Real> snew_arr = statenew[lev].array(mfi);
                const Array4<const Real> p0_arr = p0_cart[lev].const_array(mfi);
#if (AMREX_SPACEDIM == 3)
                const int klo = vbx.smallEnd(2), khi = vbx.bigEnd(2);
#else
                const int klo = 0, khi = 0;
#endif
                long q = 0;
                for (int k = klo; k <= khi; ++k) {
                    for (int j = vbx.smallEnd(1); j <= vbx.bigEnd(1); ++j) {
                        for (int i = vbx.smallEnd(0); i <= vbx.bigEnd(0); ++i, ++q) {
                            if (!mask[q]) continue;
                            eos_rh_t eos_state;
                            eos_state.rho = snew_arr(i, j, k, Rho);
                            eos_state.T = sold_arr(i, j, k, Temp);
                            eos_state.p = p0_arr(i, j, k);
                            for (auto n = 0; n < NumSpec; ++n) {
                                eos_state.xn[n] = snew_arr(i, j, k, FirstSpec + n) / eos_state.rho;
                            }
                            eos(eos_input_rp, eos_state);
                            snew_arr(i, j, k, RhoH) = snew_arr(i, j, k, Rho) * eos_state.h;
                        }
                    }
                }
            }
#endif
        }
#ifdef AMREX_USE_GPU
        if (want_mask) Gpu::streamSynchronizeAll();  // dmask goes out of scope
#endif
    }

    // :257-296, unchanged: AMR synchronisation and ghost fill
    if (reflux_type == 2) {
        for (int lev = finest_level - 1; lev >= 0; --lev) {
            flux_reg_s[lev + 1]->Reflux(statenew[lev], 1.0, start_comp, start_comp, num_comp,
                                        geom[lev]);
            if (start_comp == FirstSpec) {
                flux_reg_s[lev + 1]->Reflux(statenew[lev], 1.0, Rho, Rho, 1, geom[lev]);
            }
        }
    }
    AverageDown(statenew, start_comp, num_comp);
    FillPatch(t_old, statenew, statenew, statenew, start_comp, start_comp, num_comp, start_comp,
              bcs_s);
    if (start_comp == FirstSpec) {
        AverageDown(statenew, Rho, 1);
        FillPatch(t_old, statenew, statenew, statenew, Rho, Rho, 1, Rho, bcs_s);
    }
}
