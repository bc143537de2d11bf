#!/bin/bash
# TEST INFRASTRUCTURE (not product code): builds the whole-code parity check of north_star —
# "the fcompare diff of plotfiles after 10 full time steps must stay within 1e-10".
#
# Two executables of the reference problem Exec/test_problems/reacting_bubble are built with the
# reference's own GNU-make system, CPU build (USE_MPI=FALSE USE_OMP=TRUE COMP=gnu), from the
# sources where they lie under /root/reference (the problem directory is copied to a scratch
# path because the reference tree is read-only and make writes tmp_build_dir/ into the cwd):
#   Maestro<D>d.ref.ex  the UNMODIFIED reference
#   Maestro<D>d.mx.ex   the same build with maestroex_b200/overlay/ shadowing
#                       MaestroMakeEdgeScalars.cpp, MaestroMakeFlux.cpp, MaestroCelltoEdge.cpp,
#                       MaestroMakew0.cpp and Maestro::UpdateScal, linked against libmxadv.so
#                       exactly as INTEGRATION.md section 1 / 2b tells a maintainer to
# plus AMReX's fcompare (external/amrex/Tools/Plotfile).  Outputs go only to oracle/_ref/e2e/
# (git-ignored; travels to the GPU box, where /root/reference does not exist).  Nothing is
# copied from the reference into the repository history.
#
# usage: oracle/e2e/build_e2e.sh [2|3] [reacting_bubble|wdconvect]       (default 2, reacting_bubble)
# wdconvect (Exec/science/wdconvect, 3-D only): spherical base state, outflow on all six faces,
# BASELINE config 3; executables are named Maestro3d.wd.{ref,mx,fx}.ex.
set -euo pipefail
DIM=${1:-2}
PROB=${2:-reacting_bubble}
if [ "$PROB" = wdconvect ]; then PDIR=Exec/science/wdconvect; TAGP=wd_; SUF=.wd; else PDIR=Exec/test_problems/reacting_bubble; TAGP=rb_; SUF=; fi
REF=${MAESTROEX_REF:-/root/reference}
REPO=$(cd "$(dirname "$0")/../.." && pwd)
OUT=$REPO/oracle/_ref/e2e
WORK=${E2E_WORK:-/tmp/e2e}
J=${E2E_JOBS:-$(nproc)}
MK="make -j$J MAESTROEX_HOME=$REF DIM=$DIM USE_MPI=FALSE USE_OMP=TRUE COMP=gnu"
mkdir -p "$OUT" "$WORK"

# 1. the unmodified reference
if [ ! -d "$WORK/${TAGP}ref$DIM" ]; then
    cp -r "$REF/$PDIR" "$WORK/${TAGP}ref$DIM"
fi
(cd "$WORK/${TAGP}ref$DIM" && $MK > build.log 2>&1)
cp "$WORK/${TAGP}ref$DIM/Maestro${DIM}d.gnu.OMP.ex" "$OUT/Maestro${DIM}d$SUF.ref.ex"

# 2. the overlay build: start from the reference's object files, drop the shadowed ones
make -C "$REPO/maestroex_b200/csrc" > /dev/null
rm -rf "$WORK/${TAGP}mx$DIM"
cp -r "$WORK/${TAGP}ref$DIM" "$WORK/${TAGP}mx$DIM"
cd "$WORK/${TAGP}mx$DIM"
rm -f Maestro${DIM}d.gnu.OMP.ex
for f in MaestroMakeEdgeScalars MaestroMakeFlux MaestroCelltoEdge MaestroMakew0 MaestroAdvection; do
    find tmp_build_dir \( -name "$f.o" -o -name "$f.d" \) -delete
done
python3 - "$REPO" <<'EOF'
import sys
repo = sys.argv[1]
mk = open("GNUmakefile").read()
inc = "include $(MAESTROEX_HOME)/Exec/Make.Maestro"
add = f"""MXADV_HOME := {repo}
Bpack   += $(MXADV_HOME)/maestroex_b200/overlay/Make.package
Blocs   += $(MXADV_HOME)/maestroex_b200/overlay $(MXADV_HOME)/include
LIBRARIES         += -lmxadv
LIBRARY_LOCATIONS += $(MXADV_HOME)/maestroex_b200/csrc
CEXE_sources += MaestroUpdateScal.cpp
"""
tail = """
$(objEXETempDir)/MaestroAdvection.o: CPPFLAGS += -DUpdateScal=UpdateScal_reference
"""
assert inc in mk
open("GNUmakefile", "w").write(mk.replace(inc, add + inc + tail))
EOF
$MK > build_mx.log 2>&1
cp "Maestro${DIM}d.gnu.OMP.ex" "$OUT/Maestro${DIM}d$SUF.mx.ex"

# 2b. (3-D) the fused overlay on top: DensityAdvance / EnthalpyAdvance shadowed as well, so the
#     run goes through mx_density_edge_flux / mx_enthalpy_edge_flux, the kernels bench.py times
if [ "$DIM" = 3 ]; then
    rm -rf "$WORK/${TAGP}fx$DIM"
    cp -r "$WORK/${TAGP}mx$DIM" "$WORK/${TAGP}fx$DIM"
    cd "$WORK/${TAGP}fx$DIM"
    rm -f Maestro${DIM}d.gnu.OMP.ex
    for f in MaestroDensityAdvance MaestroEnthalpyAdvance; do
        find tmp_build_dir \( -name "$f.o" -o -name "$f.d" \) -delete
    done
    python3 - "$REPO" <<'EOF2'
import sys
repo = sys.argv[1]
mk = open("GNUmakefile").read()
key = "LIBRARIES         += -lmxadv"
add = """Bpack   += $(MXADV_HOME)/maestroex_b200/overlay_fused/Make.package
Blocs   += $(MXADV_HOME)/maestroex_b200/overlay_fused
"""
assert key in mk
open("GNUmakefile", "w").write(mk.replace(key, add + key))
EOF2
    $MK > build_fx.log 2>&1
    cp "Maestro${DIM}d.gnu.OMP.ex" "$OUT/Maestro${DIM}d$SUF.fx.ex"
fi

# 2c. the companion overlay on top: ConvertRhoXToX / ConvertRhoHToH /
#     PutInPertForm, ModifyScalForce, Put1dArrayOnCart / Addw0 / MakeW0mac / MakeS0mac, UpdateVel,
#     MakeUtrans and VelPred on
#     libmxadv.so.  They share translation units with members that stay the reference's, so those
#     reference files are compiled with the overlaid members renamed.
if [ "$DIM" = 3 ]; then BASE=fx; else BASE=mx; fi
rm -rf "$WORK/${TAGP}cx$DIM"
cp -r "$WORK/${TAGP}${BASE}$DIM" "$WORK/${TAGP}cx$DIM"
cd "$WORK/${TAGP}cx$DIM"
rm -f Maestro${DIM}d.gnu.OMP.ex
for f in MaestroConvert MaestroForce MaestroFill3dData MaestroAdvection MaestroMakeUtrans MaestroVelPred; do
    find tmp_build_dir \( -name "$f.o" -o -name "$f.d" \) -delete
done
python3 - "$REPO" <<'EOF3'
import sys
repo = sys.argv[1]
mk = open("GNUmakefile").read()
key = "LIBRARIES         += -lmxadv"
add = """Bpack   += $(MXADV_HOME)/maestroex_b200/overlay_companions/Make.package
Blocs   += $(MXADV_HOME)/maestroex_b200/overlay_companions
"""
tail = """
$(objEXETempDir)/MaestroAdvection.o: CPPFLAGS += -DUpdateVel=UpdateVel_reference
$(objEXETempDir)/MaestroConvert.o: CPPFLAGS += -DConvertRhoXToX=ConvertRhoXToX_reference -DConvertRhoHToH=ConvertRhoHToH_reference -DPutInPertForm=PutInPertForm_reference
$(objEXETempDir)/MaestroForce.o: CPPFLAGS += -DModifyScalForce=ModifyScalForce_reference
$(objEXETempDir)/MaestroFill3dData.o: CPPFLAGS += -DPut1dArrayOnCart=Put1dArrayOnCart_reference -DAddw0=Addw0_reference -DMakeW0mac=MakeW0mac_reference -DMakeS0mac=MakeS0mac_reference
$(objEXETempDir)/MaestroMakeUtrans.o: CPPFLAGS += -DMakeUtrans=MakeUtrans_reference
$(objEXETempDir)/MaestroVelPred.o: CPPFLAGS += -DVelPred=VelPred_reference
"""
assert key in mk
open("GNUmakefile", "w").write(mk.replace(key, add + key) + tail)
EOF3
$MK > build_cx.log 2>&1 || { tail -30 build_cx.log; exit 1; }
cp "Maestro${DIM}d.gnu.OMP.ex" "$OUT/Maestro${DIM}d$SUF.cx.ex"

# 3. fcompare
if [ ! -d "$WORK/plotfile" ]; then cp -r "$REF/external/amrex/Tools/Plotfile" "$WORK/plotfile"; fi
(cd "$WORK/plotfile" && make -j$J AMREX_HOME=$REF/external/amrex DIM=$DIM USE_MPI=FALSE COMP=gnu \
    programs=fcompare > build.log 2>&1)
cp "$WORK/plotfile/fcompare.gnu.ex" "$OUT/fcompare${DIM}d.ex"
# 3b. fextrema (the tool of the reference's CI goldens, .github/workflows/*-compare.yml); the
#     executable name does not carry the dimension, so force the relink
rm -f "$WORK/plotfile/fextrema.gnu.ex"
(cd "$WORK/plotfile" && make -j$J AMREX_HOME=$REF/external/amrex DIM=$DIM USE_MPI=FALSE COMP=gnu \
    programs=fextrema > build_fextrema.log 2>&1)
cp "$WORK/plotfile/fextrema.gnu.ex" "$OUT/fextrema${DIM}d.ex"

# 4. run-time data of the problem (inputs, initial model, EOS table; the table is 60 MB of
#    text, shipped gzipped)
if [ "$PROB" = wdconvect ]; then
    cp "$WORK/${TAGP}ref$DIM"/inputs_files/inputs_3d_regression "$OUT/inputs_3d_wdconvect"
    cp "$WORK/${TAGP}ref$DIM"/inputs_files/inputs_3d_regression.2levels "$OUT/inputs_3d_wdconvect_2levels"
    cp "$WORK/${TAGP}ref$DIM"/model_files/kepler_new_6.25e8.hybrid.hse.320 "$OUT/"
else
    cp "$WORK/${TAGP}ref$DIM"/inputs_${DIM}d_C "$WORK/${TAGP}ref$DIM"/model.hse.cool.coulomb "$OUT/"
    [ "$DIM" = 2 ] && cp "$WORK/${TAGP}ref$DIM"/inputs_2d_regression "$OUT/"
fi
[ -f "$OUT/helm_table.dat.gz" ] || gzip -c "$REF/external/Microphysics/EOS/helmholtz/helm_table.dat" > "$OUT/helm_table.dat.gz"
strip "$OUT"/*.ex
ls -la "$OUT"
