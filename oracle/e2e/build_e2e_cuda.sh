#!/bin/bash
# TEST / MEASUREMENT INFRASTRUCTURE (not product code).  GPU-resident variant of build_e2e.sh:
# the reference problem Exec/test_problems/reacting_bubble (3-D) built with the reference's own
# GNU-make system as a CUDA application for sm_100 (USE_CUDA=TRUE CUDA_ARCH=100
# TINY_PROFILE=TRUE, no MPI / OpenMP), twice:
#   Maestro3d.cuda.ref.ex  the UNMODIFIED reference: its own ParallelFor lambdas on the B200
#   Maestro3d.cuda.cx.ex   mx + overlay_fused + overlay_companions (every layer of INTEGRATION.md)
#   Maestro3d.cuda.mx.ex   the same build with maestroex_b200/overlay/ shadowing the hot path,
#                          linked against libmxadv.so; fabs are AMReX device memory, so the
#                          library runs on them in place on amrex::Gpu::gpuStream()
# Used by oracle/e2e/run_cuda_compare.sh for (i) the in-application timing of the hot path
# against the reference's own GPU implementation on the same B200 (TinyProfiler, regions
# synchronised) and (ii) an fcompare of the two runs.  Outputs go to oracle/_ref/e2e/ only.
#
# TinyProfiler's CUDA build includes <nvToolsExt.h> (NVTX v2), which CUDA 12.9 no longer ships;
# a one-line forwarding header to the header-only NVTX v3 is put on the include path, and
# -lnvToolsExt resolves to the copy bundled with the image's python environment.
set -euo pipefail
REF=${MAESTROEX_REF:-/root/reference}
REPO=$(cd "$(dirname "$0")/../.." && pwd)
OUT=$REPO/oracle/_ref/e2e
WORK=${E2E_WORK:-/tmp/e2e}
J=${E2E_JOBS:-$(nproc)}
NVTXLIB=${NVTXLIB:-/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nvtx/lib/libnvToolsExt.so.1}
MK="make -j$J MAESTROEX_HOME=$REF DIM=3 USE_MPI=FALSE USE_OMP=FALSE USE_CUDA=TRUE CUDA_ARCH=100 TINY_PROFILE=TRUE COMP=gnu"
mkdir -p "$OUT" "$WORK/nvtx/include" "$WORK/nvtx/lib"
# (written once: touching it would make every TinyProfiler user out of date)
[ -f "$WORK/nvtx/include/nvToolsExt.h" ] || \
    printf '#pragma once\n#include <nvtx3/nvToolsExt.h>\n' > "$WORK/nvtx/include/nvToolsExt.h"
ln -sf "$NVTXLIB" "$WORK/nvtx/lib/libnvToolsExt.so"

# 1. the unmodified reference
if [ ! -d "$WORK/rb_cuda3" ]; then
    cp -r "$REF/Exec/test_problems/reacting_bubble" "$WORK/rb_cuda3"
    python3 - "$WORK" <<'EOF'
import sys
work = sys.argv[1]
p = f"{work}/rb_cuda3/GNUmakefile"
mk = open(p).read()
inc = "include $(MAESTROEX_HOME)/Exec/Make.Maestro"
add = f"INCLUDE_LOCATIONS += {work}/nvtx/include\nLIBRARY_LOCATIONS += {work}/nvtx/lib\n"
open(p, "w").write(mk.replace(inc, add + inc))
EOF
fi
(cd "$WORK/rb_cuda3" && $MK > build.log 2>&1)
cp "$WORK/rb_cuda3/Maestro3d.gnu.TPROF.CUDA.ex" "$OUT/Maestro3d.cuda.ref.ex"

# 2. the overlay build
make -C "$REPO/maestroex_b200/csrc" > /dev/null
rm -rf "$WORK/rb_cudamx3"
cp -r "$WORK/rb_cuda3" "$WORK/rb_cudamx3"
cd "$WORK/rb_cudamx3"
rm -f Maestro3d.gnu.TPROF.CUDA.ex
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
open("GNUmakefile", "w").write(mk.replace(inc, add + inc + tail))
EOF
$MK > build_mx.log 2>&1
cp Maestro3d.gnu.TPROF.CUDA.ex "$OUT/Maestro3d.cuda.mx.ex"
# 3. the fused overlay on top (DensityAdvance / EnthalpyAdvance on the fused kernels)
rm -rf "$WORK/rb_cudafx3"
cp -r "$WORK/rb_cudamx3" "$WORK/rb_cudafx3"
cd "$WORK/rb_cudafx3"
rm -f Maestro3d.gnu.TPROF.CUDA.ex
for f in MaestroDensityAdvance MaestroEnthalpyAdvance; do
    find tmp_build_dir \( -name "$f.o" -o -name "$f.d" \) -delete
done
python3 - <<'EOF'
mk = open("GNUmakefile").read()
key = "LIBRARIES         += -lmxadv"
add = """Bpack   += $(MXADV_HOME)/maestroex_b200/overlay_fused/Make.package
Blocs   += $(MXADV_HOME)/maestroex_b200/overlay_fused
"""
open("GNUmakefile", "w").write(mk.replace(key, add + key))
EOF
$MK > build_fx.log 2>&1
cp Maestro3d.gnu.TPROF.CUDA.ex "$OUT/Maestro3d.cuda.fx.ex"
# 4. the companion overlay on top (SURVEY.md section 8f ranks 2-4 incl. MakeUtrans / VelPred), as
#    step 2c of build_e2e.sh
rm -rf "$WORK/rb_cudacx3"
cp -r "$WORK/rb_cudafx3" "$WORK/rb_cudacx3"
cd "$WORK/rb_cudacx3"
rm -f Maestro3d.gnu.TPROF.CUDA.ex
for f in MaestroConvert MaestroForce MaestroFill3dData MaestroAdvection MaestroMakeUtrans MaestroVelPred; do
    find tmp_build_dir \( -name "$f.o" -o -name "$f.d" \) -delete
done
python3 - <<'EOF'
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
EOF
$MK > build_cx.log 2>&1 || { tail -30 build_cx.log; exit 1; }
cp Maestro3d.gnu.TPROF.CUDA.ex "$OUT/Maestro3d.cuda.cx.ex"
cp "$WORK/rb_cuda3/inputs_3d_C" "$WORK/rb_cuda3/model.hse.cool.coulomb" "$OUT/"
[ -f "$OUT/helm_table.dat.gz" ] || gzip -c "$REF/external/Microphysics/EOS/helmholtz/helm_table.dat" > "$OUT/helm_table.dat.gz"
strip "$OUT"/Maestro3d.cuda.*.ex
for f in "$OUT"/Maestro3d.cuda.*.ex; do xz -T0 -3 -f "$f"; done  # 165 MB -> 36 MB each (gpurun snapshot limit)
ls -la "$OUT"
