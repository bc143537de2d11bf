#!/bin/bash
# MEASUREMENT SCRIPT for the GPU box (reads nothing under /root/reference).  Runs the two CUDA
# executables built by build_e2e_cuda.sh — the unmodified reference and the overlay build on
# libmxadv.so — on the same B200 for NSTEPS steps of reacting_bubble inputs_3d_C, then
#   * fcompare of the two step-NSTEPS plotfiles, and
#   * the TinyProfiler inclusive times of the hot-path members in both runs
#     (tiny_profiler.device_synchronize_around_region=1, so a region's time is the device time
#     of the work it launched).
# amrex.the_arena_is_managed=1: the reference fills physical boundaries on the host
# (Maestro::ScalarFill -> amrex::fab_filcc, Source/MaestroBCFill.cpp:17), which segfaults on
# AMReX 24.08's default device-only arena; with managed fabs the unmodified reference runs.
# usage: run_cuda_compare.sh OUTDIR TAG [extra ParmParse arguments...]
set -uo pipefail
OUTDIR=$1; TAG=$2; shift 2
REPO=$(cd "$(dirname "$0")/../.." && pwd)
E2E=$REPO/oracle/_ref/e2e
NSTEPS=${NSTEPS:-10}
ARMS=${ARMS:-"ref mx fx"}
NVTXDIR=/opt/prime-rl/.venv/lib/python3.12/site-packages/nvidia/nvtx/lib
export LD_LIBRARY_PATH=$REPO/maestroex_b200/csrc:$NVTXDIR:${LD_LIBRARY_PATH:-}
mkdir -p "$OUTDIR"
OUTDIR=$(cd "$OUTDIR" && pwd)
TMP=$(mktemp -d)
gunzip -c "$E2E/helm_table.dat.gz" > "$TMP/helm_table.dat"
for arm in $ARMS; do
    mkdir -p "$TMP/$arm"
    # the CUDA executables travel xz-compressed (165 MB each otherwise: gpurun's snapshot limit)
    EXE="$E2E/Maestro3d.cuda.$arm.ex"
    if [ ! -x "$EXE" ] && [ -f "$EXE.xz" ]; then xz -dc "$EXE.xz" > "$TMP/Maestro3d.cuda.$arm.ex"; chmod +x "$TMP/Maestro3d.cuda.$arm.ex"; EXE="$TMP/Maestro3d.cuda.$arm.ex"; fi
    cp "$E2E/inputs_3d_C" "$E2E/model.hse.cool.coulomb" "$TMP/$arm/"
    ln -s "$TMP/helm_table.dat" "$TMP/$arm/helm_table.dat"
    (cd "$TMP/$arm" && timeout 900 "$EXE" inputs_3d_C \
        maestro.max_step=$NSTEPS maestro.plot_int=$NSTEPS maestro.chk_int=-1 maestro.plot_deltat=-1 \
        tiny_profiler.device_synchronize_around_region=1 amrex.fpe_trap_invalid=0 \
        amrex.the_arena_is_managed=1 amrex.the_arena_init_size=4294967296 "$@" \
        > "$OUTDIR/run_${TAG}_$arm.log" 2>&1; echo "rc=$?" >> "$OUTDIR/run_${TAG}_$arm.log")
    tail -n 1 "$OUTDIR/run_${TAG}_$arm.log"
done
PLT=$(printf 'plt%07d' "$NSTEPS")
LAST=ref
for arm in $ARMS; do  # reference build against each overlay build; overlay builds against each other
    [ "$arm" = ref ] && continue
    "$E2E/fcompare3d.ex" -a "$TMP/ref/$PLT" "$TMP/$arm/$PLT" > "$OUTDIR/fcompare_cuda_${TAG}_ref_vs_$arm.txt" 2>&1
    tail -n 1 "$OUTDIR/fcompare_cuda_${TAG}_ref_vs_$arm.txt"
    if [ "$LAST" != ref ]; then
        "$E2E/fcompare3d.ex" -a "$TMP/$LAST/$PLT" "$TMP/$arm/$PLT" > "$OUTDIR/fcompare_cuda_${TAG}_${LAST}_vs_$arm.txt" 2>&1
        tail -n 1 "$OUTDIR/fcompare_cuda_${TAG}_${LAST}_vs_$arm.txt"
    fi
    LAST=$arm
done
python3 - "$OUTDIR" "$TAG" <<'EOF'
import json, re, sys
out, tag = sys.argv[1:3]
names = ["Maestro::MakeEdgeScal()", "Maestro::MakeRhoXFlux()", "Maestro::MakeRhoHFlux()",
         "Maestro::UpdateScal()", "Maestro::Makew0()", "Maestro::CelltoEdge()",
         "Maestro::DensityAdvance()", "Maestro::EnthalpyAdvance()", "Maestro::VelocityAdvance()",
         "Maestro::AdvancePremac()", "Maestro::MakeUtrans()", "Maestro::VelPred()",
         "Maestro::MakeVelForce()", "Maestro::UpdateVel()", "Maestro::ModifyScalForce()",
         "Maestro::Put1dArrayOnCart()", "Maestro::Addw0()", "Maestro::MacProj()", "Maestro::NodalProj()",
         "Maestro::React()", "Maestro::AdvanceTimeStep()", "Maestro::Evolve()"]
res = {}
import os
arms = [a for a in ("ref", "mx", "fx", "cx") if os.path.exists(f"{out}/run_{tag}_{a}.log")]
for arm in arms:
    txt = open(f"{out}/run_{tag}_{arm}.log").read()
    # second TinyProfiler table = inclusive times: name ncalls min avg max pct
    incl = txt.split("Incl.")[-1] if "Incl." in txt else ""
    d = {}
    for n in names:
        m = re.search(re.escape(n) + r"\s+(\d+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)\s+([0-9.eE+-]+)", incl)
        if m:
            d[n] = {"ncalls": int(m.group(1)), "incl_s": float(m.group(3))}
    res[arm] = d
fc = open(f"{out}/fcompare_cuda_{tag}_ref_vs_{arms[-1]}.txt").read()
worst = 0.0
for line in fc.splitlines():
    m = re.match(r"^\s(\S.*?)\s+([-+0-9.eE]+)\s+([-+0-9.eE]+)\s*$", line)
    if m and not m.group(1).startswith("variable"):
        worst = max(worst, float(m.group(3)))
res["fcompare_max_rel"] = worst
res["fcompare_agree"] = "PLOTFILE AGREE" in fc
json.dump(res, open(f"{out}/cuda_compare_{tag}.json", "w"), indent=1)
for n in names:
    a = res["ref"].get(n)
    if not a:
        continue
    line = f"{n:30s} ref {a['incl_s']:8.4f} s"
    for arm in arms[1:]:
        b = res[arm].get(n)
        if b:
            line += f" | {arm} {b['incl_s']:8.4f} s x{a['incl_s'] / max(b['incl_s'], 1e-12):5.2f}"
    print(line + f"  ({a['ncalls']} calls)")
print("fcompare max rel", worst, "agree", res["fcompare_agree"])
EOF
rm -rf "$TMP"
