"""north_star's whole-code criterion: "the fcompare diff of plotfiles after 10 full time steps
must stay within 1e-10".

oracle/e2e/build_e2e.sh builds (in the authoring container, where /root/reference exists) the
UNMODIFIED reference problem Exec/test_problems/reacting_bubble and the same problem with
maestroex_b200/overlay/ shadowing MaestroMakeEdgeScalars.cpp / MaestroMakeFlux.cpp /
MaestroCelltoEdge.cpp / MaestroMakew0.cpp / Maestro::UpdateScal on libmxadv.so, both with the
reference's own GNU-make system (CPU AMReX, OpenMP), plus AMReX's fcompare.  Here both run
`inputs_<D>d_C` for 10 steps (single level; 15 boxes of 64^2 in 2-D, 64 boxes of 32^3 in 3-D; burner, projections, EOS
are the reference's own code in both) and the step-10 plotfiles are compared variable by
variable.  The executables hold host MultiFabs, so every overlay call goes through the C ABI's
host-fab staging path onto the GPU.  Nothing here reads /root/reference at run time."""
import gzip
import os
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E2E = os.path.join(ROOT, "oracle", "_ref", "e2e")
CSRC = os.path.join(ROOT, "maestroex_b200", "csrc")
FCOMPARE_TOL = 1e-10  # north_star
NSTEPS = 10


_CACHE = {}


def _table(tmp):
    dst = os.path.join(tmp, "helm_table.dat")
    with gzip.open(os.path.join(E2E, "helm_table.dat.gz"), "rb") as f, open(dst, "wb") as g:
        shutil.copyfileobj(f, g, 1 << 24)
    return dst


# problem -> (inputs file, initial model, plotfile base name, executable tag, steps)
PROBLEMS = {
    "reacting_bubble": (None, "model.hse.cool.coulomb", "plt", "", NSTEPS),
    # Exec/science/wdconvect, inputs_files/inputs_3d_regression: 64^3, spherical base state,
    # outflow on all six faces, ppm_type 0 / slope_order 2 as shipped (BASELINE config 3 at
    # test size; its upstream CI compares 3 steps, .github/workflows/wdconvect-compare.yml)
    "wdconvect": ("inputs_3d_wdconvect", "kepler_new_6.25e8.hybrid.hse.320", "wdconvect_plt", ".wd", 3),
}


def _run(exe, rundir, dim, table, extra=(), problem="reacting_bubble"):
    inputs, model, pltbase, _, nsteps = PROBLEMS[problem]
    inputs = inputs or f"inputs_{dim}d_C"
    os.makedirs(rundir)
    for f in (inputs, model):
        shutil.copy(os.path.join(E2E, f), rundir)
    os.symlink(table, os.path.join(rundir, "helm_table.dat"))
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = CSRC + ":" + env.get("LD_LIBRARY_PATH", "")
    env.setdefault("OMP_NUM_THREADS", str(min(16, os.cpu_count() or 1)))
    cmd = [os.path.join(E2E, exe), inputs, f"maestro.max_step={nsteps}",
           f"maestro.plot_int={nsteps}", "maestro.chk_int=-1", "maestro.plot_deltat=-1",
           *extra]
    r = subprocess.run(cmd, cwd=rundir, env=env, capture_output=True, text=True, timeout=1500)
    open(os.path.join(rundir, "run.log"), "w").write(r.stdout + r.stderr)
    assert r.returncode == 0, (exe, r.stdout[-2000:], r.stderr[-2000:])
    plt = os.path.join(rundir, pltbase + "%07d" % nsteps)
    assert os.path.isdir(plt), os.listdir(rundir)
    return plt, r.stdout


def parse_fcompare(text):
    """{variable: (abs_err, rel_err)} from fcompare's table (Tools/Plotfile/fcompare.cpp)"""
    out = {}
    for line in text.splitlines():
        m = re.match(r"^\s(\S.*?)\s+([-+0-9.eE]+|nan|inf)\s+([-+0-9.eE]+|nan|inf)\s*$", line)
        if m and not m.group(1).startswith("variable"):
            out[m.group(1).strip()] = (float(m.group(2)), float(m.group(3)))
    return out


# arm "mx": maestroex_b200/overlay (member-for-member drop-in); arm "fx": overlay +
# overlay_fused (DensityAdvance / EnthalpyAdvance on mx_density_edge_flux /
# mx_enthalpy_edge_flux, the fused kernels bench.py measures)
@pytest.mark.parametrize("dim,arm,extra", [
    (2, "mx", ()),                             # inputs_2d_C as shipped: ppm_type 0
    (2, "mx", ("maestro.ppm_type=1",)),
    (2, "mx", ("maestro.ppm_type=2",)),
    (3, "mx", ()),                             # inputs_3d_C as shipped: 64x64x512, ppm_type 1
    (3, "fx", ()),
    (3, "fx", ("maestro.ppm_type=0", "amr.n_cell=32 32 256")),   # coarser grid: cheap
    (3, "fx", ("maestro.ppm_type=2", "amr.n_cell=32 32 256")),
    # boxes of 32 x 32 x 64: the host-fab z-slab pipeline runs with two slabs per box
    (3, "fx", ("amr.n_cell=32 32 256", "amr.max_grid_size=64")),
    # arm "cx": overlay (+ overlay_fused in 3-D) + overlay_companions, i.e. ALSO ConvertRhoXToX /
    # ConvertRhoHToH / PutInPertForm, ModifyScalForce, Put1dArrayOnCart, Addw0 and UpdateVel on
    # libmxadv.so inside the real time step
    (2, "cx", ()),
    (2, "cx", ("maestro.ppm_type=1",)),
    (3, "cx", ("amr.n_cell=32 32 256",)),
])
def test_fcompare_10_steps(tmp_path, tmp_path_factory, dim, arm, extra):
    for f in (f"Maestro{dim}d.ref.ex", f"Maestro{dim}d.{arm}.ex", f"fcompare{dim}d.ex",
              "helm_table.dat.gz"):
        if not os.path.exists(os.path.join(E2E, f)):
            pytest.skip(f"oracle/_ref/e2e/{f} not built (oracle/e2e/build_e2e.sh {dim})")
    tmp = str(tmp_path)
    if "table" not in _CACHE:
        _CACHE["table"] = _table(str(tmp_path_factory.mktemp("helm")))
    table = _CACHE["table"]
    if (dim, extra) not in _CACHE:  # the reference run is shared by the arms compared with it
        rdir = os.path.join(str(tmp_path_factory.mktemp("ref")), "ref")
        _CACHE[(dim, extra)] = _run(f"Maestro{dim}d.ref.ex", rdir, dim, table, extra)[0]
    pr = _CACHE[(dim, extra)]
    pm, _ = _run(f"Maestro{dim}d.{arm}.ex", os.path.join(tmp, arm), dim, table, extra)
    r = subprocess.run([os.path.join(E2E, f"fcompare{dim}d.ex"), "-a", pr, pm],
                       capture_output=True, text=True, timeout=600)
    table_txt = r.stdout
    keep = os.environ.get("MXADV_FCOMPARE_OUT")
    if keep:
        os.makedirs(keep, exist_ok=True)
        tag = f"{dim}d" + {"mx": "", "fx": "_fused", "cx": "_companions"}[arm] + "".join("_" + e.split(".")[-1].replace(" ", "x") for e in extra)
        open(os.path.join(keep, f"fcompare_{tag}.txt"), "w").write(table_txt + r.stderr)
    errs = parse_fcompare(table_txt)
    assert len(errs) >= 20, table_txt[-3000:]
    worst = max(errs.items(), key=lambda kv: kv[1][1])
    assert worst[1][1] <= FCOMPARE_TOL, (worst, table_txt[-3000:])


@pytest.mark.parametrize("arm,extra", [
    ("mx", ()),                        # as shipped: ppm_type 0, slope_order 2
    ("fx", ()),
    ("fx", ("maestro.ppm_type=1",)),
    ("mx", ("maestro.ppm_type=2",)),
    ("cx", ()),   # + MakeW0mac / MakeS0mac / spherical Put1dArrayOnCart / Addw0 / UpdateVel on the GPU
])
def test_fcompare_wdconvect(tmp_path, tmp_path_factory, arm, extra):
    """BASELINE config 3 as an application: wdconvect (spherical base state, all-outflow) with
    the overlay on libmxadv.so against the unmodified reference, 3 steps, fcompare <= 1e-10.
    Exercises Makew0Sphr, the spherical MakeRhoXFlux / MakeRhoHFlux forms (r0mac / h0mac face
    arrays) and the outflow boundary code of MakeEdgeScal inside the real time step."""
    for f in ("Maestro3d.wd.ref.ex", f"Maestro3d.wd.{arm}.ex", "fcompare3d.ex", "helm_table.dat.gz",
              "inputs_3d_wdconvect"):
        if not os.path.exists(os.path.join(E2E, f)):
            pytest.skip(f"oracle/_ref/e2e/{f} not built (oracle/e2e/build_e2e.sh 3 wdconvect)")
    if "table" not in _CACHE:
        _CACHE["table"] = _table(str(tmp_path_factory.mktemp("helm")))
    table = _CACHE["table"]
    key = ("wd", extra)
    if key not in _CACHE:
        rdir = os.path.join(str(tmp_path_factory.mktemp("wdref")), "ref")
        _CACHE[key] = _run("Maestro3d.wd.ref.ex", rdir, 3, table, extra, problem="wdconvect")[0]
    pm, _ = _run(f"Maestro3d.wd.{arm}.ex", os.path.join(str(tmp_path), arm), 3, table, extra,
                 problem="wdconvect")
    r = subprocess.run([os.path.join(E2E, "fcompare3d.ex"), "-a", _CACHE[key], pm],
                       capture_output=True, text=True, timeout=600)
    keep = os.environ.get("MXADV_FCOMPARE_OUT")
    if keep:
        os.makedirs(keep, exist_ok=True)
        tag = "wdconvect_" + arm + "".join("_" + e.split(".")[-1] for e in extra)
        open(os.path.join(keep, f"fcompare_{tag}.txt"), "w").write(r.stdout + r.stderr)
    errs = parse_fcompare(r.stdout)
    assert len(errs) >= 20, r.stdout[-3000:]
    worst = max(errs.items(), key=lambda kv: kv[1][1])
    assert worst[1][1] <= FCOMPARE_TOL, (worst, r.stdout[-3000:])
