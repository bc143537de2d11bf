"""The reference's own CI golden for this path: after 10 steps of
`inputs_2d_regression`, AMReX's `fextrema` of the step-10 plotfile must reproduce
Exec/test_problems/reacting_bubble/ci-benchmarks/reacting_bubble_2d_extrema.out TEXTUALLY
(.github/workflows/reacting_bubble-compare.yml: `diff fextrema.out ci-benchmarks/...`).

tests/golden/reacting_bubble_2d_extrema.out is that fixture (a data file of the reference's
test suite, copied verbatim).  Two checks:
  * CPU (`not gpu`): the UNMODIFIED reference executable built by oracle/e2e/build_e2e.sh
    reproduces it - i.e. the thing every fcompare test measures against really is the reference
    as its maintainers test it (skipped where oracle/_ref/e2e is not built);
  * GPU: the overlay executable (every MakeEdgeScal / MakeRhoXFlux / MakeRhoHFlux / CelltoEdge /
    Makew0 / UpdateScal call on libmxadv.so) reproduces it too.
The same for the second golden of this path, Exec/science/wdconvect/ci-benchmarks/
wdconvect_3d_extrema.out: 3 steps of inputs_3d_regression.2levels (spherical base state, TWO AMR
levels; .github/workflows/wdconvect-compare.yml) - the only multi-level run in the suite.
Nothing here reads /root/reference at run time."""
import gzip
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E2E = os.path.join(ROOT, "oracle", "_ref", "e2e")
CSRC = os.path.join(ROOT, "maestroex_b200", "csrc")
GOLDEN = os.path.join(ROOT, "tests", "golden", "reacting_bubble_2d_extrema.out")


GOLDEN_WD = os.path.join(ROOT, "tests", "golden", "wdconvect_3d_extrema.out")


def _extrema(exe, rundir, problem="rb"):
    if problem == "rb":
        inputs, model, tool, plt, args = ("inputs_2d_regression", "model.hse.cool.coulomb", "fextrema2d.ex",
                                          "reacting_bubble_2d_plt0000010", ["maestro.max_step=10"])
    else:
        inputs, model, tool, plt, args = ("inputs_3d_wdconvect_2levels", "kepler_new_6.25e8.hybrid.hse.320",
                                          "fextrema3d.ex", "wdconvect_plt0000003", [])
    for f in (exe, tool, inputs, model, "helm_table.dat.gz"):
        if not os.path.exists(os.path.join(E2E, f)):
            pytest.skip(f"oracle/_ref/e2e/{f} not built (oracle/e2e/build_e2e.sh)")
    os.makedirs(rundir, exist_ok=True)
    for f in (inputs, model):
        shutil.copy(os.path.join(E2E, f), rundir)
    with gzip.open(os.path.join(E2E, "helm_table.dat.gz"), "rb") as f, \
            open(os.path.join(rundir, "helm_table.dat"), "wb") as g:
        shutil.copyfileobj(f, g, 1 << 24)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = CSRC + ":" + env.get("LD_LIBRARY_PATH", "")
    env.setdefault("OMP_NUM_THREADS", str(min(16, os.cpu_count() or 1)))
    r = subprocess.run([os.path.join(E2E, exe), inputs, *args],
                       cwd=rundir, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    r = subprocess.run([os.path.join(E2E, tool), plt],
                       cwd=rundir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_build_reproduces_ci_extrema(tmp_path):
    assert _extrema("Maestro2d.ref.ex", str(tmp_path)) == open(GOLDEN).read()


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["Maestro2d.mx.ex", "Maestro2d.cx.ex"])
def test_overlay_build_reproduces_ci_extrema(tmp_path, exe):
    """mx: maestroex_b200/overlay; cx: + overlay_companions (conversions, forcing, base-state
    mappings and UpdateVel on the GPU as well)"""
    assert _extrema(exe, str(tmp_path)) == open(GOLDEN).read()


def test_reference_build_reproduces_wdconvect_ci_extrema(tmp_path):
    assert _extrema("Maestro3d.wd.ref.ex", str(tmp_path), "wd") == open(GOLDEN_WD).read()


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["Maestro3d.wd.mx.ex", "Maestro3d.wd.cx.ex"])
def test_overlay_build_reproduces_wdconvect_ci_extrema(tmp_path, exe):
    """two AMR levels: AverageDownFaces / flux registers / FillPatch of the reference around the
    per-box library calls; cx: every spherical mapping (MakeS0mac, MakeW0mac, Put1dArrayOnCart) and
    UpdateVel on the GPU as well"""
    assert _extrema(exe, str(tmp_path), "wd") == open(GOLDEN_WD).read()
