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


def _extrema(exe, rundir):
    for f in (exe, "fextrema2d.ex", "inputs_2d_regression", "model.hse.cool.coulomb", "helm_table.dat.gz"):
        if not os.path.exists(os.path.join(E2E, f)):
            pytest.skip(f"oracle/_ref/e2e/{f} not built (oracle/e2e/build_e2e.sh 2)")
    os.makedirs(rundir, exist_ok=True)
    for f in ("inputs_2d_regression", "model.hse.cool.coulomb"):
        shutil.copy(os.path.join(E2E, f), rundir)
    with gzip.open(os.path.join(E2E, "helm_table.dat.gz"), "rb") as f, \
            open(os.path.join(rundir, "helm_table.dat"), "wb") as g:
        shutil.copyfileobj(f, g, 1 << 24)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = CSRC + ":" + env.get("LD_LIBRARY_PATH", "")
    env.setdefault("OMP_NUM_THREADS", str(min(16, os.cpu_count() or 1)))
    r = subprocess.run([os.path.join(E2E, exe), "inputs_2d_regression", "maestro.max_step=10"],
                       cwd=rundir, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    r = subprocess.run([os.path.join(E2E, "fextrema2d.ex"), "reacting_bubble_2d_plt0000010"],
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
