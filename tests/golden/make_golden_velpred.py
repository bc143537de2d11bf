"""Writes tests/golden/velpred_<case>.npz: outputs of the REFERENCE's own Maestro::MakeUtrans /
Maestro::VelPred (Source/MaestroMakeUtrans.cpp, Source/MaestroVelPred.cpp compiled unmodified
into oracle/_ref by oracle/Makefile) on the seeded inputs of tests/test_velpred.py.
Run here (needs /root/reference):  python tests/golden/make_golden_velpred.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import backends  # noqa: E402
import test_velpred as T  # noqa: E402

for name, runner, kw in T.GOLDEN:
    ref = backends.reference(kw["dim"])
    assert ref is not None, "oracle/_ref is not built"
    outs = runner(ref, **kw)
    np.savez_compressed(T.golden_path(name), **{f"out{q}": a for q, a in enumerate(outs)})
    print(name, [a.shape for a in outs])
