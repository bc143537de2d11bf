"""MAC-velocity predictor: Maestro::MakeUtrans and Maestro::VelPred.

Checkers: (i) the REFERENCE ITSELF, Source/MaestroMakeUtrans.cpp and Source/MaestroVelPred.cpp
compiled unmodified into oracle/_ref/libmaestro_ref{2,3}d.so (oracle/Makefile); (ii) the plain-C
restatement orc_make_utrans / orc_vel_pred (oracle/mxadv_oracle.c), staged like the reference;
(iii) golden vectors generated from (i) (tests/golden/velpred_*.npz, written by
tests/golden/make_golden_velpred.py).

CPU (``not gpu``): the C oracle against the compiled reference on every case, bit for bit -> the
oracle is pinned; the committed goldens are reproduced by the compiled reference and by the oracle.
GPU: libmxadv.so (mx_velpred.cu) through the C ABI against the oracle, the compiled reference and
the goldens on every case (2-D / 3-D, ppm_type 0/1/2, slope_order 0/2/4, every phys_bc code,
planar / spherical, ppm_trace_forces, box on the domain faces and strictly inside), bit for bit,
with device- and host-resident fabs."""
import os

import numpy as np
import pytest

from maestroex_b200.abi import (BC_EXT_DIR, BC_FOEXTRAP, BC_HOEXTRAP, BC_INT_DIR, BC_REFLECT_EVEN,
                                BC_REFLECT_ODD, PHYS_INFLOW, PHYS_INTERIOR, PHYS_NOSLIPWALL,
                                PHYS_OUTFLOW, PHYS_SLIPWALL, PHYS_SYMMETRY, Fab, VelPredParams)
from test_cart import Prob, identity, same

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bcs_u_of(dim, lo, hi):
    """Maestro::BCSetup's velocity BCRecs for a phys_bc (Source/MaestroSetup.cpp:325-545)."""
    bc = np.zeros((dim, 6), dtype=np.int32)
    for side, codes in ((0, lo), (3, hi)):
        for d in range(dim):
            for n in range(dim):
                p = codes[d]
                if p == PHYS_INTERIOR:
                    v = BC_INT_DIR
                elif p == PHYS_INFLOW or p == PHYS_NOSLIPWALL:
                    v = BC_EXT_DIR
                elif p == PHYS_OUTFLOW:
                    v = BC_FOEXTRAP
                elif p == PHYS_SYMMETRY:
                    v = BC_REFLECT_ODD if n == d else BC_REFLECT_EVEN
                else:  # slipwall
                    v = BC_EXT_DIR if n == d else BC_HOEXTRAP
                bc[n, side + d] = v
    return bc


PHYS = {
    "interior": ((0, 0, 0), (0, 0, 0)),
    # reacting_bubble: periodic sides, slip wall below, outflow above
    "bubble": ((PHYS_INTERIOR, PHYS_INTERIOR, PHYS_SLIPWALL), (PHYS_INTERIOR, PHYS_INTERIOR, PHYS_OUTFLOW)),
    "outflow": ((PHYS_OUTFLOW,) * 3, (PHYS_OUTFLOW,) * 3),
    "walls": ((PHYS_SLIPWALL, PHYS_NOSLIPWALL, PHYS_SYMMETRY), (PHYS_SYMMETRY, PHYS_SLIPWALL, PHYS_NOSLIPWALL)),
    "mixed": ((PHYS_INFLOW, PHYS_SYMMETRY, PHYS_OUTFLOW), (PHYS_OUTFLOW, PHYS_INFLOW, PHYS_SLIPWALL)),
    "mixed2": ((PHYS_NOSLIPWALL, PHYS_OUTFLOW, PHYS_INFLOW), (PHYS_INFLOW, PHYS_NOSLIPWALL, PHYS_SYMMETRY)),
}


def phys_for(dim, name):
    lo, hi = PHYS[name]
    if dim == 2:  # the last direction of the 3-D sets becomes y
        lo, hi = (lo[0], lo[2]), (hi[0], hi[2])
    return tuple(lo), tuple(hi)


def setup(dim, ppm, so, phys, sph, trace, pad, seed):
    p = Prob(dim, (12, 10, 14)[:dim], seed, dom_pad=pad, spherical=sph)
    ng = 4 if ppm == 2 else 3
    utilde = p.cell(dim, ng, -1.0, 1.0)
    ufull = p.cell(dim, ng, -1.0, 1.0)
    # some exactly stagnant / tiny velocities so that every Riemann branch is taken
    flat = ufull.data.reshape(-1)
    flat[:: 17] = 0.0
    flat[5:: 23] *= 1e-12
    w0_cart = p.cell(dim, 1, -0.3, 0.3)
    w0mac = p.faces(1, 1, -0.3, 0.3) if sph else None
    force = p.cell(dim, ng if trace else 1, -1.0, 1.0)
    lo, hi = phys_for(dim, phys)
    par = VelPredParams(dt=0.013, rel_eps=1e-8, ppm_type=ppm, slope_order=so, ppm_trace_forces=trace,
                        spherical=1 if sph else 0, phys_bc_lo=lo, phys_bc_hi=hi, bc=bcs_u_of(dim, lo, hi))
    return p, utilde, ufull, w0_cart, w0mac, force, par


def one_thread(calls):
    """The compiled reference runs VelPred's tiles under OpenMP, and tiles overlap: every tile
    writes ul / ur / uimh over its GROWN box, boundary fix-ups and the |utrans| < rel_eps average
    as second writes (MaestroVelPred.cpp:815-892, :880-891).  A neighbouring tile's thread can read
    the first write.  The race is the reference's own (it shows up as a handful of zones in ~1 of
    200 runs here); the checker therefore runs it on one thread, where tiles run one after the
    other and every tile reads what it wrote itself."""
    if calls.prefix == "ref_":
        calls.lib.ref_set_num_threads(1)


def run_utrans(calls, dim, ppm, so=4, phys="bubble", sph=False, pad=(0, 0, 0), mv=identity, back=identity):
    p, utilde, ufull, w0_cart, w0mac, _, par = setup(dim, ppm, so, phys, sph, 0, pad, 31)
    utrans = p.faces(1, 1)
    d = [mv(u) for u in utrans]
    one_thread(calls)
    calls.make_utrans(p.geom, mv(utilde), mv(ufull), d, None if sph else mv(w0_cart),
                      [mv(w) for w in w0mac] if sph else None, par)
    return [back(u).data for u in d]


def run_velpred(calls, dim, ppm, so=4, phys="bubble", sph=False, trace=0, pad=(0, 0, 0), mv=identity,
                back=identity):
    p, utilde, ufull, w0_cart, w0mac, force, par = setup(dim, ppm, so, phys, sph, trace, pad, 32)
    utrans = p.faces(1, 1)
    for u in utrans:  # exact zeros / tiny values on some faces
        f = u.data.reshape(-1)
        f[:: 13] = 0.0
        f[3:: 29] *= 1e-12
    umac = p.faces(1, 1)
    d = [mv(u) for u in umac]
    one_thread(calls)
    calls.vel_pred(p.geom, mv(utilde), mv(ufull), [mv(u) for u in utrans], d,
                   [mv(w) for w in w0mac] if sph else None, None if sph else mv(w0_cart), mv(force), par)
    return [back(u).data for u in d]


CASES = []
for dim in (2, 3):
    for ppm in (0, 1, 2):
        for phys in PHYS:
            CASES.append((f"utrans-{dim}d-ppm{ppm}-{phys}", run_utrans, dict(dim=dim, ppm=ppm, phys=phys)))
            CASES.append((f"velpred-{dim}d-ppm{ppm}-{phys}", run_velpred, dict(dim=dim, ppm=ppm, phys=phys)))
    for so in (0, 2):
        CASES.append((f"utrans-{dim}d-slope{so}", run_utrans, dict(dim=dim, ppm=0, so=so, phys="walls")))
        CASES.append((f"velpred-{dim}d-slope{so}", run_velpred, dict(dim=dim, ppm=0, so=so, phys="mixed")))
    CASES.append((f"utrans-{dim}d-inner", run_utrans, dict(dim=dim, ppm=1, phys="outflow", pad=(2, 3, 4))))
    CASES.append((f"velpred-{dim}d-inner", run_velpred, dict(dim=dim, ppm=1, phys="outflow", pad=(2, 3, 4))))
for ppm in (0, 1, 2):
    CASES.append((f"utrans-sph-ppm{ppm}", run_utrans, dict(dim=3, ppm=ppm, phys="outflow", sph=True)))
    CASES.append((f"velpred-sph-ppm{ppm}", run_velpred, dict(dim=3, ppm=ppm, phys="outflow", sph=True)))
for ppm in (1, 2):
    CASES.append((f"velpred-3d-trace-ppm{ppm}", run_velpred, dict(dim=3, ppm=ppm, phys="bubble", trace=1)))
    CASES.append((f"velpred-3d-trace-ppm{ppm}-mixed", run_velpred, dict(dim=3, ppm=ppm, phys="mixed", trace=1)))

# the cases whose reference outputs are committed (tests/golden/velpred_<name>.npz)
GOLDEN = [c for c in CASES if c[0] in (
    "utrans-2d-ppm0-bubble", "utrans-2d-ppm1-walls", "utrans-3d-ppm1-bubble", "utrans-3d-ppm2-mixed",
    "utrans-sph-ppm1", "velpred-2d-ppm0-mixed", "velpred-2d-ppm1-bubble", "velpred-2d-ppm2-walls",
    "velpred-3d-ppm0-walls", "velpred-3d-ppm1-bubble", "velpred-3d-ppm2-mixed2", "velpred-sph-ppm1",
    "velpred-3d-trace-ppm1")]


def golden_path(name):
    return os.path.join(GOLD, f"velpred_{name}.npz")


def load_golden(name):
    z = np.load(golden_path(name))
    return [z[f"out{q}"] for q in range(len(z.files))]


@pytest.mark.parametrize("name,runner,kw", GOLDEN, ids=[c[0] for c in GOLDEN])
def test_golden_reproduced_by_reference(name, runner, kw, ref2d, ref3d):
    ref = ref2d if kw["dim"] == 2 else ref3d
    got = runner(ref, **kw)
    for q, (a, b) in enumerate(zip(got, load_golden(name))):
        same(a, b, f"{name}[{q}] reference vs committed golden")


@pytest.mark.parametrize("name,runner,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference(name, runner, kw, oracle_calls, ref2d, ref3d):
    ref = ref2d if kw["dim"] == 2 else ref3d
    want = runner(ref, **kw)
    got = runner(oracle_calls, **kw)
    for q, (a, b) in enumerate(zip(got, want)):
        same(a, b, f"{name}[{q}] oracle vs reference")


@pytest.mark.parametrize("name,runner,kw", GOLDEN, ids=[c[0] for c in GOLDEN])
def test_oracle_matches_golden(name, runner, kw, oracle_calls):
    got = runner(oracle_calls, **kw)
    for q, (a, b) in enumerate(zip(got, load_golden(name))):
        same(a, b, f"{name}[{q}] oracle vs committed golden")


@pytest.mark.parametrize("name,runner,kw", GOLDEN, ids=[c[0] for c in GOLDEN])
def test_golden_files_sane(name, runner, kw):
    outs = load_golden(name)
    assert len(outs) == kw["dim"]
    for a in outs:
        assert a.dtype == np.float64 and np.isfinite(a).all()


def test_unsupported_is_loud():
    """ppm_trace_forces in 2-D (the reference reads unfilled temporaries there) is refused by the
    library before any device work, like every other unsupported option."""
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.calls import MxError
    calls = mxlib.calls()
    p, utilde, ufull, w0_cart, _, force, par = setup(2, 1, 4, "bubble", False, 1, (0, 0, 0), 5)
    with pytest.raises(MxError):
        calls.vel_pred(p.geom, utilde, ufull, p.faces(1, 1), p.faces(1, 1), None, w0_cart, force, par)


@pytest.fixture(scope="module")
def mx():
    from maestroex_b200 import lib as mxlib
    lib = mxlib.load()
    assert lib.mx_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return mxlib.calls()


def _to_dev(f):
    import torch
    return Fab(torch.from_numpy(np.ascontiguousarray(f.data)).cuda(), f.lo, f.hi)


def _to_host(f):
    return Fab(f.data.cpu().numpy(), f.lo, f.hi)


def _run_mx(mx, runner, kw, residency):
    if residency == "device":
        return runner(mx, mv=_to_dev, back=_to_host, **kw)
    return runner(mx, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("residency", ["device", "host"])
@pytest.mark.parametrize("name,runner,kw", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_reference(name, runner, kw, residency, mx, ref2d, ref3d):
    ref = ref2d if kw["dim"] == 2 else ref3d
    want = runner(ref, **kw)
    got = _run_mx(mx, runner, kw, residency)
    for q, (a, b) in enumerate(zip(got, want)):
        same(a, b, f"{name}[{q}] CUDA ({residency}) vs reference")


@pytest.mark.gpu
@pytest.mark.parametrize("name,runner,kw", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_oracle(name, runner, kw, mx, oracle_calls):
    want = runner(oracle_calls, **kw)
    got = _run_mx(mx, runner, kw, "device")
    for q, (a, b) in enumerate(zip(got, want)):
        same(a, b, f"{name}[{q}] CUDA vs oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("name,runner,kw", GOLDEN, ids=[c[0] for c in GOLDEN])
def test_cuda_matches_golden(name, runner, kw, mx):
    got = _run_mx(mx, runner, kw, "device")
    for q, (a, b) in enumerate(zip(got, load_golden(name))):
        same(a, b, f"{name}[{q}] CUDA vs committed golden")
