"""The passes either side of the edge-state / flux path:
ConvertRhoXToX / ConvertRhoHToH, PutInPertForm, Put1dArrayOnCart, ModifyScalForce, Addw0,
MakeW0mac, MakeS0mac, UpdateVel.

CPU (``not gpu``): the C oracle (oracle/mxadv_oracle_cart.c) against the reference's own
MaestroConvert.cpp / MaestroFill3dData.cpp / MaestroForce.cpp / MaestroAdvection.cpp compiled
unmodified (oracle/_ref), bit for bit -> the oracle is pinned.
GPU: libmxadv.so (mx_cart.cu) against the oracle through the C ABI, bit for bit, with device-
and host-resident fabs."""
import numpy as np
import pytest

from maestroex_b200.abi import BC_FOEXTRAP, BC_INT_DIR, Base, Fab, Geom, Sphr, box_shape

NSCAL, RHO, RHOH, FIRSTSPEC, NSPEC = 7, 0, 1, 4, 3


def rnd(rng, shape, lo=0.5, hi=1.5):
    return np.ascontiguousarray(rng.uniform(lo, hi, size=shape))


class Prob:
    """One box; `dom` >= box.  Fabs are created on demand with the ghost widths the reference
    gives them."""

    def __init__(self, dim, n, seed, dom_pad=(0, 0, 0), spherical=False):
        self.dim = dim
        self.rng = np.random.default_rng(seed)
        n = tuple(n) + (1,) * (3 - len(n))
        lo = tuple(dom_pad[d] if d < dim else 0 for d in range(3))
        hi = tuple(lo[d] + n[d] - 1 if d < dim else 0 for d in range(3))
        dlo = (0, 0, 0)
        dhi = tuple(hi[d] + dom_pad[d] if d < dim else 0 for d in range(3))
        nmax = max(dhi[d] + 1 for d in range(dim))
        dx = tuple(1.0 / nmax if d < dim else 0.0 for d in range(3))
        self.geom = Geom(dim, lo[:dim], hi[:dim], dlo[:dim], dhi[:dim], dx[:dim])
        self.lo, self.hi, self.dlo, self.dhi, self.dx = lo, hi, dlo, dhi, dx
        self.nr = dhi[dim - 1] + 1
        self.sphr = None
        if spherical:
            drf = dx[0] / 5.0
            nr = int(np.sqrt(3.0) * (nmax / 2 + 4) * 5) + 4
            self.sphr = dict(nr_fine=nr, dr_fine=drf, prob_lo=(0.0, 0.0, 0.0),
                             center=tuple(0.5 * (dhi[d] + 1) * dx[d] for d in range(3)),
                             r_cc_loc=(np.arange(nr) + 0.5) * drf, r_edge_loc=np.arange(nr + 1) * drf)
            self.nr = nr

    def cell(self, ncomp, ng, lo=0.5, hi=1.5):
        blo, bhi = self.geom.cell_box(ng)
        return Fab(rnd(self.rng, (ncomp,) + box_shape(blo, bhi), lo, hi), blo, bhi)

    def face(self, d, ncomp, ng, lo=-1.0, hi=1.0):
        blo, bhi = self.geom.face_box(d, ng)
        return Fab(rnd(self.rng, (ncomp,) + box_shape(blo, bhi), lo, hi), blo, bhi)

    def faces(self, ncomp, ng, lo=-1.0, hi=1.0):
        return [self.face(d, ncomp, ng, lo, hi) for d in range(self.dim)]

    def base(self, n, lo=0.5, hi=1.5):
        return Base(rnd(self.rng, (1, n, 1), lo, hi))

    def S(self, **kw):
        return Sphr(**{**self.sphr, **kw}) if self.sphr else None


def clone(f):
    if isinstance(f, (list, tuple)):
        return [clone(x) for x in f]
    return Fab(f.data.copy(), f.lo, f.hi)


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        q = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {a.size} differ; first {q}: {a[q]!r} vs {b[q]!r}")


# ---------------------------------------------------------------------------------------------
# case runners: each takes a Calls back end (+ optional device mover) and returns output arrays
# ---------------------------------------------------------------------------------------------
def identity(x):
    return x


def run_convert(calls, dim, flag, what, mv=identity, back=identity):
    p = Prob(dim, (12, 10, 14)[:dim], 11)
    s = p.cell(NSCAL, 3)
    d = mv(s)
    if what == "X":
        calls.convert_by_rho(p.geom, d, RHO, FIRSTSPEC, NSPEC, flag)
    else:
        calls.convert_by_rho(p.geom, d, RHO, RHOH, 1, flag)
    return [back(d).data]


def run_put1d(calls, dim, edge, vec, sph, types=(3, 2), mv=identity, back=identity):
    p = Prob(dim, (12, 10, 14)[:dim], 12, spherical=sph)
    out = p.cell(dim if vec else 1, 1)
    s0 = p.base(p.nr + 1)
    d = mv(out)
    calls.put_1d_array_on_cart(p.geom, s0, 0, d, edge, vec,
                               p.S(s0_interp_type=types[0], w0_interp_type=types[1]))
    return [back(d).data]


def run_pert(calls, dim, flag, sph, s0t=3, mv=identity, back=identity):
    p = Prob(dim, (12, 10, 14)[:dim], 13, spherical=sph)
    s = p.cell(NSCAL, 3)
    s0 = p.base(p.nr + 1)
    d = mv(s)
    calls.put_in_pert_form(p.geom, d, s0, 0, RHO, flag, p.S(s0_interp_type=s0t))
    return [back(d).data]


def run_force(calls, dim, fullform, bcf, sph, s0t=3, pad=(0, 0, 0), ecfab=False, mv=identity, back=identity):
    p = Prob(dim, (12, 10, 14)[:dim], 14, dom_pad=pad, spherical=sph)
    force = p.cell(NSCAL, 1, -1.0, 1.0)
    state = p.cell(NSCAL, 3)
    umac = p.faces(1, 1)
    s0_edge = p.base(p.nr + 1)
    s0_cart = p.cell(1, 1)
    w0_cart = p.cell(dim, 2, -0.3, 0.3)
    w0 = rnd(p.rng, (p.nr + 1,), -0.3, 0.3)
    ec = p.cell(1, 1)  # the reference's s0_edge_cart temporary as an AMR caller would pass it
    d = mv(force)
    calls.modify_scal_force(p.geom, d, mv(state), [mv(u) for u in umac], s0_edge, 0, mv(s0_cart),
                            None if sph else mv(w0_cart), w0 if sph else None, RHO, fullform, bcf,
                            p.S(s0_interp_type=s0t), s0_edge_cart=mv(ec) if ecfab else None)
    return [back(d).data]


def run_addw0(calls, dim, sph, mv=identity, back=identity):
    p = Prob(dim, (12, 10, 14)[:dim], 15, spherical=sph)
    ue = p.faces(1, 1)
    wm = p.faces(1, 1) if sph else None
    w0 = p.base(p.nr + 1, -0.3, 0.3)
    d = [mv(u) for u in ue]
    calls.addw0(p.geom, d, [mv(w) for w in wm] if sph else None, w0, 0, -1.0, 1 if sph else 0)
    return [back(u).data for u in d]


def run_w0mac(calls, w0mt, pad=(0, 0, 0), mv=identity, back=identity):
    p = Prob(3, (12, 10, 14), 16, dom_pad=pad, spherical=True)
    wm = p.faces(1, 1)
    w0_cart = p.cell(3, 2, -0.3, 0.3)
    w0 = rnd(p.rng, (p.nr + 1,), -0.3, 0.3)
    d = [mv(w) for w in wm]
    calls.make_w0mac(p.geom, d, mv(w0_cart) if w0mt == 1 else None, w0, p.S(w0mac_interp_type=w0mt))
    return [back(w).data for w in d]


def run_s0mac(calls, s0mt, s0t, pad=(0, 0, 0), cartfab=False, mv=identity, back=identity):
    p = Prob(3, (12, 10, 14), 17, dom_pad=pad, spherical=True)
    sm = p.faces(1, 1)
    s0 = p.base(p.nr)
    cart = p.cell(1, 2)
    d = [mv(w) for w in sm]
    calls.make_s0mac(p.geom, s0, d, p.S(s0mac_interp_type=s0mt, s0_interp_type=s0t),
                     s0_cart=mv(cart) if cartfab else None)
    return [back(w).data for w in d]


def run_updvel(calls, dim, sponge, sph, mv=identity, back=identity):
    p = Prob(dim, (12, 10, 14)[:dim], 18, spherical=sph)
    uold = p.cell(dim, 3, -1.0, 1.0)
    unew = p.cell(dim, 3, -1.0, 1.0)
    umac = p.faces(1, 1)
    uedge = p.faces(dim, 0)
    force = p.cell(dim, 1, -1.0, 1.0)
    sp = p.cell(1, 0, 0.5, 1.0)
    w0_cart = p.cell(dim, 2, -0.3, 0.3)
    w0mac = p.faces(1, 1) if sph else None
    d = mv(unew)
    calls.update_vel(p.geom, mv(uold), d, [mv(u) for u in umac], [mv(u) for u in uedge], mv(force),
                     mv(sp) if sponge else None, None if sph else mv(w0_cart),
                     [mv(w) for w in w0mac] if sph else None, 0.37, 1 if sponge else 0, 1 if sph else 0)
    return [back(d).data]


# (name, runner, kwargs, needs 3-D reference)
CASES = []
for dim in (2, 3):
    for flag in (True, False):
        CASES.append((f"convertX-{dim}d-{flag}", run_convert, dict(dim=dim, flag=flag, what="X")))
        CASES.append((f"convertH-{dim}d-{flag}", run_convert, dict(dim=dim, flag=flag, what="H")))
        CASES.append((f"pert-{dim}d-{flag}", run_pert, dict(dim=dim, flag=flag, sph=False)))
    for edge in (False, True):
        for vec in (False, True):
            CASES.append((f"put1d-{dim}d-e{int(edge)}v{int(vec)}", run_put1d,
                          dict(dim=dim, edge=edge, vec=vec, sph=False)))
    for ff in (0, 1):
        for bcf in (BC_FOEXTRAP, BC_INT_DIR):
            CASES.append((f"force-{dim}d-ff{ff}-bc{bcf}", run_force,
                          dict(dim=dim, fullform=ff, bcf=bcf, sph=False)))
    CASES.append((f"addw0-{dim}d", run_addw0, dict(dim=dim, sph=False)))
    for sponge in (False, True):
        CASES.append((f"updvel-{dim}d-s{int(sponge)}", run_updvel, dict(dim=dim, sponge=sponge, sph=False)))
for t in (1, 2, 3):
    CASES.append((f"put1d-sph-cc-t{t}", run_put1d, dict(dim=3, edge=False, vec=False, sph=True, types=(t, 2))))
    CASES.append((f"put1d-sph-ccvec-t{t}", run_put1d, dict(dim=3, edge=False, vec=True, sph=True, types=(t, 2))))
    CASES.append((f"put1d-sph-edge-t{t}", run_put1d, dict(dim=3, edge=True, vec=False, sph=True, types=(3, t))))
    CASES.append((f"put1d-sph-edgevec-t{t}", run_put1d, dict(dim=3, edge=True, vec=True, sph=True, types=(3, t))))
    CASES.append((f"pert-sph-t{t}", run_pert, dict(dim=3, flag=True, sph=True, s0t=t)))
    for ff in (0, 1):
        CASES.append((f"force-sph-ff{ff}-t{t}", run_force, dict(dim=3, fullform=ff, bcf=BC_FOEXTRAP, sph=True, s0t=t)))
    CASES.append((f"s0mac-t1-s0t{t}", run_s0mac, dict(s0mt=1, s0t=t)))
for t in (2, 3):
    CASES.append((f"s0mac-t{t}", run_s0mac, dict(s0mt=t, s0t=3)))
for t in (1, 2, 3, 4):
    CASES.append((f"w0mac-t{t}", run_w0mac, dict(w0mt=t)))
CASES.append(("addw0-sph", run_addw0, dict(dim=3, sph=True)))
CASES.append(("updvel-sph-s0", run_updvel, dict(dim=3, sponge=False, sph=True)))
CASES.append(("updvel-sph-s1", run_updvel, dict(dim=3, sponge=True, sph=True)))

# boxes strictly inside the domain: only oracle vs CUDA (the one-box shim's ghost fill of the
# reference's internal temporaries stands for a box that reaches the domain faces)
GPU_ONLY = [
    ("force-3d-inner", run_force, dict(dim=3, fullform=0, bcf=BC_FOEXTRAP, sph=False, pad=(2, 3, 4))),
    ("force-sph-inner", run_force, dict(dim=3, fullform=0, bcf=BC_FOEXTRAP, sph=True, pad=(2, 3, 4))),
    ("s0mac-t1-inner", run_s0mac, dict(s0mt=1, s0t=3, pad=(2, 3, 4))),
    ("w0mac-t4-inner", run_w0mac, dict(w0mt=4, pad=(2, 3, 4))),
    # AMR callers hand the reference's internal temporaries in as fabs (after their own FillPatch)
    ("force-3d-ecfab", run_force, dict(dim=3, fullform=0, bcf=BC_FOEXTRAP, sph=False, pad=(2, 3, 4), ecfab=True)),
    ("force-2d-ecfab", run_force, dict(dim=2, fullform=0, bcf=BC_FOEXTRAP, sph=False, ecfab=True)),
    ("s0mac-t1-cartfab", run_s0mac, dict(s0mt=1, s0t=3, pad=(2, 3, 4), cartfab=True)),
]


@pytest.mark.parametrize("name,runner,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference(name, runner, kw, oracle_calls, ref2d, ref3d):
    ref = ref2d if kw.get("dim", 3) == 2 else ref3d
    want = runner(ref, **kw)
    got = runner(oracle_calls, **kw)
    for q, (a, b) in enumerate(zip(got, want)):
        same(a, b, f"{name}[{q}] oracle vs reference")


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


@pytest.mark.gpu
@pytest.mark.parametrize("residency", ["device", "host"])
@pytest.mark.parametrize("name,runner,kw", CASES + GPU_ONLY, ids=[c[0] for c in CASES + GPU_ONLY])
def test_cuda_matches_oracle(name, runner, kw, residency, oracle_calls, mx):
    want = runner(oracle_calls, **kw)
    if residency == "device":
        got = runner(mx, mv=_to_dev, back=_to_host, **kw)
    else:
        got = runner(mx, **kw)
    for q, (a, b) in enumerate(zip(got, want)):
        same(a, b, f"{name}[{q}] CUDA ({residency}) vs oracle")
