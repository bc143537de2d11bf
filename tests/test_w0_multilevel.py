"""Maestro::Makew0Planar with a MULTILEVEL base state (Source/MaestroMakew0.cpp:60-272: level loop,
restriction / offset :185-213, zeroing :216-236) including the reference's RestrictBase /
FillGhostBase (Source/MaestroBaseStateGeometry.cpp:236-355): oracle vs the compiled reference on
the CPU, libmxadv.so (mx_make_w0_planar_ml) vs the oracle on the GPU.  Bit-exact."""
import numpy as np
import pytest

from maestroex_b200.abi import RadialGeom, W0Params

# name -> (nr_fine, chunks per level [(lo, hi), ...] for levels 1.., is_predictor, add_pb, dpdt, cut frac)
CASES = {
    "two_levels": (64, [[(16, 47)]], 1, 0, 0.3, 0.8),
    "two_levels_pb": (64, [[(8, 39)]], 0, 1, 0.5, 1.0),
    "two_chunks": (96, [[(8, 23), (40, 71)]], 1, 0, 0.2, 0.9),
    "three_levels": (128, [[(16, 47)], [(48, 79)]], 0, 0, 0.4, 0.7),
    "three_levels_wd": (2048, [[(256, 767)], [(640, 1279)]], 1, 0, 0.2, 0.85),
}


def problem(name):
    nrf, chunks, pred, pb, dpdt, cutf = CASES[name]
    nlev = len(chunks) + 1
    rng = np.random.default_rng(len(name) * 7 + nrf)
    nr = np.array([nrf >> (nlev - 1 - n) for n in range(nlev)], dtype=np.int32)
    drf = 1.0e8 / nrf
    dr = np.array([drf * (1 << (nlev - 1 - n)) for n in range(nlev)], dtype=np.float64)
    maxch = max(len(c) for c in chunks)
    cl = maxch + 1
    ndc = np.ones(nlev, dtype=np.int32)
    rs = np.zeros((nlev, cl), dtype=np.int32)
    re = np.zeros((nlev, cl), dtype=np.int32)
    rs[0, 1], re[0, 1] = 0, nr[0] - 1
    for n, ch in enumerate(chunks, start=1):
        ndc[n] = len(ch)
        for j, (lo, hi) in enumerate(ch, start=1):
            rs[n, j], re[n, j] = lo, hi
    cut = np.array([int(cutf * nr[n]) for n in range(nlev)], dtype=np.int32)
    rg = RadialGeom(nrf, dr, nr, cut, ndc, rs, re)

    def cc(scale, base=0.0):
        return np.ascontiguousarray(base + scale * rng.random((nlev, nrf)))

    d = dict(w0_old=np.ascontiguousarray(1.0e4 * rng.standard_normal((nlev, nrf + 1))),
             Sbar=np.ascontiguousarray(1.0e-3 * rng.standard_normal((nlev, nrf))),
             p0_old=cc(1.0e22, 1.0e23), p0_new=cc(1.0e22, 1.0e23), gamma1bar_old=cc(0.2, 1.4),
             gamma1bar_new=cc(0.2, 1.4),
             p0_minus_peosbar=np.ascontiguousarray(1.0e18 * rng.standard_normal((nlev, nrf))),
             etarho_cc=np.ascontiguousarray(1.0e2 * rng.standard_normal((nlev, nrf))))
    par = W0Params(nr_fine=nrf, base_cutoff_density_coord=int(cut[0]), dr=float(dr[0]), dt=0.37, dt_in=0.4,
                   dtold_in=0.33, is_predictor=pred, add_pb=pb, dpdt_factor=dpdt, base_cutoff_density=-1.0)
    return par, rg, d


def run(calls, par, rg, d):
    nlev, nrf = d["Sbar"].shape
    w0 = np.full((nlev, nrf + 1), 7.0)
    force = np.full((nlev, nrf), 7.0)
    pb = calls.make_w0_planar_ml(par, rg, d["w0_old"], d["Sbar"], d["p0_old"], d["p0_new"], d["gamma1bar_old"],
                                 d["gamma1bar_new"], d["p0_minus_peosbar"], d["etarho_cc"], w0, force)
    return w0, force, pb


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_vs_reference(oracle_calls, ref3d, name):
    par, rg, d = problem(name)
    rw, rf, rpb = run(ref3d, par, rg, d)
    ow, of, opb = run(oracle_calls, par, rg, d)
    assert np.isfinite(rw).all() and np.abs(rw).max() > 0.0
    assert np.abs(rw[1:]).max() > 0.0, "the fine levels were not exercised"
    assert np.array_equal(rw, ow)
    assert np.array_equal(rf, of)
    assert rpb == opb


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_vs_oracle(oracle_calls, name):
    from maestroex_b200 import lib as mxlib
    par, rg, d = problem(name)
    ow, of, opb = run(oracle_calls, par, rg, d)
    gw, gf, gpb = run(mxlib.calls(), par, rg, d)
    assert np.array_equal(gw, ow)
    assert np.array_equal(gf, of)
    assert gpb == opb
