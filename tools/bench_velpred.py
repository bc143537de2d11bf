#!/usr/bin/env python
"""Throughput of mx_make_utrans + mx_vel_pred (Maestro::MakeUtrans / Maestro::VelPred, the
MAC-velocity predictor of AdvancePremac) on one GPU, device-resident inputs, CUDA-event timing,
next to the compiled reference (oracle/_ref, all host threads) on a smaller box.
Algorithmic bytes per zone (fp64, 3-D): MakeUtrans reads utilde(3) ufull(3) w0(1) and writes
utrans(3) = 80 B; VelPred reads utilde(3) ufull(3) utrans(3) force(3) w0(1) and writes umac(3)
= 128 B.  usage: python tools/bench_velpred.py [--size 256] [--ppm 1] [--cpu-size 64]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def problem(n, ppm, xp, dev=None):
    """reacting_bubble-like box covering its domain; xp = numpy or torch"""
    import numpy as np
    from maestroex_b200.abi import Fab, Geom, VelPredParams, box_shape
    from test_velpred import bcs_u_of, phys_for
    geom = Geom(3, (0, 0, 0), (n - 1,) * 3, (0, 0, 0), (n - 1,) * 3, (1.0 / n,) * 3)
    ng = 4 if ppm == 2 else 3

    def rnd(shape, lo, hi):
        if dev is None:
            return np.random.default_rng(7).uniform(lo, hi, size=shape)
        import torch
        return torch.rand(shape, dtype=torch.float64, device=dev) * (hi - lo) + lo

    def cell(nc, g, lo=-1.0, hi=1.0):
        blo, bhi = geom.cell_box(g)
        return Fab(rnd((nc,) + box_shape(blo, bhi), lo, hi), blo, bhi)

    def faces(g):
        out = []
        for d in range(3):
            blo, bhi = geom.face_box(d, g)
            out.append(Fab(rnd((1,) + box_shape(blo, bhi), -1.0, 1.0), blo, bhi))
        return out

    lo, hi = phys_for(3, "bubble")
    par = VelPredParams(dt=0.2 / n, rel_eps=1e-8, ppm_type=ppm, slope_order=4, phys_bc_lo=lo, phys_bc_hi=hi,
                        bc=bcs_u_of(3, lo, hi))
    return dict(geom=geom, utilde=cell(3, ng), ufull=cell(3, ng), w0=cell(3, 1, -0.3, 0.3), force=cell(3, 1),
                utrans=faces(1), umac=faces(1), par=par)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--cpu-size", type=int, default=64)
    ap.add_argument("--ppm", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from maestroex_b200 import lib as mxlib
    calls = mxlib.calls()
    P = problem(a.size, a.ppm, torch, "cuda")
    st = torch.cuda.current_stream().cuda_stream

    def step(ev=None):
        calls.make_utrans(P["geom"], P["utilde"], P["ufull"], P["utrans"], P["w0"], None, P["par"], stream=st)
        if ev:
            ev.record()
        calls.vel_pred(P["geom"], P["utilde"], P["ufull"], P["utrans"], P["umac"], None, P["w0"], P["force"],
                       P["par"], stream=st)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    tu = tv = 0.0
    for _ in range(a.steps):
        e0.record()
        step(e1)
        e2.record()
        torch.cuda.synchronize()
        tu += e0.elapsed_time(e1)
        tv += e1.elapsed_time(e2)
    tu, tv = tu / a.steps, tv / a.steps
    zones = a.size ** 3
    peak = 6445.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    out = {"workload": f"{a.size}^3, ppm_type {a.ppm}, reacting_bubble BCs, planar",
           "make_utrans": {"ms": tu, "zones_per_s": zones / tu * 1e3, "algorithmic_GBps": zones * 80 / tu / 1e6,
                           "frac_of_hbm_peak": zones * 80 / tu / 1e6 / peak},
           "vel_pred": {"ms": tv, "zones_per_s": zones / tv * 1e3, "algorithmic_GBps": zones * 128 / tv / 1e6,
                        "frac_of_hbm_peak": zones * 128 / tv / 1e6 / peak},
           "peak_GBps": peak}
    from oracle import backends
    ref = backends.reference(3)
    if ref is not None and a.cpu_size > 0:
        import numpy as np
        ref.lib.ref_set_num_threads(os.cpu_count() or 1)
        C = problem(a.cpu_size, a.ppm, np)
        ts = []
        for fn in ("u", "v"):
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                if fn == "u":
                    ref.make_utrans(C["geom"], C["utilde"], C["ufull"], C["utrans"], C["w0"], None, C["par"])
                else:
                    ref.vel_pred(C["geom"], C["utilde"], C["ufull"], C["utrans"], C["umac"], None, C["w0"],
                                 C["force"], C["par"])
                best = min(best, time.perf_counter() - t0)
            ts.append(best)
        cz = a.cpu_size ** 3
        out["reference_cpu"] = {"box": f"{a.cpu_size}^3", "threads": int(ref.lib.ref_num_threads()),
                                "make_utrans_zones_per_s": cz / ts[0], "vel_pred_zones_per_s": cz / ts[1]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
