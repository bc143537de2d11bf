#!/usr/bin/env python
"""Throughput of mx_update_scal (Maestro::UpdateScal zone loop) on one GPU: species + density
update followed by the enthalpy update, device-resident inputs, CUDA-event timing.
Algorithmic bytes per zone (fp64): species pass 8 (sold) + 8 (force) + 24 (one new face per
direction) + 8 (snew) = 48 per species, + 16 for rho old/new; enthalpy pass 48 + 8 (rho read)
+ 1 (mask).  usage: python tools/bench_update_scal.py [--size 512] [--nspec 13]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--nspec", type=int, default=13)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.abi import Fab, UpdateParams
    from maestroex_b200.devsynth import make_device_problem
    from maestroex_b200.synth import FIRSTSPEC, RHOH
    n, ns = a.size, a.nspec
    prob = make_device_problem(n=(n, n, n), nspec=ns, ng=3, alloc_outputs=True)
    for f in prob.sflux:
        f.data.normal_().mul_(1.0e-4)  # small against the zone content: no floor / negative-species work
        f.data[FIRSTSPEC + ns - 1].mul_(1.0e-30)  # the trace species carries a trace flux
    del prob.sedge[:]
    torch.cuda.empty_cache()
    snew = Fab(torch.zeros_like(prob.state.data), prob.state.lo, prob.state.hi)
    prob.state.data[0] = 1.0e5
    mask = torch.zeros(n ** 3, dtype=torch.uint8, device="cuda")
    calls = mxlib.calls()
    ps = UpdateParams(dt=prob.dt, start_comp=FIRSTSPEC, num_comp=ns, num_spec=ns,
                      base_cutoff_density=1.0e-3)
    ph = UpdateParams(dt=prob.dt, start_comp=RHOH, num_comp=1, num_spec=ns,
                      base_cutoff_density=1.0e-3)
    st = torch.cuda.current_stream().cuda_stream

    def step(ev=None):
        calls.update_scal(prob.geom, prob.state, snew, prob.sflux, prob.force, ps, stream=st)
        if ev:
            ev.record()
        calls.update_scal(prob.geom, prob.state, snew, prob.sflux, prob.force, ph,
                          eos_mask=mask, stream=st)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ts, th = 0.0, 0.0
    for _ in range(a.steps):
        e0.record()
        step(e1)
        e2.record()
        torch.cuda.synchronize()
        ts += e0.elapsed_time(e1)
        th += e1.elapsed_time(e2)
    ts, th = ts / a.steps, th / a.steps
    zones = n ** 3
    bytes_s = zones * (48 * ns + 16)
    bytes_h = zones * 57
    peak = 6445.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    print(json.dumps({
        "kernel": "k_update_scal<3>", "workload": f"{n}^3, nspec {ns}",
        "species_pass": {"ms": ts, "algorithmic_bytes": bytes_s, "GBps": bytes_s / ts / 1e6,
                         "frac_of_hbm_peak": bytes_s / ts / 1e6 / peak,
                         "zone_species_per_s": zones * ns / ts * 1e3},
        "enthalpy_pass": {"ms": th, "algorithmic_bytes": bytes_h, "GBps": bytes_h / th / 1e6,
                          "frac_of_hbm_peak": bytes_h / th / 1e6 / peak},
        "peak_GBps": peak}))


if __name__ == "__main__":
    main()
