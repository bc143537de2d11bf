#!/usr/bin/env python
"""bench.py — zone-species updates/s of the scalar-advection hot path (edge states + fluxes).

One "step" = what one predictor (or corrector) half of a MAESTROeX time step asks of this
path for one brick: the DENSITY pass (MakeEdgeScal on the nspec species + rho', then
MakeRhoXFlux: Source/MaestroDensityAdvance.cpp:115,128,182) followed by the ENTHALPY pass
(MakeEdgeScal on (rho h)' + MakeRhoHFlux: Source/MaestroEnthalpyAdvance.cpp:141,229), i.e.
zones x (nspec + 2) zone-species updates, through the C ABI of libmxadv.so.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--nspec 13] [--ppm 1]
  python bench.py --impl reference ...     # the reference's own CPU code (oracle/_ref)

N > 1: launched by torchrun, one rank per GPU; every rank owns one n^3 brick of a
(px*py*pz * n^3) periodic-in-xy domain (weak scaling) and, before each pass, fills the ghost
cells of the swept state components from its neighbours (NCCL send/recv of packed slabs).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "zone-species updates/s (edge-state + flux, density + enthalpy pass)"
UNIT = "zone-species/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mxadv", choices=["mxadv", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=512, help="brick edge (zones) per GPU")
    ap.add_argument("--nspec", type=int, default=13)
    ap.add_argument("--ppm", type=int, default=1, choices=[0, 1, 2])
    ap.add_argument("--bc", default="reacting_bubble")
    ap.add_argument("--spherical", action="store_true",
                    help="wdconvect-like: spherical flux forms (r0mac / h0mac face arrays)")
    ap.add_argument("--cpu-n", type=int, default=128, help="edge of the CPU-baseline sample box")
    ap.add_argument("--e2e-n", type=int, default=128, help="edge of the host-buffer e2e boxes")
    ap.add_argument("--e2e-boxes", type=int, default=8)
    ap.add_argument("--e2e-threads", type=int, default=1,
                    help="host threads (one stream each) submitting e2e boxes concurrently; "
                         "measured 0.41-1.0e9 zone-species/s at 2-8 threads vs a stable 0.69e9 "
                         "at 1 (PCIe-bound either way), hence the default")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-halo", action="store_true", help="N>1: skip the ghost exchange")
    ap.add_argument("--halo", default="p2p", choices=["nccl", "p2p", "py"],
                    help="N>1 ghost exchange: mx_halo_exchange of libmxadv.so with NVLink peer stores "
                         "(default: measured faster, 1.8 vs 3.8 ms per step at 8 x 512^3, "
                         "profiles/r2_bench_n8_*) or NCCL send/recv, or the round-1 Python / "
                         "torch.distributed loop around mx_halo_pack/unpack (A/B only)")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic uniform {a.n}^3 brick/GPU, nspec {a.nspec} (Nscal {a.nspec + 4}), "
            f"ppm_type {a.ppm}, {'spherical' if a.spherical else 'planar'} base state, BCs {a.bc}")


# ------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ------------------------------------------------------------------------------------------
class Clocks:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
        0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.ok = [], set(), False
        self.max_mhz = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# CPU reference leg (oracle/_ref = the reference's own sources; else the C oracle port)
# ------------------------------------------------------------------------------------------
def cpu_backend():
    from oracle import backends
    ref = backends.reference(3)
    if ref is not None:
        ref.lib.ref_set_num_threads(os.cpu_count() or 1)
        return ref, "reference", int(ref.lib.ref_num_threads())
    orc = backends.oracle()
    return orc, "port", int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))


def cpu_step(calls, prob, ppm, spherical=False):
    """density pass + enthalpy pass with the separate reference member functions, writing into
    the problem's own output fabs (no harness copies inside the timed region)."""
    from common import run_edge, run_rhoh
    from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH
    sedge, sflux, eta = prob.sedge, prob.sflux, prob.etarhoflux
    run_edge(calls, prob, prob.edge_params(FIRSTSPEC, prob.nspec, ppm_type=ppm), sedge=sedge)
    run_edge(calls, prob, prob.edge_params(RHO, 1, ppm_type=ppm), sedge=sedge)
    fpar = prob.flux_params(spherical=1 if spherical else 0)
    b = prob.base
    calls.make_rhoX_flux(prob.geom, sflux, eta, sedge, prob.umac, b["rho0_old"],
                         b["rho0_edge_old"], prob.mac.get("r0mac_old"), b["rho0_new"],
                         b["rho0_edge_new"], prob.mac.get("r0mac_new"),
                         b["rho0_predicted_edge"], b["w0"], fpar)
    run_edge(calls, prob, prob.edge_params(RHOH, 1, ppm_type=ppm), sedge=sedge)
    run_rhoh(calls, prob, sedge, fpar, sflux=sflux)


def cpu_measure(a, steps, warmup):
    from maestroex_b200.synth import make_problem
    calls, kind, cores = cpu_backend()
    n = a.cpu_n
    prob = make_problem(dim=3, n=(n, n, n), nspec=a.nspec, ng=4 if a.ppm == 2 else 3,
                        bc_kind=a.bc, seed=7, spherical=a.spherical)
    for _ in range(warmup):
        cpu_step(calls, prob, a.ppm, a.spherical)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(calls, prob, a.ppm, a.spherical)
    dt = (time.perf_counter() - t0) / steps
    updates = n ** 3 * (a.nspec + 2)
    return {"value": updates / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{steps} x (density + enthalpy pass) on one {n}^3 box, nspec {a.nspec}, "
                      f"ppm_type {a.ppm}, after {warmup} warm-up; {dt * 1e3:.1f} ms/pass"}, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, dt = cpu_measure(a, max(a.steps, 1), max(a.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (seeded)",
        "config": {"workload": workload_name(a), "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_mxadv(a):
    import torch
    import torch.distributed as dist

    from maestroex_b200 import lib as mxlib
    from maestroex_b200.devsynth import make_device_problem, to_pinned_host
    from maestroex_b200.fusedcalls import (density_edge_flux, enthalpy_edge_flux,
                                           enthalpy_edge_flux_sph)
    from maestroex_b200.synth import FIRSTSPEC, RHO, RHOH

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    lib = mxlib.load()
    calls = mxlib.calls()
    ng = 4 if a.ppm == 2 else 3
    n = a.n
    from maestroex_b200.decomp import (Brick, DevicePacker, HaloExchanger, LibExchanger,
                                       rank_grid)
    pg = rank_grid(world)
    brick = Brick(rank, pg)
    coord = brick.coord
    prob = make_device_problem(n=(n, n, n), nspec=a.nspec, ng=ng, bc_kind=a.bc,
                               seed=20240807 + rank, device=str(dev),
                               dom_n=tuple(n * pg[d] for d in range(3)),
                               valid_lo=tuple(n * coord[d] for d in range(3)),
                               spherical=a.spherical)
    outputs = (prob.sedge, prob.sflux, prob.etarhoflux)
    epar_d = prob.edge_params(FIRSTSPEC, a.nspec, ppm_type=a.ppm)
    epar_h = prob.edge_params(RHOH, 1, ppm_type=a.ppm)
    fpar = prob.flux_params(spherical=1 if a.spherical else 0)
    stream = torch.cuda.current_stream().cuda_stream

    def enthalpy_pass():
        if not a.spherical:
            enthalpy_edge_flux(calls, prob, epar_h, fpar, outputs=outputs, stream=stream)
        else:  # spherical flux forms read the base state from face arrays (r0mac / h0mac)
            enthalpy_edge_flux_sph(calls, prob, epar_h, fpar, outputs=outputs, stream=stream)

    # N > 1: the ghost cells of the swept components come from the face neighbours of the
    # rank's brick.  Product path: mx_halo_exchange (include/mxadv.h) - topology, packing,
    # transport and unpacking behind the C ABI.
    halo = None
    dens_comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + a.nspec))
    if world > 1 and not a.no_halo:
        if a.halo == "py":
            halo = HaloExchanger(prob.geom, prob.state, prob.bc, brick, ng, DevicePacker(calls, dev))
        else:
            halo = LibExchanger(calls, rank, world, pg,
                                transport=LibExchanger.P2P if a.halo == "p2p" else LibExchanger.NCCL,
                                dist=dist, device=dev)  # raises if peer access is unavailable
    periodic = [int(prob.bc[RHO][d]) == 0 and int(prob.bc[RHO][3 + d]) == 0 for d in range(3)]

    def exchange(comps):
        if halo is None:
            return
        if a.halo == "py":
            halo.exchange(comps, stream)
            return
        from maestroex_b200.decomp import contiguous_runs
        for c0, nc in contiguous_runs(comps):
            halo.exchange(prob.geom, prob.state, ng, c0, nc, periodic, stream)

    def step(ev=None):
        exchange(dens_comps)
        if ev is not None:
            ev[0].record()
        density_edge_flux(calls, prob, epar_d, fpar, outputs=outputs, stream=stream)
        if ev is not None:
            ev[1].record()
        exchange([RHOH])
        enthalpy_pass()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    lib.mx_reset_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(a.steps)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with Clocks(local) as clk:
        barrier()
        t_beg.record()
        for s in range(a.steps):
            step(evs[s])
        t_end.record()
        barrier()
    launches = int(lib.mx_launch_count())
    ms_total = t_beg.elapsed_time(t_end)
    ms_density = sum(e0.elapsed_time(e1) for e0, e1 in evs) / a.steps
    t = torch.tensor([ms_total, ms_density], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_density = float(t[0]), float(t[1])
    ms_step = ms_total / a.steps
    zones = n ** 3
    updates_per_step = zones * (a.nspec + 2) * world
    value = updates_per_step / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (k_stream4, density mode): algorithmic bytes / launch
    bytes_density = zones * (64 * (a.nspec + 1) + 32)  # SURVEY.md 8(d), DESIGN.md section 4
    peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            peak = float(json.load(open(pk))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    achieved = bytes_density / (ms_density * 1e-3) / 1e9
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tj):
        try:
            key = f"density_n{a.n}_nspec{a.nspec}_ppm{a.ppm}"
            traffic = json.load(open(tj)).get(key)
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm",
                "kernel": f"k_stream4<{a.ppm},1> (density pass: nspec+1 edge states + fluxes)",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_density, "ms_per_launch": ms_density}

    # ---- e2e: the same passes through the C ABI with pinned HOST fabs (library stages them)
    e2e = None
    if not a.no_e2e and not a.spherical:
        en = a.e2e_n
        dsmall = make_device_problem(n=(en, en, en), nspec=a.nspec, ng=ng, bc_kind=a.bc,
                                     seed=99 + rank, device=str(dev))
        # one pinned-host copy of the problem per worker: a caller with several boxes in
        # flight (AMReX: one OpenMP thread / stream per box) lets the H2D copies of one box
        # overlap the D2H copies of another; every call still returns with its results in
        # host memory (the C ABI is synchronous for host fabs)
        nthr = max(1, min(a.e2e_threads, a.e2e_boxes))
        hps = [to_pinned_host(dsmall) for _ in range(nthr)]
        del dsmall
        streams = [torch.cuda.Stream(device=dev) for _ in range(nthr)]
        he_d = hps[0].edge_params(FIRSTSPEC, a.nspec, ppm_type=a.ppm)
        he_h = hps[0].edge_params(RHOH, 1, ppm_type=a.ppm)

        from maestroex_b200.abi import Fab, UpdateParams
        from maestroex_b200.fusedcalls import density_advance, enthalpy_advance
        from maestroex_b200.synth import TEMP

        # host-side outputs of the fused advances: snew (valid zones of the species / density /
        # enthalpy components are written), the time-centred enthalpy force and the EOS mask
        def pinned_like(f, fill):
            t = torch.full(tuple(f.data.shape), fill, dtype=torch.float64).pin_memory()
            return Fab(t, f.lo, f.hi)

        hnew = [pinned_like(hp.state, 1.0) for hp in hps]
        hfup = [pinned_like(hp.force, 0.0) for hp in hps]
        hmask = [torch.zeros(en ** 3, dtype=torch.uint8).pin_memory() for _ in hps]
        up_d = UpdateParams(dt=hps[0].dt, start_comp=FIRSTSPEC, num_comp=a.nspec, num_spec=a.nspec,
                            base_cutoff_density=-1.0, rho_comp=RHO, rhoh_comp=RHOH, temp_comp=TEMP,
                            first_spec=FIRSTSPEC)
        up_h = UpdateParams(dt=hps[0].dt, start_comp=RHOH, num_comp=1, num_spec=a.nspec,
                            base_cutoff_density=-1.0, rho_comp=RHO, rhoh_comp=RHOH, temp_comp=TEMP,
                            first_spec=FIRSTSPEC)

        def e2e_worker(w, fused):
            torch.cuda.set_device(local)
            hp, st = hps[w], streams[w].cuda_stream
            hout = (hp.sedge, hp.sflux, hp.etarhoflux)
            for _ in range(w, a.e2e_boxes, nthr):
                if fused:
                    # DensityAdvance / EnthalpyAdvance as overlay_fused calls them: edge states
                    # and fluxes stay on the device, the new state comes back
                    density_advance(calls, hp, hnew[w], he_d, fpar, up_d, species_force_is_zero=True,
                                    outputs=hout, stream=st)
                    enthalpy_advance(calls, hp, hnew[w], hfup[w], he_h, fpar, up_h, eos_mask=hmask[w],
                                     stream=st)
                else:
                    density_edge_flux(calls, hp, he_d, fpar, outputs=hout, stream=st)
                    enthalpy_edge_flux(calls, hp, he_h, fpar, outputs=hout, stream=st)

        def e2e_step(fused):
            if nthr == 1:
                e2e_worker(0, fused)
                return
            ths = [threading.Thread(target=e2e_worker, args=(w, fused)) for w in range(nthr)]
            for t_ in ths:
                t_.start()
            for t_ in ths:
                t_.join()

        def e2e_time(fused):
            e2e_step(fused)
            barrier()
            t0 = time.perf_counter()
            for _ in range(a.steps):
                e2e_step(fused)
            torch.cuda.synchronize()
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te[0]) / a.steps

        te_adv = e2e_time(True)
        te_mem = e2e_time(False)
        C = a.nspec + 1
        cells_s = (en + 2 * ng) ** 3
        cells_f = (en + 2) ** 3
        face = (en + 1) * en * en
        uface = (en + 3) * (en + 2) ** 2
        nb = a.e2e_boxes
        # fused advances: converted state (C comps), force(Rho), 3 MAC velocities and etarhoflux
        # up; snew(C, valid zones) and etarhoflux down | state(RhoH), force(RhoH),
        # force_update(RhoH), 3 MAC velocities, snew(Rho) up; snew(RhoH, valid zones) + EOS mask down
        h2d = 8 * nb * (C * cells_s + cells_f + 3 * uface + face
                        + cells_s + 2 * cells_f + 3 * uface + cells_s)
        d2h = 8 * nb * (C * en ** 3 + face + en ** 3) + nb * en ** 3
        h2d_mem = 8 * nb * (C * cells_s + C * cells_f + 3 * uface + face
                            + cells_s + cells_f + 3 * uface + 3 * face)
        d2h_mem = 8 * nb * (3 * C * face + 3 * C * face + face + 3 * face + 3 * face)
        upd = en ** 3 * (a.nspec + 2) * nb * world
        e2e = {"value": upd / te_adv, "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "sample": f"{nb} boxes of {en}^3 per step, pinned host fabs through "
                         f"mx_density_advance + mx_enthalpy_advance (what overlay_fused's "
                         f"DensityAdvance / EnthalpyAdvance call: edge states + fluxes + UpdateScal, "
                         f"intermediates device resident, new state returned), {nthr} host "
                         f"thread(s) / stream(s) with boxes in flight",
               "members": {"value": upd / te_mem, "h2d_bytes_per_step": h2d_mem,
                           "d2h_bytes_per_step": d2h_mem,
                           "sample": "same boxes through mx_density_edge_flux + mx_enthalpy_edge_flux "
                                     "(edge states and fluxes returned to the host, z-slab pipeline)"}}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu, _ = cpu_measure(a, 3, 1)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (seeded)",
            "config": {"workload": workload_name(a), "zones_per_gpu": zones,
                       "updates_per_step": updates_per_step, "rank_grid": list(pg),
                       "halo_exchange": (a.halo if halo is not None else False),
                       "l2": "inputs (>100 GB/step at 512^3) far exceed the 126 MB L2"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_mxadv(a)


if __name__ == "__main__":
    main()
