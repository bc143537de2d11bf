"""Worker of tests/test_gpu_multi.py (run under torchrun, one rank per GPU).

ONE global seeded problem is built identically on every rank and solved by the CPU oracle on
the UNDIVIDED box.  Each rank then cuts its brick (valid cells + ghost cells) out of the global
arrays, POISONS every ghost layer that faces a neighbour brick (state: ng layers, force: 1,
MAC velocities: 1, i.e. everything SURVEY.md section 8e lists), restores them with
mx_halo_exchange of libmxadv.so (the product's C ABI: brick topology, packing, transport and
unpacking all behind it), checks the ghosts came back bit for bit, runs the fused density pass
on the brick on its GPU and compares EVERY face of the brick — including the faces on the
brick interface — with the oracle's result for the same faces of the undivided problem."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cut_np(fab, lo, hi):
    from maestroex_b200.abi import Fab
    sl = tuple(slice(lo[d] - fab.lo[d], hi[d] - fab.lo[d] + 1) for d in (2, 1, 0))
    return Fab(np.ascontiguousarray(fab.data[(slice(None),) + sl]), tuple(lo), tuple(hi))


def main():
    import torch
    import torch.distributed as dist
    from common import run_edge, run_rhoX
    from gpu_util import to_device, to_host
    from maestroex_b200 import lib as mxlib
    from maestroex_b200.abi import Geom
    from maestroex_b200.decomp import Brick, LibExchanger, rank_grid
    from maestroex_b200.fusedcalls import density_edge_flux
    from maestroex_b200.synth import FIRSTSPEC, RHO, Problem, make_problem
    from oracle import backends

    transport = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    ppm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    grid = rank_grid(world)
    brick = Brick(rank, grid)
    nb = (32, 28, 30)  # zones per brick (even ghost-grown rows: the TMA kernel runs)
    nspec, ng = 3, (4 if ppm == 2 else 3)
    n_glob = tuple(nb[d] * grid[d] for d in range(3))
    glob = make_problem(dim=3, n=n_glob, nspec=nspec, ng=ng, bc_kind="reacting_bubble", seed=20240807)
    # make the global problem EXACTLY periodic in its periodic directions (the generator
    # evaluates its fields at the ghost coordinates, which agrees with the periodic image only
    # to rounding and not at all for the noise): ghost cells := wrapped valid cells, and for an
    # array that is nodal in the direction the last face := the first
    per_glob = [int(glob.bc[RHO][d]) == 0 and int(glob.bc[RHO][3 + d]) == 0 for d in range(3)]

    def wrap(fab, nodal_dir=-1):
        t = fab.data
        for d in range(3):
            if not per_glob[d]:
                continue
            ax = 3 - d
            gl = -fab.lo[d]                      # ghost width (domain starts at 0)
            n = n_glob[d]

            def sl(a, b):
                idx = [slice(None)] * 4
                idx[ax] = slice(a, b)
                return tuple(idx)
            if nodal_dir == d:
                t[sl(gl + n, gl + n + 1)] = t[sl(gl, gl + 1)]            # face n == face 0
                if gl:
                    t[sl(0, gl)] = t[sl(n, n + gl)]                      # faces -gl..-1 = n-gl..n-1
                    t[sl(gl + n + 1, gl + n + 1 + gl)] = t[sl(gl + 1, gl + 1 + gl)]
            elif gl:
                t[sl(0, gl)] = t[sl(n, n + gl)]                          # cells -gl..-1 = n-gl..n-1
                t[sl(gl + n, gl + n + gl)] = t[sl(gl, gl + gl)]          # cells n..n+gl-1 = 0..gl-1

    wrap(glob.state)
    wrap(glob.force)
    for d in range(3):
        wrap(glob.umac[d], d)
    orc = backends.oracle()
    epar = glob.edge_params(FIRSTSPEC, nspec, ppm_type=ppm)
    want_e = run_edge(orc, glob, epar)
    run_edge(orc, glob, glob.edge_params(RHO, 1, ppm_type=ppm), sedge=want_e)
    want_f, want_eta = run_rhoX(orc, glob, want_e, glob.flux_params())

    c = brick.coord
    lo = tuple(c[d] * nb[d] for d in range(3))
    hi = tuple(lo[d] + nb[d] - 1 for d in range(3))
    g = glob.geom
    geom = Geom(3, lo, hi, g.dom_lo, g.dom_hi, g.dx)
    state = cut_np(glob.state, *geom.cell_box(ng))
    force = cut_np(glob.force, *geom.cell_box(1))
    umac = [cut_np(glob.umac[d], *geom.face_box(d, 1)) for d in range(3)]
    sedge = [cut_np(glob.sedge[d], *geom.face_box(d, 0)) for d in range(3)]
    sflux = [cut_np(glob.sflux[d], *geom.face_box(d, 0)) for d in range(3)]
    eta = cut_np(glob.etarhoflux, *geom.face_box(2, 0))
    prob = Problem(geom=geom, nspec=nspec, ng=ng, state=state, force=force, umac=umac, sedge=sedge,
                   sflux=sflux, etarhoflux=eta, base=glob.base, mac={}, dt=glob.dt,
                   rel_eps=glob.rel_eps, bc=glob.bc)
    dprob = to_device(prob)
    # the lateral directions of reacting_bubble are periodic, the vertical one is not
    periodic = [int(glob.bc[RHO][d]) == 0 and int(glob.bc[RHO][3 + d]) == 0 for d in range(3)]

    def poison(fab, depth, nodal_dir=-1):
        """NaN into the ghost layers (all components) that face a neighbour brick."""
        t = fab.data
        for d in range(3):
            if grid[d] == 1:
                continue
            ax = 3 - d  # data is (comp, z, y, x)
            n = t.shape[ax]
            for side in (0, 1):
                if brick.neighbour(d, side, periodic[d]) is None:
                    continue
                idx = [slice(None)] * 4
                idx[ax] = slice(0, depth) if side == 0 else slice(n - depth, n)
                t[tuple(idx)] = float("nan")

    keep = {"state": dprob.state.data.clone(), "force": dprob.force.data.clone(),
            "umac": [u.data.clone() for u in dprob.umac]}
    poison(dprob.state, ng)
    poison(dprob.force, 1)
    for d in range(3):
        poison(dprob.umac[d], 1)
    assert bool(torch.isnan(dprob.state.data).any()) or all(x == 1 for x in grid)

    calls = mxlib.calls()
    ex = LibExchanger(calls, rank, world, grid, transport=transport, dist=dist, device=dev)
    sent = 0
    sent += ex.exchange(geom, dprob.state, ng, 0, prob.nscal, periodic)
    sent += ex.exchange(geom, dprob.force, 1, 0, prob.nscal, periodic)
    for d in range(3):
        sent += ex.exchange(geom, dprob.umac[d], 1, 0, 1, periodic)
    torch.cuda.synchronize()

    def same(got, want, what):
        if torch.equal(got, want):
            return
        bad = (got != want) | torch.isnan(got)
        idx = bad.nonzero()
        raise AssertionError(f"rank {rank}: {what}: {int(bad.sum())} of {got.numel()} differ, "
                             f"{int(torch.isnan(got).sum())} NaN left; first (comp,k,j,i) = {idx[0].tolist()}, "
                             f"last = {idx[-1].tolist()}, shape {tuple(got.shape)}")

    same(dprob.state.data, keep["state"], "state ghosts after the exchange")
    same(dprob.force.data, keep["force"], "force ghosts after the exchange")
    for d in range(3):
        same(dprob.umac[d].data, keep["umac"][d], f"umac[{d}] ghosts after the exchange")

    # a second exchange re-uses buffers / flags (P2P epochs); then the fused density pass
    poison(dprob.state, ng)
    ex.exchange(geom, dprob.state, ng, 0, prob.nscal, periodic)
    torch.cuda.synchronize()
    assert torch.equal(dprob.state.data, keep["state"]), "second exchange"
    ge, gf, geta = density_edge_flux(calls, dprob, prob.edge_params(FIRSTSPEC, nspec, ppm_type=ppm),
                                     prob.flux_params())
    ge, gf, geta = to_host(ge), to_host(gf), to_host([geta])[0]
    comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + nspec))
    for d in range(3):
        we = cut_np(want_e[d], ge[d].lo, ge[d].hi).data
        wf = cut_np(want_f[d], gf[d].lo, gf[d].hi).data
        for cc in comps:
            assert np.array_equal(ge[d].data[cc], we[cc]), f"rank {rank}: sedge dir {d} comp {cc}"
            assert np.array_equal(gf[d].data[cc], wf[cc]), f"rank {rank}: sflux dir {d} comp {cc}"
    assert np.array_equal(geta.data, cut_np(want_eta, geta.lo, geta.hi).data), "etarhoflux"
    ex.close()
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        print(f"MULTI-GPU PARITY OK: {world} bricks {grid}, transport {'P2P' if transport else 'NCCL'}, "
              f"ppm {ppm}, {sent} bytes sent by rank 0, launches incl. halo kernels "
              f"{mxlib.load().mx_launch_count()}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
