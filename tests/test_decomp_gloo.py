"""Host logic of the multi-GPU path (maestroex_b200/decomp.py) on CPU: rank grid, neighbour
tables, and a world_size-2 gloo run of HaloExchanger with a numpy stand-in for the device
packer (same slab layout as k_halo in csrc/mx_flux.cu), checked against a single-box problem:
after the exchange every ghost cell that has an owner brick must equal that owner's value."""
import os
import socket
import sys

import numpy as np
import pytest

from maestroex_b200.decomp import Brick, contiguous_runs, rank_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_grid():
    assert rank_grid(1) == (1, 1, 1)
    assert rank_grid(2) == (2, 1, 1)
    assert rank_grid(4) == (2, 2, 1)
    assert rank_grid(8) == (2, 2, 2)
    with pytest.raises(ValueError):
        rank_grid(6)


def test_neighbours():
    g = rank_grid(8)
    seen = set()
    for r in range(8):
        b = Brick(r, g)
        assert b.rank_of(b.coord) == r
        seen.add(b.coord)
        for d in range(3):
            lo, hi = b.neighbour(d, 0, True), b.neighbour(d, 1, True)
            assert lo == hi  # two bricks per direction: both neighbours are the same rank
            assert Brick(lo, g).neighbour(d, 1, True) == r
            if b.coord[d] == 0:
                assert b.neighbour(d, 0, False) is None and b.neighbour(d, 1, False) is not None
            else:
                assert b.neighbour(d, 1, False) is None
    assert len(seen) == 8


def test_runs():
    assert contiguous_runs([0, 4, 5, 6, 9]) == [(0, 1), (4, 3), (9, 1)]
    assert contiguous_runs([]) == []


class NumpyPacker:
    """Test stand-in for DevicePacker: numpy slab copies in k_halo's order (comp, k, j, i)."""

    def count(self, geom, fab, d, ng, ncomp):
        n = ncomp
        for q in range(3):
            n *= ng if q == d else fab.hi[q] - fab.lo[q] + 1
        return n

    def alloc(self, n):
        import torch
        return torch.empty(n, dtype=torch.float64)

    @staticmethod
    def _slab(geom, fab, d, side, ng, interior):
        nghost = geom.valid_lo[d] - fab.lo[d]
        n = fab.hi[d] - fab.lo[d] + 1
        if interior:
            start = nghost if side == 0 else n - nghost - ng
        else:
            start = nghost - ng if side == 0 else n - nghost
        sl = [slice(None)] * 4
        sl[3 - d] = slice(start, start + ng)
        return tuple(sl)

    def pack(self, geom, fab, d, side, ng, c0, nc, buf, stream=0):
        s = self._slab(geom, fab, d, side, ng, True)
        buf.numpy()[:] = fab.data[c0:c0 + nc][s].ravel()

    def unpack(self, geom, fab, d, side, ng, c0, nc, buf, stream=0):
        s = self._slab(geom, fab, d, side, ng, False)
        view = fab.data[c0:c0 + nc][s]
        view[...] = buf.numpy().reshape(view.shape)


def _worker(rank, world, port, bc_kind, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from maestroex_b200.decomp import HaloExchanger
    from maestroex_b200.synth import FIRSTSPEC, RHO, make_problem
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, ng, nspec = (12, 10, 8), 3, 3
        grid = rank_grid(world)
        brick = Brick(rank, grid)
        dom = tuple(n[d] * grid[d] for d in range(3))
        # the whole domain as one box = the truth; my brick cut out of it with stale ghosts
        whole = make_problem(dim=3, n=dom, nspec=nspec, ng=ng, bc_kind=bc_kind, seed=5)
        lo = tuple(n[d] * brick.coord[d] for d in range(3))
        mine = make_problem(dim=3, n=n, nspec=nspec, ng=ng, bc_kind=bc_kind, seed=5,
                            valid_lo=lo, dom_n=dom)
        W, M = whole.state.data, mine.state.data
        # the generator fills ghost cells with arbitrary values: build the truth by indexing the
        # valid cells of the whole domain with wrapped (periodic) / clamped global indices
        V = W[:, ng:ng + dom[2], ng:ng + dom[1], ng:ng + dom[0]]
        idx = []
        for d in range(3):
            gidx = np.arange(lo[d] - ng, lo[d] + n[d] + ng)
            idx.append(gidx % dom[d] if (bc_kind == "periodic" or d < 2)
                       else np.clip(gidx, 0, dom[d] - 1))
        M[...] = V[:, idx[2][:, None, None], idx[1][None, :, None], idx[0][None, None, :]]
        truth = M.copy()
        comps = [RHO] + list(range(FIRSTSPEC, FIRSTSPEC + nspec))
        # poison the ghosts the exchange is responsible for (directions with > 1 brick)
        for d in range(3):
            if grid[d] == 1:
                continue
            for side in (0, 1):
                per = bc_kind == "periodic" or d < 2
                if brick.neighbour(d, side, per) is None:
                    continue
                s = [slice(None)] * 4
                s[3 - d] = slice(0, ng) if side == 0 else slice(n[d] + ng, n[d] + 2 * ng)
                M[tuple(s)][comps] = np.nan
        hx = HaloExchanger(mine.geom, mine.state, mine.bc, brick, ng, NumpyPacker(), dist=dist)
        sent = hx.exchange(comps)
        ok = bool(np.array_equal(M[comps], truth[comps])) and not np.isnan(M).any()
        q.put((rank, ok, sent))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("bc_kind", ["periodic", "reacting_bubble"])
def test_halo_exchange_gloo_world2(bc_kind):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bc_kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, sent in res:
        assert ok, f"rank {rank}: ghost cells differ from the single-box truth"
        assert sent == 2 * 3 * (8 + 6) * (10 + 6) * 4 * 8  # two x slabs, ng 3, 4 comps
