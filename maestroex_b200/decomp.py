"""Brick decomposition of one level across the GPUs of a box + ghost-cell exchange.

What AMReX does for this path with `FillBoundary` / `FillPatchSingleLevel`
(external/amrex/Src/Base/AMReX_FabArrayCommI.H:49-150, called from
Source/MaestroDensityAdvance.cpp:98-104 / MaestroEnthalpyAdvance.cpp:120-136 before
MakeEdgeScal): every rank owns one brick of the BoxArray and, before a sweep, fills the
`ng` ghost layers of the swept state components from its face neighbours.  Here:

* `rank_grid(world)`   bricks per direction with minimum surface: 1,2,4,8 GPUs ->
                       (1,1,1), (2,1,1), (2,2,1), (2,2,2)   (SURVEY.md section 8e);
* `Brick`              rank <-> brick coordinate, neighbours with periodic wrap;
* `HaloExchanger`      direction-by-direction exchange (x, then y including the x ghosts,
                       then z including x and y ghosts), which also fills the edge and corner
                       ghosts the corner-coupled stencil reads (simhy(i-1, j+1, k) etc.).
                       Slabs are packed / unpacked on the device by mx_halo_pack /
                       mx_halo_unpack (include/mxadv.h) and moved with batched
                       `torch.distributed` isend/irecv: NCCL send/recv over NVLink on the GPU
                       box, gloo in the CPU tests of the host logic.

The exchanger never computes on the host: `packer` is the device packer by default; the CPU
tests inject a numpy stand-in (tests/test_decomp_gloo.py) to exercise neighbour logic, message
matching and slab geometry with world_size 2 on gloo.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

from .abi import BC_INT_DIR


def rank_grid(world: int) -> Tuple[int, int, int]:
    """Bricks per direction for `world` ranks (a power of two), minimum surface."""
    p = [1, 1, 1]
    d = 0
    w = world
    while w > 1:
        if w % 2:
            raise ValueError("world size must be a power of two")
        p[d % 3] *= 2
        w //= 2
        d += 1
    return tuple(p)


@dataclass(frozen=True)
class Brick:
    """Position of one rank's brick in the rank grid (x fastest, as AMReX orders boxes)."""

    rank: int
    grid: Tuple[int, int, int]

    @property
    def coord(self) -> Tuple[int, int, int]:
        g = self.grid
        return (self.rank % g[0], (self.rank // g[0]) % g[1], self.rank // (g[0] * g[1]))

    def rank_of(self, coord: Sequence[int]) -> int:
        g = self.grid
        return coord[0] + g[0] * (coord[1] + g[1] * coord[2])

    def neighbour(self, d: int, side: int, periodic: bool):
        """Rank of the face neighbour in direction d (side 0 = low, 1 = high), or None at a
        non-periodic domain boundary (physical-BC ghosts are filled locally)."""
        c = list(self.coord)
        c[d] += -1 if side == 0 else 1
        if c[d] < 0 or c[d] >= self.grid[d]:
            if not periodic:
                return None
            c[d] %= self.grid[d]
        return self.rank_of(c)


def contiguous_runs(comps: Sequence[int]) -> List[Tuple[int, int]]:
    """[(first, count)] for a sorted component list: one packed message per run."""
    runs: List[Tuple[int, int]] = []
    for c in sorted(set(int(x) for x in comps)):
        if runs and runs[-1][0] + runs[-1][1] == c:
            runs[-1] = (runs[-1][0], runs[-1][1] + 1)
        else:
            runs.append((c, 1))
    return runs


class DevicePacker:
    """mx_halo_pack / mx_halo_unpack on torch CUDA buffers (the product path)."""

    def __init__(self, calls, device):
        import torch
        self.calls = calls
        self.lib = calls.lib
        self.torch = torch
        self.device = device
        self.lib.mx_halo_count.restype = C.c_longlong

    def count(self, geom, fab, d, ng, ncomp) -> int:
        g, f = geom.c(), fab.c()
        return int(self.lib.mx_halo_count(C.byref(g), C.byref(f), d, ng, ncomp))

    def alloc(self, n):
        return self.torch.empty(n, dtype=self.torch.float64, device=self.device)

    def _go(self, name, geom, fab, d, side, ng, c0, nc, buf, stream):
        g, f = geom.c(), fab.c()
        fn = getattr(self.lib, name)
        fn.restype = C.c_int
        rc = fn(C.byref(g), C.byref(f), d, side, ng, c0, nc, C.c_void_p(buf.data_ptr()),
                C.c_void_p(stream))
        self.calls._check(rc, name[3:])

    def pack(self, geom, fab, d, side, ng, c0, nc, buf, stream=0):
        self._go("mx_halo_pack", geom, fab, d, side, ng, c0, nc, buf, stream)

    def unpack(self, geom, fab, d, side, ng, c0, nc, buf, stream=0):
        self._go("mx_halo_unpack", geom, fab, d, side, ng, c0, nc, buf, stream)


class HaloExchanger:
    """Fills the `ng` ghost layers of components of `fab` from the face neighbours.

    Directions with a single brick (`grid[d] == 1`) are left alone: their ghost cells are
    either periodic images of the brick itself or physical-boundary cells, both of which the
    caller fills locally (no communication).  A direction is periodic when the BCRec of the
    first exchanged component is `int_dir` on both faces (Source/MaestroSetup.cpp:276-559).
    """

    def __init__(self, geom, fab, bc, brick: Brick, ng: int, packer, dist=None, group=None):
        self.geom, self.fab, self.bc, self.brick, self.ng = geom, fab, bc, brick, ng
        self.packer = packer
        if dist is None:
            import torch.distributed as dist  # noqa: PLC0415
        self.dist = dist
        self.group = group
        self._bufs = {}
        self.bytes_sent_last = 0

    def periodic(self, d: int, comp: int) -> bool:
        return int(self.bc[comp][d]) == BC_INT_DIR and int(self.bc[comp][3 + d]) == BC_INT_DIR

    def _buf(self, key, n):
        b = self._bufs.get(key)
        if b is None or b.numel() < n:
            b = self.packer.alloc(n)
            self._bufs[key] = b
        return b[:n]

    def plan(self, comps: Sequence[int]):
        """[(d, c0, nc, lo_neighbour, hi_neighbour)] for this rank, in execution order."""
        out = []
        for d in range(self.geom.dim):
            if self.brick.grid[d] == 1:
                continue
            for c0, nc in contiguous_runs(comps):
                per = self.periodic(d, c0)
                out.append((d, c0, nc, self.brick.neighbour(d, 0, per),
                            self.brick.neighbour(d, 1, per)))
        return out

    def exchange(self, comps: Sequence[int], stream: int = 0):
        dist, pk = self.dist, self.packer
        sent = 0
        plan = self.plan(comps)
        for d in range(self.geom.dim):
            steps = [s for s in plan if s[0] == d]
            if not steps:
                continue
            sends, recvs, unpacks = [], [], []
            for (_, c0, nc, nlo, nhi) in steps:
                n = pk.count(self.geom, self.fab, d, self.ng, nc)
                nbrs = (nlo, nhi)
                # my `side` interior goes to that neighbour; tags keep lo / hi apart on gloo
                # when both neighbours are the same rank (two bricks, periodic direction)
                for side in (0, 1):
                    if nbrs[side] is None:
                        continue
                    sb = self._buf(("s", d, side, c0), n)
                    pk.pack(self.geom, self.fab, d, side, self.ng, c0, nc, sb, stream)
                    sends.append(dist.P2POp(dist.isend, sb, nbrs[side], group=self.group,
                                            tag=(d * 2 + side) * 64 + c0))
                    sent += n * 8
                # NCCL ignores tags and matches the k-th send to a peer with that peer's k-th
                # recv: the peer sends its LOW interior first, which belongs in my HIGH ghosts
                for side in (1, 0):
                    if nbrs[side] is None:
                        continue
                    rb = self._buf(("r", d, side, c0), n)
                    recvs.append(dist.P2POp(dist.irecv, rb, nbrs[side], group=self.group,
                                            tag=(d * 2 + (1 - side)) * 64 + c0))
                    unpacks.append((side, c0, nc, rb))
            for w in dist.batch_isend_irecv(sends + recvs):
                w.wait()
            for side, c0, nc, rb in unpacks:
                pk.unpack(self.geom, self.fab, d, side, self.ng, c0, nc, rb, stream)
        self.bytes_sent_last = sent
        return sent


class LibExchanger:
    """The product path: mx_halo_exchange of libmxadv.so (include/mxadv.h) — brick topology,
    packing, transport (NCCL send/recv or NVLink peer stores) and unpacking all happen behind
    the C ABI; Python only hands the 128-byte communicator id from rank 0 to the other ranks
    (torch.distributed broadcast; a C++ host would use MPI_Bcast)."""

    NCCL, P2P = 0, 1

    def __init__(self, calls, rank: int, world: int, grid, transport: int = 0, dist=None,
                 device=None):
        import torch
        self.calls, self.lib = calls, calls.lib
        self.rank, self.world, self.grid = rank, world, tuple(int(x) for x in grid)
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            calls._check(self.lib.mx_comm_unique_id(idbuf), "comm_unique_id")
        if world > 1:
            if dist is None:
                import torch.distributed as dist  # noqa: PLC0415
            t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8,
                             device=device if device is not None else "cpu")
            dist.broadcast(t, src=0)
            idbuf = (C.c_ubyte * 128)(*[int(x) for x in t.cpu().tolist()])
        self.comm = C.c_void_p()
        g = (C.c_int * 3)(*self.grid)
        calls._check(self.lib.mx_comm_create(C.byref(self.comm), idbuf, rank, world, g,
                                             int(transport)), "comm_create")
        self.lib.mx_comm_bytes_sent.restype = C.c_longlong
        self.lib.mx_comm_bytes_sent.argtypes = [C.c_void_p]

    def exchange(self, geom, fab, ng: int, comp0: int, ncomp: int, periodic, stream: int = 0):
        g, f = geom.c(), fab.c()
        per = (C.c_int * 3)(*[1 if p else 0 for p in periodic])
        rc = self.lib.mx_halo_exchange(self.comm, C.byref(g), C.byref(f), int(ng), int(comp0),
                                       int(ncomp), per, C.c_void_p(stream))
        self.calls._check(rc, "halo_exchange")
        return int(self.lib.mx_comm_bytes_sent(self.comm))

    def close(self):
        if self.comm:
            self.lib.mx_comm_destroy(self.comm)
            self.comm = C.c_void_p()
