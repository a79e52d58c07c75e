"""
Scatterer - ghost-dof (halo) exchange between GPUs
==================================================

Drop-in for ``/root/reference/cuda/scatterer.py``:

* the four kernels ``pack_fwd / unpack_fwd / pack_rev / unpack_rev`` (:18-101)
  with the ``kernel[grid, block](...)`` launch syntax;
* ``scatter_forward(comm, owners_data, ghosts_data, N, float_type)`` (:191-277)
  and ``scatter_reverse(...)`` (:104-188), each returning ``scatter(buffer)``.

Semantics kept: the local vector is ``[owned (N) |
ghosts]``; *forward* overwrites every ghost copy with its owner's value,
*reverse* adds the ghost-region partial sums into the owner and leaves the
ghost region as it was.

B200 design.  The reference launches one pack and one unpack kernel per
neighbour, brackets the MPI round with two device-wide synchronisations and
moves one vector per call (2-3 calls per RK stage).  Here
``HaloExchange`` concatenates the per-neighbour index lists once, so one
launch packs *k vectors for every neighbour* into an entry-major send buffer
(``fus_pack_multi``), the per-neighbour slices go out as ONE grouped NCCL
send/recv round over NVLink (``torch.distributed.batch_isend_irecv``; no host
synchronisation - the wait is a stream dependency), and one launch unpacks.
``forward_begin / forward_end`` run the round on a side stream so interior
cells can be computed while the halo is in flight.

``comm`` is a ``torch.distributed`` process group (``None`` = world) or any
object with an ``exchange(...)`` method (see ``LocalCluster`` for emulating
several ranks inside one process on one GPU).
"""

from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import check, current_stream, dev, fn
from .operators import _Kernel

# --------------------------------------------------------------------------- #
# the four reference kernels
# --------------------------------------------------------------------------- #


class _PackFwd(_Kernel):
    """``out[i] = in[index[i]]`` - cuda/scatterer.py:18-35."""

    def __call__(self, in_, out_, index):
        a = dev(in_)
        o, ix = dev(out_, a.dtype), dev(index, np.int64)
        check(fn("fus_pack_fwd", a.dtype)(a.ptr, o.ptr, ix.ptr, ix.size, current_stream()), "fus_pack_fwd")


class _UnpackFwd(_Kernel):
    """``out[index[i] + N] = in[i]`` - cuda/scatterer.py:38-57."""

    def __call__(self, in_, out_, index, N):
        a = dev(in_)
        o, ix = dev(out_, a.dtype), dev(index, np.int64)
        check(fn("fus_unpack_fwd", a.dtype)(a.ptr, o.ptr, ix.ptr, ix.size, int(N), current_stream()),
              "fus_unpack_fwd")


class _PackRev(_Kernel):
    """``out[i] = in[index[i] + N]`` - cuda/scatterer.py:60-79."""

    def __call__(self, in_, out_, index, N):
        a = dev(in_)
        o, ix = dev(out_, a.dtype), dev(index, np.int64)
        check(fn("fus_pack_rev", a.dtype)(a.ptr, o.ptr, ix.ptr, ix.size, int(N), current_stream()),
              "fus_pack_rev")


class _UnpackRev(_Kernel):
    """``out[index[i]] += in[i]`` (atomic) - cuda/scatterer.py:82-101."""

    def __call__(self, in_, out_, index):
        a = dev(in_)
        o, ix = dev(out_, a.dtype), dev(index, np.int64)
        check(fn("fus_unpack_rev", a.dtype)(a.ptr, o.ptr, ix.ptr, ix.size, current_stream()),
              "fus_unpack_rev")


pack_fwd = _PackFwd()
unpack_fwd = _UnpackFwd()
pack_rev = _PackRev()
unpack_rev = _UnpackRev()


# --------------------------------------------------------------------------- #
# transports
# --------------------------------------------------------------------------- #


class TorchDistTransport:
    """Grouped point-to-point round over a ``torch.distributed`` process group
    (NCCL over NVLink on CUDA tensors; gloo on CPU tensors in the host tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.group = group
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    def _peer(self, r):
        import torch.distributed as dist

        if self.group is None or self.group is dist.group.WORLD:
            return int(r)
        return dist.get_global_rank(self.group, int(r))

    def exchange(self, send_chunks, send_peers, recv_chunks, recv_peers):
        """send_chunks[i] -> send_peers[i]; recv_chunks[i] <- recv_peers[i].
        Returns when the receives are ordered before later work on the current
        stream (NCCL) / complete (gloo)."""
        import torch.distributed as dist

        ops = []
        for t, p in zip(send_chunks, send_peers):
            if t.numel():
                ops.append(dist.P2POp(dist.isend, t, self._peer(p), group=self.group))
        for t, p in zip(recv_chunks, recv_peers):
            if t.numel():
                ops.append(dist.P2POp(dist.irecv, t, self._peer(p), group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class LocalCluster:
    """Several ranks emulated by threads of ONE process on ONE GPU (tests and
    single-GPU rehearsals of the multi-GPU path).  All device work is issued on
    the same stream, so host-side barriers give the required ordering."""

    def __init__(self, size: int):
        self.size = size
        self._barrier = threading.Barrier(size)
        self._mail = {}
        self._lock = threading.Lock()

    def transport(self, rank: int):
        return _LocalTransport(self, rank)

    def run(self, fn_):
        """Run ``fn_(rank, transport)`` on ``size`` threads; returns results."""
        import torch

        out = [None] * self.size
        err = []
        devno = torch.cuda.current_device() if torch.cuda.is_available() else None

        def body(r):
            try:
                if devno is not None:
                    torch.cuda.set_device(devno)
                out[r] = fn_(r, self.transport(r))
            except BaseException as e:  # pragma: no cover
                err.append(e)
                self._barrier.abort()

        th = [threading.Thread(target=body, args=(r,)) for r in range(self.size)]
        [t.start() for t in th]
        [t.join() for t in th]
        if err:
            raise err[0]
        return out


class _LocalTransport:
    def __init__(self, cluster: LocalCluster, rank: int):
        self.cluster, self.rank, self.size = cluster, rank, cluster.size

    def exchange(self, send_chunks, send_peers, recv_chunks, recv_peers):
        cl = self.cluster
        with cl._lock:
            for t, p in zip(send_chunks, send_peers):
                cl._mail[(self.rank, int(p))] = t
        cl._barrier.wait()  # every rank has posted (and enqueued its pack)
        for t, p in zip(recv_chunks, recv_peers):
            t.copy_(cl._mail[(int(p), self.rank)])
        cl._barrier.wait()  # every rank has enqueued its copies
        return None

    def barrier(self):
        self.cluster._barrier.wait()


def _as_transport(comm):
    if comm is not None and hasattr(comm, "exchange"):
        return comm
    return TorchDistTransport(comm)


# --------------------------------------------------------------------------- #
# fused multi-vector halo exchange
# --------------------------------------------------------------------------- #


def _cat_index(idx_list):
    import torch

    parts = []
    for a in idx_list:
        if isinstance(a, np.ndarray):
            parts.append(torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)))
        elif isinstance(a, torch.Tensor):
            parts.append(a.to(torch.int64).cpu())
        else:  # device array (Numba / CuPy)
            parts.append(torch.as_tensor(a, device="cuda").to(torch.int64).cpu())
    if not parts:
        return torch.zeros(0, dtype=torch.int64)
    return torch.cat(parts)


class HaloExchange:
    """Forward / reverse halo exchange of up to ``max_vecs`` vectors per round.

    ``owners_data = [idx_list, sizes, ranks]``: positions in MY ghost block of
    the dofs owned by each neighbour ``ranks[i]``;
    ``ghosts_data = [idx_list, sizes, ranks]``: MY owned dofs that are ghosts
    on each neighbour - exactly what ``utils.compute_scatterer_data`` returns
    (cuda/utils.py:8-78).
    """

    def __init__(self, comm, owners_data, ghosts_data, N, float_type, max_vecs: int = 3,
                 device=None):
        import torch

        self.transport = _as_transport(comm)
        self.N = int(N)
        self.dtype = np.dtype(float_type)
        self.tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[self.dtype]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.max_vecs = int(max_vecs)
        o_idx, o_size, o_ranks = owners_data
        g_idx, g_size, g_ranks = ghosts_data
        self.owner_ranks = [int(r) for r in np.asarray(o_ranks).ravel()]
        self.ghost_ranks = [int(r) for r in np.asarray(g_ranks).ravel()]
        self.owner_sizes = [int(s) for s in np.asarray(o_size).ravel()]
        self.ghost_sizes = [int(s) for s in np.asarray(g_size).ravel()]
        self.owner_idx = _cat_index(o_idx).to(self.device)  # into the ghost block
        self.ghost_idx = _cat_index(g_idx).to(self.device)  # owned local indices
        self.n_owner = int(self.owner_idx.numel())
        self.n_ghost = int(self.ghost_idx.numel())
        assert self.n_owner == sum(self.owner_sizes) and self.n_ghost == sum(self.ghost_sizes)
        k = self.max_vecs
        self._buf_owner = torch.empty(max(1, self.n_owner * k), dtype=self.tdtype, device=self.device)
        self._buf_ghost = torch.empty(max(1, self.n_ghost * k), dtype=self.tdtype, device=self.device)
        self._ptrs = (C.c_void_p * 4)()
        self._side = None
        self._ev = None

    # -- helpers --------------------------------------------------------------
    def _chunks(self, buf, sizes, k):
        out, off = [], 0
        for s in sizes:
            out.append(buf[off * k:(off + s) * k])
            off += s
        return out

    def _ptr_array(self, vecs):
        for i, v in enumerate(vecs):
            d = dev(v, self.dtype)
            self._ptrs[i] = d.ptr
        return self._ptrs

    def _pack(self, vecs, buf, index, n, offset):
        if n:
            check(fn("fus_pack_multi", self.dtype)(self._ptr_array(vecs), len(vecs), buf.data_ptr(),
                                                   index.data_ptr(), n, offset, current_stream()),
                  "fus_pack_multi")

    def _unpack(self, buf, vecs, index, n, offset, add):
        if n:
            check(fn("fus_unpack_multi", self.dtype)(buf.data_ptr(), self._ptr_array(vecs), len(vecs),
                                                     index.data_ptr(), n, offset, int(add),
                                                     current_stream()), "fus_unpack_multi")

    def _check(self, vecs):
        if not 1 <= len(vecs) <= self.max_vecs:
            raise ValueError(f"HaloExchange: 1..{self.max_vecs} vectors per round, got {len(vecs)}")

    # -- the two exchanges ----------------------------------------------------
    def forward(self, *vecs):
        """Owner values -> ghost copies (cuda/scatterer.py:191-277), k vectors."""
        self._check(vecs)
        k = len(vecs)
        self._pack(vecs, self._buf_ghost, self.ghost_idx, self.n_ghost, 0)
        self.transport.exchange(self._chunks(self._buf_ghost, self.ghost_sizes, k), self.ghost_ranks,
                                self._chunks(self._buf_owner, self.owner_sizes, k), self.owner_ranks)
        self._unpack(self._buf_owner, vecs, self.owner_idx, self.n_owner, self.N, add=False)

    def reverse(self, *vecs):
        """Ghost-region partial sums added into the owners
        (cuda/scatterer.py:104-188), k vectors."""
        self._check(vecs)
        k = len(vecs)
        self._pack(vecs, self._buf_owner, self.owner_idx, self.n_owner, self.N)
        self.transport.exchange(self._chunks(self._buf_owner, self.owner_sizes, k), self.owner_ranks,
                                self._chunks(self._buf_ghost, self.ghost_sizes, k), self.ghost_ranks)
        self._unpack(self._buf_ghost, vecs, self.ghost_idx, self.n_ghost, 0, add=True)

    # -- split phase: run the round on a side stream ---------------------------
    def _side_stream(self):
        import torch

        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def forward_begin(self, *vecs):
        import torch

        side = self._side_stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.forward(*vecs)

    def reverse_begin(self, *vecs):
        import torch

        side = self._side_stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.reverse(*vecs)

    def end(self):
        """Order the side-stream round before later work on the current stream."""
        import torch

        torch.cuda.current_stream().wait_stream(self._side_stream())

    forward_end = end
    reverse_end = end


# --------------------------------------------------------------------------- #
# reference-shaped factories
# --------------------------------------------------------------------------- #


def scatter_forward(comm, owners_data, ghosts_data, N, float_type):
    """``scatter(buffer)``: owner -> ghost overwrite - cuda/scatterer.py:191-277."""
    halo = HaloExchange(comm, owners_data, ghosts_data, N, float_type, max_vecs=3)

    def scatter(*buffers):
        halo.forward(*buffers)

    scatter.halo = halo
    return scatter


def scatter_reverse(comm, owners_data, ghosts_data, N, float_type):
    """``scatter(buffer)``: ghost -> owner add - cuda/scatterer.py:104-188."""
    halo = HaloExchange(comm, owners_data, ghosts_data, N, float_type, max_vecs=3)

    def scatter(*buffers):
        halo.reverse(*buffers)

    scatter.halo = halo
    return scatter


# --------------------------------------------------------------------------- #
# halo exchange over NVLink peer memory (no NCCL in the data path)
# --------------------------------------------------------------------------- #


class SymmFabric:
    """Peer-addressable device memory for one process per GPU: a symmetric
    arena (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped
    into every peer over NVLink/NVSwitch) plus its signal-pad barrier."""

    def __init__(self, arena_bytes: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.size = dist.get_world_size(self.group)
        t = torch.tensor([int(arena_bytes)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)  # same size everywhere
        self.arena_bytes = (int(t.item()) + 255) // 256 * 256
        self.arena = symm.empty(self.arena_bytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.arena, self.group)
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        self._used = 0
        self._channel = 0

    def alloc(self, numel: int, tdtype, slot_numel: int | None = None):
        """Carve a tensor out of the arena.  Every rank must make the same
        sequence of calls; ``slot_numel`` (the maximum over ranks) keeps the
        offsets identical when the ranks' sizes differ."""
        import torch

        item = torch.empty(0, dtype=tdtype).element_size()
        slot = (int(slot_numel if slot_numel is not None else numel) * item + 255) // 256 * 256
        if self._used + slot > self.arena_bytes:
            raise RuntimeError("SymmFabric: arena exhausted")
        out = self.arena[self._used:self._used + numel * item].view(tdtype)
        self._used += slot
        out.zero_()
        return out

    def peer_ptr(self, rank: int, t):
        return self.base[rank] + (t.data_ptr() - self.base[self.rank])

    def barrier(self):
        self.hdl.barrier(channel=0)

    def max_over_ranks(self, v: int) -> int:
        import torch
        import torch.distributed as dist

        t = torch.tensor([int(v)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def exchange_index_lists(self, send, dests, recv_sizes, sources):
        from . import utils

        return utils.exchange_index_lists(send, dests, recv_sizes, sources, self.group)


class _LocalFabric:
    """The same interface with the ranks emulated by threads of one process on
    one GPU: a peer's memory is just another local buffer."""

    def __init__(self, cluster: LocalCluster, rank: int, arena_bytes: int):
        import torch

        self.cluster, self.rank, self.size = cluster, rank, cluster.size
        self.arena_bytes = (self.max_over_ranks(arena_bytes) + 255) // 256 * 256
        self.arena = torch.zeros(self.arena_bytes, dtype=torch.uint8, device="cuda")
        with cluster._lock:
            cluster._mail[("arena", rank)] = self.arena
        cluster._barrier.wait()
        self.base = [cluster._mail[("arena", r)].data_ptr() for r in range(self.size)]
        cluster._barrier.wait()
        self._used = 0

    alloc = SymmFabric.alloc
    peer_ptr = SymmFabric.peer_ptr

    def barrier(self):
        # every rank has ENQUEUED its work on the shared stream: stream order does the rest
        self.cluster._barrier.wait()

    def max_over_ranks(self, v: int) -> int:
        cl = self.cluster
        with cl._lock:
            cl._mail[("max", self.rank)] = int(v)
        cl._barrier.wait()
        out = max(cl._mail[("max", r)] for r in range(self.size))
        cl._barrier.wait()
        return out

    def exchange_index_lists(self, send, dests, recv_sizes, sources):
        cl = self.cluster
        with cl._lock:
            for s, d in zip(send, dests):
                cl._mail[("idx", self.rank, int(d))] = s
        cl._barrier.wait()
        out = [cl._mail[("idx", int(s), self.rank)] for s in sources]
        cl._barrier.wait()
        return out


def local_fabric(cluster: LocalCluster, rank: int, arena_bytes: int):
    return _LocalFabric(cluster, rank, arena_bytes)


class P2PHaloExchange:
    """Halo exchange fused into two kernels over NVLink peer memory.

    forward  = ``fus_halo_put`` (owned values stored straight into the peers'
    ghost slots) + one barrier; reverse = barrier + ``fus_halo_get_add`` (ghost
    partial sums loaded straight from the peers and added) + barrier.  No
    staging buffers, no NCCL, no pack/unpack launches: 2 + 3 small launches per
    RK stage instead of 2 x (pack + grouped send/recv + unpack).

    Vectors that take part (``un``, ``vn``, ``b``, ``m``) must be allocated
    with ``alloc`` so that every peer can address them.  Because peers write
    into the ghost slots directly, the owner of a vector must not write its own
    ghost region between exchanges (the fused solvers update owned entries only
    when this exchange is used).
    """

    p2p = True

    def __init__(self, fabric, owners_data, ghosts_data, N, nghost, float_type):
        import torch

        self.fabric = fabric
        self.N, self.nghost = int(N), int(nghost)
        self.dtype = np.dtype(float_type)
        self.tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[self.dtype]
        o_idx, o_size, o_ranks = owners_data
        g_idx, g_size, g_ranks = ghosts_data
        self.ghost_ranks = [int(r) for r in np.asarray(g_ranks).ravel()]
        ghost_sizes = [int(s) for s in np.asarray(g_size).ravel()]
        dev = torch.device("cuda", torch.cuda.current_device())
        self.idx = _cat_index(g_idx).to(dev)
        self.n = int(self.idx.numel())
        # where each of my shared dofs sits in the neighbour's vector: the neighbour's
        # owners_idx list for me (same order as my ghosts_idx, cuda/utils.py:57-73) + its N
        send = [np.ascontiguousarray(np.asarray(ix, dtype=np.int64) + self.N) for ix in o_idx]
        recv = fabric.exchange_index_lists(send, [int(r) for r in np.asarray(o_ranks).ravel()],
                                           ghost_sizes, self.ghost_ranks)
        self.remote_pos = _cat_index([np.asarray(r, dtype=np.int64) for r in recv]).to(dev)
        seg = np.repeat(np.arange(len(ghost_sizes), dtype=np.int32), ghost_sizes)
        self.entry_seg = torch.from_numpy(seg).to(dev)
        self.slot = fabric.max_over_ranks(self.N + self.nghost)
        self._tables = {}
        self._ptrs = (C.c_void_p * 4)()

    @staticmethod
    def arena_bytes(ndofs_local: int, float_type, nvec: int = 12) -> int:
        """Arena size for ``nvec`` exchanged vectors of ``ndofs_local`` entries."""
        return nvec * ((int(ndofs_local) * np.dtype(float_type).itemsize + 255) // 256 * 256 + 256)

    def alloc(self):
        """A zeroed ``(N + nghost,)`` vector in peer-addressable memory."""
        return self.fabric.alloc(self.N + self.nghost, self.tdtype, self.slot)

    def _peer_table(self, vecs):
        import torch

        key = tuple(v.data_ptr() for v in vecs)
        tab = self._tables.get(key)
        if tab is None:
            rows = [[self.fabric.peer_ptr(q, v) for v in vecs] for q in self.ghost_ranks]
            # int64 view of the (unsigned) device addresses
            tab = torch.tensor(np.array(rows, dtype=np.uint64).astype(np.int64).reshape(-1)
                               if rows else np.zeros(0, np.int64), device=self.idx.device)
            self._tables[key] = tab
        return tab

    def _launch(self, name, vecs):
        if not 1 <= len(vecs) <= 4:
            raise ValueError("P2PHaloExchange: 1..4 vectors per round")
        if self.n:
            tab = self._peer_table(vecs)
            for i, v in enumerate(vecs):
                self._ptrs[i] = dev(v, self.dtype).ptr
            check(fn(name, self.dtype)(self._ptrs, len(vecs), tab.data_ptr(), self.idx.data_ptr(),
                                       self.remote_pos.data_ptr(), self.entry_seg.data_ptr(), self.n,
                                       current_stream()), name)

    def barrier(self):
        self.fabric.barrier()

    def forward(self, *vecs):
        """Owner values -> every ghost copy (cuda/scatterer.py:191-277)."""
        self._launch("fus_halo_put", vecs)
        self.fabric.barrier()

    def reverse(self, *vecs):
        """Ghost partial sums added into the owners (cuda/scatterer.py:104-188).
        The ghost region keeps its values, as in the reference."""
        self.fabric.barrier()  # every rank's ghost sums are complete
        self._launch("fus_halo_get_add", vecs)
        self.fabric.barrier()  # every rank has read them: they may be overwritten now
