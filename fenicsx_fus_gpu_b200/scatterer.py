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


SLOT = 1 << 21


class SymmFabric:
    """Peer-addressable device memory for one process per GPU: a symmetric
    arena (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped
    into every peer over NVLink/NVSwitch).  torch provides the allocation and
    the rendezvous only; signalling and data movement are this library's own
    kernels (csrc/halo.cu)."""

    emulated = False

    def __init__(self, arena_bytes: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.size = dist.get_world_size(self.group)
        t = torch.tensor([int(arena_bytes)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)  # same size everywhere
        self.arena_bytes = (int(t.item()) + SLOT - 1) // SLOT * SLOT
        self.arena = symm.empty(self.arena_bytes, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.arena.zero_()
        self.hdl = symm.rendezvous(self.arena, self.group)
        self.base = [int(p) for p in self.hdl.buffer_ptrs]
        self._used = 0
        torch.cuda.synchronize()
        dist.barrier(group=self.group)  # every arena is zeroed before anybody stores into it

    def alloc(self, numel: int, tdtype, slot_numel: int | None = None):
        """Carve a tensor out of the arena.  Every rank must make the same
        sequence of calls; ``slot_numel`` (the maximum over ranks) keeps the
        offsets identical when the ranks' sizes differ."""
        import torch

        item = torch.empty(0, dtype=tdtype).element_size()
        # Slots are multiples of 2 MiB: vectors that start at the same offset of a 2 MiB page stream
        # ~2 % faster through the 12-vector close kernel than vectors at arbitrary offsets (measured,
        # tools/close_bench.py: plain device memory carved at 256-byte granularity shows the same loss)
        slot = (int(slot_numel if slot_numel is not None else numel) * item + SLOT - 1) // SLOT * SLOT
        if self._used + slot > self.arena_bytes:
            raise RuntimeError("SymmFabric: arena exhausted")
        out = self.arena[self._used:self._used + numel * item].view(tdtype)
        self._used += slot
        out.zero_()
        return out

    def peer_ptr(self, rank: int, t):
        return self.base[rank] + (t.data_ptr() - self.base[self.rank])

    def host_sync(self):
        """Real GPUs run concurrently: nothing to do (see ``_LocalFabric.host_sync``)."""

    def host_barrier(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier(group=self.group)

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
    one GPU: a peer's memory is just another local buffer.  All ranks enqueue
    on ONE stream, so a kernel that waits for a flag must be enqueued after the
    kernel that raises it: ``host_sync`` (a thread barrier) is called by the
    exchange before it enqueues a wait."""

    emulated = True

    def __init__(self, cluster: LocalCluster, rank: int, arena_bytes: int):
        import torch

        self.cluster, self.rank, self.size = cluster, rank, cluster.size
        self.arena_bytes = (self.max_over_ranks(arena_bytes) + SLOT - 1) // SLOT * SLOT
        self.arena = torch.zeros(self.arena_bytes, dtype=torch.uint8, device="cuda")
        with cluster._lock:
            cluster._mail[("arena", rank)] = self.arena
        cluster._barrier.wait()
        self.base = [cluster._mail[("arena", r)].data_ptr() for r in range(self.size)]
        cluster._barrier.wait()
        self._used = 0

    alloc = SymmFabric.alloc
    peer_ptr = SymmFabric.peer_ptr

    def host_sync(self):
        # every rank has ENQUEUED its work on the shared stream: stream order does the rest
        self.cluster._barrier.wait()

    host_barrier = host_sync

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


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class P2PHaloExchange:
    """Halo exchange fused into kernels over NVLink peer memory, behind the
    C-ABI handle ``fus_halo_t`` (include/fus_b200.h, csrc/halo.cu).

    Data: ``put`` stores owned values straight into the neighbours' ghost
    slots; ``get_add`` loads the neighbours' ghost partial sums straight from
    their vectors and adds them.  Ordering: per-neighbour epoch flags raised
    and awaited by the kernels themselves (``st.release.sys`` /
    ``ld.acquire.sys``) - no global barrier, no NCCL, no host synchronisation,
    capturable in a CUDA graph.

    * ``forward(*vecs)`` / ``reverse(*vecs)`` are the reference's
      ``scatter_forward`` / ``scatter_reverse`` (cuda/scatterer.py:104-277),
      safe in any call sequence that all ranks make identically.
    * ``put / wait_forward / signal_reverse / get_add`` are the split-phase
      pieces the fused solvers interleave with interior-cell work; ``fork`` /
      ``join`` / ``side`` put the exchange on a second stream.

    Vectors that take part must be allocated with ``alloc`` (same offset of a
    symmetric arena on every rank).  Peers write the ghost slots directly, so
    the owner of a vector must not write its own ghost region between
    exchanges (the fused solvers update owned entries only).
    """

    p2p = True

    def __init__(self, fabric, owners_data, ghosts_data, N, nghost, float_type):
        import torch

        lib = _lib.lib()
        self.fabric = fabric
        self.N, self.nghost = int(N), int(nghost)
        self.dtype = np.dtype(float_type)
        self.tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[self.dtype]
        o_idx, o_size, o_ranks = owners_data
        g_idx, g_size, g_ranks = ghosts_data
        self.ghost_ranks = np.ascontiguousarray(np.asarray(g_ranks).ravel(), dtype=np.int32)
        self.owner_ranks = np.ascontiguousarray(np.asarray(o_ranks).ravel(), dtype=np.int32)
        ghost_sizes = [int(s) for s in np.asarray(g_size).ravel()]
        idx = _cat_index(g_idx).numpy()
        self.n = int(idx.size)
        # where each of my shared dofs sits in the neighbour's vector: the neighbour's
        # owners_idx list for me (same order as my ghosts_idx, cuda/utils.py:57-73) + its N
        send = [np.ascontiguousarray(np.asarray(ix, dtype=np.int64) + self.N) for ix in o_idx]
        recv = fabric.exchange_index_lists(send, [int(r) for r in self.owner_ranks], ghost_sizes,
                                           [int(r) for r in self.ghost_ranks])
        remote_pos = _cat_index([np.asarray(r, dtype=np.int64) for r in recv]).numpy()
        seg = np.repeat(np.arange(len(ghost_sizes), dtype=np.int32), ghost_sizes)
        self.slot = fabric.max_over_ranks(self.N + self.nghost)
        world = fabric.size
        self.pad = fabric.alloc(int(lib.fus_halo_pad_bytes(world)), torch.uint8)
        peer_pad = np.array([fabric.peer_ptr(q, self.pad) for q in range(world)], dtype=np.uint64)
        peer_delta = np.array([fabric.base[q] - fabric.base[fabric.rank] for q in range(world)], dtype=np.int64)
        keep = [np.ascontiguousarray(a) for a in (idx.astype(np.int64), remote_pos.astype(np.int64), seg)]
        d = _lib.HaloDesc()
        d.rank, d.world = fabric.rank, world
        d.n_ghost_ranks, d.ghost_ranks = int(self.ghost_ranks.size), self.ghost_ranks.ctypes.data
        d.n_owner_ranks, d.owner_ranks = int(self.owner_ranks.size), self.owner_ranks.ctypes.data
        d.n, d.idx, d.remote_pos, d.entry_seg = self.n, keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data
        d.size_local, d.num_ghosts = self.N, self.nghost
        d.signal_pad = self.pad.data_ptr()
        d.peer_pad, d.peer_delta = peer_pad.ctypes.data, peer_delta.ctypes.data
        d.close_group = 16 // self.dtype.itemsize  # one 16-byte pack of the close kernel
        torch.cuda.synchronize()  # the pad is zeroed before the handle (and any neighbour) can use it
        h = C.c_void_p()
        check(lib.fus_halo_create(C.byref(d), C.byref(h)), "fus_halo_create")
        self._h = h
        self._lib = lib
        self.nshared = int(lib.fus_halo_num_shared(h))
        self.shared_mask = lib.fus_halo_shared_mask(h)  # device address of the bitmask
        # >= 0: the shared dofs are the tail [shared_tail, N) of the owned block (utils.shared_last_numbering),
        # so the bulk close is a plain prefix and needs no mask
        self.shared_tail = int(lib.fus_halo_shared_tail(h))
        self._side = None
        self.use_side = True  # put the exchange kernels on a second stream (False: A/B measurements)
        fabric.host_barrier()  # every rank's pad and handle exist before the first signal

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.fus_halo_destroy(h)
            except Exception:  # interpreter shutdown
                pass

    @staticmethod
    def arena_bytes(ndofs_local: int, float_type, nvec: int = 12) -> int:
        """Arena size for ``nvec`` exchanged vectors of ``ndofs_local`` entries (+ the signal pad)."""
        return (nvec * ((int(ndofs_local) * np.dtype(float_type).itemsize + SLOT - 1) // SLOT) + 1) * SLOT

    def alloc(self):
        """A zeroed ``(N + nghost,)`` vector in peer-addressable memory."""
        return self.fabric.alloc(self.N + self.nghost, self.tdtype, self.slot)

    @property
    def handle(self):
        return self._h

    def _vecs(self, vecs, lo=1):
        if not lo <= len(vecs) <= 4:
            raise ValueError(f"P2PHaloExchange: {lo}..4 vectors per round")
        arr = (C.c_void_p * 4)()
        for i, v in enumerate(vecs):
            arr[i] = dev(v, self.dtype).ptr
        return arr

    # -- split-phase pieces ------------------------------------------------------
    def put(self, *vecs):
        """Owner values -> the neighbours' ghost slots, then the FWD epoch."""
        check(fn("fus_halo_put", self.dtype)(self._h, self._vecs(vecs), len(vecs), current_stream()), "fus_halo_put")

    def wait_forward(self, *zero_vecs):
        """Wait for the puts of every owner of my ghosts; then clear the ghost part of ``zero_vecs``."""
        self.fabric.host_sync()
        check(fn("fus_halo_wait_forward", self.dtype)(self._h, self._vecs(zero_vecs, 0), len(zero_vecs),
                                                      current_stream()), "fus_halo_wait_forward")

    def signal_reverse(self):
        """REV epoch -> the owners of my ghosts: my ghost partial sums are complete."""
        check(self._lib.fus_halo_signal_reverse(self._h, current_stream()), "fus_halo_signal_reverse")

    def get_add(self, *vecs):
        """Wait for every ghosting neighbour's REV epoch, then add their partial sums."""
        self.fabric.host_sync()
        check(fn("fus_halo_get_add", self.dtype)(self._h, self._vecs(vecs), len(vecs), current_stream()),
              "fus_halo_get_add")

    def wait_reverse(self):
        """Wait (one warp) for every ghosting neighbour's REV epoch - followed by a kernel that
        gathers their partial sums itself (the solvers' fused close of the shared dofs)."""
        self.fabric.host_sync()
        check(self._lib.fus_halo_wait_reverse(self._h, current_stream()), "fus_halo_wait_reverse")

    def bulk_close(self):
        """Keyword arguments for the solvers' bulk close beside ``fus_rk_close_shared``: the prefix
        ``n = shared_tail`` without a mask when the shared dofs are a contiguous tail, else the
        whole owned block with the skip mask."""
        if self.shared_tail >= 0:
            return dict(n=self.shared_tail)
        return dict(skip=self.shared_mask)

    def arm_stiffness_wait(self, first_interface_cell):
        """The next stiffness launch of this thread waits in-kernel, before its first interface
        batch, for the puts of every owner of my ghosts (``fus_stiffness_arm_halo_wait``).
        (Emulated ranks: call ``sync_point()`` once before the stage's launches.)"""
        check(self._lib.fus_stiffness_arm_halo_wait(self._h, int(first_interface_cell)), "fus_stiffness_arm_halo_wait")

    def sync_point(self):
        """Emulated ranks only (one shared stream): every rank has enqueued its signalling kernels
        before anybody enqueues a kernel that waits for them.  Nothing on real GPUs."""
        self.fabric.host_sync()

    def barrier(self):
        """Neighbour barrier on the current stream (a host barrier when the ranks are emulated
        on one stream, where a waiting kernel would block the kernel it waits for)."""
        if self.fabric.emulated:
            self.fabric.host_sync()
        else:
            check(self._lib.fus_halo_barrier(self._h, current_stream()), "fus_halo_barrier")

    def status(self):
        """Raises if a wait ran into the time-out (a neighbour died).  Synchronises."""
        check(self._lib.fus_halo_status(self._h), "fus_halo_status")

    # -- the reference's two exchanges -----------------------------------------------
    def forward(self, *vecs):
        """Owner values -> every ghost copy (cuda/scatterer.py:191-277)."""
        if self.fabric.emulated:
            self.barrier()
            self.put(*vecs)
            self.wait_forward()
        else:
            check(fn("fus_halo_forward", self.dtype)(self._h, self._vecs(vecs), len(vecs), current_stream()),
                  "fus_halo_forward")

    def reverse(self, *vecs):
        """Ghost partial sums added into the owners (cuda/scatterer.py:104-188).
        The ghost region keeps its values, as in the reference."""
        if self.fabric.emulated:
            self.signal_reverse()
            self.get_add(*vecs)
            self.barrier()
        else:
            check(fn("fus_halo_reverse", self.dtype)(self._h, self._vecs(vecs), len(vecs), current_stream()),
                  "fus_halo_reverse")

    # -- second stream for the exchange (real GPUs only) -----------------------------
    @property
    def concurrent(self):
        return self.use_side and not self.fabric.emulated

    def _side_stream(self):
        import torch

        if self._side is None:
            self._side = torch.cuda.Stream(priority=-1)
        return self._side

    def fork(self):
        """Later work under ``side()`` is ordered after everything enqueued so far."""
        import torch

        if self.concurrent:
            self._side_stream().wait_stream(torch.cuda.current_stream())

    def side(self):
        import torch

        return torch.cuda.stream(self._side_stream()) if self.concurrent else _NullCtx()

    def join(self):
        """Later work on the current stream is ordered after the work under ``side()``."""
        import torch

        if self.concurrent:
            torch.cuda.current_stream().wait_stream(self._side_stream())
