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
