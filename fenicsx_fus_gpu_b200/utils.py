"""
Utils - index maps for the halo exchange and facet integration domains
======================================================================

Drop-in for the hot-path half of ``/root/reference/cuda/utils.py``:

* ``compute_scatterer_data(index_map)`` (:8-78) - same return value, bit for
  bit, but vectorised: the reference's O(ranks * size_local) Python loop over
  ``shared_dofs.links(dof)`` (:43-47) is one ``np.unique(..., return_counts)``,
  and the ghost-index exchange runs over ``torch.distributed`` (one process
  per GPU) instead of mpi4py.
* ``facet_integration_domain(facets, mesh)`` (:81-114) - vectorised over the
  facets, same ``(cell, local facet)`` rows.
* ``compute_diffusivity_of_sound`` (:157-162).

Pure host integer work; nothing here touches the GPU.
"""

from __future__ import annotations

import numpy as np


def _owner_lists(index_map):
    """Ghost positions grouped by owner (cuda/utils.py:23-37).  ``np.argsort``
    with the default (unstable) kind on the same input, as the reference does,
    so the order inside a group is reproduced exactly."""
    owners = np.asarray(index_map.owners)
    unique_owners, owners_size = np.unique(owners, return_counts=True)
    owners_argsorted = np.argsort(owners)
    owners_offsets = np.insert(np.cumsum(owners_size), 0, 0)
    owners_idx = [owners_argsorted[owners_offsets[i]:owners_offsets[i + 1]]
                  for i in range(unique_owners.size)]
    return owners_idx, owners_size, unique_owners


def _ghost_ranks(index_map):
    """Ranks that ghost my owned dofs and how many each (cuda/utils.py:39-52).

    The reference appends ``shared_rank`` once per owned dof whose destination
    list contains it, then takes ``np.unique(..., return_counts=True)``: that is
    the histogram of ``index_to_dest_ranks().array``.
    """
    shared = index_map.index_to_dest_ranks()
    arr = np.asarray(shared.array)
    if arr.size == 0:
        # np.unique(np.array([])) in the reference: float64 empties
        return np.unique(np.array([]), return_counts=True)
    return np.unique(arr, return_counts=True)


def compute_scatterer_data(index_map, comm=None):
    """Extract scatterer data - cuda/utils.py:8-78.

    ``comm`` is a ``torch.distributed`` process group (default: the world
    group).  With a single process (or no initialised process group) there are
    no neighbours and both lists are empty.

    Returns ``owners_data = [owners_idx, owners_size, unique_owners]`` and
    ``ghosts_data = [ghosts_idx, ghosts_size, unique_ghosts]``.
    """
    owners_idx, owners_size, unique_owners = _owner_lists(index_map)
    unique_ghosts, ghosts_size = _ghost_ranks(index_map)
    ghosts_np = np.asarray(index_map.ghosts)

    send = [np.ascontiguousarray(ghosts_np[idx], dtype=np.int64) for idx in owners_idx]
    recv = exchange_index_lists(send, unique_owners, ghosts_size, unique_ghosts, comm)
    ghosts_idx = [r - index_map.local_range[0] for r in recv]

    owners_data = [owners_idx, owners_size, unique_owners]
    ghosts_data = [ghosts_idx, ghosts_size, unique_ghosts]
    return owners_data, ghosts_data


def exchange_index_lists(send, dests, recv_sizes, sources, comm=None):
    """The ``Isend/Irecv/Waitall`` round of cuda/utils.py:57-71 over
    ``torch.distributed`` point-to-point ops (gloo on CPU tensors; for NCCL
    the int64 lists are staged through the device)."""
    if len(send) == 0 and len(sources) == 0:
        return []
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("index map has neighbours but torch.distributed is not initialised")
    backend = dist.get_backend(comm)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    sbuf = [torch.from_numpy(s).to(device) for s in send]
    rbuf = [torch.empty(int(n), dtype=torch.int64, device=device) for n in recv_sizes]
    ops = []
    for t, dst in zip(sbuf, dests):
        ops.append(dist.P2POp(dist.isend, t, _global_rank(int(dst), comm), group=comm))
    for t, src in zip(rbuf, sources):
        ops.append(dist.P2POp(dist.irecv, t, _global_rank(int(src), comm), group=comm))
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    if device.type == "cuda":
        torch.cuda.synchronize()
    return [t.cpu().numpy() for t in rbuf]


def _global_rank(group_rank: int, comm):
    import torch.distributed as dist

    if comm is None or comm is dist.group.WORLD:
        return group_rank
    return dist.get_global_rank(comm, group_rank)


def compute_scatterer_data_all(index_maps):
    """All ranks at once in one process (used where ranks are emulated on one
    host: fixtures, single-process multi-partition tests)."""
    pre = []
    sent = {}
    for rank, im in enumerate(index_maps):
        owners_idx, owners_size, unique_owners = _owner_lists(im)
        unique_ghosts, ghosts_size = _ghost_ranks(im)
        ghosts_np = np.asarray(im.ghosts)
        for idx, owner in zip(owners_idx, unique_owners):
            sent[(rank, int(owner))] = np.ascontiguousarray(ghosts_np[idx], dtype=np.int64)
        pre.append((owners_idx, owners_size, unique_owners, ghosts_size, unique_ghosts))
    out = []
    for rank, im in enumerate(index_maps):
        owners_idx, owners_size, unique_owners, ghosts_size, unique_ghosts = pre[rank]
        ghosts_idx = [sent[(int(g), rank)] - im.local_range[0] for g in unique_ghosts]
        out.append(([owners_idx, owners_size, unique_owners], [ghosts_idx, ghosts_size, unique_ghosts]))
    return out


def facet_integration_domain(facets, mesh):
    """``boundary_data[i] = (cell, local facet)`` - cuda/utils.py:81-114.

    ``mesh`` is a DOLFINx mesh (``mesh.topology.connectivity``) - the same
    look-ups as the reference, vectorised over the facets through the
    adjacency arrays instead of per-facet ``links`` calls.  (Box meshes of the
    synthetic substrate use ``substrate.boundary_facets`` which yields the
    same rows directly.)
    """
    facets = np.asarray(facets, dtype=np.int32)
    tdim = mesh.topology.dim
    c2f = mesh.topology.connectivity(tdim, tdim - 1)
    f2c = mesh.topology.connectivity(tdim - 1, tdim)
    f2c_off = np.asarray(f2c.offsets)
    f2c_arr = np.asarray(f2c.array)
    cells = f2c_arr[f2c_off[facets]].astype(np.int32)  # links(facet)[0]
    c2f_off = np.asarray(c2f.offsets)
    c2f_arr = np.asarray(c2f.array)
    nfc = int(c2f_off[1] - c2f_off[0]) if c2f_off.size > 1 else 0
    boundary_data = np.zeros((facets.size, 2), dtype=np.int32)
    if facets.size == 0:
        return boundary_data
    # cells of one type have the same number of facets: rows of the cell->facet table
    rows = c2f_arr[c2f_off[cells][:, None] + np.arange(nfc)[None, :]]
    local = np.argmax(rows == facets[:, None], axis=1)  # first match, as np.where(...)[0][0]
    boundary_data[:, 0] = cells
    boundary_data[:, 1] = local
    return boundary_data


def shared_last_numbering(nlocal: int, nghost: int, ghosts_data):
    """A local renumbering that puts the owned dofs other ranks ghost (the ``ghosts_data`` index
    lists of ``compute_scatterer_data``) at the END of the owned block, contiguous and in their
    old relative order; the other owned dofs keep their order in front, ghosts keep their indices.

    Why: the fused solvers close those dofs in a separate kernel (after the reverse halo) beside
    the bulk close.  In DOLFINx' numbering they are scattered - one 8-byte access per 32-byte
    sector on a face normal to the fastest direction - and that scattered traffic is what slowed the
    bulk close on the ranks that own faces (profiles/r02_multigpu_timeline.md).  Contiguous, the
    shared block is a coalesced tail and the bulk close a plain prefix.  The index map (global
    numbering, ghost order on the neighbours) is untouched: only this rank's private indices move.

    Returns ``(perm, ghosts_data')``: ``perm[old] = new`` over ``nlocal + nghost`` entries and the
    index lists in the new numbering (a new list; the sizes / ranks arrays are shared)."""
    g_idx, g_size, g_ranks = ghosts_data
    perm = np.arange(nlocal + nghost, dtype=np.int64)
    lists = [np.asarray(ix, dtype=np.int64) for ix in g_idx]
    if not lists or sum(a.size for a in lists) == 0:
        return perm, ghosts_data
    shared = np.zeros(nlocal, dtype=bool)
    shared[np.concatenate(lists)] = True
    ns = int(shared.sum())
    perm[:nlocal][~shared] = np.arange(nlocal - ns, dtype=np.int64)
    perm[:nlocal][shared] = nlocal - ns + np.arange(ns, dtype=np.int64)
    return perm, [[perm[a] for a in lists], g_size, g_ranks]


def colour_cells(connectivity, seed: int = 0) -> np.ndarray:
    """Greedy distance-1 colouring of the cells of a conforming mesh: no two
    cells of one colour share an entry of ``connectivity``.

    ``connectivity`` is ``(Nc, k)`` int: the geometry dofmap ``x_dofs`` (8
    vertices per hexahedron - two cells share a dof iff they share a vertex, so
    this is enough and 16x cheaper at degree 4) or the dofmap itself.  The
    reference scatters with atomics only (``cuda/operators.py:190``); the
    colouring backs the deterministic ``FUS_NO_ATOMICS`` launches
    (``operators.stiffness_operator(..., colour_offsets=...)``).

    Each colour is a maximal independent set grown Luby-style: among the
    candidates, a cell joins when it holds the highest (seeded, random)
    priority on every one of its entries; its neighbours leave the candidate
    set; repeat until no candidate is left, then open the next colour.
    Vectorised over cells; returns ``(Nc,) int32`` colours ``0..ncolours-1``.
    """
    conn = np.ascontiguousarray(connectivity)
    nc, k = conn.shape
    colour = np.full(nc, -1, dtype=np.int32)
    if nc == 0:
        return colour
    nnode = int(conn.max()) + 1
    prio = np.random.default_rng(seed).permutation(nc).astype(np.int64) + 1
    remaining = np.arange(nc)
    c = 0
    while remaining.size:
        cand = remaining
        taken = np.zeros(nnode, dtype=bool)  # entries touched by this colour so far
        while cand.size:
            dm = conn[cand]
            pc = prio[cand]
            best = np.zeros(nnode, dtype=np.int64)
            np.maximum.at(best, dm.ravel(), np.repeat(pc, k))
            win = (best[dm] == pc[:, None]).all(axis=1)
            sel = cand[win]
            colour[sel] = c
            taken[conn[sel].ravel()] = True
            rest = cand[~win]
            cand = rest[~taken[conn[rest]].any(axis=1)]
        remaining = remaining[colour[remaining] < 0]
        c += 1
    return colour


def colour_order(colours):
    """``(perm, offsets)``: the stable permutation that makes every colour a
    contiguous block of cells and the block boundaries (``ncolours + 1`` int64).
    Apply ``perm`` to every per-cell array (dofmap, G, detJ, constants)."""
    colours = np.asarray(colours)
    perm = np.argsort(colours, kind="stable")
    counts = np.bincount(colours, minlength=int(colours.max()) + 1 if colours.size else 0)
    return perm, np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


def global_min(value: float, comm=None) -> float:
    """Minimum of ``value`` over the ranks - the ``comm.Allreduce(hmin, mesh_size, op=MPI.MIN)``
    of cuda/demo_linear_box.py:103-108 over ``torch.distributed`` (gloo: CPU tensor; nccl: staged
    through the device).  Without an initialised process group it is the value itself."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(comm) == 1:
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(comm) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=comm)
    return float(t.item())


def compute_eval_params(mesh, points, float_type):
    """``(points_on_proc, cells)`` of cuda/utils.py:117-154 - implemented in ``sampling.py``
    (bin grid + Newton pull-back instead of DOLFINx bounding-box trees); re-exported here because the
    reference's piston / bowl demos import it from ``utils``."""
    from .sampling import compute_eval_params as _impl

    return _impl(mesh, points, float_type)


def compute_diffusivity_of_sound(w0: float, c0: float, alpha: float) -> float:
    """``delta = 2 alpha c0^3 / w0^2`` with alpha in dB/m converted to Np/m -
    cuda/utils.py:157-162."""
    return 2.0 * (alpha / 20.0 * np.log(10.0)) * c0**3 / w0**2
