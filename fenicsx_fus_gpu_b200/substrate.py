"""
Synthetic host substrate (numpy only)
=====================================

The reference takes its mesh, dofmap, index map, quadrature rule and basis
tables from DOLFINx / Basix (``cuda/demo_linear_box.py:90-113, 167-180,
232-239, 256-299, 348-357``).  Neither library exists in this image, so this
module builds the same *arrays*, in the same layouts, for box meshes of
hexahedra.  Live DOLFINx objects can be passed to the operators unchanged;
this module is only what stands in for them when they are absent.

Layout conventions:

* 1-D point / dof order is ``[0, 1, interior...]`` (``order="basix"``, the
  Basix vertex-first convention) or plain ascending (``order="ascending"``).
* hexahedron local dof / quadrature index ``q = i*n*n + j*n + k`` with ``i``
  the x-direction (``cuda/operators.py:123``).
* P1 geometry vertices ``v = vx + 2*vy + 4*vz``.
* reference facets ordered z=0, y=0, x=0, x=1, y=1, z=1
  (``cuda/precompute.py:49-59``, ``cuda/demo_linear_box.py:291-296``).
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# Basix quadrature degree used by the reference for basis degree P
# (``cuda/demo_linear_box.py:70-80``).  GLL degree Q has (Q+3)//2 points per
# direction, which is P+1 for every entry below.
QUADRATURE_DEGREE = {2: 3, 3: 4, 4: 6, 5: 8, 6: 10, 7: 12, 8: 14, 9: 16, 10: 18}


# --------------------------------------------------------------------------- #
# 1-D Gauss-Lobatto-Legendre rule and Lagrange derivative table
# --------------------------------------------------------------------------- #


def _legendre(N: int, x: np.ndarray):
    """P_N(x) and P_N'(x) by the three-term recurrence (x in [-1, 1])."""
    p0 = np.ones_like(x)
    p1 = x.copy()
    if N == 0:
        return p0, np.zeros_like(x)
    for k in range(2, N + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    with np.errstate(divide="ignore", invalid="ignore"):
        dp = N * (p0 - x * p1) / (1.0 - x * x)
    return p1, dp


def gll_points_weights(n: int):
    """GLL rule with ``n`` points on [0, 1], ascending order, float64.

    Nodes are the roots of (1-x^2) P'_{n-1}(x); weights 2/(N(N+1) P_N(x)^2),
    both mapped from [-1, 1] to [0, 1].
    """
    N = n - 1
    x = -np.cos(np.pi * np.arange(n) / N)  # Chebyshev-Lobatto start
    for _ in range(100):
        # Newton on q(x) = P_{N+1}(x) - P_{N-1}(x)  (same roots as (1-x^2)P_N')
        pN, _ = _legendre(N, x)
        pNm1, _ = _legendre(N - 1, x)
        pNp1, _ = _legendre(N + 1, x)
        q = pNp1 - pNm1
        dq = (2 * N + 1) * pN
        dx = q / dq
        x = x - dx
        if np.max(np.abs(dx)) < 1e-16:
            break
    x[0], x[-1] = -1.0, 1.0
    x = 0.5 * (x - x[::-1])  # enforce symmetry
    pN, _ = _legendre(N, x)
    w = 2.0 / (N * (N + 1) * pN * pN)
    return 0.5 * (x + 1.0), 0.5 * w


def order_1d(n: int, order: str = "basix") -> np.ndarray:
    """Position (in ascending order) of the 1-D dof/point with index ``i``."""
    if order == "basix":
        return np.array([0, n - 1] + list(range(1, n - 1)), dtype=np.int64)
    if order == "ascending":
        return np.arange(n, dtype=np.int64)
    raise ValueError(f"unknown 1-D order {order!r}")


def lagrange_derivative_matrix(pts: np.ndarray) -> np.ndarray:
    """``D[q, i] = l_i'(pts[q])`` for the Lagrange basis on ``pts``.

    Same meaning as ``element_1D.tabulate(1, pts_1D)[1, :, :, 0]``
    (``cuda/demo_linear_box.py:356-357``).  Barycentric formula; the diagonal
    is the negative row sum so that constants differentiate to exactly 0.
    """
    pts = np.asarray(pts, dtype=np.float64)
    n = pts.size
    diff = pts[:, None] - pts[None, :]
    np.fill_diagonal(diff, 1.0)
    bw = 1.0 / np.prod(diff, axis=1)  # barycentric weights
    D = (bw[None, :] / bw[:, None]) / diff
    np.fill_diagonal(D, 0.0)
    np.fill_diagonal(D, -np.sum(D, axis=1))
    return D


# --------------------------------------------------------------------------- #
# Reference-element tables
# --------------------------------------------------------------------------- #


def p1_hex_gradients(pts: np.ndarray) -> np.ndarray:
    """``dphi[d, q, v]`` of the trilinear hex basis at ``pts`` (nq, 3).

    Stands in for ``gelement.tabulate(1, pts)[1:, :, :, 0]``
    (``cuda/demo_linear_box.py:237-239``).
    """
    pts = np.asarray(pts, dtype=np.float64)
    nq = pts.shape[0]
    out = np.zeros((3, nq, 8), dtype=np.float64)
    for v in range(8):
        bits = (v & 1, (v >> 1) & 1, (v >> 2) & 1)
        f = [pts[:, d] if bits[d] else 1.0 - pts[:, d] for d in range(3)]
        df = [1.0 if bits[d] else -1.0 for d in range(3)]
        out[0, :, v] = df[0] * f[1] * f[2]
        out[1, :, v] = f[0] * df[1] * f[2]
        out[2, :, v] = f[0] * f[1] * df[2]
    return out


def p1_hex_basis(pts: np.ndarray) -> np.ndarray:
    """``phi[q, v]`` of the trilinear hex basis at ``pts`` (nq, 3)."""
    pts = np.asarray(pts, dtype=np.float64)
    out = np.ones((pts.shape[0], 8), dtype=np.float64)
    for v in range(8):
        for d in range(3):
            out[:, v] *= pts[:, d] if (v >> d) & 1 else 1.0 - pts[:, d]
    return out


@dataclass
class ElementTables:
    """Everything the demos tabulate from Basix for one degree."""

    P: int
    order: str
    pts_1d: np.ndarray  # (n,) in dof order
    wts_1d: np.ndarray  # (n,)
    dphi_1D: np.ndarray  # (n, n)  dphi_1D[q, i]
    pts: np.ndarray  # (n^3, 3) hex quadrature points, q = i*n*n + j*n + k
    wts: np.ndarray  # (n^3,)
    dphi: np.ndarray  # (3, n^3, 8) P1 geometry gradients at pts
    pts_f: np.ndarray  # (6, n^2, 3) points on the six reference facets
    wts_f: np.ndarray  # (n^2,)
    dphi_f: np.ndarray  # (6, 3, n^2, 8)
    local_facet_dof: np.ndarray  # (6, n^2) int32, entity_closure_dofs[2]

    @property
    def n(self) -> int:
        return self.P + 1


def element_tables(P: int, order: str = "basix", dtype=np.float64) -> ElementTables:
    n = P + 1
    x_asc, w_asc = gll_points_weights(n)
    pos = order_1d(n, order)
    p1 = x_asc[pos]
    w1 = w_asc[pos]
    D = lagrange_derivative_matrix(p1)

    I, J, K = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    pts = np.stack([p1[I].ravel(), p1[J].ravel(), p1[K].ravel()], axis=1)
    wts = (w1[I] * w1[J] * w1[K]).ravel()
    dphi = p1_hex_gradients(pts)

    A, B = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a, b = p1[A].ravel(), p1[B].ravel()
    zeros, ones = np.zeros_like(a), np.ones_like(a)
    pts_f = np.zeros((6, n * n, 3))
    pts_f[0] = np.c_[a, b, zeros]  # z = 0
    pts_f[1] = np.c_[a, zeros, b]  # y = 0
    pts_f[2] = np.c_[zeros, a, b]  # x = 0
    pts_f[3] = np.c_[ones, a, b]  # x = 1
    pts_f[4] = np.c_[a, ones, b]  # y = 1
    pts_f[5] = np.c_[a, b, ones]  # z = 1
    wts_f = (w1[A] * w1[B]).ravel()
    dphi_f = np.stack([p1_hex_gradients(pts_f[f]) for f in range(6)], axis=0)

    # facet closure dofs, ascending in tensor-product numbering
    i0 = int(np.where(pos == 0)[0][0])  # 1-D index sitting at coordinate 0
    i1 = int(np.where(pos == n - 1)[0][0])  # ... at coordinate 1
    loc = np.arange(n**3).reshape(n, n, n)
    lfd = np.stack(
        [
            loc[:, :, i0].ravel(),
            loc[:, i0, :].ravel(),
            loc[i0, :, :].ravel(),
            loc[i1, :, :].ravel(),
            loc[:, i1, :].ravel(),
            loc[:, :, i1].ravel(),
        ]
    ).astype(np.int32)

    c = lambda z: np.ascontiguousarray(z, dtype=dtype)  # noqa: E731
    return ElementTables(
        P, order, c(p1), c(w1), c(D), c(pts), c(wts), c(dphi), c(pts_f), c(wts_f),
        c(dphi_f), lfd,
    )


# --------------------------------------------------------------------------- #
# Box mesh, tensor-product dofmap, facets
# --------------------------------------------------------------------------- #


class AdjacencyList:
    """Minimal ``dolfinx.graph.AdjacencyList`` stand-in."""

    def __init__(self, array: np.ndarray, offsets: np.ndarray):
        self.array = np.asarray(array)
        self.offsets = np.asarray(offsets)

    @property
    def num_nodes(self) -> int:
        return self.offsets.size - 1

    def links(self, i: int) -> np.ndarray:
        return self.array[self.offsets[i] : self.offsets[i + 1]]


@dataclass
class IndexMap:
    """The slice of ``dolfinx.common.IndexMap`` that ``compute_scatterer_data``
    reads (``cuda/utils.py:23-73``)."""

    size_local: int
    size_global: int
    local_range: tuple
    ghosts: np.ndarray  # int64 global indices
    owners: np.ndarray  # int32 ranks
    _dest: AdjacencyList = field(repr=False, default=None)

    @property
    def num_ghosts(self) -> int:
        return int(self.ghosts.size)

    def index_to_dest_ranks(self) -> AdjacencyList:
        return self._dest


@dataclass
class BoxMesh:
    """Structured hexahedral box with (optionally perturbed) vertices."""

    ncells: tuple  # (Nx, Ny, Nz) cells of THIS part
    x_dofs: np.ndarray  # (Nc, 8) int32   geometry dofmap
    x_g: np.ndarray  # (nv, 3)          vertex coordinates
    cell_origin: tuple = (0, 0, 0)  # offset of this part in the global cell grid
    global_ncells: tuple = None
    lengths: tuple = (1.0, 1.0, 1.0)
    # parts that are not boxes (partition_cells): global lexicographic id of every local cell;
    # ``ncells`` / ``cell_origin`` then carry no meaning
    cell_ids: np.ndarray = None

    @property
    def num_cells(self) -> int:
        return int(self.x_dofs.shape[0])


def create_box(
    ncells,
    lengths=(1.0, 1.0, 1.0),
    dtype=np.float64,
    perturb: float = 0.0,
    seed: int = 0,
) -> BoxMesh:
    """Serial box mesh ``[0,Lx]x[0,Ly]x[0,Lz]`` of ``Nx*Ny*Nz`` hexahedra.

    ``perturb`` is the vertex jitter as a fraction of the cell size (the
    reference perturbs by +-0.01 on a unit cube of 4-16 cells, unseeded,
    ``cuda/test_operators.py:79``); here it is seeded.
    Cell index ``cx*Ny*Nz + cy*Nz + cz``; vertex index likewise on the
    ``(N+1)^3`` grid.
    """
    if np.isscalar(ncells):
        ncells = (int(ncells),) * 3
    if np.isscalar(lengths):
        lengths = (float(lengths),) * 3
    Nx, Ny, Nz = (int(v) for v in ncells)
    gx = np.linspace(0.0, lengths[0], Nx + 1)
    gy = np.linspace(0.0, lengths[1], Ny + 1)
    gz = np.linspace(0.0, lengths[2], Nz + 1)
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    x_g = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if perturb:
        rng = np.random.default_rng(seed)
        h = np.array([lengths[0] / Nx, lengths[1] / Ny, lengths[2] / Nz])
        x_g = x_g + rng.uniform(-perturb, perturb, size=x_g.shape) * h
    x_dofs = _structured_connectivity((Nx, Ny, Nz), 1)
    return BoxMesh(
        (Nx, Ny, Nz), x_dofs, np.ascontiguousarray(x_g, dtype=dtype), (0, 0, 0),
        (Nx, Ny, Nz), tuple(float(v) for v in lengths),
    )


def _structured_connectivity(ncells, P, pos=None, origin=(0, 0, 0), grid=None):
    """(Nc, (P+1)^3) node indices of a structured grid, int64.

    ``grid`` is the node-grid shape the indices refer to (default: the grid
    spanned by ``ncells``); ``origin`` the node offset of the first cell.
    For P=1 the local order is the P1 vertex order ``vx + 2 vy + 4 vz``.
    """
    Nx, Ny, Nz = ncells
    n = P + 1
    if grid is None:
        grid = (P * Nx + 1, P * Ny + 1, P * Nz + 1)
    GY, GZ = grid[1], grid[2]
    if pos is None:  # P1 geometry: vertex order, x fastest
        v = np.arange(8)
        ox, oy, oz = v & 1, (v >> 1) & 1, (v >> 2) & 1
    else:
        I, J, K = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        ox, oy, oz = pos[I].ravel(), pos[J].ravel(), pos[K].ravel()
    # index = base(cell) + offset(local node): both separable, one broadcast add
    gx0 = (origin[0] + np.arange(Nx, dtype=np.int64) * P) * (GY * GZ)
    gy0 = (origin[1] + np.arange(Ny, dtype=np.int64) * P) * GZ
    gz0 = origin[2] + np.arange(Nz, dtype=np.int64) * P
    base = (gx0[:, None, None] + gy0[None, :, None] + gz0[None, None, :]).ravel()
    off = (ox.astype(np.int64) * GY + oy) * GZ + oz
    top = (int(base.max(initial=0)) + int(off.max(initial=0))) if base.size else 0
    if top < 2**31:
        return base.astype(np.int32)[:, None] + off.astype(np.int32)[None, :]
    return base[:, None] + off[None, :]


def tensor_dofmap(mesh: BoxMesh, P: int, order: str = "basix") -> np.ndarray:
    """Serial Q_P dofmap in tensor-product order, ``(Nc, (P+1)^3)`` int32.

    What ``V.dofmap.list[:, perm]`` is in the reference
    (``cuda/demo_linear_box.py:167-176``, ``cpp/common/permute.hpp:15-42``).
    Global dof ``(gx, gy, gz)`` -> ``(gx*GY + gy)*GZ + gz``.
    """
    dm = _structured_connectivity(mesh.ncells, P, order_1d(P + 1, order))
    if dm.dtype != np.int32:
        raise ValueError("dofmap needs more than int32 indices")
    return np.ascontiguousarray(dm)


def box_cell_colours(mesh: BoxMesh) -> np.ndarray:
    """The exact 8-colouring of a structured box: colour = parity of the cell's
    (x, y, z) position (global parity, so it is consistent across parts)."""
    Nx, Ny, Nz = mesh.ncells
    ox, oy, oz = mesh.cell_origin
    cx = (np.arange(Nx) + ox) & 1
    cy = (np.arange(Ny) + oy) & 1
    cz = (np.arange(Nz) + oz) & 1
    return (cx[:, None, None] * 4 + cy[None, :, None] * 2 + cz[None, None, :]).astype(np.int32).ravel()


def num_dofs(ncells, P: int) -> int:
    if np.isscalar(ncells):
        ncells = (ncells,) * 3
    return int(np.prod([P * int(v) + 1 for v in ncells]))


def dof_coordinates(mesh: BoxMesh, dofmap: np.ndarray, tables: ElementTables, ndofs=None):
    """Physical coordinates of every dof through the trilinear geometry map."""
    phi = p1_hex_basis(tables.pts.astype(np.float64))  # (Nd, 8)
    ndofs = int(dofmap.max()) + 1 if ndofs is None else ndofs
    xd = np.zeros((ndofs, 3), dtype=np.float64)
    # chunked to bound memory
    step = 1 << 15
    for s in range(0, mesh.num_cells, step):
        e = min(mesh.num_cells, s + step)
        coords = mesh.x_g[mesh.x_dofs[s:e]].astype(np.float64)  # (c, 8, 3)
        xd[dofmap[s:e].ravel()] = np.einsum("qv,cvd->cqd", phi, coords).reshape(-1, 3)
    return xd


_FACE_AXIS = {0: (2, 0), 1: (1, 0), 2: (0, 0), 3: (0, 1), 4: (1, 1), 5: (2, 1)}


def boundary_facets(mesh: BoxMesh, local_facet: int, predicate=None) -> np.ndarray:
    """``boundary_data`` (nf, 2) int32 = (cell, local facet) for one box face.

    Equivalent of ``locate_entities_boundary`` + ``facet_integration_domain``
    (``cuda/demo_linear_box.py:256-270``, ``cuda/utils.py:81-114``) for the
    face of the *global* box that reference facet ``local_facet`` lies on
    (0: z=0, 1: y=0, 2: x=0, 3: x=1, 4: y=1, 5: z=1).  Parts of a partitioned
    mesh only return facets that are on the global boundary.
    ``predicate(centroids (nf,3)) -> bool mask`` optionally filters them.
    """
    axis, side = _FACE_AXIS[local_facet]
    if mesh.cell_ids is not None:  # arbitrary set of cells of the global box, arbitrary order
        gN = mesh.global_ncells
        ids = np.asarray(mesh.cell_ids, dtype=np.int64)
        pos = (ids // (gN[1] * gN[2]), (ids // gN[2]) % gN[1], ids % gN[2])[axis]
        cells = np.nonzero(pos == (0 if side == 0 else gN[axis] - 1))[0]
        out = np.stack([cells, np.full_like(cells, local_facet)], axis=1).astype(np.int32)
        if predicate is not None and cells.size:
            fv = [v for v in range(8) if ((v >> axis) & 1) == side]
            cen = mesh.x_g[mesh.x_dofs[cells][:, fv]].mean(axis=1)
            out = out[np.asarray(predicate(cen), dtype=bool)]
        return np.ascontiguousarray(out.reshape(-1, 2))
    N = mesh.ncells
    org = mesh.cell_origin
    gN = mesh.global_ncells or N
    if side == 0 and org[axis] != 0:
        return np.zeros((0, 2), dtype=np.int32)
    if side == 1 and org[axis] + N[axis] != gN[axis]:
        return np.zeros((0, 2), dtype=np.int32)
    rng = [np.arange(N[0]), np.arange(N[1]), np.arange(N[2])]
    rng[axis] = np.array([0 if side == 0 else N[axis] - 1])
    cx, cy, cz = np.meshgrid(*rng, indexing="ij")
    cells = ((cx * N[1] + cy) * N[2] + cz).ravel()
    out = np.stack([cells, np.full_like(cells, local_facet)], axis=1).astype(np.int32)
    if predicate is not None:
        # the four vertices of reference facet f, in the P1 vertex numbering
        fv = [v for v in range(8) if ((v >> axis) & 1) == side]
        cen = mesh.x_g[mesh.x_dofs[cells][:, fv]].mean(axis=1)
        out = out[np.asarray(predicate(cen), dtype=bool)]
    return np.ascontiguousarray(out)


def facet_dofmap(dofmap: np.ndarray, boundary_data: np.ndarray, local_facet_dof: np.ndarray):
    """``bfacet_dofmap[i, :] = dofmap[cell][local_facet_dof[local_facet]]``
    (``cuda/demo_linear_box.py:320-333``), vectorised."""
    if boundary_data.shape[0] == 0:
        return np.zeros((0, local_facet_dof.shape[1]), dtype=np.int32)
    cells = boundary_data[:, 0]
    lf = boundary_data[:, 1]
    return np.ascontiguousarray(dofmap[cells[:, None], local_facet_dof[lf]], dtype=np.int32)


# --------------------------------------------------------------------------- #
# Block partition (stand-in for the DOLFINx partitioner, GhostMode.none)
# --------------------------------------------------------------------------- #


def block_grid(nranks: int) -> tuple:
    """Rank grid used for box meshes: 1, 2x1x1, 2x2x1, 2x2x2, then the most
    cubic factorisation."""
    best = (nranks, 1, 1)
    for a in range(1, nranks + 1):
        if nranks % a:
            continue
        for b in range(1, nranks // a + 1):
            if (nranks // a) % b:
                continue
            c = nranks // a // b
            cand = tuple(sorted((a, b, c), reverse=True))
            if max(cand) - min(cand) < max(best) - min(best):
                best = cand
    return best


def _split(N: int, parts: int):
    base, rem = divmod(N, parts)
    sizes = [base + (1 if p < rem else 0) for p in range(parts)]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    return starts


@dataclass
class Partition:
    """One rank's share of a block-partitioned box: mesh part, local dofmap
    (owned dofs first, ghosts after) and the index map."""

    rank: int
    nranks: int
    mesh: BoxMesh
    dofmap: np.ndarray  # (Nc_local, Nd) int32 local indices
    index_map: IndexMap
    local_to_global: np.ndarray  # (nlocal+nghost,) int64 *global dof ids in the new numbering*
    local_to_serial: np.ndarray  # (nlocal+nghost,) int64 lexicographic serial dof ids


def partition_box(
    ncells,
    P: int,
    nranks: int,
    lengths=(1.0, 1.0, 1.0),
    order: str = "basix",
    dtype=np.float64,
    perturb: float = 0.0,
    seed: int = 0,
    ranks=None,
    grid=None,
):
    """Block-partition a box mesh the way DOLFINx would hand it to each rank
    with ``GhostMode.none``: every rank gets its cells, the dofs those cells
    touch, and an index map in which an interface dof is owned by the lowest
    rank-block touching it.  Returns ``[Partition]`` for ``ranks`` (default
    all)."""
    if np.isscalar(ncells):
        ncells = (int(ncells),) * 3
    ncells = tuple(int(v) for v in ncells)
    R = tuple(grid) if grid is not None else block_grid(nranks)
    assert R[0] * R[1] * R[2] == nranks
    starts = [_split(ncells[d], R[d]) for d in range(3)]
    G = tuple(P * ncells[d] + 1 for d in range(3))
    pos = order_1d(P + 1, order)

    # owned dof range of block b along axis d: (lo, hi] in node units, first
    # block also owns node 0.
    def own_range(d, b):
        lo = P * starts[d][b] + (1 if b > 0 else 0)
        hi = P * starts[d][b + 1]
        return lo, hi + 1  # half-open

    def rank_of(bx, by, bz):
        return (bx * R[1] + by) * R[2] + bz

    # number of owned dofs per rank -> global offsets
    owned = np.zeros(nranks, dtype=np.int64)
    for bx in range(R[0]):
        for by in range(R[1]):
            for bz in range(R[2]):
                r = [own_range(0, bx), own_range(1, by), own_range(2, bz)]
                owned[rank_of(bx, by, bz)] = np.prod([b - a for a, b in r])
    offsets = np.concatenate([[0], np.cumsum(owned)])

    # owner block of a node coordinate along one axis
    def owner_block_1d(d, g):
        # block b owns (P*starts[b], P*starts[b+1]]; node 0 -> block 0
        b = np.searchsorted(P * starts[d][1:], g, side="left")
        return np.minimum(b, R[d] - 1)

    def new_global(gx, gy, gz):
        """global index in the rank-contiguous numbering of node (gx,gy,gz)"""
        bx, by, bz = owner_block_1d(0, gx), owner_block_1d(1, gy), owner_block_1d(2, gz)
        rk = (bx * R[1] + by) * R[2] + bz
        lox = P * starts[0][bx] + (bx > 0)
        loy = P * starts[1][by] + (by > 0)
        loz = P * starts[2][bz] + (bz > 0)
        ny = P * starts[1][by + 1] + 1 - loy
        nz = P * starts[2][bz + 1] + 1 - loz
        loc = ((gx - lox) * ny + (gy - loy)) * nz + (gz - loz)
        return offsets[rk] + loc, rk

    full = create_box(ncells, lengths, dtype=np.float64, perturb=perturb, seed=seed)
    Vg = (ncells[1] + 1, ncells[2] + 1)

    parts = []
    want = range(nranks) if ranks is None else ranks
    for rank in want:
        bz = rank % R[2]
        by = (rank // R[2]) % R[1]
        bx = rank // (R[1] * R[2])
        c0 = (starts[0][bx], starts[1][by], starts[2][bz])
        nc = (
            starts[0][bx + 1] - c0[0],
            starts[1][by + 1] - c0[1],
            starts[2][bz + 1] - c0[2],
        )
        # --- mesh part (geometry re-indexed to the part's own vertices)
        vx = np.arange(c0[0], c0[0] + nc[0] + 1)
        vy = np.arange(c0[1], c0[1] + nc[1] + 1)
        vz = np.arange(c0[2], c0[2] + nc[2] + 1)
        VX, VY, VZ = np.meshgrid(vx, vy, vz, indexing="ij")
        vid = ((VX * Vg[0] + VY) * Vg[1] + VZ).ravel()
        x_g = np.ascontiguousarray(full.x_g[vid], dtype=dtype)
        x_dofs = _structured_connectivity(nc, 1)
        mesh = BoxMesh(tuple(int(v) for v in nc), x_dofs, x_g, tuple(int(v) for v in c0),
                       ncells, tuple(float(v) for v in lengths))

        # --- dofs touched by this part: the node box [P*c0, P*(c0+nc)].  Every
        # per-node quantity is separable over the axes, so it is built from three
        # 1-D arrays and broadcast once (the node box has 1.3e8 entries at 125^3 cells).
        lo = [P * c0[d] for d in range(3)]
        ext = [P * nc[d] + 1 for d in range(3)]
        g1 = [np.arange(lo[d], lo[d] + ext[d], dtype=np.int64) for d in range(3)]
        b1 = [owner_block_1d(d, g1[d]) for d in range(3)]
        lo1 = [P * starts[d][b1[d]] + (b1[d] > 0) for d in range(3)]
        n1 = [P * starts[d][b1[d] + 1] + 1 - lo1[d] for d in range(3)]
        bc = lambda v, d: v.reshape([-1 if e == d else 1 for e in range(3)])  # noqa: E731
        own = ((bc(b1[0], 0) * R[1] + bc(b1[1], 1)) * R[2] + bc(b1[2], 2)).astype(np.int32).ravel()
        gnew = (((bc(g1[0] - lo1[0], 0) * bc(n1[1], 1) + bc(g1[1] - lo1[1], 1)) * bc(n1[2], 2)
                 + bc(g1[2] - lo1[2], 2)).ravel())
        gnew += offsets[own]
        is_owned = own == rank
        nlocal = int(is_owned.sum())
        assert nlocal == owned[rank]
        nnode = gnew.size
        # local numbering: owned -> gnew - offset (lexicographic in the owned
        # box), ghosts -> appended in ascending global order
        local = gnew - offsets[rank]
        gh_pos = np.where(~is_owned)[0]
        gh_sort = gh_pos[np.argsort(gnew[gh_pos], kind="stable")]
        local[gh_sort] = nlocal + np.arange(gh_sort.size)
        ghosts = gnew[gh_sort].astype(np.int64)
        ghost_owners = own[gh_sort].astype(np.int32)

        # local dofmap through the part's node box
        boxmap = _structured_connectivity(nc, P, pos)  # indices into the node box
        local32 = local.astype(np.int32) if nnode < 2**31 else local
        dofmap = np.ascontiguousarray(local32[boxmap], dtype=np.int32)
        del boxmap, local32

        l2g = np.empty(nnode, dtype=np.int64)
        l2g[local] = gnew
        l2s = np.empty(nnode, dtype=np.int64)
        l2s[local] = ((bc(g1[0], 0) * G[1] + bc(g1[1], 1)) * G[2] + bc(g1[2], 2)).ravel()

        # --- index_to_dest_ranks: for each owned dof, ranks that ghost it.
        # A rank-block touches node g along axis d if P*s_b <= g <= P*s_{b+1}:
        # the owner block and, when g is on an interface (g == P*starts[b+1],
        # b+1 < R), block b+1.  Owned nodes form a sub-box of the node box, so the
        # (owned dof, destination rank) pairs are Cartesian products of per-axis sets.
        mine = (bx, by, bz)
        own1 = [np.nonzero(b1[d] == mine[d])[0] for d in range(3)]  # owned positions per axis
        if1 = [g1[d][own1[d]] == P * starts[d][mine[d] + 1] if mine[d] + 1 < R[d]
               else np.zeros(own1[d].size, bool) for d in range(3)]
        pair_dof, pair_rank = [], []
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    if dx == dy == dz == 0:
                        continue
                    sel = [own1[d][if1[d]] if dd else own1[d] for d, dd in enumerate((dx, dy, dz))]
                    if all(v.size for v in sel):
                        flat = ((bc(sel[0], 0) * ext[1] + bc(sel[1], 1)) * ext[2] + bc(sel[2], 2)).ravel()
                        pair_dof.append(local[flat])
                        pair_rank.append(np.full(flat.size, rank_of(bx + dx, by + dy, bz + dz), np.int64))
        if pair_dof:
            pd = np.concatenate(pair_dof)
            pr = np.concatenate(pair_rank)
            order = np.lexsort((pr, pd))  # by dof, destinations ascending
            pd, pr = pd[order], pr[order]
        else:
            pd = np.zeros(0, np.int64)
            pr = np.zeros(0, np.int64)
        counts = np.bincount(pd, minlength=nlocal)
        d_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        d_arr = pr.astype(np.int32)
        imap = IndexMap(
            nlocal, int(np.prod(G)), (int(offsets[rank]), int(offsets[rank + 1])),
            ghosts, ghost_owners, AdjacencyList(d_arr, d_off),
        )
        parts.append(Partition(rank, nranks, mesh, dofmap, imap, l2g, l2s))
    return parts


def blob_cell_ranks(ncells, nranks: int, seed: int = 0) -> np.ndarray:
    """An irregular cell -> rank assignment for ``partition_cells``: every cell goes to the nearest
    of ``nranks`` random seed points (in a randomly stretched metric), so the parts are connected
    blobs with staircase interfaces, different sizes and different numbers of neighbours - the shape
    of a graph partitioner's output, not a block grid.  Every rank gets at least one cell."""
    if np.isscalar(ncells):
        ncells = (int(ncells),) * 3
    rng = np.random.default_rng(seed)
    cx, cy, cz = np.meshgrid(*[np.arange(n) + 0.5 for n in ncells], indexing="ij")
    cen = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1)
    for _ in range(100):
        pts = rng.uniform(0, 1, (nranks, 3)) * np.array(ncells)
        w = rng.uniform(0.6, 1.6, (nranks, 3))
        d2 = (((cen[:, None, :] - pts[None, :, :]) * w[None, :, :]) ** 2).sum(axis=2)
        ranks = np.argmin(d2, axis=1).astype(np.int32)
        if np.unique(ranks).size == nranks:
            return ranks
    raise RuntimeError("blob_cell_ranks: could not give every rank a cell")


def partition_cells(
    ncells,
    P: int,
    cell_rank: np.ndarray,
    lengths=(1.0, 1.0, 1.0),
    order: str = "basix",
    dtype=np.float64,
    perturb: float = 0.0,
    seed: int = 0,
    shuffle_seed=None,
    owner_rule: str = "lowest",
):
    """Partition a box mesh by an ARBITRARY cell -> rank map (``cell_rank[c]`` for the global
    lexicographic cell ``c``), the way a graph partitioner hands an unstructured mesh to DOLFINx
    with ``GhostMode.none``: nothing below knows that the cells form a grid.  Each rank gets its
    cells, the dofs they touch (owned first, ghosts after) and an ``IndexMap`` (ghost global
    indices, ghost owners, ``index_to_dest_ranks``) in a rank-contiguous global numbering.

    ``owner_rule``: a dof shared by several ranks is owned by the ``"lowest"`` of them (as in
    ``partition_box``) or by a pseudo-random one (``"hash"``, so that ownership is not monotone
    in the rank).  ``shuffle_seed`` (int): the local cell order, the order of the owned dofs and
    the order of the ghosts are random permutations instead of the serial order - no run of
    consecutive dof indices, no neighbouring consecutive cells survive.
    Returns ``[Partition]`` for every rank."""
    if np.isscalar(ncells):
        ncells = (int(ncells),) * 3
    ncells = tuple(int(v) for v in ncells)
    cell_rank = np.asarray(cell_rank, dtype=np.int64).ravel()
    assert cell_rank.size == int(np.prod(ncells))
    nranks = int(cell_rank.max()) + 1
    rng = np.random.default_rng(shuffle_seed) if shuffle_seed is not None else None
    full = create_box(ncells, lengths, dtype=np.float64, perturb=perturb, seed=seed)
    gdm = tensor_dofmap(full, P, order).astype(np.int64)  # serial dof ids
    ntot = num_dofs(ncells, P)
    Nd = gdm.shape[1]

    # (dof, rank) incidences, unique
    pairs = np.unique(gdm.ravel() * nranks + np.repeat(cell_rank, Nd))
    pd, pr = pairs // nranks, pairs % nranks  # sorted by dof, ranks ascending within a dof
    first = np.concatenate([[0], np.cumsum(np.bincount(pd, minlength=ntot))])  # CSR over dofs
    cnt = np.diff(first)
    if owner_rule == "lowest":
        pick = np.zeros(ntot, dtype=np.int64)
    elif owner_rule == "hash":
        pick = (np.arange(ntot, dtype=np.int64) * 2654435761 >> 7) % cnt
    else:
        raise ValueError("owner_rule must be 'lowest' or 'hash'")
    owner = pr[first[:-1] + pick]

    # rank-contiguous global numbering: owned dofs of rank r in serial (or shuffled) order
    new_global = np.empty(ntot, dtype=np.int64)
    owned_lists, offsets = [], [0]
    for r in range(nranks):
        mine = np.nonzero(owner == r)[0]
        if rng is not None:
            mine = mine[rng.permutation(mine.size)]
        owned_lists.append(mine)
        new_global[mine] = offsets[-1] + np.arange(mine.size)
        offsets.append(offsets[-1] + mine.size)

    parts = []
    for r in range(nranks):
        cells = np.nonzero(cell_rank == r)[0]
        if rng is not None:
            cells = cells[rng.permutation(cells.size)]
        mine = owned_lists[r]
        nlocal = mine.size
        touched = np.unique(gdm[cells])
        gh = touched[owner[touched] != r]
        gh = gh[rng.permutation(gh.size)] if rng is not None else gh[np.argsort(new_global[gh], kind="stable")]
        local = np.full(ntot, -1, dtype=np.int64)
        local[mine] = np.arange(nlocal)
        local[gh] = nlocal + np.arange(gh.size)
        dofmap = np.ascontiguousarray(local[gdm[cells]], dtype=np.int32)
        assert dofmap.min() >= 0
        l2s = np.concatenate([mine, gh]).astype(np.int64)
        # destinations of every owned dof: the other ranks touching it, ascending
        d_cnt = cnt[mine] - 1
        d_off = np.concatenate([[0], np.cumsum(d_cnt)]).astype(np.int32)
        sh = np.nonzero(d_cnt > 0)[0]  # shared dofs only (a surface's worth)
        c_sh = cnt[mine[sh]]  # incidences of each shared dof (its own rank included)
        run0 = np.concatenate([[0], np.cumsum(c_sh)])[:-1]
        at = np.repeat(first[mine[sh]] - run0, c_sh) + np.arange(int(c_sh.sum()))  # CSR rows, concatenated
        rk = pr[at]
        d_arr = rk[rk != r].astype(np.int32)  # rows stay in dof order, ranks ascending within a row
        assert d_arr.size == int(d_off[-1])
        imap = IndexMap(nlocal, ntot, (int(offsets[r]), int(offsets[r + 1])), new_global[gh].astype(np.int64),
                        owner[gh].astype(np.int32), AdjacencyList(d_arr, d_off))
        # geometry of the part: its vertices, renumbered
        verts, inv = np.unique(full.x_dofs[cells], return_inverse=True)
        mesh = BoxMesh((int(cells.size), 1, 1), np.ascontiguousarray(inv.reshape(-1, 8), dtype=np.int32),
                       np.ascontiguousarray(full.x_g[verts], dtype=dtype), (0, 0, 0), ncells,
                       tuple(float(v) for v in lengths), cells.astype(np.int64))
        parts.append(Partition(r, nranks, mesh, dofmap, imap, new_global[l2s], l2s))
    return parts


def serial_index_map(ndofs: int) -> IndexMap:
    """Index map of a single-rank run: everything owned, no ghosts."""
    return IndexMap(
        ndofs, ndofs, (0, ndofs), np.zeros(0, np.int64), np.zeros(0, np.int32),
        AdjacencyList(np.zeros(0, np.int32), np.zeros(ndofs + 1, np.int32)),
    )
