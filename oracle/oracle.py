"""
ORACLE - TEST INFRASTRUCTURE ONLY.

numpy-facing wrapper around ``oracle/libfus_oracle.so`` (the plain-C
restatement of the reference's CPU algorithm, ``oracle/fus_oracle_impl.h``)
and, when built, ``oracle/_ref/libfus_ref.so`` (the reference's own C++
sum-factorisation templates under a thin driver, ``oracle/ref_driver.cpp``).
Host-side pieces that the reference writes in Python (index maps, the RK
loops) are restated here in numpy / plain loops with the reference lines
cited.

Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product
package never imports this module.

Parity status: PINNED for operators, geometry, pack/unpack, index maps and
halo semantics against outputs of the reference's own Python
(``numba-cpu/*.py``, ``cuda/utils.py``, ``cuda/precompute.py``) run in the
build container - fixtures in ``tests/golden/`` (generator:
``tests/golden/make_golden.py``).  UNPINNED at the Basix/DOLFINx boundary
(GLL tables, dof ordering, partitioning): those libraries are absent, see
DESIGN.md.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> None:
    """Compile the C oracle (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "libfus_oracle.so")
    ref = os.path.join(_HERE, "_ref", "libfus_ref.so")
    have_ref_src = os.path.exists("/root/reference/cpp/common/sum_factorisation.hpp")
    if force or not os.path.exists(so) or (have_ref_src and not os.path.exists(ref)):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.path.join(_HERE, "libfus_oracle.so"))
    return _LIB


def ref_lib():
    """The compiled reference templates, or None when not built."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libfus_ref.so")
        if not os.path.exists(p):
            try:
                build()
            except Exception:
                pass
        if os.path.exists(p):
            _REF = C.CDLL(p)
    return _REF


def _sfx(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(dtype)


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def _ct(dtype):
    return C.c_double if np.dtype(dtype) == np.float64 else C.c_float


def _chk(a, dtype=None):
    assert isinstance(a, np.ndarray) and a.flags.c_contiguous, "need C-contiguous numpy"
    if dtype is not None:
        assert a.dtype == np.dtype(dtype), (a.dtype, dtype)
    return a


# --------------------------------------------------------------------------- #
# operators
# --------------------------------------------------------------------------- #


def mass_operator(x, coeff, y, detJ, dofmap):
    """numba-cpu/operators.py:19-68.  Accumulates into ``y``."""
    dt = x.dtype
    for a in (x, coeff, y, detJ):
        _chk(a, dt)
    _chk(dofmap, np.int32)
    nent, ncols = dofmap.shape
    assert detJ.shape == dofmap.shape and coeff.size == nent
    getattr(lib(), "orc_mass_" + _sfx(dt))(
        _fp(x), _fp(coeff), _fp(y), _fp(detJ), _fp(dofmap), C.c_int64(nent), C.c_int(ncols)
    )


def stiffness_operator(P, x, coeff, y, G, dofmap, dphi):
    """numba-cpu/operators.py:71-227.  Accumulates into ``y``."""
    dt = x.dtype
    dphi = np.ascontiguousarray(dphi, dtype=dt)
    for a in (x, coeff, y, G):
        _chk(a, dt)
    _chk(dofmap, np.int32)
    n = P + 1
    assert dofmap.shape[1] == n**3 and G.shape == (dofmap.shape[0], n**3, 6)
    getattr(lib(), "orc_stiffness_" + _sfx(dt))(
        _fp(x), _fp(coeff), _fp(y), _fp(G), _fp(dofmap), _fp(dphi),
        C.c_int64(dofmap.shape[0]), C.c_int(P),
    )


def ref_stiffness_operator(P, x, coeff, y, G, dofmap, dphi):
    """Same action through the reference's compiled C++ templates."""
    r = ref_lib()
    assert r is not None, "oracle/_ref/libfus_ref.so not built"
    dt = x.dtype
    dphi = np.ascontiguousarray(dphi, dtype=dt)
    rc = getattr(r, "ref_stiffness_" + _sfx(dt))(
        _fp(x), _fp(coeff), _fp(y), _fp(G), _fp(dofmap), _fp(dphi),
        C.c_int64(dofmap.shape[0]), C.c_int(P),
    )
    assert rc == 0


def ref_mass_operator(x, coeff, y, detJ, dofmap):
    r = ref_lib()
    assert r is not None
    getattr(r, "ref_mass_" + _sfx(x.dtype))(
        _fp(x), _fp(coeff), _fp(y), _fp(detJ), _fp(dofmap),
        C.c_int64(dofmap.shape[0]), C.c_int(dofmap.shape[1]),
    )


def stiffness_ranks(P, x, coeff, ybuf, G, dofmap, dphi, nranks, use_ref=True):
    """CPU-baseline helper: ``nranks`` threads, each a contiguous share of the
    cells and a private output row of ``ybuf`` (nranks, nd) - emulates
    ``mpirun -n k`` of the serial reference kernel without halo cost."""
    dt = x.dtype
    dphi = np.ascontiguousarray(dphi, dtype=dt)
    assert ybuf.shape[0] >= nranks and ybuf.flags.c_contiguous
    r = ref_lib() if use_ref else None
    if r is not None:
        fn, kind = getattr(r, "ref_stiffness_ranks_" + _sfx(dt)), "reference"
    else:
        fn, kind = getattr(lib(), "orc_stiffness_ranks_" + _sfx(dt)), "port"
    fn(
        _fp(x), _fp(coeff), _fp(ybuf), C.c_int64(ybuf.shape[1]), _fp(G), _fp(dofmap),
        _fp(dphi), C.c_int64(dofmap.shape[0]), C.c_int(P), C.c_int(nranks),
    )
    return kind


def mass_ranks(x, coeff, ybuf, detJ, dofmap, nranks, use_ref=True):
    dt = x.dtype
    r = ref_lib() if use_ref else None
    if r is not None:
        fn, kind = getattr(r, "ref_mass_ranks_" + _sfx(dt)), "reference"
    else:
        fn, kind = getattr(lib(), "orc_mass_ranks_" + _sfx(dt)), "port"
    fn(
        _fp(x), _fp(coeff), _fp(ybuf), C.c_int64(ybuf.shape[1]), _fp(detJ), _fp(dofmap),
        C.c_int64(dofmap.shape[0]), C.c_int(dofmap.shape[1]), C.c_int(nranks),
    )
    return kind


def axpy(alpha, x, y):
    getattr(lib(), "orc_axpy_" + _sfx(x.dtype))(_ct(x.dtype)(alpha), _fp(x), _fp(y), C.c_int64(x.size))


def copy(a, b):
    getattr(lib(), "orc_copy_" + _sfx(a.dtype))(_fp(a), _fp(b), C.c_int64(a.size))


def fill(alpha, x):
    getattr(lib(), "orc_fill_" + _sfx(x.dtype))(_ct(x.dtype)(alpha), _fp(x), C.c_int64(x.size))


def pointwise_divide(a, b, c):
    getattr(lib(), "orc_pointwise_divide_" + _sfx(a.dtype))(_fp(a), _fp(b), _fp(c), C.c_int64(c.size))


def square(a, b):
    getattr(lib(), "orc_square_" + _sfx(a.dtype))(_fp(a), _fp(b), C.c_int64(a.size))


# --------------------------------------------------------------------------- #
# pack / unpack   (cuda/scatterer.py:18-101)
# --------------------------------------------------------------------------- #


def pack_fwd(in_, out_, index):
    _chk(index, np.int64)
    getattr(lib(), "orc_pack_fwd_" + _sfx(in_.dtype))(_fp(in_), _fp(out_), _fp(index), C.c_int64(index.size))


def unpack_fwd(in_, out_, index, N):
    _chk(index, np.int64)
    getattr(lib(), "orc_unpack_fwd_" + _sfx(in_.dtype))(
        _fp(in_), _fp(out_), _fp(index), C.c_int64(index.size), C.c_int64(N))


def pack_rev(in_, out_, index, N):
    _chk(index, np.int64)
    getattr(lib(), "orc_pack_rev_" + _sfx(in_.dtype))(
        _fp(in_), _fp(out_), _fp(index), C.c_int64(index.size), C.c_int64(N))


def unpack_rev(in_, out_, index):
    _chk(index, np.int64)
    getattr(lib(), "orc_unpack_rev_" + _sfx(in_.dtype))(_fp(in_), _fp(out_), _fp(index), C.c_int64(index.size))


# --------------------------------------------------------------------------- #
# geometry   (cuda/precompute.py)
# --------------------------------------------------------------------------- #


def compute_scaled_jacobian_determinant(detJ, mesh, num_cell, dphi, weights):
    """cuda/precompute.py:76-112"""
    x_dofs, x_g = mesh
    dt = detJ.dtype
    getattr(lib(), "orc_detJ_" + _sfx(dt))(
        _fp(detJ), _fp(_chk(x_dofs, np.int32)), _fp(np.ascontiguousarray(x_g, dt)),
        C.c_int64(num_cell), _fp(np.ascontiguousarray(dphi, dt)),
        _fp(np.ascontiguousarray(weights, dt)), C.c_int(weights.size),
    )


def compute_scaled_geometrical_factor(G, mesh, num_cell, dphi, weights):
    """cuda/precompute.py:115-163"""
    x_dofs, x_g = mesh
    dt = G.dtype
    getattr(lib(), "orc_G_" + _sfx(dt))(
        _fp(G), _fp(_chk(x_dofs, np.int32)), _fp(np.ascontiguousarray(x_g, dt)),
        C.c_int64(num_cell), _fp(np.ascontiguousarray(dphi, dt)),
        _fp(np.ascontiguousarray(weights, dt)), C.c_int(weights.size),
    )


def compute_boundary_facets_scaled_jacobian_determinant(detJ_f, mesh, boundary_data, dphi_f, weights):
    """cuda/precompute.py:17-73"""
    x_dofs, x_g = mesh
    dt = detJ_f.dtype
    getattr(lib(), "orc_detJ_facet_" + _sfx(dt))(
        _fp(detJ_f), _fp(_chk(x_dofs, np.int32)), _fp(np.ascontiguousarray(x_g, dt)),
        _fp(_chk(boundary_data, np.int32)), C.c_int64(boundary_data.shape[0]),
        _fp(np.ascontiguousarray(dphi_f, dt)), _fp(np.ascontiguousarray(weights, dt)),
        C.c_int(weights.size),
    )


# --------------------------------------------------------------------------- #
# index maps and halo exchange, all ranks in one process
# --------------------------------------------------------------------------- #


def compute_scatterer_data_all(index_maps):
    """cuda/utils.py:8-78 restated for ALL ranks at once: the reference's
    ``Isend/Irecv`` of ghost global indices (57-71) becomes a dictionary
    look-up.  Loops kept as written (O(ranks * size_local), small cases
    only).  Returns ``[(owners_data, ghosts_data)]`` per rank."""
    per_rank = []
    sent = {}
    for rank, im in enumerate(index_maps):
        nlocal = im.size_local
        owners = im.owners
        unique_owners, owners_size = np.unique(owners, return_counts=True)
        owners_argsorted = np.argsort(owners)
        owners_offsets = np.insert(np.cumsum(owners_size), 0, 0)
        owners_idx = []
        for i, owner in enumerate(unique_owners):
            begin, end = owners_offsets[i], owners_offsets[i + 1]
            owners_idx.append(owners_argsorted[begin:end])
            sent[(rank, int(owner))] = im.ghosts[owners_argsorted[begin:end]]

        shared_dofs = im.index_to_dest_ranks()
        shared_ranks = np.unique(shared_dofs.array)
        ghosts = []
        for shared_rank in shared_ranks:
            for dof in range(nlocal):
                if shared_rank in shared_dofs.links(dof):
                    ghosts.append(shared_rank)
        ghosts = np.array(ghosts, dtype=np.int32)
        unique_ghosts, ghosts_size = np.unique(ghosts, return_counts=True)
        per_rank.append((owners_idx, owners_size, unique_owners, ghosts_size, unique_ghosts))

    out = []
    for rank, im in enumerate(index_maps):
        owners_idx, owners_size, unique_owners, ghosts_size, unique_ghosts = per_rank[rank]
        ghosts_idx = []
        for i, g in enumerate(unique_ghosts):
            recv = sent[(int(g), rank)]
            assert recv.size == ghosts_size[i]
            ghosts_idx.append(recv - im.local_range[0])
        out.append(([owners_idx, owners_size, unique_owners], [ghosts_idx, ghosts_size, unique_ghosts]))
    return out


def scatter_forward_all(scatter_data, nlocals, buffers):
    """cuda/scatterer.py:191-277 over all ranks in-process: owner value ->
    every ghost copy.  ``buffers[r]`` is rank r's [owned | ghost] vector."""
    mail = {}
    for r, (owners_data, ghosts_data) in enumerate(scatter_data):
        ghosts_idx, _, ghosts = ghosts_data
        for i, dest in enumerate(ghosts):
            sb = np.zeros(ghosts_idx[i].size, dtype=buffers[r].dtype)
            pack_fwd(buffers[r], sb, np.ascontiguousarray(ghosts_idx[i], dtype=np.int64))
            mail[(r, int(dest))] = sb
    for r, (owners_data, ghosts_data) in enumerate(scatter_data):
        owners_idx, _, owners = owners_data
        for i, src in enumerate(owners):
            unpack_fwd(mail[(int(src), r)], buffers[r],
                       np.ascontiguousarray(owners_idx[i], dtype=np.int64), nlocals[r])


def scatter_reverse_all(scatter_data, nlocals, buffers):
    """cuda/scatterer.py:104-188: ghost-region partial sums -> added into the
    owner; the ghost region is left as it was."""
    mail = {}
    for r, (owners_data, ghosts_data) in enumerate(scatter_data):
        owners_idx, _, owners = owners_data
        for i, dest in enumerate(owners):
            sb = np.zeros(owners_idx[i].size, dtype=buffers[r].dtype)
            pack_rev(buffers[r], sb, np.ascontiguousarray(owners_idx[i], dtype=np.int64), nlocals[r])
            mail[(r, int(dest))] = sb
    for r, (owners_data, ghosts_data) in enumerate(scatter_data):
        ghosts_idx, _, ghosts = ghosts_data
        for i, src in enumerate(ghosts):
            unpack_rev(mail[(int(src), r)], buffers[r],
                       np.ascontiguousarray(ghosts_idx[i], dtype=np.int64))


# --------------------------------------------------------------------------- #
# RK4 loops
# --------------------------------------------------------------------------- #

A_RUNGE = np.array([0.0, 0.5, 0.5, 1.0])
B_RUNGE = np.array([1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0])
C_RUNGE = np.array([0.0, 0.5, 0.5, 1.0])


def linear_source(t, f0, p0, c0, alpha=4.0):
    """numba-cpu/demo_linear_box.py:341-358 (cuda/demo_linear_box.py:511-530)."""
    T = 1.0 / f0
    if t < T * alpha:
        window = 0.5 * (1.0 - np.cos(f0 * np.pi * t / alpha))
    else:
        window = 1.0
    return window * p0 * 2.0 * np.pi * f0 / c0 * np.cos(2.0 * np.pi * f0 * t)


class LinearProblem:
    """Arrays of one rank of the linear demo (cuda/demo_linear_box.py:336-385)."""

    def __init__(self, P, dofmap, G, dphi_1D, cell_coeff2, m, bfacet_dofmap1, detJ_f1,
                 facet_coeff1, bfacet_dofmap2, detJ_f2, facet_coeff2, f0, p0, c0):
        self.__dict__.update(locals())


def linear_rk4(prob: LinearProblem, u_, v_, t, dt, nsteps, scatter_fwd=None, scatter_rev=None):
    """The loop of numba-cpu/demo_linear_box.py:425-459 with f() of :322-382
    (source evaluated at the stage time tn, quirk Q1).  Single rank by
    default; ``scatter_fwd/rev`` are callables on the local vector."""
    dtp = u_.dtype
    nd = u_.size
    un, vn, u0, v0 = (np.zeros(nd, dtp) for _ in range(4))
    ku, kv = u_.copy(), v_.copy()  # ku = u0.copy(), kv = v0.copy()  (:405-406)
    g, u_n, v_n, b = (np.zeros(nd, dtp) for _ in range(4))
    for _ in range(nsteps):
        u0[:] = u_
        v0[:] = v_
        for i in range(4):
            un[:] = u0
            vn[:] = v0
            un += dtp.type(A_RUNGE[i] * dt) * ku
            vn += dtp.type(A_RUNGE[i] * dt) * kv
            tn = t + C_RUNGE[i] * dt
            ku[:] = vn
            g[:] = linear_source(tn, prob.f0, prob.p0, prob.c0)
            u_n[:] = un
            v_n[:] = vn
            if scatter_fwd is not None:
                scatter_fwd(u_n)
                scatter_fwd(v_n)
            b[:] = 0.0
            stiffness_operator(prob.P, u_n, prob.cell_coeff2, b, prob.G, prob.dofmap, prob.dphi_1D)
            if prob.bfacet_dofmap1.shape[0]:
                mass_operator(g, prob.facet_coeff1, b, prob.detJ_f1, prob.bfacet_dofmap1)
            if prob.bfacet_dofmap2.shape[0]:
                mass_operator(v_n, prob.facet_coeff2, b, prob.detJ_f2, prob.bfacet_dofmap2)
            if scatter_rev is not None:
                scatter_rev(b)
            kv[:] = b / prob.m
            u_ += dtp.type(B_RUNGE[i] * dt) * ku
            v_ += dtp.type(B_RUNGE[i] * dt) * kv
        t += dt
    return t


def linear_leapfrog(prob: LinearProblem, u_, v_, t, dt, nsteps, absd=None):
    """Leapfrog (Stoermer-Verlet) restatement for the product's ``LinearLeapfrog3D`` - NOT in
    the reference (which implements RK4 only); same assembly as ``linear_rk4``'s f(), operation
    for operation what the CUDA step does:
        kick-off  : v <- v - dt/2 * b(t, u, v) / m
        each step : b = K u + g(t) src + absb v ; v += dt * b / (m - dt/2 absd) ; u += dt * v
    ``absd`` = the absorbing facet mass applied to ones (default: computed here).  On return
    ``u_`` is u(t + nsteps dt) and ``v_`` is v at the last half step."""
    dt_type = u_.dtype.type
    nd = u_.size
    b = np.zeros(nd, u_.dtype)
    g = np.zeros(nd, u_.dtype)
    if absd is None:
        absd = np.zeros(nd, u_.dtype)
        if prob.bfacet_dofmap2.shape[0]:
            mass_operator(np.ones(nd, u_.dtype), prob.facet_coeff2, absd, prob.detJ_f2, prob.bfacet_dofmap2)

    def assemble(tt):
        fill(dt_type(linear_source(tt, prob.f0, prob.p0, prob.c0)), g)
        fill(dt_type(0.0), b)
        stiffness_operator(prob.P, u_, prob.cell_coeff2, b, prob.G, prob.dofmap, prob.dphi_1D)
        if prob.bfacet_dofmap1.shape[0]:
            mass_operator(g, prob.facet_coeff1, b, prob.detJ_f1, prob.bfacet_dofmap1)
        if prob.bfacet_dofmap2.shape[0]:
            mass_operator(v_, prob.facet_coeff2, b, prob.detJ_f2, prob.bfacet_dofmap2)

    assemble(t)
    v_ += dt_type(-0.5 * dt) * (b / prob.m)
    mlf = prob.m - dt_type(0.5 * dt) * absd
    for _ in range(nsteps):
        assemble(t)
        v_ += dt_type(dt) * (b / mlf)
        u_ += dt_type(dt) * v_
        t += dt
    return t


def westervelt_source(t, f0, p0, c0, alpha=4.0):
    """cuda/demo_nonlinear_bowl.py:560-594: g and dg/dt."""
    T = 1.0 / f0
    w0 = 2.0 * np.pi * f0
    if t < T * alpha:
        window = 0.5 * (1.0 - np.cos(f0 * np.pi * t / alpha))
        dwindow = 0.5 * np.pi * f0 / alpha * np.sin(f0 * np.pi * t / alpha)
    else:
        window, dwindow = 1.0, 0.0
    g = window * 2.0 * p0 * w0 / c0 * np.cos(w0 * t)
    dg = dwindow * 2.0 * p0 * w0 / c0 * np.cos(w0 * t) - window * 2.0 * p0 * w0**2 / c0 * np.sin(w0 * t)
    return g, dg


class WesterveltProblem:
    """Arrays of the Westervelt demo (cuda/demo_nonlinear_bowl.py:358-421).
    cell_coeff2..5, facet coefficients as named there; m0 the steady LHS
    (:459-469)."""

    def __init__(self, P, dofmap, G, detJ, dphi_1D, cell_coeff2, cell_coeff3, cell_coeff4,
                 cell_coeff5, m0, bfacet_dofmap1, detJ_f1, facet_coeff1_1, facet_coeff2_1,
                 bfacet_dofmap2, detJ_f2, facet_coeff2_2, f0, p0, c0):
        self.__dict__.update(locals())


def westervelt_rk4(prob: WesterveltProblem, u_, v_, t, dt, nsteps, scatter_fwd=None,
                   scatter_rev=None, source_at_stage_time=True):
    """cuda/demo_nonlinear_bowl.py:529-657 with the numba-cpu operators (the
    reference has no CPU twin of this loop).  The CUDA demo evaluates the
    source at t (quirk Q1); ``source_at_stage_time`` selects tn as the
    CPU/C++ linear paths do."""
    dtp = u_.dtype
    nd = u_.size
    un, vn, u0, v0 = (np.zeros(nd, dtp) for _ in range(4))
    ku, kv = u_.copy(), v_.copy()
    g, dg, u_n, v_n, w_n, b, m = (np.zeros(nd, dtp) for _ in range(7))
    for _ in range(nsteps):
        u0[:] = u_
        v0[:] = v_
        for i in range(4):
            un[:] = u0
            vn[:] = v0
            un += dtp.type(A_RUNGE[i] * dt) * ku
            vn += dtp.type(A_RUNGE[i] * dt) * kv
            tn = t + C_RUNGE[i] * dt
            ku[:] = vn
            gv, dgv = westervelt_source(tn if source_at_stage_time else t, prob.f0, prob.p0, prob.c0)
            g[:] = gv
            dg[:] = dgv
            u_n[:] = un
            v_n[:] = vn
            w_n[:] = vn * vn
            if scatter_fwd is not None:
                scatter_fwd(u_n)
                scatter_fwd(v_n)
                scatter_fwd(w_n)
            m[:] = 0.0
            mass_operator(u_n, prob.cell_coeff2, m, prob.detJ, prob.dofmap)
            if scatter_rev is not None:
                scatter_rev(m)
            m += prob.m0
            b[:] = 0.0
            stiffness_operator(prob.P, u_n, prob.cell_coeff3, b, prob.G, prob.dofmap, prob.dphi_1D)
            stiffness_operator(prob.P, v_n, prob.cell_coeff4, b, prob.G, prob.dofmap, prob.dphi_1D)
            mass_operator(w_n, prob.cell_coeff5, b, prob.detJ, prob.dofmap)
            if prob.bfacet_dofmap1.shape[0]:
                mass_operator(g, prob.facet_coeff1_1, b, prob.detJ_f1, prob.bfacet_dofmap1)
                mass_operator(dg, prob.facet_coeff2_1, b, prob.detJ_f1, prob.bfacet_dofmap1)
            if prob.bfacet_dofmap2.shape[0]:
                mass_operator(v_n, prob.facet_coeff2_2, b, prob.detJ_f2, prob.bfacet_dofmap2)
            if scatter_rev is not None:
                scatter_rev(b)
            kv[:] = b / m
            u_ += dtp.type(B_RUNGE[i] * dt) * ku
            v_ += dtp.type(B_RUNGE[i] * dt) * kv
        t += dt
    return t


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb if nb > 0 else 1.0))


def eval_points(u, dofmap, cells, phi):
    """``out[p] = sum_ijk phi[p,0,i] phi[p,1,j] phi[p,2,k] u[dofmap[cells[p], i n^2 + j n + k]]``:
    the tensor-product Lagrange interpolant at sample points, i.e. what
    ``u_n_.eval(x_eval, cell_eval)`` computes in cuda/demo_linear_piston.py:569 (DOLFINx is
    absent: pinned by polynomial known answers only).  float64 accumulation."""
    phi = np.asarray(phi, dtype=np.float64)
    n = phi.shape[2]
    ue = np.asarray(u, dtype=np.float64)[np.asarray(dofmap)[np.asarray(cells, dtype=np.int64)]]
    return np.einsum("mi,mj,mk,mijk->m", phi[:, 0], phi[:, 1], phi[:, 2], ue.reshape(-1, n, n, n))
