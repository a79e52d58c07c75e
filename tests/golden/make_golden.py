"""
Golden-vector generator.  Run ONCE in the build container (it needs
/root/reference and numba):

    python tests/golden/make_golden.py [--only-irregular]

It imports the reference's own Python - ``numba-cpu/operators.py``,
``numba-cpu/sum_factorisation.py``, ``numba-cpu/scatterer.py``,
``cuda/precompute.py`` and ``cuda/utils.py`` - UNMODIFIED from
/root/reference, feeds them seeded synthetic inputs, and stores inputs and
outputs as small ``.npz`` fixtures next to this file.  ``mpi4py`` and
``dolfinx`` (absent in this image) are replaced by stub modules: a threaded
in-process mailbox that implements ``Isend/Irecv/Waitall`` and empty
``dolfinx`` shells (only imported for type annotations).

The fixtures are what pins ``oracle/`` (tests/test_oracle_golden.py) and the
CUDA path (tests/test_gpu_*.py) to the reference; nothing on the GPU box
reads /root/reference.
"""

import os
import queue
import sys
import threading
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from fenicsx_fus_gpu_b200 import substrate as S  # noqa: E402

# --------------------------------------------------------------------------- #
# stub modules: mpi4py (threaded mailbox) and dolfinx shells
# --------------------------------------------------------------------------- #

_tls = threading.local()
_mail = {}
_mail_lock = threading.Lock()


def _box(src, dst):
    with _mail_lock:
        return _mail.setdefault((src, dst), queue.Queue())


class _Req:
    def __init__(self, fn=None):
        self.fn = fn

    def Wait(self):
        if self.fn:
            self.fn()


class _Comm:
    @property
    def rank(self):
        return _tls.rank

    @property
    def size(self):
        return _tls.size

    def Isend(self, buf, dest):
        _box(self.rank, int(dest)).put(np.array(buf, copy=True))
        return _Req()

    def Irecv(self, buf, source):
        me = self.rank

        def done():
            buf[...] = _box(int(source), me).get(timeout=60)

        return _Req(done)


class _Request:
    @staticmethod
    def Waitall(reqs):
        for r in reqs:
            r.Wait()


def install_stubs():
    mpi4py = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")
    MPI.COMM_WORLD = _Comm()
    MPI.Comm = _Comm
    MPI.Request = _Request
    mpi4py.MPI = MPI
    sys.modules["mpi4py"] = mpi4py
    sys.modules["mpi4py.MPI"] = MPI
    dolfinx = types.ModuleType("dolfinx")
    mesh = types.ModuleType("dolfinx.mesh")
    mesh.Mesh = object
    geom = types.ModuleType("dolfinx.geometry")
    geom.bb_tree = geom.compute_collisions_points = geom.compute_colliding_cells = None
    dolfinx.mesh, dolfinx.geometry = mesh, geom
    sys.modules["dolfinx"] = dolfinx
    sys.modules["dolfinx.mesh"] = mesh
    sys.modules["dolfinx.geometry"] = geom
    return MPI


def run_ranks(nranks, fn):
    """Run fn(rank) on nranks threads with the stub communicator."""
    out = [None] * nranks
    err = []

    def body(r):
        _tls.rank, _tls.size = r, nranks
        try:
            out[r] = fn(r)
        except Exception as e:  # pragma: no cover
            err.append(e)

    th = [threading.Thread(target=body, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    return out


def load_ref(subdir, name, alias):
    """Import /root/reference/<subdir>/<name>.py as module ``alias``."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(alias, os.path.join(REF, subdir, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def scatter_case(name, parts, ncells, P, nranks, MPI, utils, scat):
    """Index maps of ``parts`` through the reference's cuda/utils.py compute_scatterer_data and one
    forward / reverse halo through its numba-cpu/scatterer.py -> scatter_<name>.npz."""
    imaps = [p.index_map for p in parts]
    res = run_ranks(nranks, lambda r: utils.compute_scatterer_data(imaps[r]))
    store = {"nranks": nranks, "P": P, "ncells": np.array(ncells)}
    for r, (od, gd) in enumerate(res):
        im = imaps[r]
        store[f"r{r}_size_local"] = im.size_local
        store[f"r{r}_local_range"] = np.array(im.local_range)
        store[f"r{r}_ghosts"] = im.ghosts
        store[f"r{r}_owners"] = im.owners
        store[f"r{r}_dest_array"] = im.index_to_dest_ranks().array
        store[f"r{r}_dest_offsets"] = im.index_to_dest_ranks().offsets
        store[f"r{r}_owners_ranks"] = np.asarray(od[2])
        store[f"r{r}_owners_size"] = np.asarray(od[1])
        store[f"r{r}_ghosts_ranks"] = np.asarray(gd[2])
        store[f"r{r}_ghosts_size"] = np.asarray(gd[1])
        for i, a_ in enumerate(od[0]):
            store[f"r{r}_owners_idx{i}"] = np.asarray(a_, dtype=np.int64)
        for i, a_ in enumerate(gd[0]):
            store[f"r{r}_ghosts_idx{i}"] = np.asarray(a_, dtype=np.int64)

    # halo exchange through numba-cpu/scatterer.py (4-tuple, flat indices)
    rng = np.random.default_rng(11)
    vecs = [rng.standard_normal(im.size_local + im.num_ghosts) for im in imaps]

    def cpu_data(od, gd):
        def flat(d):
            idx = np.concatenate([np.asarray(v, np.int64) for v in d[0]]) if len(d[0]) else np.zeros(0, np.int64)
            size = np.asarray(d[1], dtype=np.int64)
            return [idx, size, np.insert(np.cumsum(size), 0, 0), np.asarray(d[2])]
        return flat(od), flat(gd)

    def do_rev(r):
        od, gd = cpu_data(*res[r])
        v = vecs[r].copy()
        scat.scatter_reverse(MPI.COMM_WORLD, od, gd, imaps[r].size_local, np.float64)(v)
        return v

    def do_fwd(r):
        od, gd = cpu_data(*res[r])
        v = vecs[r].copy()
        scat.scatter_forward(MPI.COMM_WORLD, od, gd, imaps[r].size_local, np.float64)(v)
        return v

    rev = run_ranks(nranks, do_rev)
    fwd = run_ranks(nranks, do_fwd)
    for r in range(nranks):
        store[f"r{r}_vec"] = vecs[r]
        store[f"r{r}_rev"] = rev[r]
        store[f"r{r}_fwd"] = fwd[r]
    np.savez_compressed(os.path.join(HERE, f"scatter_{name}.npz"), **store)
    print(f"scatter {name}: ranks={nranks} ghosts={[im.num_ghosts for im in imaps]}")


def irregular_scatter_cases(MPI, utils, scat):
    """Unstructured-like partitions (substrate.partition_cells): blob-shaped parts, shuffled cell /
    dof / ghost order, lowest-rank and pseudo-random ownership of the shared dofs."""
    for name, ncells, P, nranks, rule in (("u5", (5, 4, 3), 2, 5, "hash"), ("u4", (4, 4, 3), 3, 4, "lowest")):
        cr = S.blob_cell_ranks(ncells, nranks, seed=4)
        parts = S.partition_cells(ncells, P, cr, shuffle_seed=7, owner_rule=rule)
        scatter_case(name, parts, ncells, P, nranks, MPI, utils, scat)


def main():
    MPI = install_stubs()
    sys.path.insert(0, os.path.join(REF, "numba-cpu"))  # `from sum_factorisation import ...`
    ops = load_ref("numba-cpu", "operators", "ref_operators")
    pre = load_ref("cuda", "precompute", "ref_precompute")
    utils = load_ref("cuda", "utils", "ref_utils")
    scat = load_ref("numba-cpu", "scatterer", "ref_scatterer")
    if "--only-irregular" in sys.argv:  # add the unstructured-like fixtures without touching the others
        irregular_scatter_cases(MPI, utils, scat)
        return

    # ---------------- operators + geometry, P = 2..7, f32/f64 ---------------- #
    for P in range(2, 8):
        for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
            tb = S.element_tables(P, "basix", dt)
            mesh = S.create_box((2, 2, 3), (1.0, 0.8, 1.2), dtype=dt, perturb=0.15, seed=P)
            dofmap = S.tensor_dofmap(mesh, P)
            nd = int(dofmap.max()) + 1
            Nc, Nd = dofmap.shape
            rng = np.random.default_rng(100 + P)
            x = rng.standard_normal(nd).astype(dt)
            coeff = rng.uniform(0.5, 2.0, Nc).astype(dt)

            detJ = np.zeros((Nc, Nd), dt)
            pre.compute_scaled_jacobian_determinant(detJ, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
            G = np.zeros((Nc, Nd, 6), dt)
            pre.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)

            bdata = np.concatenate([S.boundary_facets(mesh, f) for f in range(6)])
            detJ_f = np.zeros((bdata.shape[0], tb.n**2), dt)
            pre.compute_boundary_facets_scaled_jacobian_determinant(
                detJ_f, (mesh.x_dofs, mesh.x_g), bdata, tb.dphi_f, tb.wts_f)
            bdofmap = S.facet_dofmap(dofmap, bdata, tb.local_facet_dof)
            fcoeff = rng.uniform(0.5, 2.0, bdata.shape[0]).astype(dt)

            y_mass = np.zeros(nd, dt)
            ops.mass_operator(Nd, dt)(x, coeff, y_mass, detJ, dofmap)
            y_stiff = np.zeros(nd, dt)
            ops.stiffness_operator(P, tb.dphi_1D.flatten(), dt)(x, coeff, y_stiff, G, dofmap)
            y_fmass = np.zeros(nd, dt)
            ops.mass_operator(tb.n**2, dt)(x, fcoeff, y_fmass, detJ_f, bdofmap)

            np.savez_compressed(
                os.path.join(HERE, f"operators_P{P}_{tag}.npz"),
                P=P, x_dofs=mesh.x_dofs, x_g=mesh.x_g, dofmap=dofmap, dphi_1D=tb.dphi_1D,
                dphi=tb.dphi, wts=tb.wts, dphi_f=tb.dphi_f, wts_f=tb.wts_f, x=x, coeff=coeff,
                detJ=detJ, G=G, bdata=bdata, bdofmap=bdofmap, detJ_f=detJ_f, fcoeff=fcoeff,
                y_mass=y_mass, y_stiff=y_stiff, y_fmass=y_fmass,
            )
            print(f"operators P={P} {tag}: nd={nd} |y_stiff|={np.linalg.norm(y_stiff):.6e}")

    # vector kernels (numba-cpu/operators.py:230-300)
    rng = np.random.default_rng(7)
    a = rng.standard_normal(1000)
    b = rng.uniform(0.5, 2.0, 1000)
    y = rng.standard_normal(1000)
    y_axpy = y.copy()
    ops.axpy(1000)(0.37, a, y_axpy)
    c_div = np.zeros(1000)
    ops.pointwise_divide(a, b, c_div)
    b_copy = np.zeros(1000)
    ops.copy(a, b_copy)
    f_fill = np.zeros(1000)
    ops.fill(2.5, f_fill)
    np.savez_compressed(os.path.join(HERE, "vector_ops.npz"), a=a, b=b, y=y, alpha=0.37,
                        y_axpy=y_axpy, c_div=c_div, b_copy=b_copy, f_fill=f_fill)

    # ---------------- index maps + halo, 2x2x2 and 3x1x1 partitions ---------- #
    for name, ncells, P, nranks, grid in (("r8", (4, 4, 4), 2, 8, None), ("r3", (6, 2, 2), 3, 3, (3, 1, 1)),
                                          ("r2", (4, 3, 2), 4, 2, None)):
        scatter_case(name, S.partition_box(ncells, P, nranks, grid=grid), ncells, P, nranks, MPI, utils, scat)
    irregular_scatter_cases(MPI, utils, scat)

    # ---------------- linear RK4 loop with the reference operators ----------- #
    # Statement sequence of numba-cpu/demo_linear_box.py:322-382, 425-459,
    # single rank (scatters are no-ops), reference numba kernels unmodified.
    P, dt_ = 3, np.float64
    tb = S.element_tables(P, "basix", dt_)
    L = 0.012
    mesh = S.create_box((3, 3, 3), (L, L, L), dtype=dt_, perturb=0.1, seed=3)
    dofmap = S.tensor_dofmap(mesh, P)
    nd = int(dofmap.max()) + 1
    Nc, Nd = dofmap.shape
    rho, c0, f0, p0 = 1000.0, 1500.0, 0.5e6, 60000.0
    detJ = np.zeros((Nc, Nd))
    pre.compute_scaled_jacobian_determinant(detJ, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    G = np.zeros((Nc, Nd, 6))
    pre.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    bd1, bd2 = S.boundary_facets(mesh, 2), S.boundary_facets(mesh, 3)
    dJ1 = np.zeros((bd1.shape[0], tb.n**2))
    dJ2 = np.zeros((bd2.shape[0], tb.n**2))
    pre.compute_boundary_facets_scaled_jacobian_determinant(dJ1, (mesh.x_dofs, mesh.x_g), bd1, tb.dphi_f, tb.wts_f)
    pre.compute_boundary_facets_scaled_jacobian_determinant(dJ2, (mesh.x_dofs, mesh.x_g), bd2, tb.dphi_f, tb.wts_f)
    fd1 = S.facet_dofmap(dofmap, bd1, tb.local_facet_dof)
    fd2 = S.facet_dofmap(dofmap, bd2, tb.local_facet_dof)
    cc1 = np.full(Nc, 1.0 / rho / c0 / c0)
    cc2 = np.full(Nc, -1.0 / rho)
    fc1 = np.full(bd1.shape[0], 1.0 / rho)
    fc2 = np.full(bd2.shape[0], -1.0 / rho / c0)
    mass_c = ops.mass_operator(Nd, dt_)
    mass_f = ops.mass_operator(tb.n**2, dt_)
    stiff = ops.stiffness_operator(P, tb.dphi_1D.flatten(), dt_)
    axpy = ops.axpy(nd)
    m = np.zeros(nd)
    mass_c(np.ones(nd), cc1, m, detJ, dofmap)
    h = L / 3
    dt = 0.65 * h / (c0 * P**2)
    u_, v_ = np.zeros(nd), np.zeros(nd)
    un, vn, u0, v0 = (np.zeros(nd) for _ in range(4))
    ku, kv = u0.copy(), v0.copy()
    g, u_n, v_n, b = (np.zeros(nd) for _ in range(4))
    a_r = np.array([0.0, 0.5, 0.5, 1.0])
    b_r = np.array([1 / 6, 1 / 3, 1 / 3, 1 / 6])
    t = 0.0
    nsteps = 20
    for _ in range(nsteps):
        ops.copy(u_, u0)
        ops.copy(v_, v0)
        for i in range(4):
            ops.copy(u0, un)
            ops.copy(v0, vn)
            axpy(a_r[i] * dt, ku, un)
            axpy(a_r[i] * dt, kv, vn)
            tn = t + a_r[i] * dt
            ops.copy(vn, ku)
            T_, alpha = 1 / f0, 4
            window = 0.5 * (1 - np.cos(f0 * np.pi * tn / alpha)) if tn < T_ * alpha else 1.0
            ops.fill(window * p0 * 2 * np.pi * f0 / c0 * np.cos(2 * np.pi * f0 * tn), g)
            ops.copy(un, u_n)
            ops.copy(vn, v_n)
            ops.fill(0.0, b)
            stiff(u_n, cc2, b, G, dofmap)
            mass_f(g, fc1, b, dJ1, fd1)
            mass_f(v_n, fc2, b, dJ2, fd2)
            ops.pointwise_divide(b, m, kv)
            axpy(b_r[i] * dt, ku, u_)
            axpy(b_r[i] * dt, kv, v_)
        t += dt
    np.savez_compressed(
        os.path.join(HERE, "linear_rk4_P3.npz"), P=P, L=L, ncells=3, perturb=0.1, seed=3,
        rho=rho, c0=c0, f0=f0, p0=p0, dt=dt, nsteps=nsteps, x_dofs=mesh.x_dofs, x_g=mesh.x_g,
        dofmap=dofmap, G=G, detJ=detJ, m=m, u=u_, v=v_, t_final=t,
    )
    print(f"linear rk4: nd={nd} |u|={np.linalg.norm(u_):.6e} |v|={np.linalg.norm(v_):.6e}")


if __name__ == "__main__":
    main()
