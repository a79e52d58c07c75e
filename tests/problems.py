"""Shared problem builders for the tests: small box problems in the array
layouts of the reference demos (cuda/demo_linear_box.py:336-385,
cuda/demo_nonlinear_bowl.py:358-421), built on the synthetic substrate with the
geometry tables from the CPU oracle."""

import numpy as np

from fenicsx_fus_gpu_b200 import substrate as S
from oracle import oracle as orc


class Data(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def geometry(mesh, tb, dt):
    Nc = mesh.num_cells
    Nd = tb.n**3
    detJ = np.zeros((Nc, Nd), dt)
    G = np.zeros((Nc, Nd, 6), dt)
    orc.compute_scaled_jacobian_determinant(detJ, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    orc.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    return G, detJ


def facets(mesh, dofmap, tb, dt, local_facets):
    bd = np.concatenate([S.boundary_facets(mesh, f) for f in local_facets]) if local_facets else np.zeros((0, 2), np.int32)
    dJ = np.zeros((bd.shape[0], tb.n**2), dt)
    if bd.shape[0]:
        orc.compute_boundary_facets_scaled_jacobian_determinant(dJ, (mesh.x_dofs, mesh.x_g), bd, tb.dphi_f, tb.wts_f)
    return bd, dJ, S.facet_dofmap(dofmap, bd, tb.local_facet_dof)


def linear_problem(P, ncells, L, dt=np.float64, perturb=0.1, seed=3, mesh=None, dofmap=None,
                   ndofs=None, rho=1000.0, c0=1500.0, f0=0.5e6, p0=60000.0,
                   src_facets=(2,), abs_facets=(3,)):
    """Arrays of the linear demo on one (part of a) box: source on x=0,
    absorbing on x=L (cuda/demo_linear_box.py:256-263)."""
    tb = S.element_tables(P, "basix", dt)
    if mesh is None:
        mesh = S.create_box(ncells, L, dtype=dt, perturb=perturb, seed=seed)
        dofmap = S.tensor_dofmap(mesh, P)
        ndofs = int(dofmap.max()) + 1
    Nc = mesh.num_cells
    G, detJ = geometry(mesh, tb, dt)
    bd1, dJ1, fd1 = facets(mesh, dofmap, tb, dt, list(src_facets))
    bd2, dJ2, fd2 = facets(mesh, dofmap, tb, dt, list(abs_facets))
    return Data(
        P=P, tb=tb, mesh=mesh, dofmap=dofmap, ndofs=ndofs, G=G, detJ=detJ,
        cell_coeff1=np.full(Nc, 1.0 / rho / c0 / c0, dt), cell_coeff2=np.full(Nc, -1.0 / rho, dt),
        bfacet_dofmap1=fd1, detJ_f1=dJ1, facet_coeff1=np.full(bd1.shape[0], 1.0 / rho, dt),
        bfacet_dofmap2=fd2, detJ_f2=dJ2, facet_coeff2=np.full(bd2.shape[0], -1.0 / rho / c0, dt),
        rho=rho, c0=c0, f0=f0, p0=p0,
    )


def westervelt_problem(P, ncells, L, dt=np.float64, perturb=0.1, seed=5, mesh=None, dofmap=None,
                       ndofs=None, rho=1000.0, c0=1480.0, f0=1.1e6, p0=1.0e6, beta=3.5,
                       alpha_dB=20.0, src_facets=(2,), abs_facets=(0, 1, 3, 4, 5)):
    """Arrays of the Westervelt demo (cuda/demo_nonlinear_bowl.py:358-374)."""
    from fenicsx_fus_gpu_b200.utils import compute_diffusivity_of_sound

    d = linear_problem(P, ncells, L, dt, perturb, seed, mesh, dofmap, ndofs, rho, c0, f0, p0,
                       src_facets, abs_facets)
    Nc = d.mesh.num_cells
    delta = compute_diffusivity_of_sound(2.0 * np.pi * f0, c0, alpha_dB)
    n1, n2 = d.bfacet_dofmap1.shape[0], d.bfacet_dofmap2.shape[0]
    d.update(
        delta=delta, beta=beta,
        cell_coeff2=np.full(Nc, -2.0 * beta / rho / rho / c0**4, dt),
        cell_coeff3=np.full(Nc, -1.0 / rho, dt),
        cell_coeff4=np.full(Nc, -delta / rho / c0 / c0, dt),
        cell_coeff5=np.full(Nc, 2.0 * beta / rho / rho / c0**4, dt),
        facet_coeff1_1=np.full(n1, 1.0 / rho, dt), facet_coeff2_1=np.full(n1, delta / rho / c0 / c0, dt),
        facet_coeff1_2=np.full(n2, delta / rho / c0**3, dt), facet_coeff2_2=np.full(n2, -1.0 / rho / c0, dt),
    )
    return d


def cfl_dt(P, h, c0, f0, cfl=0.65):
    """cuda/demo_linear_box.py:116-120."""
    dt = cfl * h / (c0 * P**2)
    period = 1.0 / f0
    return period / (int(period / dt) + 1)


def fake_dolfinx(mesh, dofmap, perm, index_map=None):
    """Stand-ins for a DOLFINx mesh and function space exposing exactly the attributes the
    reference demos (and dolfinx_bridge.setup_from_dolfinx / utils.facet_integration_domain)
    read: mesh.topology.{dim, index_map(d).size_local, connectivity(a, b)},
    mesh.geometry.{dofmap, x}, V.dofmap.{list, index_map}.  ``V.dofmap.list[:, perm]`` is the
    tensor-product ``dofmap``.  Returns ``(mesh_like, V_like, facet_id(cell, local_facet))``."""
    Nx, Ny, Nz = mesh.ncells
    ncell = Nx * Ny * Nz
    key_of, f2c = {}, {}
    c2f = np.zeros((ncell, 6), np.int32)
    for cx in range(Nx):
        for cy in range(Ny):
            for cz in range(Nz):
                c = (cx * Ny + cy) * Nz + cz
                for lf in range(6):
                    axis, side = S._FACE_AXIS[lf]
                    pos = [cx, cy, cz]
                    pos[axis] += side
                    fid = key_of.setdefault((axis, tuple(pos)), len(key_of))
                    c2f[c, lf] = fid
                    f2c.setdefault(fid, []).append(c)

    class _Sized:
        def __init__(self, n):
            self.size_local = n

    class Topo:
        dim = 3

        def index_map(self, d):
            return _Sized(ncell if d == 3 else len(f2c))

        def connectivity(self, a, b):
            if (a, b) == (3, 2):
                return S.AdjacencyList(c2f.ravel(), np.arange(0, 6 * ncell + 1, 6))
            offs = np.cumsum([0] + [len(f2c[f]) for f in range(len(f2c))])
            return S.AdjacencyList(np.concatenate([f2c[f] for f in range(len(f2c))]), offs)

    class Geo:
        pass

    class M:
        topology = Topo()
        geometry = Geo()

    M.geometry.dofmap, M.geometry.x = mesh.x_dofs, mesh.x_g

    class DM:
        pass

    class V:
        dofmap = DM()

    nd = int(dofmap.max()) + 1
    V.dofmap.list = np.ascontiguousarray(dofmap[:, np.argsort(perm)])
    V.dofmap.index_map = index_map if index_map is not None else S.serial_index_map(nd)
    return M(), V(), lambda c, lf: c2f[c, lf]
