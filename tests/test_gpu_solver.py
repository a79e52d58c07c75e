"""GPU parity of the fused RK4 loops and the halo exchange.

* linear loop against the fixture produced by the reference's own numba-cpu
  kernels (tests/golden/linear_rk4_P3.npz) and against the oracle loop;
* Westervelt loop against the oracle restatement of
  cuda/demo_nonlinear_bowl.py:529-657;
* the multi-rank path (block partition, index maps, forward/reverse halo,
  ghost handling) with the ranks emulated by threads on one GPU, against the
  single-rank oracle;
* halo exchange against the reference-generated scatter fixtures.

Tolerances: rel-L2 <= 1e-12 (float64), <= 1e-5 (float32) - BASELINE.json.
"""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _linear_solver(d, dt_type, halo=None, **kw):
    from fenicsx_fus_gpu_b200.solver import LinearSpectral3D, linear_source

    return LinearSpectral3D(
        d.P, dt_type, d.ndofs, d.dofmap, d.G, d.detJ, d.tb.dphi_1D, d.cell_coeff1, d.cell_coeff2,
        d.bfacet_dofmap1, d.detJ_f1, d.facet_coeff1, d.bfacet_dofmap2, d.detJ_f2, d.facet_coeff2,
        halo=halo, source=lambda t: linear_source(t, d.f0, d.p0, d.c0), **kw)


def _oracle_linear(d, dt, nsteps, dt_type):
    from oracle import oracle as orc

    m = np.zeros(d.ndofs, dt_type)
    orc.mass_operator(np.ones(d.ndofs, dt_type), d.cell_coeff1, m, d.detJ, d.dofmap)
    prob = orc.LinearProblem(d.P, d.dofmap, d.G, d.tb.dphi_1D, d.cell_coeff2, m, d.bfacet_dofmap1,
                             d.detJ_f1, d.facet_coeff1, d.bfacet_dofmap2, d.detJ_f2, d.facet_coeff2,
                             d.f0, d.p0, d.c0)
    u, v = np.zeros(d.ndofs, dt_type), np.zeros(d.ndofs, dt_type)
    orc.linear_rk4(prob, u, v, 0.0, dt, nsteps)
    return u, v


def test_linear_rk4_vs_reference_golden(golden_dir):
    """20 RK4 steps, every kernel of the stage, against the loop run with the
    reference's unmodified numba-cpu operators."""
    import problems
    from fenicsx_fus_gpu_b200 import substrate as S

    g = np.load(os.path.join(golden_dir, "linear_rk4_P3.npz"))
    P = int(g["P"])
    mesh = S.BoxMesh((3, 3, 3), g["x_dofs"], g["x_g"], (0, 0, 0), (3, 3, 3), (float(g["L"]),) * 3)
    d = problems.linear_problem(P, 3, float(g["L"]), mesh=mesh, dofmap=g["dofmap"],
                                ndofs=int(g["dofmap"].max()) + 1, rho=float(g["rho"]),
                                c0=float(g["c0"]), f0=float(g["f0"]), p0=float(g["p0"]))
    assert rel_l2(d.G, g["G"]) < 1e-12  # oracle geometry == reference geometry
    d.G, d.detJ = g["G"], g["detJ"]
    for use_graph in (True, False):
        s = _linear_solver(d, np.float64, use_graph=use_graph)
        assert rel_l2(s.m.cpu().numpy(), g["m"]) < 1e-13
        s.init()
        t = s.rk4(0.0, float(g["dt"]), int(g["nsteps"]))
        assert abs(t - float(g["t_final"])) < 1e-18
        assert rel_l2(s.u.cpu().numpy(), g["u"]) < 1e-12
        assert rel_l2(s.v.cpu().numpy(), g["v"]) < 1e-12


@pytest.mark.parametrize("P,N,tag,nsteps", [(4, 5, "f64", 12), (4, 5, "f32", 12), (2, 6, "f64", 10),
                                            (6, 3, "f64", 8)])
def test_linear_rk4_vs_oracle(P, N, tag, nsteps):
    import problems

    dtt = np.float64 if tag == "f64" else np.float32
    L = 0.01
    d = problems.linear_problem(P, N, L, dtt, perturb=0.12, seed=P)
    dt = problems.cfl_dt(P, L / N, d.c0, d.f0)
    u_ref, v_ref = _oracle_linear(d, dt, nsteps, dtt)
    s = _linear_solver(d, dtt)
    s.init()
    s.rk4(0.0, dt, nsteps)
    tol = 1e-12 if tag == "f64" else 1e-5
    assert rel_l2(s.u.cpu().numpy(), u_ref) < tol
    assert rel_l2(s.v.cpu().numpy(), v_ref) < tol
    # eager stepping with host-evaluated source scalars gives the same answer
    e = _linear_solver(d, dtt, use_graph=False)
    e.init()
    for _ in range(nsteps):
        e.step_eager(dt)
    assert rel_l2(e.u.cpu().numpy(), s.u.cpu().numpy()) < (1e-13 if tag == "f64" else 1e-6)
    # continuing the same solver (second rk4 call) == one long run
    s2 = _linear_solver(d, dtt)
    s2.init()
    s2.rk4(0.0, dt, nsteps // 2)
    s2.rk4(s2.t, dt, nsteps - nsteps // 2)
    assert rel_l2(s2.u.cpu().numpy(), s.u.cpu().numpy()) < (1e-13 if tag == "f64" else 1e-6)


@pytest.mark.parametrize("P,N,tag", [(4, 4, "f64"), (3, 5, "f64"), (4, 4, "f32")])
def test_westervelt_rk4_vs_oracle(P, N, tag):
    import problems
    from fenicsx_fus_gpu_b200.solver import WesterveltSpectral3D, westervelt_source
    from oracle import oracle as orc

    dtt = np.float64 if tag == "f64" else np.float32
    L = 0.006
    d = problems.westervelt_problem(P, N, L, dtt, perturb=0.1, seed=7)
    dt = problems.cfl_dt(P, L / N, d.c0, d.f0, cfl=0.4)
    nsteps = 10
    ones = np.ones(d.ndofs, dtt)
    m0 = np.zeros(d.ndofs, dtt)
    orc.mass_operator(ones, d.cell_coeff1, m0, d.detJ, d.dofmap)
    orc.mass_operator(ones, d.facet_coeff1_2, m0, d.detJ_f2, d.bfacet_dofmap2)
    prob = orc.WesterveltProblem(d.P, d.dofmap, d.G, d.detJ, d.tb.dphi_1D, d.cell_coeff2, d.cell_coeff3,
                                 d.cell_coeff4, d.cell_coeff5, m0, d.bfacet_dofmap1, d.detJ_f1,
                                 d.facet_coeff1_1, d.facet_coeff2_1, d.bfacet_dofmap2, d.detJ_f2,
                                 d.facet_coeff2_2, d.f0, d.p0, d.c0)
    u_ref, v_ref = np.zeros(d.ndofs, dtt), np.zeros(d.ndofs, dtt)
    orc.westervelt_rk4(prob, u_ref, v_ref, 0.0, dt, nsteps)
    assert np.linalg.norm(u_ref) > 0

    tol = 1e-12 if tag == "f64" else 1e-5
    # "pointwise": the cell-mass pair as un*m2, vn^2*m5 in the close kernel (default);
    # "cells": recomputed over the cells every stage, the reference's data flow
    for mass_form in ("pointwise", "cells"):
        s = WesterveltSpectral3D(
            d.P, dtt, d.ndofs, d.dofmap, d.G, d.detJ, d.tb.dphi_1D, d.cell_coeff1, d.cell_coeff2,
            d.cell_coeff3, d.cell_coeff4, d.cell_coeff5, d.bfacet_dofmap1, d.detJ_f1, d.facet_coeff1_1,
            d.facet_coeff2_1, d.bfacet_dofmap2, d.detJ_f2, d.facet_coeff1_2, d.facet_coeff2_2,
            source=lambda t: westervelt_source(t, d.f0, d.p0, d.c0), mass_form=mass_form)
        assert rel_l2(s.m0.cpu().numpy(), m0) < (1e-13 if tag == "f64" else 1e-6)
        s.init()
        s.rk4(0.0, dt, nsteps)
        assert rel_l2(s.u.cpu().numpy(), u_ref) < tol, mass_form
        assert rel_l2(s.v.cpu().numpy(), v_ref) < tol, mass_form


@pytest.mark.parametrize("name", ["r2", "r3", "r8", "u4", "u5"])
def test_halo_exchange_vs_reference_fixture(golden_dir, name):
    """Forward and reverse halo through HaloExchange (ranks emulated by threads
    on one GPU) against what numba-cpu/scatterer.py produced."""
    from fenicsx_fus_gpu_b200.scatterer import HaloExchange, LocalCluster

    with np.load(os.path.join(golden_dir, f"scatter_{name}.npz")) as z:
        g = {k: z[k] for k in z.files}  # read everything before the rank threads start
    R = int(g["nranks"])

    def lists(r, which):
        ranks = g[f"r{r}_{which}_ranks"]
        return [[g[f"r{r}_{which}_idx{i}"] for i in range(ranks.size)], g[f"r{r}_{which}_size"], ranks]

    def body(r, transport):
        N = int(g[f"r{r}_size_local"])
        halo = HaloExchange(transport, lists(r, "owners"), lists(r, "ghosts"), N, np.float64)
        v = torch.from_numpy(g[f"r{r}_vec"]).cuda()
        f = v.clone()
        halo.forward(f)
        rv = v.clone()
        halo.reverse(rv)
        # two vectors in one round == two rounds
        a, b2 = v.clone(), (2.0 * v).clone()
        halo.forward(a, b2)
        torch.cuda.synchronize()
        return f.cpu().numpy(), rv.cpu().numpy(), a.cpu().numpy(), b2.cpu().numpy()

    out = LocalCluster(R).run(body)
    for r in range(R):
        f, rv, a, b2 = out[r]
        assert np.array_equal(f, g[f"r{r}_fwd"])  # forward moves values: bit exact
        assert rel_l2(rv, g[f"r{r}_rev"]) < 1e-15  # reverse adds: order of atomics may differ
        assert np.array_equal(a, f)
        N = int(g[f"r{r}_size_local"])
        assert np.array_equal(b2[N:], 2.0 * f[N:])


@pytest.mark.parametrize("kind", ["nccl-shaped", "p2p"])
@pytest.mark.parametrize("R,N,P", [(2, (4, 3, 3), 3), (8, (4, 4, 4), 2), (4, (4, 4, 2), 4)])
def test_linear_rk4_partitioned_vs_serial_oracle(R, N, P, kind):
    """The multi-GPU algorithm end to end on one GPU: R partitions, each with
    its own solver + halo exchange, against the single-rank oracle run."""
    import problems
    from fenicsx_fus_gpu_b200 import substrate as S
    from fenicsx_fus_gpu_b200 import utils
    from fenicsx_fus_gpu_b200.scatterer import HaloExchange, LocalCluster, P2PHaloExchange, local_fabric

    dtt = np.float64
    L = (0.012, 0.01, 0.011)
    nsteps = 8
    serial = problems.linear_problem(P, N, L, dtt, perturb=0.1, seed=11)
    h = min(L[i] / N[i] for i in range(3))
    dt = problems.cfl_dt(P, h, serial.c0, serial.f0)
    u_ref, v_ref = _oracle_linear(serial, dt, nsteps, dtt)

    parts = S.partition_box(N, P, R, lengths=L, dtype=dtt, perturb=0.1, seed=11)
    sdata = utils.compute_scatterer_data_all([p.index_map for p in parts])

    def body(r, transport):
        p = parts[r]
        nd = p.index_map.size_local + p.index_map.num_ghosts
        d = problems.linear_problem(P, None, None, dtt, mesh=p.mesh, dofmap=p.dofmap, ndofs=nd)
        if kind == "p2p":  # halo fused into put / get_add kernels over peer-addressable memory
            ndmax = max(q.index_map.size_local + q.index_map.num_ghosts for q in parts)
            fab = local_fabric(transport.cluster, r, P2PHaloExchange.arena_bytes(ndmax, dtt))
            halo = P2PHaloExchange(fab, sdata[r][0], sdata[r][1], p.index_map.size_local,
                                   p.index_map.num_ghosts, dtt)
        else:  # pack -> grouped send/recv -> unpack
            halo = HaloExchange(transport, sdata[r][0], sdata[r][1], p.index_map.size_local, dtt)
        s = _linear_solver(d, dtt, halo=halo, use_graph=False)
        s.init()
        s.rk4(0.0, dt, nsteps)
        torch.cuda.synchronize()
        return s.u.cpu().numpy(), s.v.cpu().numpy()

    out = LocalCluster(R).run(body)
    u = np.zeros_like(u_ref)
    v = np.zeros_like(v_ref)
    for r, p in enumerate(parts):
        nl = p.index_map.size_local
        u[p.local_to_serial[:nl]] = out[r][0][:nl]
        v[p.local_to_serial[:nl]] = out[r][1][:nl]
    assert rel_l2(u, u_ref) < 1e-12
    assert rel_l2(v, v_ref) < 1e-12


def test_problem_builder_piston_vs_oracle():
    """problem.box_setup / linear_solver (device geometry, facet groups filtered by a
    centroid predicate - the piston demo's set-up) against the oracle loop on the same
    arrays built on the host."""
    import problems
    from fenicsx_fus_gpu_b200 import problem
    from fenicsx_fus_gpu_b200 import substrate as S
    from oracle import oracle as orc

    P, N, L, dtt, nsteps = 3, (4, 4, 3), (0.008, 0.008, 0.006), np.float64, 6
    su = problem.box_setup(P, N, L, dtt, perturb=0.08, seed=4)
    piston = problem.disc(0, 1, (0.004, 0.004), 0.0025)
    keep = lambda cen: ~(piston(cen) & (cen[:, 2] < 0.0005))  # noqa: E731
    s = problem.linear_solver(su, [0], [0, 1, 2, 3, 4, 5], source_predicate=piston, absorbing_predicate=keep)
    # device geometry == oracle geometry
    d = problems.linear_problem(P, N, L, dtt, perturb=0.08, seed=4)
    assert rel_l2(su.dev["G"].cpu().numpy(), d.G) < 1e-12
    assert rel_l2(su.dev["detJ"].cpu().numpy(), d.detJ) < 1e-13
    # host-side facet groups for the oracle
    tb = d.tb
    bd1 = S.boundary_facets(d.mesh, 0, piston)
    bd2 = np.concatenate([S.boundary_facets(d.mesh, f, keep) for f in range(6)])
    assert 0 < bd1.shape[0] < N[0] * N[1]
    assert bd2.shape[0] == 2 * (N[0] * N[1] + N[1] * N[2] + N[0] * N[2]) - bd1.shape[0]

    def grp(bd):
        dJ = np.zeros((bd.shape[0], tb.n**2))
        orc.compute_boundary_facets_scaled_jacobian_determinant(dJ, (d.mesh.x_dofs, d.mesh.x_g), bd, tb.dphi_f, tb.wts_f)
        return S.facet_dofmap(d.dofmap, bd, tb.local_facet_dof), dJ

    d.bfacet_dofmap1, d.detJ_f1 = grp(bd1)
    d.bfacet_dofmap2, d.detJ_f2 = grp(bd2)
    d.facet_coeff1 = np.full(bd1.shape[0], 1.0 / d.rho)
    d.facet_coeff2 = np.full(bd2.shape[0], -1.0 / d.rho / d.c0)
    dt = problem.cfl_time_step(P, su.h, d.c0, d.f0, 0.65)
    u_ref, v_ref = _oracle_linear(d, dt, nsteps, dtt)
    s.init()
    s.rk4(0.0, dt, nsteps)
    assert rel_l2(s.u.cpu().numpy(), u_ref) < 1e-12
    assert rel_l2(s.v.cpu().numpy(), v_ref) < 1e-12


def test_full_size_properties_demo_linear_box():
    """BASELINE.json configs[1] at full size (80^3 cells, degree 4, 33 M dofs, f64),
    where the CPU oracle would need minutes: size-independent properties.
    K 1 = 0, sum(M 1) = |Omega|, v.Ku = u.Kv, and the fused step conserves the
    state when there is no source (u = v = 0 stays 0)."""
    from fenicsx_fus_gpu_b200 import operators as ops
    from fenicsx_fus_gpu_b200 import problem

    P, N, L = 4, 80, 0.12
    su = problem.box_setup(P, N, L, np.float64)
    nd, nc = su.ndofs, su.mesh.num_cells
    assert nd == 33076161
    dm, G, dJ = su.dev["dofmap"], su.dev["G"], su.dev["detJ"]
    D = torch.from_numpy(su.tables.dphi_1D).cuda()
    c = torch.ones(nc, dtype=torch.float64, device="cuda")
    K = ops.stiffness_operator(P, np.float64)

    def apply(x):
        y = torch.zeros(nd, dtype=torch.float64, device="cuda")
        K[nc, (5, 5, 5)](x, c, y, G, dm, D)
        return y

    gen = torch.Generator(device="cuda").manual_seed(3)
    u = torch.randn(nd, dtype=torch.float64, device="cuda", generator=gen)
    v = torch.randn(nd, dtype=torch.float64, device="cuda", generator=gen)
    Ku, Kv = apply(u), apply(v)
    ones = torch.ones(nd, dtype=torch.float64, device="cuda")
    assert float(apply(ones).norm() / Ku.norm()) < 1e-12
    assert abs(float(v @ Ku - u @ Kv)) / abs(float(v @ Ku)) < 1e-10
    m = torch.zeros(nd, dtype=torch.float64, device="cuda")
    ops.mass_operator[1, 128](ones, c, m, dJ, dm)
    assert abs(float(m.sum()) - L**3) / L**3 < 1e-12
    assert float(m.min()) > 0
    s = problem.linear_solver(su, [2], [3], p0=0.0)
    s.init()
    s.rk4(0.0, 1e-8, 2)
    assert float(s.u.abs().max()) == 0.0 and float(s.v.abs().max()) == 0.0


def test_westervelt_rk4_partitioned_p2p_vs_serial_oracle():
    """Westervelt stage over 4 emulated ranks with the peer-memory halo (forward of
    (un, vn), reverse of (b, m) in one get_add) against the single-rank oracle."""
    import problems
    from fenicsx_fus_gpu_b200 import substrate as S
    from fenicsx_fus_gpu_b200 import utils
    from fenicsx_fus_gpu_b200.scatterer import LocalCluster, P2PHaloExchange, local_fabric
    from fenicsx_fus_gpu_b200.solver import WesterveltSpectral3D, westervelt_source
    from oracle import oracle as orc

    dtt, P, N, L, R, nsteps = np.float64, 3, (4, 4, 3), (0.006, 0.006, 0.0045), 4, 6
    d = problems.westervelt_problem(P, N, L, dtt, perturb=0.1, seed=9)
    dt = problems.cfl_dt(P, 0.0015, d.c0, d.f0, cfl=0.4)
    ones = np.ones(d.ndofs)
    m0 = np.zeros(d.ndofs)
    orc.mass_operator(ones, d.cell_coeff1, m0, d.detJ, d.dofmap)
    orc.mass_operator(ones, d.facet_coeff1_2, m0, d.detJ_f2, d.bfacet_dofmap2)
    prob = orc.WesterveltProblem(d.P, d.dofmap, d.G, d.detJ, d.tb.dphi_1D, d.cell_coeff2, d.cell_coeff3,
                                 d.cell_coeff4, d.cell_coeff5, m0, d.bfacet_dofmap1, d.detJ_f1,
                                 d.facet_coeff1_1, d.facet_coeff2_1, d.bfacet_dofmap2, d.detJ_f2,
                                 d.facet_coeff2_2, d.f0, d.p0, d.c0)
    u_ref, v_ref = np.zeros(d.ndofs), np.zeros(d.ndofs)
    orc.westervelt_rk4(prob, u_ref, v_ref, 0.0, dt, nsteps)

    parts = S.partition_box(N, P, R, lengths=L, dtype=dtt, perturb=0.1, seed=9)
    sdata = utils.compute_scatterer_data_all([p.index_map for p in parts])
    ndmax = max(q.index_map.size_local + q.index_map.num_ghosts for q in parts)

    def body(r, transport):
        p = parts[r]
        nd = p.index_map.size_local + p.index_map.num_ghosts
        w = problems.westervelt_problem(P, None, None, dtt, mesh=p.mesh, dofmap=p.dofmap, ndofs=nd)
        fab = local_fabric(transport.cluster, r, P2PHaloExchange.arena_bytes(ndmax, dtt))
        halo = P2PHaloExchange(fab, sdata[r][0], sdata[r][1], p.index_map.size_local, p.index_map.num_ghosts, dtt)
        s = WesterveltSpectral3D(
            P, dtt, nd, w.dofmap, w.G, w.detJ, w.tb.dphi_1D, w.cell_coeff1, w.cell_coeff2, w.cell_coeff3,
            w.cell_coeff4, w.cell_coeff5, w.bfacet_dofmap1, w.detJ_f1, w.facet_coeff1_1, w.facet_coeff2_1,
            w.bfacet_dofmap2, w.detJ_f2, w.facet_coeff1_2, w.facet_coeff2_2, halo=halo,
            source=lambda t: westervelt_source(t, w.f0, w.p0, w.c0), use_graph=False)
        s.init()
        s.rk4(0.0, dt, nsteps)
        torch.cuda.synchronize()
        return s.u.cpu().numpy(), s.v.cpu().numpy()

    out = LocalCluster(R).run(body)
    u, v = np.zeros_like(u_ref), np.zeros_like(v_ref)
    for r, p in enumerate(parts):
        nl = p.index_map.size_local
        u[p.local_to_serial[:nl]] = out[r][0][:nl]
        v[p.local_to_serial[:nl]] = out[r][1][:nl]
    assert np.linalg.norm(u_ref) > 0
    assert rel_l2(u, u_ref) < 1e-12
    assert rel_l2(v, v_ref) < 1e-12
