"""GPU parity, second batch (round-1 review items):

* the four reference-named halo kernels ``pack_fwd / unpack_fwd / pack_rev / unpack_rev``
  (cuda/scatterer.py:18-101) and the ``scatter_forward / scatter_reverse`` factories
  (:104-277) against what numba-cpu/scatterer.py produced (tests/golden/scatter_r*.npz);
* ``fus_geometry_* / fus_facet_geometry_*`` against the G / detJ / detJ_f the reference's
  cuda/precompute.py produced, degree 2..7 x f32/f64;
* BASELINE.json's configs at their real degree x size against the CPU oracle: configs[1]
  (degree 4, 80^3 cells, 33 M dofs) one stiffness + one mass action + two RK4 steps; the piston
  (degree 5) and Westervelt (degree 4 and 6) loops at >= 1 M dofs; Westervelt float32 partitioned
  over emulated ranks with the peer-memory halo.

Tolerances are BASELINE.json's: rel-L2 <= 1e-12 (float64), <= 1e-5 (float32); moves are bit exact.
(the oracle is built with -ffast-math like numba's fastmath=True, so the float64 bound is against
a reference that may reassociate; observed differences are ~1e-16.)
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


def d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _fixture(golden_dir, name):
    with np.load(os.path.join(golden_dir, f"scatter_{name}.npz")) as z:
        return {k: z[k] for k in z.files}


def _lists(g, r, which):
    ranks = g[f"r{r}_{which}_ranks"]
    return [[g[f"r{r}_{which}_idx{i}"] for i in range(ranks.size)], g[f"r{r}_{which}_size"], ranks]


@pytest.mark.parametrize("name", ["r2", "r3", "r8", "u4", "u5"])
def test_pack_unpack_kernels_vs_reference_fixture(golden_dir, name):
    """The reference's scatter, kernel for kernel (cuda/scatterer.py:139-188, 226-277), with the MPI
    round replaced by handing the send buffer to the receiver: pack_fwd -> unpack_fwd must reproduce
    the forward fixture bit for bit, pack_rev -> unpack_rev the reverse one."""
    from fenicsx_fus_gpu_b200.scatterer import pack_fwd, pack_rev, unpack_fwd, unpack_rev

    g = _fixture(golden_dir, name)
    R = int(g["nranks"])
    own = [_lists(g, r, "owners") for r in range(R)]
    gho = [_lists(g, r, "ghosts") for r in range(R)]
    N = [int(g[f"r{r}_size_local"]) for r in range(R)]
    # forward: owner r packs the dofs rank q ghosts; q unpacks them into its ghost block
    fwd = [d(g[f"r{r}_vec"]) for r in range(R)]
    mail = {}
    for r in range(R):
        src = d(g[f"r{r}_vec"])
        for i, q in enumerate(gho[r][2]):
            ix = d(gho[r][0][i].astype(np.int64))
            buf = torch.empty(ix.numel(), dtype=torch.float64, device="cuda")
            pack_fwd[(ix.numel() + 127) // 128, 128](src, buf, ix)
            mail[(r, int(q))] = buf
    for q in range(R):
        for i, r in enumerate(own[q][2]):
            ix = d(own[q][0][i].astype(np.int64))
            unpack_fwd[(ix.numel() + 127) // 128, 128](mail[(int(r), q)], fwd[q], ix, N[q])
    for r in range(R):
        assert np.array_equal(fwd[r].cpu().numpy(), g[f"r{r}_fwd"])
    # reverse: ghost holder q packs its ghost block entries owned by r; r adds them in
    rev = [d(g[f"r{r}_vec"]) for r in range(R)]
    mail = {}
    for q in range(R):
        src = d(g[f"r{q}_vec"])
        for i, r in enumerate(own[q][2]):
            ix = d(own[q][0][i].astype(np.int64))
            buf = torch.empty(ix.numel(), dtype=torch.float64, device="cuda")
            pack_rev[(ix.numel() + 127) // 128, 128](src, buf, ix, N[q])
            mail[(q, int(r))] = buf
    for r in range(R):
        for i, q in enumerate(gho[r][2]):
            ix = d(gho[r][0][i].astype(np.int64))
            unpack_rev[(ix.numel() + 127) // 128, 128](mail[(int(q), r)], rev[r], ix)
    for r in range(R):
        assert rel_l2(rev[r].cpu().numpy(), g[f"r{r}_rev"]) < 1e-15
    # float32 instantiations move the same values
    x32 = d(g["r0_vec"].astype(np.float32))
    ix = d(gho[0][0][0].astype(np.int64))
    b32 = torch.empty(ix.numel(), dtype=torch.float32, device="cuda")
    pack_fwd[1, 128](x32, b32, ix)
    assert np.array_equal(b32.cpu().numpy(), g["r0_vec"].astype(np.float32)[gho[0][0][0]])


@pytest.mark.parametrize("name", ["r2", "r3", "r8", "u4", "u5"])
def test_scatter_factories_vs_reference_fixture(golden_dir, name):
    """scatter_forward / scatter_reverse as the demos call them (cuda/demo_linear_box.py:206-207):
    ``scatter(buffer)`` modifies the vector in place."""
    from fenicsx_fus_gpu_b200.scatterer import LocalCluster, scatter_forward, scatter_reverse

    g = _fixture(golden_dir, name)
    R = int(g["nranks"])

    def body(r, transport):
        N = int(g[f"r{r}_size_local"])
        fwd = scatter_forward(transport, _lists(g, r, "owners"), _lists(g, r, "ghosts"), N, np.float64)
        rev = scatter_reverse(transport, _lists(g, r, "owners"), _lists(g, r, "ghosts"), N, np.float64)
        f = d(g[f"r{r}_vec"])
        fwd(f)
        b = d(g[f"r{r}_vec"])
        rev(b)
        torch.cuda.synchronize()
        return f.cpu().numpy(), b.cpu().numpy()

    out = LocalCluster(R).run(body)
    for r in range(R):
        assert np.array_equal(out[r][0], g[f"r{r}_fwd"])
        assert rel_l2(out[r][1], g[f"r{r}_rev"]) < 1e-15


@pytest.mark.parametrize("name", ["r2", "r3", "r8", "u4", "u5"])
def test_p2p_halo_handle_vs_reference_fixture(golden_dir, name):
    """The fus_halo_* handle (epoch flags, put / get_add) on emulated ranks: forward twice in a
    row and reverse, against the numba-cpu fixture."""
    from fenicsx_fus_gpu_b200.scatterer import LocalCluster, P2PHaloExchange, local_fabric

    g = _fixture(golden_dir, name)
    R = int(g["nranks"])
    ndmax = max(g[f"r{r}_vec"].size for r in range(R))

    def body(r, transport):
        N = int(g[f"r{r}_size_local"])
        nd = g[f"r{r}_vec"].size
        fab = local_fabric(transport.cluster, r, P2PHaloExchange.arena_bytes(ndmax, np.float64, 4))
        halo = P2PHaloExchange(fab, _lists(g, r, "owners"), _lists(g, r, "ghosts"), N, nd - N, np.float64)
        a, b = halo.alloc(), halo.alloc()
        a.copy_(d(g[f"r{r}_vec"]))
        b.copy_(d(g[f"r{r}_vec"]))
        halo.forward(a)
        halo.forward(a)
        halo.reverse(b)
        torch.cuda.synchronize()
        halo.status()
        return a.cpu().numpy(), b.cpu().numpy(), halo.nshared

    out = LocalCluster(R).run(body)
    for r in range(R):
        assert np.array_equal(out[r][0], g[f"r{r}_fwd"])
        assert rel_l2(out[r][1], g[f"r{r}_rev"]) < 1e-15
        shared = np.unique(np.concatenate([np.zeros(0, np.int64)] + _lists(g, r, "ghosts")[0]))
        assert shared.size <= out[r][2] <= 2 * shared.size  # shared dofs + the rest of their 16-byte packs


@pytest.mark.parametrize("P", range(2, 8))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_device_geometry_vs_reference_golden(golden_dir, P, tag):
    """fus_geometry_* / fus_facet_geometry_* against the tables cuda/precompute.py:17-163 produced
    (host arrays in, host arrays out - the reference's calling convention - and device arrays)."""
    from fenicsx_fus_gpu_b200 import precompute as pre

    g = np.load(os.path.join(golden_dir, f"operators_P{P}_{tag}.npz"))
    tol = 1e-12 if tag == "f64" else 2e-4  # f32: LAPACK inverse (reference) vs adjugate on jittered cells
    mesh = (g["x_dofs"], g["x_g"])
    Nc = g["dofmap"].shape[0]
    detJ = np.zeros_like(g["detJ"])
    pre.compute_scaled_jacobian_determinant(detJ, mesh, Nc, g["dphi"], g["wts"])
    assert rel_l2(detJ, g["detJ"]) < tol
    G = np.zeros_like(g["G"])
    pre.compute_scaled_geometrical_factor(G, mesh, Nc, g["dphi"], g["wts"])
    assert rel_l2(G, g["G"]) < tol
    dJf = np.zeros_like(g["detJ_f"])
    pre.compute_boundary_facets_scaled_jacobian_determinant(dJf, mesh, g["bdata"], g["dphi_f"], g["wts_f"])
    assert rel_l2(dJf, g["detJ_f"]) < tol
    # device in / device out, both tables in one pass
    Gd, Jd = torch.zeros_like(d(g["G"])), torch.zeros_like(d(g["detJ"]))
    pre.compute_geometry(Gd, Jd, (d(g["x_dofs"]), d(g["x_g"])), Nc, d(g["dphi"]), d(g["wts"]))
    assert rel_l2(Gd.cpu().numpy(), g["G"]) < tol and rel_l2(Jd.cpu().numpy(), g["detJ"]) < tol
    assert rel_l2(Gd.cpu().numpy(), G) < (1e-14 if tag == "f64" else 1e-6)
    # a tensor of the wrong dtype is refused, not reinterpreted
    with pytest.raises(Exception):
        pre.compute_geometry(Gd, Jd, (d(g["x_dofs"]).to(torch.int64), d(g["x_g"])), Nc, d(g["dphi"]), d(g["wts"]))


def _oracle_linear(dd, dt, nsteps, dtt):
    from oracle import oracle as orc

    m = np.zeros(dd.ndofs, dtt)
    orc.mass_operator(np.ones(dd.ndofs, dtt), dd.cell_coeff1, m, dd.detJ, dd.dofmap)
    prob = orc.LinearProblem(dd.P, dd.dofmap, dd.G, dd.tb.dphi_1D, dd.cell_coeff2, m, dd.bfacet_dofmap1,
                             dd.detJ_f1, dd.facet_coeff1, dd.bfacet_dofmap2, dd.detJ_f2, dd.facet_coeff2,
                             dd.f0, dd.p0, dd.c0)
    u, v = np.zeros(dd.ndofs, dtt), np.zeros(dd.ndofs, dtt)
    orc.linear_rk4(prob, u, v, 0.0, dt, nsteps)
    return u, v


def test_full_size_demo_linear_box_vs_oracle():
    """BASELINE.json configs[1] at FULL size (degree 4, 80^3 cells, 33 076 161 dofs, float64):
    one stiffness action, one mass action and two fused RK4 steps against the CPU oracle
    (about a minute of single-core oracle time)."""
    import problems
    from fenicsx_fus_gpu_b200 import operators as ops
    from fenicsx_fus_gpu_b200.solver import LinearSpectral3D, linear_source
    from oracle import oracle as orc

    P, N, L = 4, 80, 0.12
    dd = problems.linear_problem(P, N, L, np.float64, perturb=0.0)
    nd, nc = dd.ndofs, dd.dofmap.shape[0]
    assert nd == 33076161 and nc == 512000
    rng = np.random.default_rng(1)
    x = rng.standard_normal(nd)
    y_ref = np.zeros(nd)
    orc.stiffness_operator(P, x, dd.cell_coeff2, y_ref, dd.G, dd.dofmap, dd.tb.dphi_1D)
    m_ref = np.zeros(nd)
    orc.mass_operator(x, dd.cell_coeff1, m_ref, dd.detJ, dd.dofmap)
    xd, Gd, dmd = d(x), d(dd.G), d(dd.dofmap)
    y = torch.zeros(nd, dtype=torch.float64, device="cuda")
    ops.stiffness_operator(P, np.float64)[nc, (5, 5, 5)](xd, d(dd.cell_coeff2), y, Gd, dmd, d(dd.tb.dphi_1D))
    assert rel_l2(y.cpu().numpy(), y_ref) < 1e-12
    y.zero_()
    dJd = d(dd.detJ)
    ops.mass_operator[1, 128](xd, d(dd.cell_coeff1), y, dJd, dmd)
    assert rel_l2(y.cpu().numpy(), m_ref) < 1e-12
    del y_ref, m_ref, x, xd
    dt = problems.cfl_dt(P, L / N, dd.c0, dd.f0)
    u_ref, v_ref = _oracle_linear(dd, dt, 2, np.float64)
    s = LinearSpectral3D(P, np.float64, nd, dmd, Gd, dJd, dd.tb.dphi_1D, dd.cell_coeff1, dd.cell_coeff2,
                         dd.bfacet_dofmap1, dd.detJ_f1, dd.facet_coeff1, dd.bfacet_dofmap2, dd.detJ_f2,
                         dd.facet_coeff2, source=lambda t: linear_source(t, dd.f0, dd.p0, dd.c0))
    s.init()
    s.rk4(0.0, dt, 2)
    assert np.linalg.norm(v_ref) > 0
    assert rel_l2(s.u.cpu().numpy(), u_ref) < 1e-12
    assert rel_l2(s.v.cpu().numpy(), v_ref) < 1e-12


def test_piston_degree5_1m_dofs_vs_oracle():
    """configs[2]'s discretisation (degree 5, piston disc on z=0, every other exterior facet
    absorbing) at 1.19 M dofs: the fused loop (graph replay) against the oracle loop."""
    import problems
    from fenicsx_fus_gpu_b200 import problem
    from fenicsx_fus_gpu_b200 import substrate as S
    from oracle import oracle as orc

    P, N, nsteps = 5, 21, 3
    h = 0.12 / 93
    L = h * N
    su = problem.box_setup(P, N, L, np.float64, perturb=0.05, seed=4)
    assert su.ndofs == (P * N + 1) ** 3 and su.ndofs > 1_000_000
    piston = problem.disc(0, 1, (0.5 * L, 0.5 * L), 0.3 * L)
    keep = lambda cen: ~(piston(cen) & (cen[:, 2] < 0.5 * h))  # noqa: E731
    s = problem.linear_solver(su, [0], [0, 1, 2, 3, 4, 5], source_predicate=piston, absorbing_predicate=keep)
    dd = problems.linear_problem(P, N, L, np.float64, perturb=0.05, seed=4)
    tb = dd.tb

    def grp(bd):
        dJ = np.zeros((bd.shape[0], tb.n**2))
        orc.compute_boundary_facets_scaled_jacobian_determinant(dJ, (dd.mesh.x_dofs, dd.mesh.x_g), bd, tb.dphi_f, tb.wts_f)
        return S.facet_dofmap(dd.dofmap, bd, tb.local_facet_dof), dJ

    bd1 = S.boundary_facets(dd.mesh, 0, piston)
    bd2 = np.concatenate([S.boundary_facets(dd.mesh, f, keep) for f in range(6)])
    dd.bfacet_dofmap1, dd.detJ_f1 = grp(bd1)
    dd.bfacet_dofmap2, dd.detJ_f2 = grp(bd2)
    dd.facet_coeff1 = np.full(bd1.shape[0], 1.0 / dd.rho)
    dd.facet_coeff2 = np.full(bd2.shape[0], -1.0 / dd.rho / dd.c0)
    dt = problem.cfl_time_step(P, h * 0.9, dd.c0, dd.f0, 0.65)
    u_ref, v_ref = _oracle_linear(dd, dt, nsteps, np.float64)
    s.init()
    s.rk4(0.0, dt, nsteps)
    assert np.linalg.norm(u_ref) > 0
    assert rel_l2(s.u.cpu().numpy(), u_ref) < 1e-12
    assert rel_l2(s.v.cpu().numpy(), v_ref) < 1e-12


@pytest.mark.parametrize("P,N", [(4, 26), (6, 17)])
def test_westervelt_1m_dofs_vs_oracle(P, N):
    """configs[3]'s discretisation (Westervelt, degree 4; and degree 6 as in
    cuda/demo_nonlinear_box.py:81) at > 1 M dofs, both mass forms, against the oracle restatement
    of cuda/demo_nonlinear_bowl.py:529-657."""
    import problems
    from fenicsx_fus_gpu_b200.solver import WesterveltSpectral3D, westervelt_source
    from oracle import oracle as orc

    dtt, nsteps = np.float64, 2
    L = 0.08 / 198 * N
    w = problems.westervelt_problem(P, N, L, dtt, perturb=0.05, seed=6)
    assert w.ndofs > 1_000_000
    dt = problems.cfl_dt(P, 0.9 * L / N, w.c0, w.f0, cfl=0.4)
    ones = np.ones(w.ndofs)
    m0 = np.zeros(w.ndofs)
    orc.mass_operator(ones, w.cell_coeff1, m0, w.detJ, w.dofmap)
    orc.mass_operator(ones, w.facet_coeff1_2, m0, w.detJ_f2, w.bfacet_dofmap2)
    prob = orc.WesterveltProblem(w.P, w.dofmap, w.G, w.detJ, w.tb.dphi_1D, w.cell_coeff2, w.cell_coeff3,
                                 w.cell_coeff4, w.cell_coeff5, m0, w.bfacet_dofmap1, w.detJ_f1,
                                 w.facet_coeff1_1, w.facet_coeff2_1, w.bfacet_dofmap2, w.detJ_f2,
                                 w.facet_coeff2_2, w.f0, w.p0, w.c0)
    u_ref, v_ref = np.zeros(w.ndofs), np.zeros(w.ndofs)
    orc.westervelt_rk4(prob, u_ref, v_ref, 0.0, dt, nsteps)
    assert np.linalg.norm(u_ref) > 0
    for mass_form in ("pointwise", "cells"):
        s = WesterveltSpectral3D(
            P, dtt, w.ndofs, w.dofmap, w.G, w.detJ, w.tb.dphi_1D, w.cell_coeff1, w.cell_coeff2, w.cell_coeff3,
            w.cell_coeff4, w.cell_coeff5, w.bfacet_dofmap1, w.detJ_f1, w.facet_coeff1_1, w.facet_coeff2_1,
            w.bfacet_dofmap2, w.detJ_f2, w.facet_coeff1_2, w.facet_coeff2_2,
            source=lambda t: westervelt_source(t, w.f0, w.p0, w.c0), mass_form=mass_form)
        s.init()
        s.rk4(0.0, dt, nsteps)
        assert rel_l2(s.u.cpu().numpy(), u_ref) < 1e-12, mass_form
        assert rel_l2(s.v.cpu().numpy(), v_ref) < 1e-12, mass_form


@pytest.mark.parametrize("mass_form", ["pointwise", "cells"])
def test_westervelt_f32_partitioned_p2p_vs_serial_oracle(mass_form):
    """Westervelt in float32 over 4 emulated ranks with the peer-memory halo (interior / interface
    split, masked close, fused close + put) against the single-rank float32 oracle: rel-L2 <= 1e-5."""
    import problems
    from fenicsx_fus_gpu_b200 import substrate as S
    from fenicsx_fus_gpu_b200 import utils
    from fenicsx_fus_gpu_b200.scatterer import LocalCluster, P2PHaloExchange, local_fabric
    from fenicsx_fus_gpu_b200.solver import WesterveltSpectral3D, westervelt_source
    from oracle import oracle as orc

    dtt, P, N, L, R, nsteps = np.float32, 4, (4, 4, 3), (0.006, 0.006, 0.0045), 4, 6
    w = problems.westervelt_problem(P, N, L, dtt, perturb=0.1, seed=9)
    dt = problems.cfl_dt(P, 0.0015, w.c0, w.f0, cfl=0.4)
    ones = np.ones(w.ndofs, dtt)
    m0 = np.zeros(w.ndofs, dtt)
    orc.mass_operator(ones, w.cell_coeff1, m0, w.detJ, w.dofmap)
    orc.mass_operator(ones, w.facet_coeff1_2, m0, w.detJ_f2, w.bfacet_dofmap2)
    prob = orc.WesterveltProblem(w.P, w.dofmap, w.G, w.detJ, w.tb.dphi_1D, w.cell_coeff2, w.cell_coeff3,
                                 w.cell_coeff4, w.cell_coeff5, m0, w.bfacet_dofmap1, w.detJ_f1,
                                 w.facet_coeff1_1, w.facet_coeff2_1, w.bfacet_dofmap2, w.detJ_f2,
                                 w.facet_coeff2_2, w.f0, w.p0, w.c0)
    u_ref, v_ref = np.zeros(w.ndofs, dtt), np.zeros(w.ndofs, dtt)
    orc.westervelt_rk4(prob, u_ref, v_ref, 0.0, dt, nsteps)
    parts = S.partition_box(N, P, R, lengths=L, dtype=dtt, perturb=0.1, seed=9)
    sdata = utils.compute_scatterer_data_all([p.index_map for p in parts])
    ndmax = max(q.index_map.size_local + q.index_map.num_ghosts for q in parts)

    def body(r, transport):
        p = parts[r]
        nd = p.index_map.size_local + p.index_map.num_ghosts
        q = problems.westervelt_problem(P, None, None, dtt, mesh=p.mesh, dofmap=p.dofmap, ndofs=nd)
        fab = local_fabric(transport.cluster, r, P2PHaloExchange.arena_bytes(ndmax, dtt, 14))
        halo = P2PHaloExchange(fab, sdata[r][0], sdata[r][1], p.index_map.size_local, p.index_map.num_ghosts, dtt)
        s = WesterveltSpectral3D(
            P, dtt, nd, q.dofmap, q.G, q.detJ, q.tb.dphi_1D, q.cell_coeff1, q.cell_coeff2, q.cell_coeff3,
            q.cell_coeff4, q.cell_coeff5, q.bfacet_dofmap1, q.detJ_f1, q.facet_coeff1_1, q.facet_coeff2_1,
            q.bfacet_dofmap2, q.detJ_f2, q.facet_coeff1_2, q.facet_coeff2_2, halo=halo,
            source=lambda t: westervelt_source(t, q.f0, q.p0, q.c0), use_graph=False, mass_form=mass_form)
        assert r == 0 or (s.ninterface > 0 and s._segs[-1].ninterior < s._segs[-1].n)  # rank 0 owns all it touches
        s.init()
        s.rk4(0.0, dt, nsteps)
        torch.cuda.synchronize()
        halo.status()
        return s.u.cpu().numpy(), s.v.cpu().numpy()

    out = LocalCluster(R).run(body)
    u, v = np.zeros_like(u_ref), np.zeros_like(v_ref)
    for r, p in enumerate(parts):
        nl = p.index_map.size_local
        u[p.local_to_serial[:nl]] = out[r][0][:nl]
        v[p.local_to_serial[:nl]] = out[r][1][:nl]
    assert np.linalg.norm(u_ref) > 0
    assert rel_l2(u, u_ref) < 1e-5
    assert rel_l2(v, v_ref) < 1e-5


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("order", ["mesh", "shuffled"])
def test_host_buffer_stiffness_pipeline_vs_oracle(tag, order):
    """fus_stiffness_host_*: x and y in pinned HOST memory, chunked upload / action / download on
    three streams.  Accumulate and FUS_HOST_Y_ZERO modes, on a mesh whose cell order follows the
    dof order (real pipeline) and on a shuffled one (pieces collapse: upload -> action -> download)."""
    from fenicsx_fus_gpu_b200 import _lib
    from fenicsx_fus_gpu_b200 import substrate as S
    from oracle import oracle as orc

    dt = np.float64 if tag == "f64" else np.float32
    tol = 1e-12 if tag == "f64" else 1e-5
    P, N = 4, (12, 11, 10)
    tb = S.element_tables(P, "basix", dt)
    mesh = S.create_box(N, (1.0, 0.9, 0.8), dtype=dt, perturb=0.15, seed=3)
    dofmap = S.tensor_dofmap(mesh, P)
    nd, nc = int(dofmap.max()) + 1, dofmap.shape[0]
    G = np.zeros((nc, tb.n**3, 6), dt)
    orc.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), nc, tb.dphi, tb.wts)
    rng = np.random.default_rng(2)
    coeff = rng.uniform(0.5, 2.0, nc).astype(dt)
    if order == "shuffled":
        perm = rng.permutation(nc)
        dofmap, G, coeff = np.ascontiguousarray(dofmap[perm]), np.ascontiguousarray(G[perm]), coeff[perm]
    x = rng.standard_normal(nd).astype(dt)
    y0 = rng.standard_normal(nd).astype(dt)
    y_ref = np.zeros(nd, dt)
    orc.stiffness_operator(P, x, coeff, y_ref, G, dofmap, tb.dphi_1D)
    xh = torch.from_numpy(x).pin_memory()
    Gd, dmd, cd, Dd = d(G), d(dofmap), d(coeff), d(tb.dphi_1D)
    xd = torch.empty(nd, dtype=Gd.dtype, device="cuda")
    yd = torch.empty(nd, dtype=Gd.dtype, device="cuda")
    fh = _lib.fn("fus_stiffness_host", dt)
    st = torch.cuda.current_stream().cuda_stream
    for flags, start in ((0, y0), (_lib.FUS_HOST_Y_ZERO, np.zeros(nd, dt))):
        yh = torch.from_numpy(start.copy()).pin_memory()
        for rep in range(2):  # the second call reuses the cached streams / events
            yh.copy_(torch.from_numpy(start))
            rc = fh(xh.data_ptr(), yh.data_ptr(), nd, xd.data_ptr(), yd.data_ptr(), cd.data_ptr(), Gd.data_ptr(),
                    dmd.data_ptr(), Dd.data_ptr(), nc, P, flags, st)
            assert rc == 0, _lib.lib().fus_last_error()
            assert rel_l2(yh.numpy() - start, y_ref) < tol, (flags, rep)
    # pageable host memory works too (copies degrade to staged ones)
    yp = np.zeros(nd, dt)
    rc = fh(x.ctypes.data, yp.ctypes.data, nd, xd.data_ptr(), yd.data_ptr(), cd.data_ptr(), Gd.data_ptr(),
            dmd.data_ptr(), Dd.data_ptr(), nc, P, _lib.FUS_HOST_Y_ZERO, st)
    assert rc == 0 and rel_l2(yp, y_ref) < tol
    # a dofmap entry beyond nd is refused
    rc = fh(xh.data_ptr(), yh.data_ptr(), nd - 5, xd.data_ptr(), yd.data_ptr(), cd.data_ptr(), Gd.data_ptr(),
            dmd.data_ptr(), Dd.data_ptr(), nc, P, 0, st)
    assert rc == _lib.lib().fus_abi_version() * 0 + 100002
