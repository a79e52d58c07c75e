"""The leapfrog integrator (north star: "RK4/leapfrog"; the reference has RK4 only, so there is no
reference twin): the CUDA step against the oracle's restatement of the same scheme (rel-L2 <= 1e-12
in float64, <= 1e-5 in float32), its convergence order against the RK4 solution of the same
problem, and the partitioned run (peer-memory halo, emulated ranks) against the serial oracle."""

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


def _solver(cls, d, dtt, halo=None, **kw):
    from fenicsx_fus_gpu_b200.solver import linear_source

    return cls(d.P, dtt, d.ndofs, d.dofmap, d.G, d.detJ, d.tb.dphi_1D, d.cell_coeff1, d.cell_coeff2,
               d.bfacet_dofmap1, d.detJ_f1, d.facet_coeff1, d.bfacet_dofmap2, d.detJ_f2, d.facet_coeff2,
               halo=halo, source=lambda t: linear_source(t, d.f0, d.p0, d.c0), **kw)


def _oracle(d, dt, nsteps, dtt):
    from oracle import oracle as orc

    m = np.zeros(d.ndofs, dtt)
    orc.mass_operator(np.ones(d.ndofs, dtt), d.cell_coeff1, m, d.detJ, d.dofmap)
    prob = orc.LinearProblem(d.P, d.dofmap, d.G, d.tb.dphi_1D, d.cell_coeff2, m, d.bfacet_dofmap1, d.detJ_f1,
                             d.facet_coeff1, d.bfacet_dofmap2, d.detJ_f2, d.facet_coeff2, d.f0, d.p0, d.c0)
    u, v = np.zeros(d.ndofs, dtt), np.zeros(d.ndofs, dtt)
    orc.linear_leapfrog(prob, u, v, 0.0, dt, nsteps)
    return u, v


@pytest.mark.parametrize("P,N,tag", [(4, 5, "f64"), (3, 6, "f32"), (6, 3, "f64")])
def test_leapfrog_vs_oracle(P, N, tag):
    import problems
    from fenicsx_fus_gpu_b200.solver import LinearLeapfrog3D

    dtt = np.float64 if tag == "f64" else np.float32
    L, nsteps = 0.01, 24
    d = problems.linear_problem(P, N, L, dtt, perturb=0.12, seed=P, abs_facets=(3, 5))
    dt = problems.cfl_dt(P, L / N, d.c0, d.f0, cfl=0.3)
    u_ref, v_ref = _oracle(d, dt, nsteps, dtt)
    assert np.linalg.norm(u_ref) > 0
    tol = 1e-12 if tag == "f64" else 1e-5
    for use_graph in (True, False):
        s = _solver(LinearLeapfrog3D, d, dtt, use_graph=use_graph)
        s.init()
        t = s.rk4(0.0, dt, nsteps // 2)
        t = s.steps(t, dt, nsteps - nsteps // 2)  # a second call continues (no second kick-off)
        assert abs(t - nsteps * dt) < 1e-15
        assert rel_l2(s.u.cpu().numpy(), u_ref) < tol, use_graph
        assert rel_l2(s.v.cpu().numpy(), v_ref) < tol, use_graph


def test_leapfrog_is_second_order_and_agrees_with_rk4():
    """Same problem, same final time: the leapfrog solution converges to the RK4 one with dt^2."""
    import problems
    from fenicsx_fus_gpu_b200.solver import LinearLeapfrog3D, LinearSpectral3D

    P, N, L, dtt = 3, 6, 0.012, np.float64
    d = problems.linear_problem(P, N, L, dtt, perturb=0.1, seed=2)
    dt0 = problems.cfl_dt(P, L / N, d.c0, d.f0, cfl=0.4)
    T = 64 * dt0
    ref = _solver(LinearSpectral3D, d, dtt)
    ref.init()
    ref.rk4(0.0, dt0 / 4, 256)  # RK4 at dt0/4: time error ~ (1/4)^4 of an already small one
    u_ref = ref.u.cpu().numpy()
    assert np.linalg.norm(u_ref) > 0
    errs = []
    for k in (1, 2, 4):
        s = _solver(LinearLeapfrog3D, d, dtt)
        s.init()
        s.rk4(0.0, dt0 / k, 64 * k)
        assert abs(s.t - T) < 1e-14
        errs.append(rel_l2(s.u.cpu().numpy(), u_ref))
    assert errs[0] < 0.3 and errs[2] < 0.02
    assert 3.5 < errs[0] / errs[1] < 4.6 and 3.5 < errs[1] / errs[2] < 4.6, errs


def test_leapfrog_partitioned_p2p_vs_serial_oracle():
    import problems
    from fenicsx_fus_gpu_b200 import substrate as S
    from fenicsx_fus_gpu_b200 import utils
    from fenicsx_fus_gpu_b200.scatterer import HaloExchange, LocalCluster, P2PHaloExchange, local_fabric
    from fenicsx_fus_gpu_b200.solver import LinearLeapfrog3D

    dtt, P, N, L, R, nsteps = np.float64, 3, (4, 4, 3), (0.012, 0.01, 0.011), 4, 16
    serial = problems.linear_problem(P, N, L, dtt, perturb=0.1, seed=11, abs_facets=(3, 4))
    dt = problems.cfl_dt(P, min(L[i] / N[i] for i in range(3)), serial.c0, serial.f0, cfl=0.3)
    u_ref, v_ref = _oracle(serial, dt, nsteps, dtt)
    parts = S.partition_box(N, P, R, lengths=L, dtype=dtt, perturb=0.1, seed=11)
    sdata = utils.compute_scatterer_data_all([p.index_map for p in parts])
    ndmax = max(q.index_map.size_local + q.index_map.num_ghosts for q in parts)
    for kind in ("p2p", "nccl-shaped"):
        def body(r, transport):
            p = parts[r]
            nd = p.index_map.size_local + p.index_map.num_ghosts
            d = problems.linear_problem(P, None, None, dtt, mesh=p.mesh, dofmap=p.dofmap, ndofs=nd, abs_facets=(3, 4))
            if kind == "p2p":
                fab = local_fabric(transport.cluster, r, P2PHaloExchange.arena_bytes(ndmax, dtt))
                halo = P2PHaloExchange(fab, sdata[r][0], sdata[r][1], p.index_map.size_local, p.index_map.num_ghosts, dtt)
            else:
                halo = HaloExchange(transport, sdata[r][0], sdata[r][1], p.index_map.size_local, dtt)
            s = _solver(LinearLeapfrog3D, d, dtt, halo=halo, use_graph=False)
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
        assert rel_l2(u, u_ref) < 1e-12, kind
        assert rel_l2(v, v_ref) < 1e-12, kind
