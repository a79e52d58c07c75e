"""The DOLFINx bridge (dolfinx_bridge.setup_from_dolfinx) on stand-in DOLFINx objects: the
extraction of cuda/demo_linear_box.py:99-113, 175-180, 232-270 through duck typing must give
the same arrays - and the same RK4 result - as the box problem builder."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_bridge_matches_box_builder_and_oracle():
    import problems
    import test_gpu_solver as tgs
    from fenicsx_fus_gpu_b200 import dolfinx_bridge as br, problem, substrate as S, utils

    P, N, L = 3, 4, 0.01
    tb = S.element_tables(P)
    mesh = S.create_box(N, L, perturb=0.1, seed=4)
    dofmap = S.tensor_dofmap(mesh, P)
    perm = np.random.default_rng(0).permutation(tb.n**3)  # stands in for argsort(tp.dof_ordering)
    M, V, facet_id = problems.fake_dolfinx(mesh, dofmap, perm)

    su = br.setup_from_dolfinx(M, V, P, np.float64, tables=tb, perm=perm)
    ref = problem.box_setup(P, N, L, np.float64, perturb=0.1, seed=4)
    assert np.array_equal(su.dofmap, ref.dofmap) and su.ndofs == ref.ndofs == su.nlocal
    assert torch.equal(su.dev["G"], ref.dev["G"]) and torch.equal(su.dev["detJ"], ref.dev["detJ"])
    assert su.h == pytest.approx(np.sqrt(3) * L / N, rel=0.5)

    # facets: DOLFINx facet ids -> (cell, local facet) rows -> solver
    bd_src = S.boundary_facets(mesh, 2)
    bd_abs = S.boundary_facets(mesh, 3)
    f_src = np.array([facet_id(c, lf) for c, lf in bd_src], np.int32)
    f_abs = np.array([facet_id(c, lf) for c, lf in bd_abs], np.int32)
    got_src, got_abs = utils.facet_integration_domain(f_src, M), utils.facet_integration_domain(f_abs, M)
    assert np.array_equal(got_src, bd_src) and np.array_equal(got_abs, bd_abs)
    s = problem.linear_solver(su, got_src, got_abs)
    d = problems.linear_problem(P, N, L, np.float64, perturb=0.1, seed=4)
    dt = problems.cfl_dt(P, L / N, d.c0, d.f0)
    u_ref, v_ref = tgs._oracle_linear(d, dt, 10, np.float64)
    s.init()
    s.rk4(0.0, dt, 10)
    assert tgs.rel_l2(s.u.cpu().numpy(), u_ref) < 1e-12
    assert tgs.rel_l2(s.v.cpu().numpy(), v_ref) < 1e-12
    # sampling works on the bridged mesh object too
    from fenicsx_fus_gpu_b200 import sampling as sp

    xp, cells = sp.compute_eval_params(su.mesh, np.array([[0.5 * L], [0.4 * L], [0.3 * L]]), np.float64)
    assert len(cells) == 1
    ev = sp.PointEvaluator(P, np.float64, su.dev["dofmap"], su.mesh, xp, cells, tb.pts_1d)
    assert np.isfinite(ev.to_host(s.u)).all()
