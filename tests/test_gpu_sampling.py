"""GPU parity of the device-side field sampling (csrc/sampling.cu, sampling.PointEvaluator)
- the replacement of `copy_to_host` + `Function.eval` in cuda/demo_linear_piston.py:564-570 -
against the oracle interpolant and polynomial known answers."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("P,tag,perturb", [(2, "f64", 0.0), (4, "f64", 0.15), (4, "f32", 0.15), (5, "f64", 0.0),
                                           (7, "f32", 0.1)])
def test_point_evaluator_vs_oracle(P, tag, perturb):
    from fenicsx_fus_gpu_b200 import sampling as sp, substrate as S
    from oracle import oracle as orc

    dt = np.float64 if tag == "f64" else np.float32
    tb = S.element_tables(P, "basix", dt)
    mesh = S.create_box((4, 3, 5), (1.0, 0.8, 1.2), dtype=dt, perturb=perturb, seed=P)
    dm = S.tensor_dofmap(mesh, P)
    nd = int(dm.max()) + 1
    rng = np.random.default_rng(P)
    pts = rng.uniform([-0.05, -0.05, -0.05], [1.05, 0.85, 1.25], (777, 3))
    xp, cells = sp.compute_eval_params(mesh, pts.T, dt)
    assert len(cells) > 500
    u = rng.standard_normal(nd).astype(dt)
    ev = sp.PointEvaluator(P, dt, torch.from_numpy(dm).cuda(), mesh, xp, cells, tb.pts_1d)
    ref = orc.eval_points(u, dm, cells, ev.phi.cpu().numpy())
    ud = torch.from_numpy(u).cuda()
    got = ev(ud).cpu().numpy()
    tol = 1e-13 if tag == "f64" else 1e-5
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < tol
    assert np.array_equal(ev.to_host(ud), got)  # same launch, bit-reproducible reduction order
    if perturb == 0.0:  # polynomial known answer on the affine mesh
        xd = S.dof_coordinates(mesh, dm, tb)
        f = lambda x: 2 - x[:, 0] ** 2 + x[:, 1] * x[:, 2] ** 2 + x[:, 0] * x[:, 1]  # noqa: E731
        val = ev(torch.from_numpy(f(xd).astype(dt)).cuda()).cpu().numpy()
        assert np.abs(val - f(xp.astype(np.float64))).max() < 1e-12
    with pytest.raises(Exception):
        ev(u)  # host array: no CPU path


def test_point_evaluator_empty_and_linear_solver_plane():
    """No local points -> empty result; and the demo use: sample the pressure on a plane
    while the solver runs, without copying the whole vector."""
    import problems
    import test_gpu_solver as tgs
    from fenicsx_fus_gpu_b200 import sampling as sp
    from oracle import oracle as orc

    P, N, L = 3, 4, 0.01
    d = problems.linear_problem(P, N, L, np.float64, perturb=0.1, seed=1)
    none = sp.PointEvaluator(P, np.float64, d.dofmap, d.mesh, np.zeros((0, 3)), [], d.tb.pts_1d)
    assert none(torch.zeros(d.ndofs, dtype=torch.float64, device="cuda")).shape == (0,)
    g = np.linspace(0.1 * L, 0.9 * L, 9)
    X, Z = np.meshgrid(g, g)
    pts = np.stack([X.ravel(), np.full(X.size, 0.5 * L), Z.ravel()])
    xp, cells = sp.compute_eval_params(d.mesh, pts, np.float64)
    assert len(cells) == 81
    ev = sp.PointEvaluator(P, np.float64, d.dofmap, d.mesh, xp, cells, d.tb.pts_1d)
    s = tgs._linear_solver(d, np.float64)
    s.init()
    s.rk4(0.0, problems.cfl_dt(P, L / N, d.c0, d.f0), 15)
    vals = ev.to_host(s.u)
    ref = orc.eval_points(s.u.cpu().numpy(), d.dofmap, cells, ev.phi.cpu().numpy())
    assert np.linalg.norm(ref) > 0 and np.linalg.norm(vals - ref) / np.linalg.norm(ref) < 1e-13
