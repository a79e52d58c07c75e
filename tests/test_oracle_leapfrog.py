"""CPU: the oracle's leapfrog restatement (oracle.linear_leapfrog - the scheme the product's
LinearLeapfrog3D implements; the reference has RK4 only) converges to the oracle's RK4 solution
(numba-cpu/demo_linear_box.py:425-459 restated) with dt^2, absorbing boundary included."""

import numpy as np

import problems
from oracle import oracle as orc


def _prob(d):
    m = np.zeros(d.ndofs)
    orc.mass_operator(np.ones(d.ndofs), d.cell_coeff1, m, d.detJ, d.dofmap)
    return orc.LinearProblem(d.P, d.dofmap, d.G, d.tb.dphi_1D, d.cell_coeff2, m, d.bfacet_dofmap1, d.detJ_f1,
                             d.facet_coeff1, d.bfacet_dofmap2, d.detJ_f2, d.facet_coeff2, d.f0, d.p0, d.c0)


def test_oracle_leapfrog_second_order_against_oracle_rk4():
    P, N, L = 2, 4, 0.012
    d = problems.linear_problem(P, N, L, np.float64, perturb=0.1, seed=2)
    prob = _prob(d)
    dt0 = problems.cfl_dt(P, L / N, d.c0, d.f0, cfl=0.4)
    u_ref, v_ref = np.zeros(d.ndofs), np.zeros(d.ndofs)
    orc.linear_rk4(prob, u_ref, v_ref, 0.0, dt0 / 4, 48 * 4)
    assert np.linalg.norm(u_ref) > 0
    errs = []
    for k in (1, 2, 4):
        u, v = np.zeros(d.ndofs), np.zeros(d.ndofs)
        t = orc.linear_leapfrog(prob, u, v, 0.0, dt0 / k, 48 * k)
        assert abs(t - 48 * dt0) < 1e-14
        errs.append(orc.rel_l2(u, u_ref))
    assert errs[0] < 0.3 and errs[2] < 0.02
    assert 3.5 < errs[0] / errs[1] < 4.6 and 3.5 < errs[1] / errs[2] < 4.6, errs
