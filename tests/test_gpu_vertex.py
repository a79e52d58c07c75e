"""On-the-fly geometry (csrc/stiffness_vertex.cu): the stiffness action with G = w |det J| J^-1 J^-T
rebuilt in the kernel from 36 trilinear coefficients per cell instead of streamed
(cuda/operators.py:154-164 reads it; cuda/precompute.py:115-163 fills it).

* host side (no GPU): the bilinear expansion of the P1 derivative table and the per-point
  reconstruction, against the oracle's G and detJ tables;
* CUDA: the operator against the reference-generated fixtures (numba-cpu/operators.py outputs),
  against the streamed kernel on ragged, strongly perturbed meshes, and the dual form.
"""

import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from fenicsx_fus_gpu_b200 import precompute as pre  # noqa: E402
from fenicsx_fus_gpu_b200 import substrate as S  # noqa: E402

TOL = {"f64": 1e-12, "f32": 1e-5}


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _coeffs_host(tb, mesh):
    M = pre.trilinear_expansion(tb.dphi, tb.pts)
    return np.einsum("dmv,cvk->cdmk", M, mesh.x_g[mesh.x_dofs].astype(np.float64))  # (Nc, 3, 4, 3)


@pytest.mark.parametrize("P", range(2, 8))
def test_trilinear_reconstruction_vs_oracle_tables(P):
    """The arithmetic the kernel does per quadrature point, in numpy, against the oracle's
    restatement of cuda/precompute.py:76-163."""
    from oracle import oracle as orc

    tb = S.element_tables(P)
    n = P + 1
    mesh = S.create_box((2, 2, 2), (1.0, 0.8, 1.2), perturb=0.2, seed=P)
    Nc, Nd = mesh.num_cells, n**3
    G, detJ = np.zeros((Nc, Nd, 6)), np.zeros((Nc, Nd))
    orc.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    orc.compute_scaled_jacobian_determinant(detJ, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    Tc = _coeffs_host(tb, mesh)
    X, W = tb.pts_1d, tb.wts_1d
    I, J, K = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    xi, eta, zeta = X[I].ravel(), X[J].ravel(), X[K].ravel()  # q = (i n + j) n + k
    w = (W[I] * W[J] * W[K]).ravel()
    bil = lambda T, u, v: (T[:, None, 0] + u[None, :, None] * T[:, None, 1] + v[None, :, None] * T[:, None, 2]  # noqa: E731
                           + (u * v)[None, :, None] * T[:, None, 3])
    t0, t1, t2 = bil(Tc[:, 0], eta, zeta), bil(Tc[:, 1], xi, zeta), bil(Tc[:, 2], xi, eta)  # (Nc, nq, 3)
    c0, c1, c2 = np.cross(t1, t2), np.cross(t2, t0), np.cross(t0, t1)
    det = (t0 * c0).sum(-1)
    s = w[None, :] / np.abs(det)
    G2 = np.stack([(c0 * c0).sum(-1), (c0 * c1).sum(-1), (c0 * c2).sum(-1), (c1 * c1).sum(-1), (c1 * c2).sum(-1),
                   (c2 * c2).sum(-1)], axis=-1) * s[..., None]
    assert np.abs(G2 - G).max() < 1e-13 * np.abs(G).max()
    assert np.abs(w[None, :] * np.abs(det) - detJ).max() < 1e-13 * np.abs(detJ).max()


def test_trilinear_expansion_rejects_other_tables():
    from fenicsx_fus_gpu_b200._lib import FusError

    tb = S.element_tables(3)
    bad = tb.dphi.copy()
    bad[0] += (tb.pts[:, 1] ** 2)[:, None]  # quadratic in eta: not a trilinear geometry
    with pytest.raises(FusError):
        pre.trilinear_expansion(bad, tb.pts)
    with pytest.raises(FusError):
        pre.trilinear_expansion(tb.dphi[:, :, :4], tb.pts)


# --------------------------------------------------------------------------- #
# CUDA
# --------------------------------------------------------------------------- #
def _gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def d(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("P", range(2, 8))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_vertex_operator_vs_reference_golden(golden_dir, P, tag):
    torch = _gpu()
    from fenicsx_fus_gpu_b200 import operators as ops

    g = np.load(os.path.join(golden_dir, f"operators_P{P}_{tag}.npz"))
    dt = g["x"].dtype
    tb = S.element_tables(P, "basix", dt)
    assert np.array_equal(tb.dphi, g["dphi"])  # the fixture's tables are the substrate's
    n, nc = P + 1, g["dofmap"].shape[0]
    Tc = pre.trilinear_coefficients((g["x_dofs"], g["x_g"]), nc, g["dphi"], tb.pts)
    ref = np.einsum("dmv,cvk->cdmk", pre.trilinear_expansion(g["dphi"], tb.pts), g["x_g"][g["x_dofs"]].astype(np.float64))
    assert rel_l2(Tc.cpu().numpy().reshape(nc, 3, 4, 3), ref) < (1e-14 if tag == "f64" else 1e-6)
    x, coeff, dofmap = d(g["x"]), d(g["coeff"]), d(g["dofmap"])
    y = torch.zeros(g["x"].size, dtype=x.dtype, device="cuda")
    K = ops.stiffness_operator_vertex(P, dt)
    K[nc, (n, n, n)](x, coeff, y, Tc, tb.pts_1d, tb.wts_1d, dofmap, g["dphi_1D"])
    assert rel_l2(y.cpu().numpy(), g["y_stiff"]) < TOL[tag]
    K[nc, (n, n, n)](x, coeff, y, Tc, tb.pts_1d, tb.wts_1d, dofmap, g["dphi_1D"])  # accumulates
    assert rel_l2(0.5 * y.cpu().numpy(), g["y_stiff"]) < TOL[tag]


@pytest.mark.gpu
@pytest.mark.parametrize("P,N,tag", [(4, (7, 5, 3), "f64"), (4, (9, 7, 6), "f32"), (2, (11, 9, 5), "f64"), (5, (5, 4, 3), "f64"),
                                     (7, (3, 3, 2), "f64"), (3, (8, 3, 5), "f32"), (6, (4, 3, 3), "f32")])
def test_vertex_operator_vs_streamed_kernel(P, N, tag):
    """Ragged cell counts (the last batch of a CTA is partial), vertices jittered by a quarter of a
    cell, random per-cell constants; single and dual form; a second launch on a sub-range of cells."""
    torch = _gpu()
    from fenicsx_fus_gpu_b200 import operators as ops
    from fenicsx_fus_gpu_b200._lib import check, current_stream, fn

    dt = np.float64 if tag == "f64" else np.float32
    tb = S.element_tables(P, "basix", dt)
    n = P + 1
    mesh = S.create_box(N, (1.0, 0.7, 0.9), dtype=dt, perturb=0.25, seed=P)
    dofmap_h = S.tensor_dofmap(mesh, P)
    nc, nd = mesh.num_cells, S.num_dofs(N, P)
    tdt = torch.float64 if tag == "f64" else torch.float32
    G = torch.empty((nc, n**3, 6), dtype=tdt, device="cuda")
    detJ = torch.empty((nc, n**3), dtype=tdt, device="cuda")
    pre.compute_geometry(G, detJ, (d(mesh.x_dofs), d(mesh.x_g)), nc, d(tb.dphi), d(tb.wts))
    Tc = pre.trilinear_coefficients((mesh.x_dofs, mesh.x_g), nc, tb.dphi, tb.pts)
    rng = np.random.default_rng(P)
    x, x2 = d(rng.standard_normal(nd).astype(dt)), d(rng.standard_normal(nd).astype(dt))
    c, c2 = d(rng.uniform(0.5, 2.0, nc).astype(dt)), d(rng.uniform(-1.0, 1.0, nc).astype(dt))
    dm = d(dofmap_h)
    y_ref = torch.zeros(nd, dtype=tdt, device="cuda")
    ops.stiffness_operator(P, dt)[nc, (n, n, n)](x, c, y_ref, G, dm, d(tb.dphi_1D))
    y = torch.zeros(nd, dtype=tdt, device="cuda")
    K = ops.stiffness_operator_vertex(P, dt)
    K[nc, (n, n, n)](x, c, y, Tc, tb.pts_1d, tb.wts_1d, dm, tb.dphi_1D)
    tol = 1e-13 if tag == "f64" else 2e-6
    assert rel_l2(y.cpu().numpy(), y_ref.cpu().numpy()) < tol
    # dual form through the C ABI (tables resident from the call above): y += K(c; x) + K(c2; x2)
    ops.stiffness_operator(P, dt)[nc, (n, n, n)](x2, c2, y_ref, G, dm, d(tb.dphi_1D))
    y.zero_()
    check(fn("fus_stiffness2_vertex", dt)(x.data_ptr(), c.data_ptr(), x2.data_ptr(), c2.data_ptr(), y.data_ptr(),
                                          Tc.data_ptr(), dm.data_ptr(), None, nc, P, 1, current_stream()),
          "fus_stiffness2_vertex")
    assert rel_l2(y.cpu().numpy(), y_ref.cpu().numpy()) < tol
    # a sub-range of cells (pointers offset by whole cells, as the solvers' launch ranges do)
    c0, m = nc // 3, nc - nc // 3 - 1
    ya, yb = torch.zeros_like(y), torch.zeros_like(y)
    ops.stiffness_operator(P, dt)[m, (n, n, n)](x, c[c0:c0 + m], ya, G[c0:c0 + m], dm[c0:c0 + m], d(tb.dphi_1D))
    K[m, (n, n, n)](x, c[c0:c0 + m], yb, Tc[c0:c0 + m], tb.pts_1d, tb.wts_1d, dm[c0:c0 + m], tb.dphi_1D)
    assert rel_l2(yb.cpu().numpy(), ya.cpu().numpy()) < tol


@pytest.mark.gpu
def test_vertex_operator_argument_checks():
    torch = _gpu()
    from fenicsx_fus_gpu_b200 import operators as ops
    from fenicsx_fus_gpu_b200._lib import FusError

    P, dt = 3, np.float64
    tb = S.element_tables(P)
    mesh = S.create_box((2, 2, 2))
    dm = d(S.tensor_dofmap(mesh, P))
    nd = S.num_dofs(2, P)
    x, y, c = torch.zeros(nd, dtype=torch.float64, device="cuda"), torch.zeros(nd, dtype=torch.float64, device="cuda"), \
        torch.ones(8, dtype=torch.float64, device="cuda")
    Tc = pre.trilinear_coefficients((mesh.x_dofs, mesh.x_g), 8, tb.dphi, tb.pts)
    K = ops.stiffness_operator_vertex(P, dt)
    with pytest.raises(FusError):
        K[8, (4, 4, 4)](x, c, y, Tc[:4], tb.pts_1d, tb.wts_1d, dm, tb.dphi_1D)  # Tc too short
    with pytest.raises(FusError):
        K[8, (4, 4, 4)](x, c, y, Tc, tb.pts_1d[:2], tb.wts_1d, dm, tb.dphi_1D)  # wrong table size
    with pytest.raises(ValueError):
        ops.stiffness_operator_vertex(9, dt)
    with pytest.raises(FusError):
        pre.trilinear_coefficients((mesh.x_dofs, mesh.x_g.astype(np.float32)), 8, tb.dphi, tb.pts, np.float64)
