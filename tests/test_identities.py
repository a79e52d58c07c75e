"""CPU checks (oracle only, no GPU) of the algebraic identities the compressed / pointwise
device paths rest on, so that their parity with the reference algorithm is pinned on the CPU
side as well:

* affine cells:        G[c, q, :] = wq[q] * Gc[c, :],  detJ[c, q] = wq[q] * detJc[c]
* rectilinear cells:   K u = c * sum_d g_dd (w1 x w1) (x) K1 u,   K1 = D^T diag(w1) D
* lumped mass:         M(c; w) = w * M(c; 1)   (the pointwise Westervelt cell-mass pair)
"""

import numpy as np

from fenicsx_fus_gpu_b200 import substrate as S
from oracle import oracle as orc

SHEAR = np.array([[1.0, 0.15, -0.1], [0.05, 0.9, 0.2], [-0.12, 0.07, 1.1]])


def _geometry(mesh, tb):
    nc, nq = mesh.num_cells, tb.n**3
    G = np.zeros((nc, nq, 6))
    detJ = np.zeros((nc, nq))
    orc.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), nc, tb.dphi, tb.wts)
    orc.compute_scaled_jacobian_determinant(detJ, (mesh.x_dofs, mesh.x_g), nc, tb.dphi, tb.wts)
    return G, detJ


def test_affine_cells_factor_exactly():
    tb = S.element_tables(4)
    mesh = S.create_box((3, 2, 2), (1.0, 0.7, 0.9))
    mesh.x_g = np.ascontiguousarray(mesh.x_g @ SHEAR.T)
    G, detJ = _geometry(mesh, tb)
    Gn = G / tb.wts[None, :, None]
    Gc = Gn.mean(axis=1)
    assert np.abs(Gn - Gc[:, None, :]).max() <= 64 * np.finfo(float).eps * np.abs(Gc).max()
    dn = detJ / tb.wts[None, :]
    assert np.abs(dn - dn.mean(axis=1, keepdims=True)).max() <= 64 * np.finfo(float).eps * np.abs(dn).max()
    assert np.abs(Gc[:, [1, 2, 4]]).min() > 0  # sheared: genuinely non-diagonal
    # a perturbed vertex breaks it
    bad = S.create_box((2, 2, 2), 1.0, perturb=0.2, seed=1)
    Gb, _ = _geometry(bad, tb)
    Gbn = Gb / tb.wts[None, :, None]
    assert np.abs(Gbn - Gbn.mean(axis=1)[:, None, :]).max() > 1e-3 * np.abs(Gbn).max()


def test_rectilinear_stiffness_is_three_1d_products():
    for P in (2, 4, 5):
        tb = S.element_tables(P)
        n = tb.n
        mesh = S.create_box((3, 2, 2), (1.0, 0.7, 1.9))
        x = mesh.x_g.copy()
        for ax in range(3):
            x[:, ax] = x[:, ax] + 0.08 * np.sin(2.5 * x[:, ax])  # graded, still axis-aligned
        mesh.x_g = x
        dofmap = S.tensor_dofmap(mesh, P)
        nd, nc = int(dofmap.max()) + 1, mesh.num_cells
        G, _ = _geometry(mesh, tb)
        Gc = (G / tb.wts[None, :, None]).mean(axis=1)
        assert np.abs(Gc[:, [1, 2, 4]]).max() <= 1e-12 * np.abs(Gc[:, [0, 3, 5]]).max()
        rng = np.random.default_rng(P)
        u = rng.standard_normal(nd)
        coeff = rng.uniform(0.5, 2.0, nc)
        y_ref = np.zeros(nd)
        orc.stiffness_operator(P, u, coeff, y_ref, G, dofmap, tb.dphi_1D)
        # the separable form the rectilinear kernel evaluates
        D, w1 = tb.dphi_1D.astype(float), tb.wts_1d.astype(float)
        assert np.allclose(w1[:, None, None] * w1[None, :, None] * w1[None, None, :], tb.wts.reshape(n, n, n), rtol=1e-13)
        K1 = (D.T * w1[None, :]) @ D
        y = np.zeros(nd)
        ww = w1[:, None] * w1[None, :]
        for c in range(nc):
            ue = u[dofmap[c]].reshape(n, n, n)
            ye = (Gc[c, 0] * np.einsum("il,ljk->ijk", K1, ue) * ww[None, :, :]
                  + Gc[c, 3] * np.einsum("jl,ilk->ijk", K1, ue) * ww[:, None, :]
                  + Gc[c, 5] * np.einsum("kl,ijl->ijk", K1, ue) * ww[:, :, None])
            np.add.at(y, dofmap[c], coeff[c] * ye.ravel())
        assert np.linalg.norm(y - y_ref) / np.linalg.norm(y_ref) < 1e-13


def test_lumped_mass_is_pointwise():
    P = 3
    tb = S.element_tables(P)
    mesh = S.create_box((3, 3, 2), 1.0, perturb=0.15, seed=3)
    dofmap = S.tensor_dofmap(mesh, P)
    nd, nc = int(dofmap.max()) + 1, mesh.num_cells
    _, detJ = _geometry(mesh, tb)
    rng = np.random.default_rng(0)
    w = rng.standard_normal(nd)
    c = rng.uniform(0.5, 2.0, nc)
    direct, diag = np.zeros(nd), np.zeros(nd)
    orc.mass_operator(w, c, direct, detJ, dofmap)           # M(c; w), cell by cell as the reference does
    orc.mass_operator(np.ones(nd), c, diag, detJ, dofmap)   # M(c; 1), assembled once
    assert np.linalg.norm(direct - w * diag) / np.linalg.norm(direct) < 1e-14
    # and with w = vn^2 (the b += M(c5; vn^2) term)
    direct[:] = 0
    orc.mass_operator(w * w, c, direct, detJ, dofmap)
    assert np.linalg.norm(direct - w * w * diag) / np.linalg.norm(direct) < 1e-14
