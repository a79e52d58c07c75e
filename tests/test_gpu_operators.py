"""GPU parity: the sm_100a operators (through the C ABI / the reference's
Python operator surface) against the reference-generated golden fixtures and
against the CPU oracle on seeded inputs.

Tolerances are BASELINE.json's: rel-L2 <= 1e-12 (float64), <= 1e-5 (float32).
"""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}


def _skip_no_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(scope="module")
def ops():
    _skip_no_gpu()
    from fenicsx_fus_gpu_b200 import operators

    return operators


def d(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("P", range(2, 8))
@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_operators_vs_reference_golden(ops, golden_dir, P, tag):
    g = np.load(os.path.join(golden_dir, f"operators_P{P}_{tag}.npz"))
    dt = g["x"].dtype
    nd = g["x"].size
    n = P + 1
    x, coeff, dofmap = d(g["x"]), d(g["coeff"]), d(g["dofmap"])

    y = torch.zeros(nd, dtype=x.dtype, device="cuda")
    ops.mass_operator[(dofmap.numel() + 127) // 128, 128](x, coeff, y, d(g["detJ"]), dofmap)
    assert rel_l2(y.cpu().numpy(), g["y_mass"]) < TOL[tag]

    y.zero_()
    stiff = ops.stiffness_operator(P, dt)
    stiff[dofmap.shape[0], (n, n, n)](x, coeff, y, d(g["G"]), dofmap, d(g["dphi_1D"]))
    assert rel_l2(y.cpu().numpy(), g["y_stiff"]) < TOL[tag]

    # accumulate semantics: a second application doubles the result
    stiff[dofmap.shape[0], (n, n, n)](x, coeff, y, d(g["G"]), dofmap, d(g["dphi_1D"]))
    assert rel_l2(0.5 * y.cpu().numpy(), g["y_stiff"]) < TOL[tag]

    y.zero_()
    bd = d(g["bdofmap"])
    ops.mass_operator[(bd.numel() + 127) // 128, 128](x, d(g["fcoeff"]), y, d(g["detJ_f"]), bd)
    assert rel_l2(y.cpu().numpy(), g["y_fmass"]) < TOL[tag]


@pytest.mark.parametrize("P,N,tag", [(2, 9, "f64"), (3, 7, "f64"), (4, 11, "f64"), (4, 11, "f32"),
                                     (5, 6, "f64"), (6, 5, "f32"), (7, 4, "f64"), (7, 4, "f32")])
def test_stiffness_vs_oracle_many_batches(ops, P, N, tag):
    """Enough cells for several persistent-CTA iterations, a ragged last batch
    and (odd cell counts) TMA batches that start at unaligned addresses."""
    from fenicsx_fus_gpu_b200 import substrate as S
    from oracle import oracle as orc

    dt = np.float64 if tag == "f64" else np.float32
    tb = S.element_tables(P, "basix", dt)
    mesh = S.create_box((N, N + 1, N + 2), (1.0, 1.1, 1.3), dtype=dt, perturb=0.2, seed=P)
    dofmap = S.tensor_dofmap(mesh, P)
    nd = int(dofmap.max()) + 1
    Nc = dofmap.shape[0]
    rng = np.random.default_rng(P * 17 + N)
    x = rng.standard_normal(nd).astype(dt)
    coeff = rng.uniform(0.5, 2.0, Nc).astype(dt)
    G = np.zeros((Nc, tb.n**3, 6), dt)
    orc.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    y_ref = np.zeros(nd, dt)
    orc.stiffness_operator(P, x, coeff, y_ref, G, dofmap, tb.dphi_1D)

    y = torch.zeros(nd, dtype=d(x).dtype, device="cuda")
    n = P + 1
    ops.stiffness_operator(P, dt)[Nc, (n, n, n)](d(x), d(coeff), y, d(G), d(dofmap), d(tb.dphi_1D))
    assert rel_l2(y.cpu().numpy(), y_ref) < TOL[tag]

    # a sub-range of cells whose G / dofmap slices start at an odd cell (base
    # pointers not 16-byte aligned for odd n in f32 -> non-TMA path)
    c0 = 3
    y2 = torch.zeros(nd, dtype=y.dtype, device="cuda")
    Gd, dmd, cd = d(G), d(dofmap), d(coeff)
    ops.stiffness_operator(P, dt)[Nc - c0, (n, n, n)](d(x), cd[c0:], y2, Gd[c0:], dmd[c0:], tb.dphi_1D)
    y_ref2 = np.zeros(nd, dt)
    orc.stiffness_operator(P, x, coeff[c0:].copy(), y_ref2, G[c0:].copy(), dofmap[c0:].copy(), tb.dphi_1D)
    assert rel_l2(y2.cpu().numpy(), y_ref2) < TOL[tag]


def test_stiffness_properties_large(ops):
    """Size-independent checks at a size the CPU oracle would not finish in
    seconds: K(const) = 0, symmetry v.Ku = u.Kv, linearity."""
    from fenicsx_fus_gpu_b200 import precompute as pre
    from fenicsx_fus_gpu_b200 import substrate as S

    P, N, dt = 4, 40, np.float64
    tb = S.element_tables(P, "basix", dt)
    mesh = S.create_box(N, 1.0, dtype=dt, perturb=0.15, seed=1)
    dofmap = S.tensor_dofmap(mesh, P)
    nd = int(dofmap.max()) + 1
    Nc = dofmap.shape[0]
    G = torch.zeros((Nc, tb.n**3, 6), dtype=torch.float64, device="cuda")
    pre.compute_scaled_geometrical_factor(G, (d(mesh.x_dofs), d(mesh.x_g)), Nc, d(tb.dphi), d(tb.wts))
    dm, D = d(dofmap), d(tb.dphi_1D)
    c = torch.ones(Nc, dtype=torch.float64, device="cuda")
    K = ops.stiffness_operator(P, dt)

    def apply(v):
        y = torch.zeros(nd, dtype=torch.float64, device="cuda")
        K[Nc, (5, 5, 5)](v, c, y, G, dm, D)
        return y

    gen = torch.Generator(device="cuda").manual_seed(0)
    u = torch.randn(nd, dtype=torch.float64, device="cuda", generator=gen)
    v = torch.randn(nd, dtype=torch.float64, device="cuda", generator=gen)
    Ku, Kv = apply(u), apply(v)
    ones = torch.ones(nd, dtype=torch.float64, device="cuda")
    assert float(apply(ones).norm() / Ku.norm()) < 1e-12
    assert abs(float(v @ Ku - u @ Kv)) / abs(float(v @ Ku)) < 1e-11
    Kuv = apply(2.0 * u - 3.0 * v)
    assert float((Kuv - (2.0 * Ku - 3.0 * Kv)).norm() / Kuv.norm()) < 1e-12
    assert float(u @ Ku) > 0.0


def test_vector_ops_vs_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "vector_ops.npz"))
    a, b = d(g["a"]), d(g["b"])
    y = d(g["y"])
    ops.axpy[1, 1024](float(g["alpha"]), a, y)
    assert rel_l2(y.cpu().numpy(), g["y_axpy"]) < 1e-15
    c = torch.zeros_like(a)
    ops.pointwise_divide[1, 1024](a, b, c)
    assert rel_l2(c.cpu().numpy(), g["c_div"]) < 1e-15
    ops.copy[1, 1024](a, c)
    assert np.array_equal(c.cpu().numpy(), g["b_copy"])
    ops.fill[1, 1024](2.5, c)
    assert np.array_equal(c.cpu().numpy(), g["f_fill"])
    ops.square[1, 1024](a, c)
    assert np.array_equal(c.cpu().numpy(), g["a"] * g["a"])


@pytest.mark.parametrize("n", [1, 2, 3, 5, 1023, 100003])
@pytest.mark.parametrize("tdt", ["float64", "float32"])
def test_vector_ops_ragged_and_unaligned(ops, n, tdt):
    tdt = getattr(torch, tdt)
    gen = torch.Generator(device="cuda").manual_seed(n)
    base_a = torch.randn(n + 3, dtype=tdt, device="cuda", generator=gen)
    base_y = torch.randn(n + 3, dtype=tdt, device="cuda", generator=gen)
    for off in (0, 1):  # off=1: pointers not 16-byte aligned
        a, y = base_a[off:off + n], base_y[off:off + n].clone()
        ref = 0.25 * a + y
        ops.axpy[1, 1](0.25, a, y)
        assert torch.equal(y, ref) or float((y - ref).abs().max()) <= 1e-6 * float(ref.abs().max())
        out = torch.empty(n + 1, dtype=tdt, device="cuda")[off:][: n] if off else torch.empty(n, dtype=tdt, device="cuda")
        ops.copy[1, 1](a, out)
        assert torch.equal(out, a)
        ops.square[1, 1](a, out)
        assert torch.equal(out, a * a)
        ops.fill[1, 1](-1.5, out)
        assert bool((out == -1.5).all())


def test_empty_launch_and_host_arrays_raise(ops):
    x = torch.zeros(8, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError):
        ops.mass_operator[1, 128](x, x[:0], x, torch.zeros((0, 9), dtype=torch.float64, device="cuda"),
                                  torch.zeros((0, 9), dtype=torch.int32, device="cuda"))
    with pytest.raises(Exception):
        ops.axpy[1, 1](1.0, np.zeros(4), np.zeros(4))  # host arrays: no CPU path
    with pytest.raises(ValueError):
        ops.stiffness_operator(8, np.float64)


@pytest.mark.parametrize("P,tag,kind", [(2, "f64", "box"), (4, "f64", "greedy"), (4, "f32", "box"), (5, "f64", "greedy"),
                                        (7, "f32", "greedy")])
def test_coloured_stiffness_is_exact_and_reproducible(ops, P, tag, kind):
    """Atomics-free launches over colour classes (FUS_NO_ATOMICS): same result as the
    oracle, bit-identical from run to run, and equal to the atomic path to rounding."""
    from fenicsx_fus_gpu_b200 import substrate as S, utils
    from oracle import oracle as orc

    dt = np.float64 if tag == "f64" else np.float32
    tb = S.element_tables(P, "basix", dt)
    mesh = S.create_box((7, 6, 5), (1.0, 0.9, 0.8), dtype=dt, perturb=0.2, seed=11)
    dofmap = S.tensor_dofmap(mesh, P)
    nd, Nc, n = int(dofmap.max()) + 1, dofmap.shape[0], P + 1
    colours = S.box_cell_colours(mesh) if kind == "box" else utils.colour_cells(mesh.x_dofs, seed=1)
    perm, off = utils.colour_order(colours)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(nd).astype(dt)
    coeff = rng.uniform(0.5, 2.0, Nc).astype(dt)
    G = np.zeros((Nc, tb.n**3, 6), dt)
    orc.compute_scaled_geometrical_factor(G, (mesh.x_dofs, mesh.x_g), Nc, tb.dphi, tb.wts)
    y_ref = np.zeros(nd, dt)
    orc.stiffness_operator(P, x, coeff, y_ref, G, dofmap, tb.dphi_1D)

    K = ops.stiffness_operator(P, dt, colour_offsets=off)
    xd, cd, Gd, dmd = d(x), d(coeff[perm]), d(G[perm]), d(dofmap[perm])
    runs = []
    for _ in range(3):
        y = torch.zeros(nd, dtype=xd.dtype, device="cuda")
        K[Nc, (n, n, n)](xd, cd, y, Gd, dmd, tb.dphi_1D)
        runs.append(y.cpu().numpy())
    assert rel_l2(runs[0], y_ref) < TOL[tag]
    assert np.array_equal(runs[0], runs[1]) and np.array_equal(runs[0], runs[2])
    with pytest.raises(Exception):
        ops.stiffness_operator(P, dt, colour_offsets=[0, Nc - 1])[Nc, (n, n, n)](xd, cd, y, Gd, dmd, tb.dphi_1D)
