"""Developer tool: time individual operators on one GPU with CUDA events.

    python tools/bench_ops.py --P 4 --N 80 --dtype f64 --reps 20 [--op stiffness|mass|stage]

Prints one JSON line per operator: time, GDoF/s, algorithmic GB/s and the
fraction of the measured HBM peak (MEASURED_PEAKS.json).
"""

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenicsx_fus_gpu_b200 import operators as ops  # noqa: E402
from fenicsx_fus_gpu_b200 import precompute as pre  # noqa: E402
from fenicsx_fus_gpu_b200 import substrate as S  # noqa: E402


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def time_op(fn, reps, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.mean(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=4)
    ap.add_argument("--N", type=int, default=80)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--op", default="stiffness,mass")
    ap.add_argument("--perturb", type=float, default=0.0)
    a = ap.parse_args()
    dt = np.float64 if a.dtype == "f64" else np.float32
    tdt = torch.float64 if a.dtype == "f64" else torch.float32
    s = np.dtype(dt).itemsize
    P, N = a.P, a.N
    n = P + 1
    Nd = n**3
    tb = S.element_tables(P, "basix", dt)
    mesh = S.create_box(N, 1.0, dtype=dt, perturb=a.perturb, seed=0)
    dofmap_h = S.tensor_dofmap(mesh, P)
    nd = S.num_dofs(N, P)
    Nc = N**3
    dofmap = torch.from_numpy(dofmap_h).cuda()
    x_dofs, x_g = torch.from_numpy(mesh.x_dofs).cuda(), torch.from_numpy(mesh.x_g).cuda()
    G = torch.empty((Nc, Nd, 6), dtype=tdt, device="cuda")
    detJ = torch.empty((Nc, Nd), dtype=tdt, device="cuda")
    pre.compute_geometry(G, detJ, (x_dofs, x_g), Nc, torch.from_numpy(tb.dphi).cuda(),
                         torch.from_numpy(tb.wts).cuda())
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(nd, dtype=tdt, device="cuda", generator=gen)
    y = torch.zeros(nd, dtype=tdt, device="cuda")
    c = torch.ones(Nc, dtype=tdt, device="cuda")
    D = torch.from_numpy(tb.dphi_1D).cuda()
    pk, how = peak_gbs()
    K = ops.stiffness_operator(P, dt)
    res = []
    for op in a.op.split(","):
        if op == "stiffness":
            f = lambda: K[Nc, (n, n, n)](x, c, y, G, dofmap, D)  # noqa: E731
            bytes_ = Nc * (Nd * 4 + 6 * Nd * s + s) + 2 * s * nd
        elif op in ("coloured", "coloured_greedy"):
            # the same action as one atomics-free launch per colour (8 colours on a box; the
            # greedy colouring of an unstructured mesh needs more): bit-reproducible, and the
            # evidence behind "atomics are faster on B200" in DESIGN.md
            from fenicsx_fus_gpu_b200 import utils
            col = S.box_cell_colours(mesh) if op == "coloured" else utils.colour_cells(mesh.x_dofs)
            perm, off = utils.colour_order(col)
            pt = torch.from_numpy(perm).cuda()
            Gp, dmp = G[pt].contiguous(), dofmap[pt].contiguous()
            Kc = ops.stiffness_operator(P, dt, colour_offsets=off)
            f = lambda: Kc[Nc, (n, n, n)](x, c, y, Gp, dmp, D)  # noqa: E731
            bytes_ = Nc * (Nd * 4 + 6 * Nd * s + s) + 2 * s * nd
        elif op == "affine":
            # affine cells: 6 factors per cell instead of 6 n^3 (fus_stiffness_affine)
            aff, Gc, _ = pre.compress_geometry(G, None, torch.from_numpy(tb.wts).cuda())
            assert bool(aff.all()), "mesh has non-affine cells (use --perturb 0)"
            Ka = ops.stiffness_operator_affine(P, dt)
            wq = torch.from_numpy(tb.wts).cuda()
            f = lambda: Ka[Nc, (n, n, n)](x, c, y, Gc, wq, dofmap, D)  # noqa: E731
            bytes_ = Nc * (Nd * 4 + 7 * s) + 2 * s * nd
        elif op == "rect":
            # rectilinear cells: three decoupled 1-D stiffness products (fus_stiffness_rect)
            aff, Gc, _ = pre.compress_geometry(G, None, torch.from_numpy(tb.wts).cuda())
            assert bool(aff.all()), "mesh has non-affine cells (use --perturb 0)"
            Kr = ops.stiffness_operator_rect(P, dt)
            f = lambda: Kr[Nc, (n, n, n)](x, c, y, Gc, tb.wts, dofmap, tb.dphi_1D)  # noqa: E731
            bytes_ = Nc * (Nd * 4 + 7 * s) + 2 * s * nd
        elif op == "vertex":
            # geometry recomputed in the kernel from 36 trilinear coefficients per cell (fus_stiffness_vertex)
            Tc = pre.trilinear_coefficients((x_dofs, x_g), Nc, tb.dphi, tb.pts)
            Kv = ops.stiffness_operator_vertex(P, dt)
            f = lambda: Kv[Nc, (n, n, n)](x, c, y, Tc, tb.pts_1d, tb.wts_1d, dofmap, tb.dphi_1D)  # noqa: E731
            bytes_ = Nc * (Nd * 4 + 37 * s) + 2 * s * nd
        elif op == "mass":
            f = lambda: ops.mass_operator[1, 128](x, c, y, detJ, dofmap)  # noqa: E731
            bytes_ = Nc * (Nd * (4 + s) + s) + 2 * s * nd
        elif op == "close":
            from fenicsx_fus_gpu_b200._lib import fn, current_stream
            vs = [torch.randn(nd, dtype=tdt, device="cuda", generator=gen) for _ in range(8)]
            m = torch.rand(nd, dtype=tdt, device="cuda", generator=gen) + 1.0
            u_, v_, u0_, v0_, ku_, kv_, un_, b_ = vs
            cl = fn("fus_rk_close", dt)
            f = lambda: cl(u_.data_ptr(), v_.data_ptr(), u0_.data_ptr(), v0_.data_ptr(), ku_.data_ptr(), None,  # noqa: E731
                           un_.data_ptr(), b_.data_ptr(), m.data_ptr(), 1e-9, 0.5e-9, 1, nd, None, current_stream())
            bytes_ = 12 * s * nd
        elif op in ("wstage", "lstage", "wstage_auto", "lstage_auto"):
            from fenicsx_fus_gpu_b200 import problem
            su = problem.box_setup(P, N, 0.0015 * N, dt)
            geo = "auto" if op.endswith("_auto") else "stream"
            if op.startswith("wstage"):
                sol = problem.westervelt_solver(su, [2], [0, 1, 2, 3, 4, 5], p0=1e5, geometry=geo)
                dtm = problem.cfl_time_step(P, su.h, 1480.0, 1.1e6, 0.4)
            else:
                sol = problem.linear_solver(su, [2], [3], geometry=geo)
                dtm = problem.cfl_time_step(P, su.h, 1500.0, 0.5e6, 0.65)
            sol.init()
            sol.rk4(0.0, dtm, 3)
            f = lambda: sol.rk4(sol.t, dtm, 1)  # noqa: E731
            bytes_ = 4 * sol.stage_bytes()
        elif op == "copy":
            a_ = torch.randn(nd * 4, dtype=tdt, device="cuda", generator=gen)
            b_ = torch.empty_like(a_)
            f = lambda: ops.copy[1, 1](a_, b_)  # noqa: E731
            bytes_ = 2 * s * nd * 4
        elif op == "torchcopy":
            a_ = torch.randn(nd * 4, dtype=tdt, device="cuda", generator=gen)
            b_ = torch.empty_like(a_)
            f = lambda: b_.copy_(a_)  # noqa: E731
            bytes_ = 2 * s * nd * 4
        else:
            continue
        mean, mn = time_op(f, a.reps)
        res.append({"op": op, "P": P, "N": N, "dtype": a.dtype, "ndofs": nd, "ms_mean": mean * 1e3,
                    "ms_min": mn * 1e3, "gdofs": nd / mean / 1e9, "alg_GBs": bytes_ / mean / 1e9,
                    "frac": bytes_ / mean / 1e9 / pk, "peak": pk, "peak_kind": how})
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
