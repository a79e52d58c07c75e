"""Operator scaling sweep (BASELINE.json configs[4]): degree 2-7, f32/f64, at a
given dof budget on one GPU.  Prints one JSON line per (degree, dtype) with the
stiffness / mass / fused-stage throughput and the fraction of the measured HBM
roofline (algorithmic bytes of SURVEY.md section 8d).

    python tools/sweep.py [--dofs 1e6,8e6,30e6,125e6] [--degrees 2,3,4,5,6,7] [--dtypes f64,f32] [--stage] \
        > profiles/rNN_sweep.jsonl

``--stage`` adds the fused linear RK stage, streamed and with geometry="auto" (rectilinear
cells on these boxes: 6 factors per cell instead of 6 n^3).
"""

import argparse
import gc
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fenicsx_fus_gpu_b200 import operators as ops  # noqa: E402
from fenicsx_fus_gpu_b200 import problem  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(f, reps):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dofs", default="30e6", help="comma-separated dof budgets")
    ap.add_argument("--degrees", default="2,3,4,5,6,7")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--stage", action="store_true", help="also time the fused linear RK stage")
    a = ap.parse_args()
    pk = peak()
    for dofs, P in [(float(d), int(x)) for d in a.dofs.split(",") for x in a.degrees.split(",")]:
        for tag in a.dtypes.split(","):
            dt = np.float64 if tag == "f64" else np.float32
            s = np.dtype(dt).itemsize
            N = max(2, int(round((dofs ** (1.0 / 3.0) - 1) / P)))
            su = problem.box_setup(P, N, 0.0015 * N, dt)
            n, nd3, nc, nd = P + 1, (P + 1) ** 3, su.mesh.num_cells, su.ndofs
            tdt = torch.float64 if tag == "f64" else torch.float32
            gen = torch.Generator(device="cuda").manual_seed(1)
            x = torch.randn(nd, dtype=tdt, device="cuda", generator=gen)
            y = torch.zeros(nd, dtype=tdt, device="cuda")
            c = torch.ones(nc, dtype=tdt, device="cuda")
            D = torch.from_numpy(su.tables.dphi_1D).cuda()
            K = ops.stiffness_operator(P, dt)
            ts = timeit(lambda: K[nc, (n, n, n)](x, c, y, su.dev["G"], su.dev["dofmap"], D), a.reps)
            tm = timeit(lambda: ops.mass_operator[1, 128](x, c, y, su.dev["detJ"], su.dev["dofmap"]), a.reps)
            bs = nc * (nd3 * 4 + 6 * nd3 * s + s) + 2 * s * nd
            bm = nc * (nd3 * (4 + s) + s) + 2 * s * nd
            rec = {"degree": P, "dtype": tag, "cells": nc, "dofs": nd,
                   "stiffness_ms": ts * 1e3, "stiffness_gdofs": nd / ts / 1e9, "stiffness_frac": bs / ts / 1e9 / pk,
                   "mass_ms": tm * 1e3, "mass_gdofs": nd / tm / 1e9, "mass_frac": bm / tm / 1e9 / pk}
            if a.stage:
                sol = problem.linear_solver(su, [2], [3])
                sol.init()
                dtm = problem.cfl_time_step(P, su.h, 1500.0, 0.5e6, 0.65)
                sol.rk4(0.0, dtm, 3)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                sol.rk4(sol.t, dtm, a.reps)
                e1.record()
                torch.cuda.synchronize()
                tst = e0.elapsed_time(e1) * 1e-3 / (4 * a.reps)
                rec.update({"stage_ms": tst * 1e3, "stage_gdofs": nd / tst / 1e9,
                            "stage_frac": sol.stage_bytes() / tst / 1e9 / pk, "steps_per_s": 1.0 / (4 * tst)})
                del sol
                sol = problem.linear_solver(su, [2], [3], geometry="auto")
                sol.init()
                sol.rk4(0.0, dtm, 3)
                torch.cuda.synchronize()
                e0.record()
                sol.rk4(sol.t, dtm, a.reps)
                e1.record()
                torch.cuda.synchronize()
                tau = e0.elapsed_time(e1) * 1e-3 / (4 * a.reps)
                rec.update({"auto_stage_ms": tau * 1e3, "auto_stage_gdofs": nd / tau / 1e9,
                            "auto_rect_cells": sol.nrect, "auto_steps_per_s": 1.0 / (4 * tau)})
                del sol
            print(json.dumps(rec), flush=True)
            del su, x, y, c, K, D
            gc.collect()  # solvers hold captured graphs (private pools) through reference cycles
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
