"""The reference's CPU path on this host's cores - the baseline arm of ``bench.py``.

Runs the ACTUAL benchmark problem (``demo_linear_box``: degree-4 hexahedra, 80^3 cells,
33 076 161 dofs, float64 by default), split into ``k`` blocks of cells, one PROCESS per core
(``k = len(os.sched_getaffinity(0))``), each process stepping its block through the RK4 loop
of ``/root/reference/numba-cpu/demo_linear_box.py:322-382, 425-459`` with the reference's own
UNMODIFIED ``numba-cpu/operators.py`` (shipped as ``baseline/_ref/numba_cpu/`` by
``baseline/make_ref.py``).  The reference's kernels are serial ``@njit`` loops that scale
through MPI ranks only; mpi4py / DOLFINx are not in this image, so this emulates
``mpirun -n k`` WITHOUT the halo cost (the blocks do not exchange interface values: the
figure is an upper bound of what the reference reaches on these cores).

kinds: ``numba``  the reference's numba-cpu operators (headline; ``cpu_baseline.kind`` "reference")
       ``cpp``    the reference's C++ sum-factorisation templates (oracle/_ref, compiled in place)
       ``port``   the oracle's plain-C restatement (fallback when neither reference build exists)

Workers are processes (no GIL), pinned one per core, numba / OpenMP threading off.  A solo
step of worker 0 (everybody else idle) is timed before the parallel phase: the ratio of the
two is the memory-contention loss and is reported, so a figure that does not scale with the
core count is visible as such.
"""

from __future__ import annotations

import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "baseline", "_ref")

A_RUNGE = (0.0, 0.5, 0.5, 1.0)
B_RUNGE = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)
C_RUNGE = (0.0, 0.5, 0.5, 1.0)


def available_kind(prefer="numba"):
    order = [prefer] + [k for k in ("numba", "cpp", "port") if k != prefer]
    for k in order:
        if k == "numba" and os.path.exists(os.path.join(REFDIR, "numba_cpu", "operators.py")):
            try:
                import numba  # noqa: F401

                return "numba"
            except Exception:
                continue
        if k == "cpp" and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libfus_ref.so")):
            return "cpp"
        if k == "port":
            return "port"
    return "port"


def _split(N, parts):
    base, rem = divmod(N, parts)
    out = [0]
    for i in range(parts):
        out.append(out[-1] + base + (1 if i < rem else 0))
    return out


class _Block:
    """One process's block of the box: arrays of numba-cpu/demo_linear_box.py:232-300."""

    def __init__(self, cfg, idx):
        sys.path.insert(0, ROOT)
        from fenicsx_fus_gpu_b200 import substrate as S
        from oracle import oracle as orc  # set-up only (geometry tables); never timed

        P, dtype, kind = cfg["P"], np.dtype(cfg["dtype"]), cfg["kind"]
        grid, n, h = cfg["grid"], cfg["n"], cfg["h"]
        b = (idx // (grid[1] * grid[2]), (idx // grid[2]) % grid[1], idx % grid[2])
        cuts = [_split(n[d], grid[d]) for d in range(3)]
        nc = tuple(cuts[d][b[d] + 1] - cuts[d][b[d]] for d in range(3))
        self.P, self.dtype, self.kind = P, dtype, kind
        tb = S.element_tables(P, "basix", dtype)
        mesh = S.create_box(nc, tuple(h * c for c in nc), dtype=dtype)
        self.dofmap = S.tensor_dofmap(mesh, P)
        self.nd = S.num_dofs(nc, P)
        ncell = mesh.num_cells
        self.G = np.zeros((ncell, tb.n**3, 6), dtype)
        detJ = np.zeros((ncell, tb.n**3), dtype)
        orc.compute_scaled_geometrical_factor(self.G, (mesh.x_dofs, mesh.x_g), ncell, tb.dphi, tb.wts)
        orc.compute_scaled_jacobian_determinant(detJ, (mesh.x_dofs, mesh.x_g), ncell, tb.dphi, tb.wts)
        rho, c0 = cfg["rho"], cfg["c0"]
        self.f0, self.p0, self.c0 = cfg["f0"], cfg["p0"], c0
        self.cell_coeff1 = np.full(ncell, 1.0 / rho / c0 / c0, dtype)
        self.cell_coeff2 = np.full(ncell, -1.0 / rho, dtype)
        # source facets on x = 0 (first block layer), absorbing on x = L (last): demo_linear_box.py:256-263
        e = np.zeros((0, 2), np.int32)
        bd1 = S.boundary_facets(mesh, 2) if b[0] == 0 else e
        bd2 = S.boundary_facets(mesh, 3) if b[0] == grid[0] - 1 else e
        self.fd1 = S.facet_dofmap(self.dofmap, bd1, tb.local_facet_dof)
        self.fd2 = S.facet_dofmap(self.dofmap, bd2, tb.local_facet_dof)
        self.dJ1 = np.zeros((bd1.shape[0], tb.n**2), dtype)
        self.dJ2 = np.zeros((bd2.shape[0], tb.n**2), dtype)
        if bd1.shape[0]:
            orc.compute_boundary_facets_scaled_jacobian_determinant(self.dJ1, (mesh.x_dofs, mesh.x_g), bd1, tb.dphi_f, tb.wts_f)
        if bd2.shape[0]:
            orc.compute_boundary_facets_scaled_jacobian_determinant(self.dJ2, (mesh.x_dofs, mesh.x_g), bd2, tb.dphi_f, tb.wts_f)
        self.fc1 = np.full(bd1.shape[0], 1.0 / rho, dtype)
        self.fc2 = np.full(bd2.shape[0], -1.0 / rho / c0, dtype)
        n3, n2 = tb.n**3, tb.n**2
        if kind == "numba":
            sys.path.insert(0, os.path.join(REFDIR, "numba_cpu"))
            import operators as ref_ops  # the reference's numba-cpu/operators.py, unmodified

            self.mass_cell = ref_ops.mass_operator(n3, dtype.type)
            self.mass_facet = ref_ops.mass_operator(n2, dtype.type)
            # numba-cpu/time_operators.py:213,222: dphi flattened, closure constant
            self.stiff = ref_ops.stiffness_operator(P, np.ascontiguousarray(tb.dphi_1D.flatten()), dtype.type)
        elif kind == "cpp":
            D = tb.dphi_1D
            self.mass_cell = self.mass_facet = orc.ref_mass_operator
            self.stiff = lambda x, c, y, G, dm: orc.ref_stiffness_operator(P, x, c, y, G, dm, D)
        else:
            D = tb.dphi_1D
            self.mass_cell = self.mass_facet = orc.mass_operator
            self.stiff = lambda x, c, y, G, dm: orc.stiffness_operator(P, x, c, y, G, dm, D)
        z = lambda: np.zeros(self.nd, dtype)  # noqa: E731
        self.m, self.b, self.g = z(), z(), z()
        self.u_, self.v_, self.un, self.vn, self.u0, self.v0, self.ku, self.kv = (z() for _ in range(8))
        self.mass_cell(np.ones(self.nd, dtype), self.cell_coeff1, self.m, detJ, self.dofmap)  # :300-307
        self.t = 0.0
        self.ncell = ncell

    def f(self, t, u, v, result):
        """numba-cpu/demo_linear_box.py:322-382 (single rank: the scatters are no-ops)."""
        T, alpha = 1.0 / self.f0, 4
        window = 0.5 * (1 - np.cos(self.f0 * np.pi * t / alpha)) if t < T * alpha else 1.0
        w0 = 2.0 * np.pi * self.f0
        self.g[:] = window * self.p0 * w0 / self.c0 * np.cos(w0 * t)
        self.b[:] = 0.0
        self.stiff(u, self.cell_coeff2, self.b, self.G, self.dofmap)
        if self.fd1.shape[0]:
            self.mass_facet(self.g, self.fc1, self.b, self.dJ1, self.fd1)
        if self.fd2.shape[0]:
            self.mass_facet(v, self.fc2, self.b, self.dJ2, self.fd2)
        result[:] = self.b[:] / self.m[:]

    def step(self, dt):
        """numba-cpu/demo_linear_box.py:425-459."""
        self.u0[:] = self.u_[:]
        self.v0[:] = self.v_[:]
        for i in range(4):
            self.un[:] = self.u0[:]
            self.vn[:] = self.v0[:]
            self.un += A_RUNGE[i] * dt * self.ku
            self.vn += A_RUNGE[i] * dt * self.kv
            tn = self.t + C_RUNGE[i] * dt
            self.ku[:] = self.vn[:]
            self.f(tn, self.un, self.vn, self.kv)
            self.u_ += B_RUNGE[i] * dt * self.ku
            self.v_ += B_RUNGE[i] * dt * self.kv
        self.t += dt


def _worker(idx, cfg, cores, barrier, q):
    try:
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS", "NUMBA_NUM_THREADS"):
            os.environ[k] = "1"
        if cores:
            try:
                os.sched_setaffinity(0, {cores[idx % len(cores)]})
            except Exception:
                pass
        blk = _Block(cfg, idx)
        dt = cfg["dt"]
        t0 = time.perf_counter()
        for _ in range(max(1, cfg["warmup"])):  # includes the JIT compilation
            blk.step(dt)
        t_warm = time.perf_counter() - t0
        barrier.wait()
        solo = None
        if idx == 0:  # one step with every other core idle: the contention-free rate
            t0 = time.perf_counter()
            blk.step(dt)
            solo = time.perf_counter() - t0
        barrier.wait()
        barrier.wait()  # the parent's clock starts between these two
        t0 = time.perf_counter()
        for _ in range(cfg["steps"]):
            blk.step(dt)
        busy = time.perf_counter() - t0
        barrier.wait()
        q.put((idx, dict(busy=busy, solo=solo, warm=t_warm, ncell=blk.ncell, nd=blk.nd,
                         norm=float(np.linalg.norm(blk.v_)))))
    except BaseException as e:  # pragma: no cover
        q.put((idx, dict(error=repr(e))))
        try:
            barrier.abort()
        except Exception:
            pass


def run(n=80, P=4, dtype="float64", steps=3, warmup=1, kind=None, cores=None, h=0.12 / 80,
        rho=1000.0, c0=1500.0, f0=0.5e6, p0=60000.0, cfl=0.65):
    """Time ``steps`` RK4 steps of the n^3-cell box split over all cores.  Returns a dict."""
    sys.path.insert(0, ROOT)
    from fenicsx_fus_gpu_b200 import substrate as S

    avail = sorted(os.sched_getaffinity(0))
    k = int(cores) if cores else len(avail)
    k = max(1, min(k, n**3))
    kind = available_kind(kind or "numba")
    grid = S.block_grid(k)
    if any(g > n for g in grid):
        raise ValueError(f"cannot split {n}^3 cells over a {grid} grid")
    dt_ = cfl * h / (c0 * P**2)
    period = 1.0 / f0
    dt_ = period / (int(period / dt_) + 1)
    cfg = dict(P=P, dtype=np.dtype(dtype).name, kind=kind, grid=tuple(grid), n=(n, n, n), h=h, rho=rho, c0=c0,
               f0=f0, p0=p0, dt=dt_, steps=int(steps), warmup=int(warmup))
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(k + 1)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(i, cfg, avail, barrier, q), daemon=True) for i in range(k)]
    t_setup = time.perf_counter()
    for p in procs:
        p.start()
    try:
        barrier.wait(timeout=1800)  # every block built, JIT done
        barrier.wait(timeout=600)  # worker 0's solo step done
        t0 = time.perf_counter()
        barrier.wait(timeout=600)
        barrier.wait(timeout=3600)
        elapsed = time.perf_counter() - t0
    except Exception as e:
        errs = []
        while not q.empty():
            errs.append(q.get()[1])
        for p in procs:
            p.terminate()
        raise RuntimeError(f"cpu_arm: a worker failed: {errs or e!r}")
    res = dict(q.get(timeout=120) for _ in range(k))
    for p in procs:
        p.join(timeout=30)
    bad = {i: r["error"] for i, r in res.items() if "error" in r}
    if bad:
        raise RuntimeError(f"cpu_arm: worker errors {bad}")
    gdofs = (P * n + 1) ** 3
    solo = res[0]["solo"]
    busy = max(r["busy"] for r in res.values())
    # the solo step covers worker 0's block only; scaled to the whole box over k ideal cores
    ideal = solo * sum(r["ncell"] for r in res.values()) / res[0]["ncell"] / k
    manifest = None
    if kind == "numba":
        try:
            manifest = json.load(open(os.path.join(REFDIR, "MANIFEST.json")))["numba_cpu/operators.py"]["sha256"][:16]
        except Exception:
            pass
    kind_label = {"numba": "reference", "cpp": "reference", "port": "port"}[kind]
    what = {"numba": "the reference's numba-cpu/operators.py (unmodified, baseline/_ref)",
            "cpp": "the reference's C++ sum-factorisation templates (oracle/_ref)",
            "port": "the oracle's plain-C port"}[kind]
    return dict(
        value=gdofs * 4 * steps / elapsed / 1e9, unit="GDoF/s", cores=k, kind=kind_label, impl=kind,
        ms_per_step=elapsed / steps * 1e3, steps_per_s=steps / elapsed, steps=int(steps), warmup=int(max(1, warmup)),
        elapsed_s=elapsed, setup_s=t0 - t_setup, global_dofs=gdofs, grid=list(grid), dtype=np.dtype(dtype).name,
        sample=(f"the whole {n}^3-cell degree-{P} box ({gdofs} dofs, {np.dtype(dtype).name}), {steps} RK4 steps "
                f"(numba-cpu/demo_linear_box.py:425-459), {what}; {k} processes (one per core, grid "
                f"{'x'.join(map(str, grid))} of cell blocks, mpirun -n {k} emulation without halo exchange)"),
        scaling_check=dict(solo_step_s_worker0=solo, ideal_ms_per_step=ideal * 1e3,
                           parallel_efficiency=ideal / (elapsed / steps), slowest_worker_busy_s=busy,
                           note="ideal = worker 0's step alone on an idle host, scaled to the box over k cores; "
                                "efficiency < 1 is memory-bandwidth contention between the processes"),
        operators_sha256=manifest, norm_v=float(np.sqrt(sum(r["norm"] ** 2 for r in res.values()))))


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=80)
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--kind", default=None, choices=[None, "numba", "cpp", "port"])
    ap.add_argument("--cores", type=int, default=0)
    ap.add_argument("--dtype", default="float64")
    a = ap.parse_args()
    print(json.dumps(run(a.n, a.degree, a.dtype, a.steps, a.warmup, a.kind, a.cores or None)))
