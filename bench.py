#!/usr/bin/env python
"""bench.py - the headline benchmark of the hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "demo_linear_box"): linear acoustic wave,
explicit RK4, degree-4 hexahedra, 80^3 cells = 33 076 161 dofs per GPU, float64,
source facets on x=0 and absorbing facets on x=L (cuda/demo_linear_box.py:53-122,
256-263).  A *step* is one RK4 time step = 4 fused stages.  With N > 1 the box is
block-partitioned, every GPU keeps an 80^3 block (weak scaling) and the ghost
dofs are exchanged over NCCL each stage.

The JSON line: ``value`` = fused-RK-stage throughput in GDoF/s (global dofs x 4
stages x K / time, K graph-replayed steps, state resident in HBM), the metric
BASELINE.json names ("operator GDoF/s and RK time-steps/s"); ``steps_per_s`` and the
stand-alone operator GDoF/s ride along.  ``e2e`` = the same steps with, per step, the
H2D copy of the step's source amplitudes from pinned memory and the D2H read of a
sampled pressure plane; ``e2e_operator_host_buffers`` = one stiffness action through
the C ABI with x and y in pinned HOST memory.  ``roofline`` = the dominant kernel (the
stiffness kernel; the dual-stiffness stage kernel for the Westervelt workload):
algorithmic bytes / its mean duration inside the RK steps (CUDA events
around each launch), against the measured HBM peak, with the stand-alone back-to-back
figure beside it.  ``stage_roofline`` = the whole fused stage on the bytes it moves.
``affine_geometry`` = the same steps with ``geometry="auto"`` (cells with a constant
Jacobian keep 6 geometric factors; extra, not the headline; ``--geometry auto`` makes it
the measured path and says so in ``config``).  ``cpu_baseline`` is the reference's CPU
path (its own C++ sum-factorisation templates when oracle/_ref is built, else the C
port) timed on this host's cores on a bounded sample of the same workload.
``--workload linear_piston|nonlinear_bowl`` time BASELINE.json's other demo configs.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload constants: cuda/demo_linear_box.py:53-66, 116
P = 4
F0, P0, C0, RHO = 0.5e6, 60000.0, 1500.0, 1000.0
DOMAIN_LENGTH = 0.12
N_PER_GPU = 80  # int(2 * 0.12 / (1500 / 0.5e6)) = 80 cells per direction
CFL = 0.65
CPU_SAMPLE_N = 24  # cells per direction of the bounded CPU sample (P=4: 912 673 dofs)


def peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cfl_dt(h):
    dt = CFL * h / (C0 * P**2)
    period = 1.0 / F0
    return period / (int(period / dt) + 1)


# --------------------------------------------------------------------------- #
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------- #


class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln for ts, ln in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [ln for _, ln in self.lines]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nme, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- #
# problem set-up (device-side geometry; host substrate for mesh / dofmap)
# --------------------------------------------------------------------------- #


WORKLOADS = {
    # name: degree, c0, rho, f0, reference cell size h, CFL, BASELINE.json config it stands for
    "linear_box": dict(P=4, c0=1500.0, rho=1000.0, f0=0.5e6, h=0.12 / 80, cfl=0.65, n=80, nonlinear=False,
                       label="demo_linear_box: linear wave, RK4 (BASELINE.json configs[1])"),
    # cuda/demo_linear_piston.py:53-66: degree 5, piston of radius 10 mm on z=0, every other exterior facet absorbing
    "linear_piston": dict(P=5, c0=1500.0, rho=1000.0, f0=0.5e6, h=0.12 / 93, cfl=0.65, n=58, nonlinear=False,
                          label="demo_linear_piston: planar piston + absorbing facets, RK4 (BASELINE.json configs[2])"),
    # cuda/demo_nonlinear_bowl.py:61-75,122: Westervelt, c=1480, f=1.1 MHz, beta=3.5, alpha=0.2 dB, CFL 0.4
    "nonlinear_bowl": dict(P=4, c0=1480.0, rho=1000.0, f0=1.1e6, h=0.08 / 198, cfl=0.4, n=99, nonlinear=True,
                           label="demo_nonlinear_bowl: Westervelt, source disc + all facets absorbing, RK4 "
                                 "(BASELINE.json configs[3])"),
}


def build_problem(rank, world, n_per_gpu, dtype, halo_kind, workload="linear_box", degree=None, geometry="stream"):
    """The demo's preamble through fenicsx_fus_gpu_b200.problem (device geometry,
    block partition, halo).  Returns the solver and an info dict."""
    from fenicsx_fus_gpu_b200 import problem
    from fenicsx_fus_gpu_b200 import substrate as S

    W = WORKLOADS[workload]
    deg = degree or W["P"]
    grid = S.block_grid(world)
    ncells = tuple(n_per_gpu * g for g in grid)
    h = W["h"]  # the demo's cell size; the box grows with the rank grid
    lengths = tuple(h * n for n in ncells)
    used = halo_kind if world > 1 else "none"
    try:
        su = problem.box_setup(deg, ncells, lengths, dtype, rank, world, grid=grid, halo_kind=halo_kind)
    except Exception as e:  # peer memory unavailable on this box: NCCL send/recv round instead
        if halo_kind != "p2p" or world == 1:
            raise
        print(f"[bench] peer-memory halo unavailable ({e!r}); using the NCCL halo", file=sys.stderr, flush=True)
        used = "nccl"
        su = problem.box_setup(deg, ncells, lengths, dtype, rank, world, grid=grid, halo_kind="nccl")

    def make_solver(geometry="stream"):
        if workload == "linear_box":
            return problem.linear_solver(su, source_facets=[2], absorbing_facets=[3], rho=W["rho"], c0=W["c0"],
                                         f0=W["f0"], p0=P0, geometry=geometry)
        if workload == "linear_piston":
            piston = problem.disc(0, 1, (0.5 * lengths[0], 0.5 * lengths[1]), 0.01)
            return problem.linear_solver(
                su, source_facets=[0], absorbing_facets=[0, 1, 2, 3, 4, 5], rho=W["rho"], c0=W["c0"], f0=W["f0"],
                p0=P0, source_predicate=piston, geometry=geometry,
                absorbing_predicate=lambda cen: ~(piston(cen) & (cen[:, 2] < 0.5 * h)))
        centre = (0.5 * lengths[1], 0.5 * lengths[2])
        return problem.westervelt_solver(
            su, source_facets=[2], absorbing_facets=[0, 1, 2, 3, 4, 5], rho=W["rho"], c0=W["c0"], f0=W["f0"],
            beta=3.5, alpha_dB=0.2, source_predicate=problem.disc(1, 2, centre, 0.3 * lengths[1]), geometry=geometry)

    solver = make_solver(geometry)
    info = dict(make_solver=make_solver, ncells_local=su.mesh.num_cells, ndofs_local=su.ndofs, nlocal=su.nlocal, global_cells=ncells,
                global_dofs=su.global_dofs, grid=grid, h=h, detJ=su.dev["detJ"], tb=su.tables,
                dofmap=su.dev["dofmap"], halo=used, degree=deg,
                dt=problem.cfl_time_step(deg, h, W["c0"], W["f0"], W["cfl"]))
    return solver, info


# --------------------------------------------------------------------------- #
# CPU arm: the reference's CPU path on a bounded sample
# --------------------------------------------------------------------------- #


class CpuRK:
    """numba-cpu/demo_linear_box.py:425-459 with the reference operators
    (oracle/_ref C++ templates, else the oracle C port), ``cores`` threads each
    owning 1/k of the cells - the reference scales through MPI ranks only."""

    def __init__(self, n, dtype, cores):
        from fenicsx_fus_gpu_b200 import substrate as S
        from oracle import oracle as orc

        self.orc, self.cores, self.dtype = orc, cores, dtype
        self.kind = "reference" if orc.ref_lib() is not None else "port"
        tb = S.element_tables(P, "basix", dtype)
        h = DOMAIN_LENGTH / N_PER_GPU
        mesh = S.create_box(n, h * n, dtype=dtype)
        self.dofmap = S.tensor_dofmap(mesh, P)
        self.nd = S.num_dofs(n, P)
        nc = mesh.num_cells
        self.G = np.zeros((nc, tb.n**3, 6), dtype)
        detJ = np.zeros((nc, tb.n**3), dtype)
        orc.compute_scaled_geometrical_factor(self.G, (mesh.x_dofs, mesh.x_g), nc, tb.dphi, tb.wts)
        orc.compute_scaled_jacobian_determinant(detJ, (mesh.x_dofs, mesh.x_g), nc, tb.dphi, tb.wts)
        self.c2 = np.full(nc, -1.0 / RHO, dtype)
        self.m = np.zeros(self.nd, dtype)
        orc.mass_operator(np.ones(self.nd, dtype), np.full(nc, 1.0 / RHO / C0 / C0, dtype), self.m, detJ, self.dofmap)
        bd1, bd2 = S.boundary_facets(mesh, 2), S.boundary_facets(mesh, 3)
        self.fd1 = S.facet_dofmap(self.dofmap, bd1, tb.local_facet_dof)
        self.fd2 = S.facet_dofmap(self.dofmap, bd2, tb.local_facet_dof)
        self.dJ1 = np.zeros((bd1.shape[0], tb.n**2), dtype)
        self.dJ2 = np.zeros((bd2.shape[0], tb.n**2), dtype)
        orc.compute_boundary_facets_scaled_jacobian_determinant(self.dJ1, (mesh.x_dofs, mesh.x_g), bd1, tb.dphi_f, tb.wts_f)
        orc.compute_boundary_facets_scaled_jacobian_determinant(self.dJ2, (mesh.x_dofs, mesh.x_g), bd2, tb.dphi_f, tb.wts_f)
        self.fc1 = np.full(bd1.shape[0], 1.0 / RHO, dtype)
        self.fc2 = np.full(bd2.shape[0], -1.0 / RHO / C0, dtype)
        self.dphi = tb.dphi_1D
        self.dt = cfl_dt(h)
        z = lambda: np.zeros(self.nd, dtype)  # noqa: E731
        self.u, self.v, self.u0, self.v0, self.un, self.vn = z(), z(), z(), z(), z(), z()
        self.ku, self.kv, self.g, self.b = z(), z(), z(), z()
        self.ybuf = np.zeros((cores, self.nd), dtype)
        self.t = 0.0
        self.sample = f"{n}^3 cells, degree {P}, {self.nd} dofs, {np.dtype(dtype).name}"

    def step(self):
        from fenicsx_fus_gpu_b200.solver import A_RUNGE, B_RUNGE, C_RUNGE, linear_source

        o, dt = self.orc, self.dt
        self.u0[:] = self.u
        self.v0[:] = self.v
        for i in range(4):
            self.un[:] = self.u0
            self.vn[:] = self.v0
            o.axpy(A_RUNGE[i] * dt, self.ku, self.un)
            o.axpy(A_RUNGE[i] * dt, self.kv, self.vn)
            o.copy(self.vn, self.ku)
            o.fill(linear_source(self.t + C_RUNGE[i] * dt, F0, P0, C0)[0], self.g)
            o.fill(0.0, self.b)
            self.ybuf[:] = 0
            o.stiffness_ranks(P, self.un, self.c2, self.ybuf, self.G, self.dofmap, self.dphi, self.cores)
            self.b += self.ybuf.sum(axis=0)
            o.mass_operator(self.g, self.fc1, self.b, self.dJ1, self.fd1)
            o.mass_operator(self.vn, self.fc2, self.b, self.dJ2, self.fd2)
            o.pointwise_divide(self.b, self.m, self.kv)
            o.axpy(B_RUNGE[i] * dt, self.ku, self.u)
            o.axpy(B_RUNGE[i] * dt, self.kv, self.v)
        self.t += dt


def time_cpu(steps, warmup, budget_s=15.0):
    """``cores`` threads, each stepping its own box of 1/cores of the sample
    (the reference kernels are serial per MPI rank; this is ``mpirun -n cores``
    on this host without the halo cost).  Runs ``steps`` RK4 steps, extended
    until about ``budget_s`` seconds of work have been timed."""
    cores = len(os.sched_getaffinity(0))
    n = max(4, int(round(CPU_SAMPLE_N / cores ** (1.0 / 3.0))))
    rks = [CpuRK(n, np.float64, 1) for _ in range(cores)]
    done = [0] * cores
    t_end = [0.0] * cores
    go = threading.Barrier(cores + 1)

    def body(i):
        rk = rks[i]
        for _ in range(max(1, warmup)):
            rk.step()
        go.wait()
        t0 = time.perf_counter()
        while done[i] < steps or time.perf_counter() - t0 < budget_s:
            rk.step()
            done[i] += 1
            if time.perf_counter() - t0 > 2.0 * budget_s:
                break
        t_end[i] = time.perf_counter()

    th = [threading.Thread(target=body, args=(i,)) for i in range(cores)]
    [t.start() for t in th]
    go.wait()
    t0 = time.perf_counter()
    [t.join() for t in th]
    el = max(t_end) - t0
    work = sum(rk.nd * 4 * d for rk, d in zip(rks, done))
    nsteps = min(done)
    return dict(value=work / el / 1e9, unit="GDoF/s", cores=cores, kind=rks[0].kind,
                sample=f"{cores} threads x {nsteps}+ RK4 steps, each thread its own {rks[0].sample} box "
                       f"(mpirun -n {cores} emulation without halo cost), {el:.1f} s",
                ms_per_step=el / max(1, nsteps) * 1e3, steps_per_s=nsteps / el, steps=nsteps)


# --------------------------------------------------------------------------- #
# main
# --------------------------------------------------------------------------- #


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=0, help="cells per direction per GPU (default: the workload's)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--workload", default="linear_box", choices=sorted(WORKLOADS),
                    help="which reference demo to time (default: the headline demo_linear_box)")
    ap.add_argument("--degree", type=int, default=0, help="override the workload's polynomial degree (2..7)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-affine", action="store_true", help="skip the extra affine-geometry measurement")
    ap.add_argument("--geometry", default="stream", choices=["stream", "auto"],
                    help="stream (default, the reference's data flow: G read in full every stage) or auto "
                         "(cells with a constant Jacobian keep 6 factors; the JSON line says so in config)")
    ap.add_argument("--halo", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU halo: fused put/get kernels over NVLink peer memory, or NCCL send/recv")
    ap.add_argument("--no-graph", action="store_true", help="launch the steps eagerly instead of replaying a CUDA graph")
    ap.add_argument("--watchdog", type=int, default=0, help="dump all Python stacks to stderr after this many seconds")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    if a.watchdog > 0:
        import faulthandler

        faulthandler.dump_traceback_later(a.watchdog, exit=True)
    t_start = time.time()

    def log(msg):
        if a.verbose:
            print(f"[bench r{os.environ.get('RANK', '0')} +{time.time() - t_start:6.1f}s] {msg}", file=sys.stderr, flush=True)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, a.warmup)
    W = WORKLOADS[a.workload]
    deg = a.degree or W["P"]
    n_per_gpu = a.n_per_gpu or W["n"]
    config = {"workload": f"{W['label']}, degree {deg} hexahedra, {n_per_gpu}^3 cells per GPU, {a.dtype}",
              "degree": deg, "cells_per_gpu": n_per_gpu**3, "parallelism": f"block partition x{world}",
              "l2": "working set per GPU (G + dofmap + 9 vectors, GBs) >> 126 MB L2: no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = time_cpu(a.steps, a.warmup, budget_s=10.0)
        line = {"impl": "reference", "metric": "fused RK4 stage throughput (global dofs x stages / s)",
                "value": r["value"], "unit": "GDoF/s", "n_gpus": a.gpus, "steps": r["steps"], "warmup": max(1, a.warmup),
                "ms_per_step": r["ms_per_step"], "steps_per_s": r["steps_per_s"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "GDoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # stdout carries exactly one JSON line: anything libraries print there meanwhile (NCCL's
    # version banner, ...) is sent to stderr by pointing fd 1 at fd 2 until the line is ready
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from fenicsx_fus_gpu_b200 import _lib
    from fenicsx_fus_gpu_b200 import operators as ops

    dtype = np.float64 if a.dtype == "f64" else np.float32
    s = np.dtype(dtype).itemsize
    log("building the problem")
    solver, info = build_problem(rank, world, n_per_gpu, dtype, a.halo, a.workload, deg, a.geometry)
    config["geometry"] = ("G streamed (reference data flow)" if a.geometry == "stream" else
                          f"auto: {solver.nrect} rectilinear + {solver.naff - solver.nrect} affine of {solver.ncells} "
                          "cells per GPU keep 6 geometric factors instead of 6 n^3")
    config["parallelism"] = (f"block partition x{world}, halo: " +
                             {"p2p": "fused put/get kernels over NVLink peer memory", "nccl": "NCCL send/recv",
                              "none": "none (1 GPU)"}[info["halo"]])
    solver.use_graph = not a.no_graph
    log(f"problem built: {info['ndofs_local']} local dofs, {info['ncells_local']} cells")
    dt = info["dt"]
    lib = _lib.lib()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing: K graph-replayed RK4 steps ------------------
    solver.init()
    solver.rk4(0.0, dt, warmup)  # warm-up (captures the graph)
    torch.cuda.synchronize()
    log(f"warm-up done (graph={'yes' if solver._graph is not None else 'no: ' + str(solver.graph_error)})")
    lib.fus_reset_launch_count()
    barrier()
    cvd = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
    clocks = Clocks(int(cvd[local_rank]) if local_rank < len(cvd) else local_rank)  # nvidia-smi index of this rank's GPU
    clocks.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    barrier()
    e0.record()
    solver.rk4(solver.t, dt, a.steps)
    e1.record()
    barrier()
    w1 = time.time()
    elapsed = maxr(e0.elapsed_time(e1) * 1e-3)
    clk = clocks.stop(w0, w1)
    gdofs_global = info["global_dofs"]
    value = gdofs_global * 4 * a.steps / elapsed / 1e9
    # kernels per step: counted once in eager mode (a graph replay bypasses the C entry points)
    lib.fus_reset_launch_count()
    log(f"timed region done: {elapsed * 1e3 / a.steps:.3f} ms/step")
    solver.use_graph = False
    solver.rk4(solver.t, dt, 1)
    torch.cuda.synchronize()
    per_step = int(lib.fus_launch_count())
    # the dominant kernel IN SITU: the same steps launched eagerly with a CUDA-event pair around
    # every stiffness launch (4 per step) - its duration inside the real stage sequence
    solver.probe = []
    solver.rk4(solver.t, dt, min(a.steps, 10))
    torch.cuda.synchronize()
    t_in_step = float(np.mean([e_a.elapsed_time(e_b) for e_a, e_b in solver.probe])) * 1e-3
    n_probed = len(solver.probe)
    solver.probe = None
    solver.use_graph = not a.no_graph

    # ---- end to end: per step, H2D of the step's source amplitudes from pinned host,
    #      the step, D2H of the sampled pressure plane (the demo's 100x100 probe grid) -------
    nsample = min(10000, info["nlocal"])
    sample_idx = torch.linspace(0, info["nlocal"] - 1, nsample, device="cuda").to(torch.int64)
    sample_dev = torch.empty(nsample, dtype=solver.T, device="cuda")
    sample_host = torch.empty(nsample, dtype=solver.T).pin_memory()
    tab_host = torch.from_numpy(solver.source_table(solver.t, dt, a.steps + warmup)).pin_memory()
    solver.gtab = torch.zeros((1, 8), dtype=solver.T, device="cuda")
    solver.step_dev.zero_()
    if solver.use_graph:
        solver._capture(dt)  # graph bound to the 1-row device table
    log("e2e graph ready")
    row = solver.gtab

    def e2e_step(k):
        row.copy_(tab_host[k:k + 1], non_blocking=True)  # H2D: this step's inputs
        solver.step_dev.zero_()
        solver.replay_step(dt)
        check = _lib.fn("fus_pack_fwd", dtype)(solver.u.data_ptr(), sample_dev.data_ptr(), sample_idx.data_ptr(),
                                               nsample, torch.cuda.current_stream().cuda_stream)
        assert check == 0
        sample_host.copy_(sample_dev, non_blocking=True)  # D2H: the step's result
        torch.cuda.current_stream().synchronize()
        return float(sample_host[0])

    for k in range(warmup):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for k in range(a.steps):
        e2e_step(warmup + k)
    e1.record()
    barrier()
    e2e_elapsed = maxr(max(e0.elapsed_time(e1) * 1e-3, 0.0))
    e2e_wall = maxr(time.perf_counter() - t0)
    e2e_value = gdofs_global * 4 * a.steps / max(e2e_elapsed, e2e_wall) / 1e9
    h2d = 8 * s
    d2h = nsample * s

    log("e2e done")
    # ---- the dominant kernel alone, under CUDA events on the launching stream (roofline):
    #      3 warm-ups, then `reps` back-to-back launches between one event pair (mean) and
    #      `reps` individually bracketed launches (min), SURVEY.md 8(d) protocol ----
    nc, nd = info["ncells_local"], info["ndofs_local"]
    n = deg + 1
    nd3 = n**3
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(nd, dtype=solver.T, device="cuda", generator=gen)
    y = torch.zeros(nd, dtype=solver.T, device="cuda")
    D = torch.from_numpy(info["tb"].dphi_1D).cuda()
    tname = "double" if a.dtype == "f64" else "float"
    if a.geometry == "auto":
        # the solver's own stage-kernel launches (rectilinear / affine / streamed ranges) on its own vectors
        solver._set_tables()
        launch = solver._stage_kernel if W["nonlinear"] else solver._stiffness
        geo = "2" if solver.nrect == nc else ("1" if solver.naff == nc else "mixed")
        kname = f"stiffness_kernel<{tname},{n},{1 if W['nonlinear'] else 0},1,GEO={geo}> (geometry=auto)"
        nstream = nc - solver.naff
        per_dof = 3 if W["nonlinear"] else 2
        bytes_stiff = nstream * (nd3 * 4 + 6 * nd3 * s + 2 * s) + solver.naff * (nd3 * 4 + 8 * s) + per_dof * s * nd
    elif W["nonlinear"]:
        # the Westervelt stage kernel: both stiffness terms in one pass over G (the cell-mass pair is
        # pointwise in the close kernel), launched as the solver launches it, on its own vectors
        solver._set_tables()
        launch = solver._stage_kernel
        kname = f"stiffness_kernel<{tname},{n},1,1,0> (fus_stiffness2_{a.dtype}: K(c3; un) + K(c4; vn))"
        # per cell: dofmap + G + 2 coefficients; per dof: read un, vn, write b
        bytes_stiff = nc * (nd3 * 4 + 6 * nd3 * s + 2 * s) + 3 * s * nd
    else:
        K = ops.stiffness_operator(deg, dtype)

        def launch():
            K[nc, (n, n, n)](x, solver.cell_coeff2, y, solver.G, solver.dofmap, D)
        kname = f"stiffness_kernel<{tname},{n},0,1> (fus_stiffness_{a.dtype})"
        bytes_stiff = nc * (nd3 * 4 + 6 * nd3 * s + s) + 2 * s * nd  # B_K per cell + 2s per dof
    reps = max(20, a.steps)
    for _ in range(3):
        launch()
    barrier()
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
    torch.cuda.synchronize()
    t_stiff = e0.elapsed_time(e1) * 1e-3 / reps
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for ea, eb in evs:
        y.zero_()
        ea.record()
        launch()
        eb.record()
    torch.cuda.synchronize()
    t_stiff_min = float(np.min([ea.elapsed_time(eb) for ea, eb in evs])) * 1e-3
    peak, peak_kind = peaks()
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "stiffness_traffic.json")))
        if (n_per_gpu == tr.get("n_per_gpu") and a.dtype == tr.get("dtype") and a.workload == "linear_box"
                and deg == 4 and world == 1 and a.geometry == "stream"):
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    # achieved = algorithmic bytes / the kernel's average duration inside the RK steps (CUDA events
    # around each of its launches); the stand-alone figure (the same kernel launched back to
    # back, which runs into the power cap sooner) rides along
    roofline = {"kernel": kname, "bound": "hbm", "achieved": bytes_stiff / t_in_step / 1e9, "peak": peak,
                "unit": "GB/s", "frac": bytes_stiff / t_in_step / 1e9 / peak, "traffic": traffic,
                "peak_kind": peak_kind, "algorithmic_bytes_per_launch": bytes_stiff, "launch_ms": t_in_step * 1e3,
                "launches_timed": n_probed, "timing": "CUDA events around every launch of the kernel inside the RK steps",
                "standalone": {"launch_ms": t_stiff * 1e3, "launch_ms_min": t_stiff_min * 1e3, "launches_timed": reps,
                               "achieved": bytes_stiff / t_stiff / 1e9, "frac": bytes_stiff / t_stiff / 1e9 / peak,
                               "timing": "back-to-back launches between one event pair"}}
    # the mass operator (cells) for the "operators" block
    evm = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    c1 = torch.ones(nc, dtype=solver.T, device="cuda")
    for ea, eb in evm:
        y.zero_()
        ea.record()
        ops.mass_operator[1, 128](x, c1, y, info["detJ"], solver.dofmap)
        eb.record()
    torch.cuda.synchronize()
    t_mass = float(np.mean([ea.elapsed_time(eb) for ea, eb in evm])) * 1e-3
    stage_bytes = solver.stage_bytes()
    stage_t = elapsed / (4 * a.steps)

    # ---- the operator through the C ABI with HOST buffers (fus_stiffness_host_*): x from pinned
    #      host memory, y back to pinned host memory, both copies inside the call ----
    op_host = None
    if not W["nonlinear"] and solver.G is not None and a.geometry == "stream":
        xh = torch.randn(nd, dtype=solver.T).pin_memory()
        yh = torch.zeros(nd, dtype=solver.T).pin_memory()
        fh = _lib.fn("fus_stiffness_host", dtype)
        st = torch.cuda.current_stream().cuda_stream

        def host_call():
            rc = fh(xh.data_ptr(), yh.data_ptr(), nd, x.data_ptr(), y.data_ptr(), solver.cell_coeff2.data_ptr(),
                    solver.G.data_ptr(), solver.dofmap.data_ptr(), D.data_ptr(), nc, deg, 0, st)
            assert rc == 0
        host_call()
        barrier()
        t0h = time.perf_counter()
        nh = 5
        for _ in range(nh):
            host_call()  # synchronous on return
        th = maxr((time.perf_counter() - t0h) / nh)
        op_host = {"what": "fus_stiffness_host: y_host += K x_host (H2D x, y; kernel; D2H y) per call",
                   "value": gdofs_global / th / 1e9, "unit": "GDoF/s", "ms_per_call": th * 1e3,
                   "h2d_bytes_per_call": 2 * nd * s, "d2h_bytes_per_call": nd * s}
        del xh, yh

    # ---- the same steps with geometry="auto" (extra, not the headline): cells with a constant
    #      Jacobian keep 6 geometric factors instead of 6 n^3 (results equal to rounding) ----
    affine = None
    if not a.no_affine and a.geometry == "stream" and (world == 1 or info["halo"] == "nccl"):  # (the peer-memory arena holds one solver's vectors)
        del x, y
        sol2 = info["make_solver"]("auto")
        sol2.use_graph = not a.no_graph
        sol2.init()
        sol2.rk4(0.0, dt, warmup)
        barrier()
        e0.record()
        sol2.rk4(sol2.t, dt, a.steps)
        e1.record()
        barrier()
        el2 = maxr(e0.elapsed_time(e1) * 1e-3)
        affine = {"value": gdofs_global * 4 * a.steps / el2 / 1e9, "unit": "GDoF/s", "ms_per_step": el2 / a.steps * 1e3,
                  "steps_per_s": a.steps / el2, "affine_cells_per_gpu": sol2.naff, "cells_per_gpu": sol2.ncells,
                  "algorithmic_bytes_per_stage": sol2.stage_bytes(),
                  "note": "solver(geometry='auto'): G = wq x Gc on cells with a constant Jacobian (all cells of "
                          "this box); not the headline because unstructured meshes stream G"}
        if world > 1:
            sol2._graph = None

    line = {
        "metric": "fused RK4 stage throughput (global dofs x stages / s)", "value": value, "unit": "GDoF/s",
        "workload": a.workload, "n_gpus": world, "steps": a.steps, "warmup": warmup, "ms_per_step": elapsed / a.steps * 1e3,
        "steps_per_s": a.steps / elapsed, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.dtype, "data": "synthetic", "config": {**config, "global_dofs": gdofs_global,
                                                          "global_cells": list(info["global_cells"])},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "GDoF/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps_per_s": a.steps / max(e2e_elapsed, e2e_wall)},
        "gpu_launches": per_step * a.steps,
        "roofline": roofline,
        "stage_roofline": {"bound": "hbm", "achieved": stage_bytes / stage_t / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": stage_bytes / stage_t / 1e9 / peak,
                           "algorithmic_bytes_per_stage": stage_bytes, "stage_ms": stage_t * 1e3,
                           "bytes_model": "stage kernel bytes + 10.25 vector passes per stage (what the fused "
                                          "ping-pong stages move: 9 + 12 + 12 + 8 per step)",
                           "survey_8d_model": {"algorithmic_bytes_per_stage": solver.stage_bytes_survey(),
                                               "frac": solver.stage_bytes_survey() / stage_t / 1e9 / peak,
                                               "note": "SURVEY.md 8(d): 6s + stiffness + 8s per dof, more than "
                                                       "this implementation moves"}},
        "e2e_operator_host_buffers": op_host,
        "affine_geometry": affine,
        "operators": {("westervelt_stage_kernel_gdofs" if W["nonlinear"] else "stiffness_gdofs"): info["ndofs_local"] / t_stiff / 1e9, "mass_gdofs": info["ndofs_local"] / t_mass / 1e9,
                      "per_gpu_dofs": info["ndofs_local"]},
    }
    if rank == 0:
        if world == 1 and not a.no_cpu and a.workload == "linear_box" and deg == P:
            r = time_cpu(3, 1, budget_s=12.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down: ncclCommDestroy blocks while captured graphs still hold NCCL
        # kernels, so drop the graphs first and do not wait on communicator
        # destruction (the process is about to exit anyway).
        dist.barrier()
        torch.cuda.synchronize()
        solver._graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
