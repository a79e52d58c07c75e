"""Where does a multi-GPU RK4 stage spend its time?  (no nsys in the image)

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
        tools/mgpu_timeline.py [--n 80] [--steps 10] [--out gpurun_out/timeline.json]

Runs the bench workload eagerly with a CUDA-event pair around every launch of a step (both
streams), prints per-kernel mean durations, the per-stage sum, and the graph-replayed step time of
several variants (side stream on / off, cell split on / off) for comparison.  Also works on 1 GPU.
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=80)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--split-mode", default="none")
    ap.add_argument("--rank-grid", default=None, help="AxBxC rank grid (default: the most cubic one)")
    ap.add_argument("--no-renumber", action="store_true", help="keep the index map's local numbering")
    ap.add_argument("--fake-p2p", action="store_true",
                    help="1 GPU: a one-rank peer-memory halo (no neighbours) - vectors in the symmetric arena and the "
                         "skip mask passed to the close kernel, nothing exchanged: their cost in situ")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    import bench

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_graph(solver, dt, steps):
        solver.init()
        solver.rk4(0.0, dt, 3)
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        solver.rk4(solver.t, dt, steps)
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"n_gpus": world, "n": a.n, "rank_grid": a.rank_grid, "renumber_shared": not a.no_renumber}
    if a.fake_p2p:
        from fenicsx_fus_gpu_b200 import problem
        from fenicsx_fus_gpu_b200.scatterer import P2PHaloExchange, SymmFabric

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29578")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        h = 0.12 / 80
        su = problem.box_setup(4, a.n, h * a.n, np.float64)
        fab = SymmFabric(P2PHaloExchange.arena_bytes(su.ndofs, np.float64))
        e = [[], np.zeros(0, np.int32), np.zeros(0, np.int32)]
        su.halo = P2PHaloExchange(fab, e, e, su.ndofs, 0, np.float64)
        solver = problem.linear_solver(su, source_facets=[2], absorbing_facets=[3], p0=60000.0)
        info = {"dt": problem.cfl_time_step(4, h, 1500.0, 0.5e6, 0.65)}
    else:
        solver, info = bench.build_problem(rank, world, a.n, np.float64, "p2p", "linear_box", 4, "stream",
                                           rank_grid=[int(v) for v in a.rank_grid.split("x")] if a.rank_grid else None,
                                           renumber_shared=not a.no_renumber)
    dt = info["dt"]
    solver.split_mode = a.split_mode
    out["graph_ms_per_step"] = timed_graph(solver, dt, a.steps)
    out["graph_ms_per_step_again"] = timed_graph(solver, dt, a.steps)
    if world > 1:
        # same captured work, exchange kernels on the main stream (no fork / join)
        solver.halo.use_side = False
        solver._graph = None
        solver._graph_dt = None
        out["graph_ms_per_step_no_side_stream"] = timed_graph(solver, dt, a.steps)
        solver.halo.use_side = True
        solver._graph = None
        solver._graph_dt = None

    # eager, instrumented
    rec = []

    def wrap(obj, name, label):
        f = getattr(obj, name)

        def g(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*args, **kw)
            e1.record()
            rec.append((label(*args, **kw) if callable(label) else label, e0, e1))
            return r

        setattr(obj, name, g)

    wrap(solver, "_assemble", "stiffness")
    wrap(solver, "_boundary", "boundary")
    wrap(solver, "_close", lambda *x, **k: f"close{'(masked)' if solver.p2p else ''}[stage {x[0]}]")
    if solver.p2p:
        wrap(solver, "_close_shared", "close_shared+put")
        for nm in ("put", "wait_forward", "wait_reverse", "barrier"):
            wrap(solver.halo, nm, "halo." + nm)
    solver.use_graph = False
    solver.init()
    solver.rk4(0.0, dt, 2)
    sync()
    rec.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    solver.rk4(solver.t, dt, a.steps)
    e1.record()
    sync()
    out["eager_ms_per_step"] = e0.elapsed_time(e1) / a.steps
    agg = {}
    for label, s0, s1 in rec:
        agg.setdefault(label, []).append(s0.elapsed_time(s1))
    out["kernels_ms"] = {k: {"mean": float(np.mean(v)), "min": float(np.min(v)), "per_step": len(v) / a.steps,
                             "ms_per_step": float(np.sum(v)) / a.steps} for k, v in agg.items()}
    out["sum_of_kernels_ms_per_step"] = float(sum(v["ms_per_step"] for v in out["kernels_ms"].values()))
    allo = [None] * world
    if world > 1:
        dist.all_gather_object(allo, out)
    else:
        allo = [out]
    if rank == 0:
        for r, o in enumerate(allo):
            print(f"--- rank {r}: graph {o['graph_ms_per_step']:.3f} / {o['graph_ms_per_step_again']:.3f} ms/step"
                  + (f", no side stream {o['graph_ms_per_step_no_side_stream']:.3f}" if world > 1 else "")
                  + f", eager {o['eager_ms_per_step']:.3f}, sum of kernels {o['sum_of_kernels_ms_per_step']:.3f}")
            for k, v in o["kernels_ms"].items():
                print(f"    {k:24s} mean {v['mean'] * 1e3:8.1f} us  min {v['min'] * 1e3:8.1f} us  x{v['per_step']:.0f}/step"
                      f"  = {v['ms_per_step']:.3f} ms/step")
        if a.out:
            json.dump(allo, open(a.out, "w"), indent=1)
    if world > 1 or a.fake_p2p:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
