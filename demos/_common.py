"""Shared launcher bits of the demo scripts: one process per GPU under torchrun
(``python -m torch.distributed.run --nproc-per-node N demos/demo_*.py``) or a
plain ``python demos/demo_*.py`` on one GPU."""

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def init():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world


def parser(desc, degree, cells):
    ap = argparse.ArgumentParser(description=desc)
    ap.add_argument("--degree", type=int, default=degree)
    ap.add_argument("--cells", type=int, default=cells, help="cells per direction PER GPU (default: the reference demo's mesh)")
    ap.add_argument("--steps", type=int, default=0, help="number of steps (default: run to the demo's final time)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    return ap


def run(solver, t0, dt, nsteps, rank, report=100):
    """The demos' loop: print u[0] every ``report`` steps (cuda/demo_linear_box.py:569-570),
    then "Solve time" / "Solve time per step" (:579-581)."""
    import torch

    solver.init()
    torch.cuda.synchronize()
    t_start = time.perf_counter()
    done = 0
    while done < nsteps:
        k = min(report, nsteps - done)
        solver.rk4(solver.t if done else t0, dt, k)
        done += k
        u0 = float(solver.u[0])  # device -> host read of one value (synchronises)
        if rank == 0:
            print(f"t: {solver.t:5.5},\t Steps: {done}/{nsteps}, \t u[0] = {u0}", flush=True)
    torch.cuda.synchronize()
    el = time.perf_counter() - t_start
    if rank == 0:
        print(f"Solve time: {el}")
        print(f"Solve time per step: {el / max(1, nsteps)}")
    return el


def finish(world):
    if world > 1:
        import torch
        import torch.distributed as dist

        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)
