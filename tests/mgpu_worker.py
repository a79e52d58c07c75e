"""Worker of tests/test_multigpu.py: run under torchrun on N >= 2 real GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29611 tests/mgpu_worker.py --out /tmp/mgpu.json [--quick]

Every case solves one global box partitioned over the N GPUs (peer-memory halo with
interior / interface overlap, or the NCCL halo) and on rank 0 alone, and records the
rel-L2 difference (fenicsx_fus_gpu_b200.selfcheck).  The scatter cases push the
reference-generated fixture tests/golden/scatter_r2.npz (numba-cpu/scatterer.py outputs)
through the product's scatter_forward / scatter_reverse factories (NCCL) and through the
peer-memory exchange on the real GPUs.
"""

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def scatter_cases(rank, world):
    import torch

    from fenicsx_fus_gpu_b200.scatterer import P2PHaloExchange, SymmFabric, scatter_forward, scatter_reverse

    out = []
    if world != 2:
        return out
    with np.load(os.path.join(ROOT, "tests", "golden", "scatter_r2.npz")) as z:
        g = {k: z[k] for k in z.files}

    def lists(which):
        ranks = g[f"r{rank}_{which}_ranks"]
        return [[g[f"r{rank}_{which}_idx{i}"] for i in range(ranks.size)], g[f"r{rank}_{which}_size"], ranks]

    N = int(g[f"r{rank}_size_local"])
    vec = g[f"r{rank}_vec"]
    nd = vec.size
    od, gd = lists("owners"), lists("ghosts")
    # the reference-shaped factories (NCCL grouped send/recv)
    fwd = scatter_forward(None, od, gd, N, np.float64)
    rev = scatter_reverse(None, od, gd, N, np.float64)
    f = torch.from_numpy(vec).cuda()
    fwd(f)
    r = torch.from_numpy(vec).cuda()
    rev(r)
    torch.cuda.synchronize()
    out.append(dict(case="scatter_forward factory (NCCL) vs numba-cpu fixture", rank=rank,
                    ok=bool(np.array_equal(f.cpu().numpy(), g[f"r{rank}_fwd"]))))
    err = float(np.linalg.norm(r.cpu().numpy() - g[f"r{rank}_rev"]) / np.linalg.norm(g[f"r{rank}_rev"]))
    out.append(dict(case="scatter_reverse factory (NCCL) vs numba-cpu fixture", rank=rank, rel_l2=err, ok=err < 1e-15))
    # the peer-memory exchange on the same lists, twice in a row (no barrier needed in between)
    fab = SymmFabric(P2PHaloExchange.arena_bytes(nd, np.float64, 4))
    halo = P2PHaloExchange(fab, od, gd, N, nd - N, np.float64)
    a, b = halo.alloc(), halo.alloc()
    a.copy_(torch.from_numpy(vec))
    b.copy_(torch.from_numpy(vec))
    halo.forward(a)
    halo.forward(a)
    halo.reverse(b)
    torch.cuda.synchronize()
    halo.status()
    out.append(dict(case="P2PHaloExchange.forward x2 vs numba-cpu fixture", rank=rank,
                    ok=bool(np.array_equal(a.cpu().numpy(), g[f"r{rank}_fwd"]))))
    err = float(np.linalg.norm(b.cpu().numpy() - g[f"r{rank}_rev"]) / np.linalg.norm(g[f"r{rank}_rev"]))
    out.append(dict(case="P2PHaloExchange.reverse vs numba-cpu fixture", rank=rank, rel_l2=err, ok=err < 1e-15))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default=None, help="run only the cases with this partition kind (e.g. blob), full list")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from fenicsx_fus_gpu_b200.selfcheck import multi_gpu_parity

    cases = [
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8),  # graph replay, overlap
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8, use_graph=False),
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8, split_cells=False),
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8, split_mode="two"),
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8, split_mode="fused"),
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8, halo_kind="nccl"),
        dict(workload="westervelt", P=4, n_per_rank=5, dtype="float64", nsteps=6),
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=12, integrator="leapfrog"),
        # unstructured-like partitions: irregular parts, shuffled numbering, hash ownership
        dict(workload="linear", P=4, n_per_rank=5, dtype="float64", nsteps=8, partition="blob"),
        dict(workload="linear", P=3, n_per_rank=6, dtype="float64", nsteps=8, partition="blob", halo_kind="nccl"),
        dict(workload="westervelt", P=3, n_per_rank=5, dtype="float64", nsteps=6, partition="blob"),
        # the index map's own numbering (no shared-last renumbering); ranks split along z: the shared
        # face is the scattered one
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8, renumber_shared=False),
        dict(workload="linear", P=4, n_per_rank=6, dtype="float64", nsteps=8, grid=(1, 1, world)),
    ]
    if not a.quick:
        cases += [
            dict(workload="linear", P=4, n_per_rank=6, dtype="float32", nsteps=8),
            dict(workload="linear", P=4, n_per_rank=8, dtype="float64", nsteps=6, geometry="auto", perturb=0.0),
            dict(workload="piston", P=5, n_per_rank=4, dtype="float64", nsteps=6),
            dict(workload="westervelt_cells", P=3, n_per_rank=5, dtype="float64", nsteps=6),
            dict(workload="westervelt_cells", P=3, n_per_rank=5, dtype="float64", nsteps=6, split_mode="fused"),
            dict(workload="linear", P=4, n_per_rank=8, dtype="float64", nsteps=6, geometry="auto", perturb=0.0, split_mode="two"),
            dict(workload="westervelt", P=4, n_per_rank=5, dtype="float32", nsteps=6),
            dict(workload="westervelt", P=6, n_per_rank=3, dtype="float64", nsteps=4),
            dict(workload="linear", P=2, n_per_rank=9, dtype="float64", nsteps=8),
            dict(workload="linear", P=7, n_per_rank=3, dtype="float64", nsteps=4),
            dict(workload="linear", P=4, n_per_rank=5, dtype="float64", nsteps=12, integrator="leapfrog", partition="blob"),
            dict(workload="linear", P=4, n_per_rank=5, dtype="float32", nsteps=8, partition="blob", split_mode="fused"),
            dict(workload="westervelt_cells", P=3, n_per_rank=5, dtype="float64", nsteps=6, partition="blob"),
        ]
    if a.only:
        cases = [c for c in cases if c.get("partition", "block") == a.only]
    results = []
    for c in cases:
        c = dict(c)
        c["dtype"] = np.dtype(c["dtype"])
        r = multi_gpu_parity(**c)
        results.append(r)
        if rank == 0:
            print(f"[mgpu] {r['workload']} P{r['degree']} {r['dtype']} {r['partition']} halo={r['halo']} geometry={r['geometry']} "
                  f"{r['integrator']} graph={r['graph']} split={r['split_mode']} iface={r['interface_cells']}: rel-L2 u {r['rel_l2_u']:.2e} v {r['rel_l2_v']:.2e} "
                  f"{'ok' if r['ok'] else 'FAIL'}", flush=True)
    sc = scatter_cases(rank, world)
    allsc = [None] * world
    dist.all_gather_object(allsc, sc)
    if rank == 0:
        flat = [x for part in allsc for x in part]
        for x in flat:
            print(f"[mgpu] {x['case']} rank {x['rank']}: {'ok' if x['ok'] else 'FAIL'}", flush=True)
        json.dump(dict(n_gpus=world, parity=results, scatter=flat,
                       ok=all(r["ok"] for r in results) and all(x["ok"] for x in flat)), open(a.out, "w"), indent=1)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)  # captured graphs hold the communicators: skip the blocking tear-down


if __name__ == "__main__":
    main()
