"""A/B of the stage-closing kernel (fus_rk_close_f64) at the bench size: plain device memory vs
a symmetric (peer-mappable VMM) arena, with and without the shared-dof skip mask."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fenicsx_fus_gpu_b200 import _lib  # noqa: E402

n = 33076161
torch.cuda.set_device(0)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29577")
import torch.distributed as dist  # noqa: E402

dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
import torch.distributed._symmetric_memory as symm  # noqa: E402


def vecs(kind):
    if kind == "torch":
        return [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(9)]
    slot = (n * 8 + 255) // 256 * 256
    if kind.startswith("plain"):  # one cudaMalloc'ed block carved like the arena
        arena = torch.zeros(9 * slot + (1 << 22), dtype=torch.uint8, device="cuda")
    else:
        arena = symm.empty(9 * slot + (1 << 22), dtype=torch.uint8, device=torch.device("cuda", 0))
        symm.rendezvous(arena, dist.group.WORLD)
        arena.zero_()
    skew = (lambda i: i * 37 * 1024) if kind.endswith("skew") else (lambda i: 0)
    return [arena[i * slot + skew(i):i * slot + skew(i) + n * 8].view(torch.float64) for i in range(9)]


mask0 = torch.zeros((n + 7) // 8 + 16, dtype=torch.uint8, device="cuda")
maskp = mask0.clone()
maskp[::401] = 0xF0  # a sprinkling of skipped groups
f = _lib.fn("fus_rk_close", np.float64)
st = torch.cuda.current_stream().cuda_stream
for kind in ("torch", "plain", "plain-skew", "symm", "symm-skew"):
    u, v, u0, v0, ku, un, b, m, _ = vecs(kind)
    m.fill_(1.0)
    print(kind, [hex(t.data_ptr() % (1 << 21)) for t in (u, v, u0, v0, ku, un, b, m)], flush=True)
    for label, mk in (("no mask", None), ("sparse mask", maskp.data_ptr())):
        for mode in (3, 1, 4):
            def go():
                rc = f(u.data_ptr(), v.data_ptr(), u0.data_ptr(), v0.data_ptr(), ku.data_ptr(), None, un.data_ptr(),
                       b.data_ptr(), m.data_ptr(), 0.1, 0.05, mode, n, None, mk, st)
                assert rc == 0
            for _ in range(3):
                go()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                go()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 20
            passes = {3: 9, 1: 12, 4: 8}[mode]
            print(f"{kind:6s} {label:12s} mode {mode}: {t * 1e3:7.1f} us  {passes * n * 8 / t / 1e6:7.1f} GB/s", flush=True)
os._exit(0)
