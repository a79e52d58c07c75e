"""Probe: is torch symmetric memory (CUDA P2P over NVLink) usable on this box?"""
import faulthandler
import os
import sys
import time

faulthandler.dump_traceback_later(70, exit=True)
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t0 = time.time()
buf = symm.empty(1 << 20, dtype=torch.float64, device=f"cuda:{lr}")
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(f"r{rank}: rendezvous ok in {time.time()-t0:.2f}s; multicast={hdl.has_multicast_support}; ptrs={[hex(p) for p in hdl.buffer_ptrs]}", flush=True)
buf.fill_(float(rank))
hdl.barrier(channel=0)
peer = (rank + 1) % world
pb = hdl.get_buffer(peer, (1 << 20,), torch.float64)
mine = torch.full((1024,), 100.0 + rank, dtype=torch.float64, device="cuda")
pb[:1024].copy_(mine)          # P2P store into the peer
hdl.barrier(channel=0)
torch.cuda.synchronize()
print(f"r{rank}: buf[0]={buf[0].item()} (expect {100.0 + (rank - 1) % world}) buf[2000]={buf[2000].item()}", flush=True)
# graph capture of the barrier + a p2p copy
g = torch.cuda.CUDAGraph()
try:
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        pb[:1024].copy_(mine); hdl.barrier(channel=1)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        pb[:1024].copy_(mine + 1.0)
        hdl.barrier(channel=1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(f"r{rank}: graph replay ok buf[0]={buf[0].item()}", flush=True)
except Exception as e:
    print(f"r{rank}: graph capture failed: {e!r}", flush=True)
# time barrier
torch.cuda.synchronize(); dist.barrier()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    hdl.barrier(channel=0)
e1.record(); torch.cuda.synchronize()
print(f"r{rank}: symm barrier {e0.elapsed_time(e1)*10:.1f} us each", flush=True)
sys.stdout.flush()
os._exit(0)
