"""The multi-GPU path on >= 2 REAL GPUs (skipped on a single-GPU box): the peer-memory halo
(fus_halo_* handle: cross-GPU stores / loads, epoch flags, interior / interface overlap) under
CUDA-graph replay, the NCCL halo and the reference-shaped scatter factories, against the
single-GPU solve and the reference-generated scatter fixture.  Runs tests/mgpu_worker.py
under torchrun.

    gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu -q
"""

import json
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitioned_solve_on_real_gpus_matches_single_gpu(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    n = 2 if torch.cuda.device_count() < 4 else 4
    out = tmp_path / "mgpu.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py"),
           "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = json.load(open(out))
    bad = [c for c in res["parity"] if not c["ok"]] + [c for c in res["scatter"] if not c["ok"]]
    assert not bad, bad
    assert res["ok"] and len(res["parity"]) >= 10
