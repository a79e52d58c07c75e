"""Every script under demos/ runs for two steps on one GPU (the reference's demos are its
user-facing entry points: cuda/demo_*.py, cuda/time_operators.py)."""

import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DEMOS = [
    ("demo_linear_box.py", ["--cells", "6", "--steps", "2"]),
    ("demo_linear_box.py", ["--cells", "6", "--steps", "2", "--integrator", "leapfrog"]),
    ("demo_linear_piston.py", ["--cells", "5", "--steps", "2"]),
    ("demo_nonlinear_bowl.py", ["--cells", "5", "--steps", "2"]),
    ("demo_nonlinear_box.py", ["--cells", "4", "--steps", "2"]),
    ("time_operators.py", ["--N", "8"]),
]


@pytest.mark.parametrize("script,args", DEMOS)
def test_demo_runs(script, args, tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cmd = [sys.executable, os.path.join(ROOT, "demos", script), *args]
    if script.startswith("demo_linear_piston") or script.startswith("demo_nonlinear_bowl"):
        cmd += ["--sample-dir", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "nan" not in r.stdout.lower()
