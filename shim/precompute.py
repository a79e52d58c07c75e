"""Drop-in for /root/reference/cuda/precompute.py: put this directory first on sys.path and the
reference scripts' `from precompute import ...` resolve to the B200 path."""
from fenicsx_fus_gpu_b200.precompute import *  # noqa: F401,F403
