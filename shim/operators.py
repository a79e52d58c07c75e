"""Drop-in for /root/reference/cuda/operators.py: put this directory first on sys.path and the
reference scripts' `from operators import ...` resolve to the B200 path."""
from fenicsx_fus_gpu_b200.operators import *  # noqa: F401,F403
