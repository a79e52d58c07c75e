"""Drop-in for /root/reference/cuda/utils.py: put this directory first on sys.path and the
reference scripts' `from utils import ...` resolve to the B200 path."""
from fenicsx_fus_gpu_b200.utils import *  # noqa: F401,F403
