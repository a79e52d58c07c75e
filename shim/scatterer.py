"""Drop-in for /root/reference/cuda/scatterer.py: put this directory first on sys.path and the
reference scripts' `from scatterer import ...` resolve to the B200 path."""
from fenicsx_fus_gpu_b200.scatterer import *  # noqa: F401,F403
