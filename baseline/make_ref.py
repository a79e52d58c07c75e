"""Recipe for ``baseline/_ref/`` - the UNMODIFIED reference modules the baseline arms run.

    python baseline/make_ref.py

``/root/reference`` has no packaging (no setup.py / pyproject.toml: ``pip install
/root/reference`` has nothing to install) and exists only in the build container, so the
modules the CPU arm (``bench.py --impl reference``: numba-cpu operators) and the GPU-vs-GPU
arm (``--impl reference-cuda``: the reference's own Numba-CUDA kernels) import are placed,
byte for byte, under ``baseline/_ref/`` - git-ignored (never in history), not
gpurun-ignored (travels to the GPU box).  ``MANIFEST.json`` records the sha256 of every file
so ``baseline/cpu_arm.py`` can state that what it timed is the reference's own code.
"""

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FUS_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = {
    "numba_cpu": ["numba-cpu/operators.py", "numba-cpu/sum_factorisation.py", "numba-cpu/precompute.py"],
    "cuda": ["cuda/operators.py"],
}


def main() -> int:
    if not os.path.isdir(REF):
        print(f"make_ref: {REF} not present (GPU box): keeping {DEST} as shipped")
        return 0 if os.path.exists(os.path.join(DEST, "MANIFEST.json")) else 1
    manifest = {}
    for sub, files in FILES.items():
        os.makedirs(os.path.join(DEST, sub), exist_ok=True)
        for rel in files:
            dst = os.path.join(DEST, sub, os.path.basename(rel))
            shutil.copyfile(os.path.join(REF, rel), dst)
            manifest[os.path.join(sub, os.path.basename(rel))] = {
                "source": rel, "sha256": hashlib.sha256(open(dst, "rb").read()).hexdigest()}
    json.dump(manifest, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    print(f"make_ref: {len(manifest)} reference modules -> {DEST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
