"""Developer tool: SASS evidence for the hot kernels of libfus_b200.so (kept under profiles/).

    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt

Per kernel: a histogram of the memory / synchronisation / FP64 mnemonics and the lines that show
the Blackwell-native data movement (UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS.*TRANS64 =
mbarrier complete_tx), the fire-and-forget reductions (REDG) and the cross-GPU signalling
(LDG/STG .STRONG.SYS, MEMBAR.SC.SYS, CCTL.IVALL)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fenicsx_fus_gpu_b200", "libfus_b200.so")
KERNELS = [
    ("stiffness_kernelIdLi5ELi0ELb1ELi0ELb0E", "stiffness_kernel<double, n=5, MODE 0, ATOMIC, GEO 0 (G streamed), no halo wait>: the headline kernel"),
    ("stiffness_kernelIdLi5ELi1ELb1ELi0ELb0E", "stiffness_kernel<double, 5, MODE 1 (dual: Westervelt stage)>"),
    ("stiffness_kernelIfLi5ELi0ELb1ELi0ELb0E", "stiffness_kernel<float, 5, MODE 0>"),
    ("stiffness_kernelIdLi5ELi0ELb1ELi2ELb0E", "stiffness_kernel<double, 5, GEO 2 (rectilinear cells)>"),
    ("stiffness_kernelIdLi5ELi0ELb1ELi0ELb1E", "stiffness_kernel<double, 5, WAIT = true (in-kernel halo wait, split_mode 'fused')>"),
    ("rk_close_kernelIdLb1ELi0E", "rk_close_kernel<double, VEC, linear>"),
    ("rk_close_shared_kernelIdLi0E", "rk_close_shared_kernel<double, linear>: gather ghost sums + close shared dofs + put + FWD signal"),
    ("halo_wait_kernel", "halo_wait_kernel: one-warp ld.acquire.sys spin"),
    ("halo_put_kernel", "halo_put_kernel<double>: owner values -> peers' ghost slots + FWD signal"),
    ("boundary_kernelIdLb1E", "boundary_kernel<double, SIGNAL>: boundary terms + REV signal"),
    ("mass4_kernelIdLb0E", "mass4_kernel<double>"),
]
PAT = re.compile(r"\b(UBLKCP[A-Z0-9.]*|SYNCS[A-Z0-9.]*|REDG?[A-Z0-9.]*|ATOMG?[A-Z0-9.]*|DFMA|DMUL|DADD|FFMA|LDS[A-Z0-9.]*|STS[A-Z0-9.]*|LDG[A-Z0-9.]*|STG[A-Z0-9.]*|"
                 r"BAR[A-Z0-9.]*|MEMBAR[A-Z0-9.]*|CCTL[A-Z0-9.]*|FENCE[A-Z0-9.]*|NANOSLEEP|LDC[A-Z0-9.]*|UTC[A-Z0-9.]*MMA[A-Z0-9.]*)\b")
SHOW = re.compile(r"UBLKCP|SYNCS|REDG|STRONG\.SYS|MEMBAR\.SC\.SYS|CCTL|NANOSLEEP")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(ln)
    print(f"# SASS excerpt of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; sm_100a)\n")
    print("# no UTC*MMA (tcgen05) anywhere: the contractions are n x n with n <= 8 in FP64/FP32, see DESIGN.md 3.1\n")
    for key, title in KERNELS:
        hit = [f for f in funcs if key in f]
        if not hit:
            print(f"## {title}\n(not found: {key})\n")
            continue
        body = funcs[hit[0]]
        hist = {}
        for ln in body:
            for op in PAT.findall(ln):
                hist[op] = hist.get(op, 0) + 1
        print(f"## {title}\n{hit[0]}\n{len(body)} SASS lines")
        print("   " + "  ".join(f"{k} x{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])))
        shown = 0
        for ln in body:
            if SHOW.search(ln) and shown < 14:
                print("   " + re.sub(r"/\*[0-9a-fx]+\*/\s*$", "", ln).strip())
                shown += 1
        print()


if __name__ == "__main__":
    sys.exit(main())
