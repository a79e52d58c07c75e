"""Developer tool: brute-force the shared-memory tile strides of the stiffness
kernel that minimise bank-conflict wavefronts.

Tile address (in elements of T):  addr(c,i,j,k) = c*Sc + i*Sp + j*Sj + k.
Thread tid = c*n^2 + a*n + b plays three roles:
  P1: (j,k) = (a,b), one access per i        (x pencils, transform, final)
  P2: (i,k) = (a,b) or (b,a), one access per j (y pencils)
  P3: (i,j) = (a,b) or (b,a), one access per k (z pencils)
Wavefront model: 32 banks x 4 B; 4-byte accesses are resolved per warp,
8-byte per half-warp, 16-byte per quarter-warp; lanes reading the same word
broadcast; cost = max over banks of distinct words.
"""

import itertools
import sys


def wavefronts(addrs_bytes, size):
    """addrs_bytes: list of (lane, byte address) for active lanes of one warp."""
    group = {4: 32, 8: 16, 16: 8}[size]
    total = 0
    for g0 in range(0, 32, group):
        banks = {}
        for lane, ad in addrs_bytes:
            if g0 <= lane < g0 + group:
                for w in range(ad // 4, (ad + size) // 4):
                    banks.setdefault(w % 32, set()).add(w)
        if banks:
            total += max(len(v) for v in banks.values())
    return total


def pattern_cost(n, s, B, threads, Sc, Sp, Sj, p2swap, p3swap):
    N2 = n * n
    cost = {"P1": 0, "P2": 0, "P3": 0}
    nwarps = (B * N2 + 31) // 32
    for w in range(nwarps):
        lanes = [(l, w * 32 + l) for l in range(32) if w * 32 + l < B * N2]
        for m in range(n):
            a1, a2, a3 = [], [], []
            for lane, tid in lanes:
                c, t2 = divmod(tid, N2)
                a, b = divmod(t2, n)
                a1.append((lane, s * (c * Sc + m * Sp + a * Sj + b)))
                i, k = (b, a) if p2swap else (a, b)
                a2.append((lane, s * (c * Sc + i * Sp + m * Sj + k)))
                i, j = (b, a) if p3swap else (a, b)
                a3.append((lane, s * (c * Sc + i * Sp + j * Sj + m)))
            cost["P1"] += wavefronts(a1, s)
            cost["P2"] += wavefronts(a2, s)
            cost["P3"] += wavefronts(a3, s)
    return cost


def ideal(n, s, B):
    nw = (B * n * n + 31) // 32
    return nw * n * (s // 4)


def search(n, s, B):
    best = None
    for Sj in (n, n + 1):
        for Sp in range(n * Sj, n * Sj + 17):
            for Sc in range(n * Sp, n * Sp + 33):
                for p2, p3 in itertools.product((0, 1), (0, 1)):
                    c = pattern_cost(n, s, B, 0, Sc, Sp, Sj, p2, p3)
                    # weights: P1 x7 accesses, P2 x4, P3 x4 per stage of the algorithm
                    tot = 7 * c["P1"] + 4 * c["P2"] + 4 * c["P3"]
                    key = (tot, Sc)
                    if best is None or key < best[0]:
                        best = (key, dict(Sj=Sj, Sp=Sp, Sc=Sc, p2swap=p2, p3swap=p3, **c))
    return best


if __name__ == "__main__":
    cfgs = {(3, 8): 14, (4, 8): 8, (5, 8): 5, (6, 8): 3, (7, 8): 2, (8, 8): 2,
            (3, 4): 14, (4, 4): 8, (5, 4): 5, (6, 4): 3, (7, 4): 2, (8, 4): 2}
    for (n, s), B in cfgs.items():
        if len(sys.argv) > 1 and int(sys.argv[1]) != n:
            continue
        key, b = search(n, s, B)
        idl = ideal(n, s, B)
        base = pattern_cost(n, s, B, 0, n**3, n * n, n, 0, 0)
        print(f"n={n} s={s} B={B}: ideal/pattern={idl}  unpadded={base}  best={b}")
