"""Mint large-segment herding KATs from the reference's OWN ExemplarGenerator.herding (util.py:401-434, unmodified):
YOOCHOOSE has labels with up to 3 342 candidates, far beyond the 40 small cases of herding.npz.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_herding_big.py

Only seeds, sizes and the reference's picks are stored (herding_big.npz, a few KB): the test regenerates the float32
reps from the seed with the same NumPy calls.  Per SURVEY A.10 float32 herding on n >~ 1000 candidates can flip at a
near-tie of the arg-max (top-1 / top-2 gap 2e-7 .. 8e-6); `safe_prefix` is the number of leading picks before the first
step whose gap in a float64 replay is below 1e-5 -- the prefix every faithful implementation must reproduce."""
import math
import os
import sys
from collections import defaultdict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refstub  # noqa: E402

MAXLEN = 50
CASES = [(101, 400, 200), (102, 1000, 500), (103, 1500, 750), (104, 2799, 900), (105, 3100, 1000), (106, 640, 640)]


def make_rep(seed, n, d=150):
    rng = np.random.RandomState(seed)
    center = rng.randn(d).astype(np.float32)
    return (center[None] + 0.6 * rng.randn(n, d)).astype(np.float32)


def safe_prefix(rep, m, picks):
    D = (rep.T / np.linalg.norm(rep.T, axis=0)).astype(np.float64)
    mu = D.mean(axis=1); w = mu.copy(); sel = []
    for _ in range(int(math.ceil(1.1 * m))):
        s = w @ D
        t = np.partition(s, -2)[-2:]
        if abs(t[1] - t[0]) < 1e-5:
            return len(sel)
        i = int(np.argmax(s))
        w = w + mu - D[:, i]
        if i not in sel:
            sel.append(i)
        if len(sel) == m:
            break
    return len(sel)


def main():
    util = refstub.load_reference_util()
    gen = util.ExemplarGenerator.__new__(util.ExemplarGenerator)
    gen.exemplars = defaultdict(list)
    seeds, ns, ms, offs, allp, safe = [], [], [], [0], [], []
    for ci, (seed, n, m) in enumerate(CASES):
        rep = make_rep(seed, n)
        seq = np.zeros((n, MAXLEN + 1), np.int64)
        seq[:, -1] = 3
        seq[:, -2] = np.arange(n) % 30000 + 1
        seq[:, -3] = np.arange(n) // 30000 + 1              # (seq[-3] - 1) * 30000 + seq[-2] - 1 = candidate index
        logits = np.zeros((n, 1), np.float32)
        saved = gen.herding(rep, logits, seq, ci, m)
        chosen = [(int(e[0][-3]) - 1) * 30000 + int(e[0][-2]) - 1 for e in gen.exemplars[ci]]
        assert saved == len(chosen)
        seeds.append(seed); ns.append(n); ms.append(m)
        allp += chosen; offs.append(offs[-1] + len(chosen))
        safe.append(safe_prefix(rep, m, chosen))
        print("case n=%d m=%d: %d picks, safe prefix %d" % (n, m, len(chosen), safe[-1]))
    np.savez_compressed(os.path.join(HERE, "herding_big.npz"), seed=np.array(seeds), n=np.array(ns), m=np.array(ms),
                        picks=np.array(allp, np.int32), pick_off=np.array(offs, np.int64), safe_prefix=np.array(safe))


if __name__ == "__main__":
    main()
