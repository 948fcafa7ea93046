"""Elimination-tree statistics of the batch (throughput) preset: front-size histogram, update-block traffic per size
class, subtree coverage. Host only (jgb_selfcheck_tree). Usage: python scripts/tree_stats.py [case] [preset]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jgb200  # noqa: E402
from jgb200._lib import ptr  # noqa: E402
from oracle import nr as onr  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import oracle_system  # noqa: E402


def tree(case, preset=2):
    o = onr.newton_raphson(oracle_system(case))
    n = o.dim
    grp = np.zeros(n, dtype=np.int64)
    for i in range(o.mdl.n):
        if o.pvpq[i] >= 0:
            grp[o.pvpq[i]] = i
        if o.pq[i] >= 0:
            grp[o.pq[i]] = i
    cp, rv = (o.j_colptr + 1).astype(np.int64), (o.j_rowval + 1).astype(np.int64)
    cap = n
    nf_ = np.zeros(1, dtype=np.int64)
    k, nf, par, nasm = (np.zeros(cap, dtype=np.int32) for _ in range(4))
    rc = jgb200.load().jgb_selfcheck_tree(n, ptr(cp, C.c_int64), ptr(rv, C.c_int64), ptr(grp, C.c_int64), preset, cap,
                                          ptr(nf_, C.c_int64), ptr(k, C.c_int32), ptr(nf, C.c_int32),
                                          ptr(par, C.c_int32), ptr(nasm, C.c_int32))
    assert rc == 0
    F = int(nf_[0])
    return k[:F].copy(), nf[:F].copy(), par[:F].copy(), nasm[:F].copy()


if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "synthetic10k"
    preset = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    k, nf, par, nasm = tree(case, preset)
    F = len(k)
    u = nf - k
    upd = u * (u + 1)
    usz = k * (nf + 1) - k * (k - 1) // 2
    flops = np.array([sum(2.0 * (n_ - p - 1) * (n_ - p) + 1 for p in range(k_)) for k_, n_ in zip(k, nf)])
    print(f"{case}: fronts {F}, u_size {usz.sum()}, upd_size {upd.sum()}, flops {flops.sum():.3g}, nasm {nasm.sum()}")
    edges = [0, 4, 8, 12, 16, 20, 24, 32, 48, 64, 96, 128, 1 << 30]
    print("class      fronts   k_sum   usz%   upd%  flops%  asm%  mean_k")
    for a, b in zip(edges[:-1], edges[1:]):
        m = (nf > a) & (nf <= b)
        if m.any():
            print(f"{a + 1:4d}-{min(b, 9999):4d} {m.sum():7d} {k[m].sum():7d} {100 * usz[m].sum() / usz.sum():6.1f} "
                  f"{100 * upd[m].sum() / upd.sum():6.1f} {100 * flops[m].sum() / flops.sum():6.1f} "
                  f"{100 * nasm[m].sum() / nasm.sum():5.1f} {k[m].mean():6.2f}")
    # height / levels
    h = np.zeros(F, dtype=int)
    for f in range(F):
        if par[f] >= 0:
            h[par[f]] = max(h[par[f]], h[f] + 1)
    print("levels", h.max() + 1)
    for lim in (12, 16, 20, 24, 32):
        # maximal subtrees whose fronts are all <= lim rows
        sub_ok = nf <= lim
        for f in range(F):          # children precede parents
            if par[f] >= 0 and not sub_ok[f]:
                sub_ok[par[f]] = False
        # propagate: a front is in a fused subtree iff it and all descendants ok; root of subtree = ok front whose
        # parent is not ok
        roots = [f for f in range(F) if sub_ok[f] and (par[f] < 0 or not sub_ok[par[f]])]
        inside = sub_ok
        print(f"lim {lim}: subtrees {len(roots)}, fronts {inside.sum()} ({100 * inside.mean():.1f}%), "
              f"flops {100 * flops[inside].sum() / flops.sum():.1f}%, usz {100 * usz[inside].sum() / usz.sum():.1f}%, "
              f"asm {100 * nasm[inside].sum() / nasm.sum():.1f}%, "
              f"upd kept on chip {100 * (upd[inside].sum() - upd[roots].sum()) / upd.sum():.1f}%, "
              f"upd of roots {100 * upd[roots].sum() / upd.sum():.1f}%; outside: fronts {F - inside.sum()}, levels {h[~inside].max() - h[~inside].min() + 1 if (~inside).any() else 0}")
    # subtree size distribution and stack peaks for lim 16 / 12 / 8
    ch = [[] for _ in range(F)]
    for f in range(F):
        if par[f] >= 0:
            ch[par[f]].append(f)
    for lim in (8, 12, 16):
        ok = nf <= lim
        for f in range(F):
            if par[f] >= 0 and not ok[f]:
                ok[par[f]] = False
        peak = np.zeros(F, dtype=int)
        size = np.ones(F, dtype=int)
        for f in range(F):
            if not ok[f]:
                continue
            base = 0
            pk = 0
            for c in ch[f]:
                pk = max(pk, base + peak[c])
                base += upd[c]
                size[f] += size[c]
            peak[f] = max(pk, base, upd[f])
        roots = [f for f in range(F) if ok[f] and (par[f] < 0 or not ok[par[f]])]
        sz = np.array([size[r] for r in roots])
        pk = np.array([peak[r] for r in roots])
        print(f"lim {lim}: subtrees {len(roots)} sizes: " + " ".join(f"{q}:{int(np.percentile(sz, q))}" for q in (10, 50, 90, 99, 100)),
              " stack peak: " + " ".join(f"{q}:{int(np.percentile(pk, q))}" for q in (50, 90, 99, 100)),
              f" fronts in subtrees of size>=8: {sz[sz >= 8].sum()}, size<=3: {sz[sz <= 3].sum()}")
