"""Development tool (CPU): where do the lanes of k_node_tables / k_lpnf_rank spend their steps?

Builds SA / LCP of a scaled-down copy of the headline text with the oracle, replays the per-rank work of the two stage-3
kernels in numpy / Python (interval extents, ancestor climbs, rc hops with and without the pruning bound) and reports
per-lane means against per-warp maxima -- the quantity a SIMT machine pays.  Not part of the product or the tests."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np

import oracle_py as orc
from nolzss_b200 import workloads as wl

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
scale = 50.0 * n / 250e6
x = wl.planted_dna_big(n, 4, scale)
S = wl.prepare_w_rc_single(x.tobytes())
N = len(S) // 2 - 1
t0 = time.time()
sa, lcp = orc.gpu_order_sa_lcp(S)
sa = np.asarray(sa, dtype=np.int64)
lcp = np.asarray(lcp, dtype=np.int64)
n1 = len(sa)
print(f"n={n} N={N} n1={n1} lcp entries={len(lcp)} sa/lcp {time.time() - t0:.1f} s")
NONE = 1 << 40
F0 = np.where(sa < N, sa, NONE)
R0 = np.where((sa > N) & (sa <= 2 * N), sa - N, 0)
LCP = np.zeros(n1 + 1, dtype=np.int64)
LCP[: len(lcp)] = lcp[: n1 + 1]
LCP[0] = 0
LCP[n1] = 0

# PSV / NSV (strictly smaller) by stacks
t0 = time.time()
psv = np.zeros(n1 + 1, dtype=np.int64)
nsv = np.full(n1 + 1, n1, dtype=np.int64)
st = []
L = LCP.tolist()
for k in range(n1 + 1):
    v = L[k]
    while st and L[st[-1]] >= v:
        st.pop()
    psv[k] = st[-1] if st else 0
    st.append(k)
st = []
for k in range(n1, -1, -1):
    v = L[k]
    while st and L[st[-1]] >= v:
        st.pop()
    nsv[k] = st[-1] if st else n1
    st.append(k)
print(f"psv/nsv {time.time() - t0:.1f} s")
ks = np.arange(n1 + 1)
named = LCP > 0
extL = (ks - psv)[named]          # steps of the left scan until the boundary (>= 1)
extR = (nsv - ks)[named]
for lim in (8, 12, 16, 24, 32):
    slow = (extL > lim) | (extR > lim)
    full = np.zeros(n1 + 1, dtype=bool)
    full[named] = slow
    pad = (-len(full)) % 32
    f = np.concatenate([full, np.zeros(pad, dtype=bool)])
    t8 = f.reshape(-1, 8).sum(1)
    w32 = t8.reshape(-1, 4)
    print(f"node_tables scan limit {lim}: slow lanes {slow.mean():.3f}; per 8-tile mean {t8.mean():.2f}; per warp: max-tile mean {w32.max(1).mean():.2f}, "
          f"sum mean {w32.sum(1).mean():.2f}, warps with none {np.mean(w32.sum(1) == 0):.3f}")
sz = (nsv - psv)[named]
print("interval size percentiles (named ranks):", np.percentile(sz, [50, 90, 99, 99.9, 100]).astype(int))
# scalar cost proxy of the present kernel: lanes that leave the 12-step scan pay ~32 * levels entries
lev = np.ceil(np.log(np.maximum(sz, 2)) / np.log(32))
print("levels percentiles:", np.percentile(lev, [50, 90, 99, 100]))

# ---- node table: parent + fmin (sparse table range-min over F0)
t0 = time.time()
tbl = [F0.copy()]
j = 1
while (1 << j) <= n1:
    p = tbl[-1]
    h = 1 << (j - 1)
    tbl.append(np.minimum(p[: len(p) - h], p[h:]))
    j += 1


def range_min(a, b):          # inclusive, vectors
    ln = b - a + 1
    k = np.floor(np.log2(ln)).astype(np.int64)
    out = np.empty(len(a), dtype=np.int64)
    for kk in np.unique(k):
        m = k == kk
        t = tbl[kk]
        out[m] = np.minimum(t[a[m]], t[b[m] - (1 << kk) + 1])
    return out


node_parent = np.arange(n1 + 1)
node_f = np.full(n1 + 1, NONE)
node_d = LCP.copy()
kk = ks[named]
a_, b_ = psv[named], nsv[named]
node_f[kk] = range_min(a_, b_ - 1)
node_parent[kk] = np.where(LCP[a_] >= LCP[b_], a_, b_)
print(f"node table {time.time() - t0:.1f} s")

# nearest rc rank left/right with min LCP on the way
isr = R0 > 0
PR = np.where(isr, np.arange(n1), -1)
PR = np.maximum.accumulate(PR)
NR = np.where(isr, np.arange(n1), n1 + 5)
NR = np.minimum.accumulate(NR[::-1])[::-1]

fw = np.nonzero(F0 < N)[0]
fw = fw[(F0[fw] < N)]
print("forward ranks:", len(fw))
rng = np.random.default_rng(1)
# sample warps: 32 consecutive list items
nw = min(4000, len(fw) // 32)
starts = rng.choice(len(fw) // 32, nw, replace=False) * 32
Ll = LCP
P_, F_, D_ = node_parent.tolist(), node_f.tolist(), node_d.tolist()
R0l = R0.tolist()
F0l = F0.tolist()
PRl, NRl = PR.tolist(), NR.tolist()
Lc = LCP.tolist()


def hops(r, thr, stop, direction, cap=10**9):
    """returns (hops, result) of rc_side_depth with pruning bound `stop` (0 = none)"""
    k = r
    run = NONE
    h = 0
    while True:
        if direction == 0:
            if k == 0:
                return h, 0
            nb = k - 1
            tgt = PRl[nb]
            if tgt < 0:
                return h, 0
            run = min(run, min(Lc[tgt + 1: k + 1]))
            k = tgt
        else:
            if k + 1 >= n1:
                return h, 0
            nb = k + 1
            tgt = NRl[nb]
            if tgt >= n1:
                return h, 0
            run = min(run, min(Lc[k + 1: tgt + 1]))
            k = tgt
        h += 1
        if run <= stop:
            return h, 0
        if R0l[k] > thr:
            return h, run
        if h >= cap:
            return h, -1


stats = {k: [] for k in ("climb", "hopL0", "hopR0", "hopL1", "hopR1", "rcwin", "rcext")}
wstats = {k: [] for k in stats}
t0 = time.time()
for s0 in starts:
    row = {k: [] for k in stats}
    for r in fw[s0: s0 + 32].tolist():
        i = F0l[r]
        k = r if Lc[r] >= Lc[r + 1] else r + 1
        steps = 0
        have_f = False
        childF = i
        dF = jF = 0
        belowF = i
        while True:
            d = D_[k]
            if d == 0:
                break
            if steps == 64:
                break
            steps += 1
            m = F_[k]
            if m != NONE and m + d <= i:
                have_f = True; dF = d; jF = m; belowF = childF
                break
            childF = m
            k = P_[k]
        fwd_len = ((i - jF) if belowF == jF else dF) if have_f else 0
        thr = N - i
        hl0, dl = hops(r, thr, 0, 0, 200)
        hr0, dr = hops(r, thr, 0, 1, 200)
        stop = max(fwd_len, 1)
        hl1, dl1 = hops(r, thr, stop, 0, 200)
        hr1, dr1 = hops(r, thr, max(stop, dl1 if dl1 > 0 else 0), 1, 200)
        dR = max(dl, dr)
        rcwin = (dR > fwd_len) if have_f else (dR > 1)
        ext = 0
        if rcwin:
            lo = r
            while Lc[lo] >= dR:
                lo -= 1
            hi = r
            while Lc[hi + 1] >= dR:
                hi += 1
            ext = max(r - lo, hi - r)
        for kname, v in (("climb", steps), ("hopL0", hl0), ("hopR0", hr0), ("hopL1", hl1), ("hopR1", hr1), ("rcwin", int(rcwin)), ("rcext", ext)):
            row[kname].append(v)
    for kname in stats:
        stats[kname].extend(row[kname])
        wstats[kname].append(max(row[kname]))
print(f"walk sample {time.time() - t0:.1f} s over {nw} warps")
for kname in stats:
    v = np.array(stats[kname]); w = np.array(wstats[kname])
    print(f"{kname:6s} lane mean {v.mean():8.2f}  p50 {np.percentile(v, 50):6.0f} p99 {np.percentile(v, 99):6.0f}   per-warp max: mean {w.mean():8.2f} p50 {np.percentile(w, 50):6.0f} p90 {np.percentile(w, 90):6.0f}")
for name in ("hopL0", "hopR0", "hopL1", "hopR1"):
    w = np.array(wstats[name])
    print(f"{name}: warps whose max exceeds 24 hops: {np.mean(w > 24):.3f}; lanes: {np.mean(np.array(stats[name]) > 24):.4f}")
