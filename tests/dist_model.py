"""CPU model of the boundary exchange of the distributed path (k_dist_edges in csrc/dist.cuh, the virtual-rank merge of run_dist2 in csrc/dist2_host.cuh).

The suffix array is cut into G rank ranges at boundaries whose LCP is < K (what the bucket partition of the
GPU path guarantees).  Every range publishes its two edge staircases; every range appends its neighbours'
staircases as virtual ranks and evaluates the factor rule on [virtual | real | virtual] with the ordinary
stage-3 model.  The merge below mirrors the host code of run_dist2 line by line (the model keeps S-positions where the device keeps the equivalent leaf values F0 / R0)."""
import gpu_algorithm_model as gm

NONE_MIN = gm.NONE_MIN


def _fval(s, rc, N):
    return (s if s < N else NONE_MIN) if rc else s


def _rval(s, N):
    return s if N < s <= 2 * N else 0


def edge_staircases(SA, LCP, lo, hi, K, rc, N):
    """k_dist_edges, by definition.  Real ranks [lo, hi); returns out[side][v] = (count, Fmin, Rmax, whole), v = 1..K."""
    m = hi - lo
    lcp = [0] + [LCP[lo + t] for t in range(1, m)] + [0]          # local LCP with guards at both ends
    out = [dict(), dict()]

    def agg(a, b):
        F, R = NONE_MIN, 0
        for t in range(a, b + 1):
            F = min(F, _fval(SA[lo + t], rc, N))
            if rc:
                R = max(R, _rval(SA[lo + t], N))
        return F, R

    def q_of(v):                                                    # largest index <= m-1 with lcp < v
        t = m - 1
        while lcp[t] >= v:
            t -= 1
        return t

    def p_of(v):                                                    # smallest index >= 1 with lcp < v (m if m == 1)
        if m == 1:
            return m
        t = 1
        while lcp[t] >= v:
            t += 1
        return t

    for v in range(1, K + 1):
        qv, qn = q_of(v), (m if v == K else q_of(v + 1))
        a, b = qv, qn - 1
        F, R = agg(a, b) if a <= b else (NONE_MIN, 0)
        out[0][v] = (max(0, b - a + 1), F, R, qv == 0)
        pv, pn = p_of(v), (0 if v == K else p_of(v + 1))
        a, b = pn, pv - 1
        F, R = agg(a, b) if a <= b else (NONE_MIN, 0)
        out[1][v] = (max(0, b - a + 1), F, R, pv == m)
    return out


def virtual_ranks(me, bounds, c0, edges, K):
    """Host merge of run_dist2: (left SA, left LCP, right SA, right LCP incl. guard) for range `me`."""
    G = len(bounds) - 1
    M = [bounds[g + 1] - bounds[g] for g in range(G)]

    def minint(g):
        r = 0
        for v in range(1, K + 1):
            if edges[g][0][v][3]:
                r = v
        return r

    def prev_ne(g):
        g -= 1
        while g >= 0 and not M[g]:
            g -= 1
        return g

    def next_ne(g):
        g += 1
        while g < G and not M[g]:
            g += 1
        return g if g < G else -1

    def merge(blk, g, side, cur):
        for v in range(1, K + 1):
            cnt, F, R, _ = edges[g][side][v]
            if not cnt:
                continue
            eff = min(v, cur)
            b = blk[eff]
            b[0] += cnt
            b[1] = min(b[1], F)
            b[2] = max(b[2], R)

    left = {e: [0, NONE_MIN, 0] for e in range(K + 1)}
    right = {e: [0, NONE_MIN, 0] for e in range(K + 1)}
    cur = min(c0[me], K)
    g = prev_ne(me)
    while g >= 0 and cur > 0:
        merge(left, g, 0, cur)
        cur = min(cur, minint(g), c0[g])
        g = prev_ne(g)
    g = next_ne(me)
    cur = min(c0[g], K) if g >= 0 else 0
    while g >= 0 and cur > 0:
        merge(right, g, 1, cur)
        cur = min(cur, minint(g))
        g = next_ne(g)
        if g >= 0:
            cur = min(cur, c0[g])

    def reps(b):
        out = []
        if b[1] != NONE_MIN:
            out.append(b[1])
        if b[2] != 0:
            out.append(b[2])
        return out or [NONE_MIN]

    sl, ll, prev_eff = [], [], 0
    for eff in range(1, K + 1):
        if not left[eff][0]:
            continue
        for j, rep in enumerate(reps(left[eff])):
            sl.append(rep)
            ll.append(prev_eff if j == 0 else eff)
        prev_eff = eff
    sr, lr = [], []
    for eff in range(K, 0, -1):
        if not right[eff][0]:
            continue
        for rep in reps(right[eff]):
            sr.append(rep)
            lr.append(eff)
    lr.append(0)                                                    # right guard
    return sl, ll, sr, lr


def cut_points(LCP, n1, G, K):
    """G rank ranges whose boundaries all have LCP < K (ranges may be empty)."""
    bounds = [0]
    for g in range(1, G):
        r = max(bounds[-1], n1 * g // G)
        while r < n1 and LCP[r] >= K:
            r += 1
        bounds.append(r)
    bounds.append(n1)
    return bounds


def factorize_dist_model(data: bytes, mode: str, G: int, K: int = 3, chunk=1024, k_lin=gm.K_LIN, bounds=None):
    data = bytes(data)
    if mode == "general":
        if not data:
            return []
        rc, N, nfac = False, 0, len(data)
    else:
        N = len(data) // 2 - 1
        rc, nfac = True, N
    n1 = len(data) + 1
    SA, RANK, _ = gm.suffix_array(data)
    LCP = gm.lcp_array(data, SA, RANK)
    bounds = bounds or cut_points(LCP, n1, G, K)
    G = len(bounds) - 1
    c0 = [LCP[bounds[g]] if bounds[g + 1] > bounds[g] else 0 for g in range(G)]
    assert all(c < K for c in c0)
    edges = [edge_staircases(SA, LCP, bounds[g], bounds[g + 1], K, rc, N) if bounds[g + 1] > bounds[g] else
             [{v: (0, NONE_MIN, 0, True) for v in range(1, K + 1)}] * 2 for g in range(G)]
    LR = [None] * nfac
    for me in range(G):
        lo, hi = bounds[me], bounds[me + 1]
        if hi == lo:
            continue
        sl, ll, sr, lr = virtual_ranks(me, bounds, c0, edges, K)
        SAx = sl + SA[lo:hi] + sr
        LCPx = ll + [c0[me]] + [LCP[r] for r in range(lo + 1, hi)] + lr
        if ll:
            LCPx[0] = 0
        T = gm.Trees(LCPx, SAx, rc, N)
        # round-2 stage 3 (tabulated climb) and the search-based model of round 1 must agree on the virtual-rank arrays too
        part = gm.walk_tables(T, len(SAx), nfac, None, k_lin, 16, real=(len(sl), len(sl) + hi - lo))
        assert part == gm.walk(T, len(SAx), nfac, None, k_lin, 16, real=(len(sl), len(sl) + hi - lo))
        for i, v in enumerate(part):
            if v is not None:
                assert LR[i] is None
                LR[i] = v
    assert all(v is not None for v in LR)
    return gm.chain(LR, nfac, 0, rc, chunk)
