"""Executable, line-by-line Python model of the GPU pipeline's ALGORITHMS (not of CUDA mechanics).

Test infrastructure: mirrors nolzss_b200/csrc/{sa,lcp,lpnf,chain}.cuh closely enough that a logic
error in the design (symbol order with sentinel classes, prefix doubling with discarding, the 32-ary
summary-tree searches, the per-rank interval climb, the chunked chain extraction) shows up on the
CPU, against the oracle, before any GPU time is spent.  Kept deliberately small-scale and slow.
"""
from __future__ import annotations

RC_MASK = 1 << 63
SENT = 0xFF
NONE_MIN = 0xFFFFFFFF
LR_RC_FLAG = 0x80000000


# ------------------------------------------------------------------ stage 1: suffix array
def layout_symbols(sigma: int, b: int, R: int, D: int, n1: int) -> int:
    """api.cu layout_symbols: the smallest window whose sigma^W key space is >= 64 n', widened to fill whole 8-bit passes."""
    wmax = min(29, (64 - R - D) // b)
    space, wmin = 1.0, 0
    while wmin < wmax and space < 64.0 * n1:
        space *= max(sigma, 2)
        wmin += 1
    wmin = max(wmin, 1)
    passes = (R + wmin * b + D + 7) // 8
    return max(min((passes * 8 - R - D) // b, wmax), wmin)


def choose_layout(data: bytes, n1: int, force_bits=None, force_w=None):
    hist = [0] * 256
    for c in data:
        hist[c] += 1
    cls = [SENT] * 256
    sigma = 0
    for c in range(256):
        if hist[c] >= 2:
            cls[c] = sigma
            sigma += 1
    b = 1
    while (1 << b) < sigma:
        b += 1
    w32 = min(14, 28 // b)
    use32 = w32 >= 1 and float(max(sigma, 2)) ** w32 >= 16.0 * n1
    if force_bits == 32:
        use32 = True
    if force_bits == 64:
        use32 = False
    if use32:
        lay = dict(key_bits=32, b=b, W=w32, D=4, R=0)
    else:
        lay = dict(key_bits=64, b=b, W=layout_symbols(sigma, b, 0, 5, n1), D=5, R=0)
    if force_w:
        lay["W"] = min(force_w, lay["W"])
    lay["dshift"] = lay["key_bits"] - lay["R"] - lay["W"] * b - lay["D"]     # KeyLayout::dshift(): [R][W*b symbols][D][zeros]
    return cls, lay


def build_keys(data: bytes, cls, lay):
    L = len(data)
    n1 = L + 1
    kb, b, W, D = lay["key_bits"], lay["b"], lay["W"], lay["D"]
    none = (1 << D) - 1
    keys = []
    for p in range(n1):
        key, dist, sh = 0, none, kb - b
        for t in range(W):
            c = cls[data[p + t]] if p + t < L else SENT
            if c == SENT:
                dist = t
                break
            key |= c << sh
            sh -= b
        keys.append(key | (dist << lay["dshift"]))
    return keys


def build_keys_rolling(data: bytes, cls, lay, Q=8):
    """sa.cuh kb_tile_keys: every thread builds the keys of Q consecutive suffixes -- the first window symbol by symbol,
    the following ones by shifting a symbol out and one in; the sentinel flags of the window roll along in a bit mask
    (bit t: symbol t of the window is a sentinel) and everything from the first sentinel on is dropped from the key."""
    L = len(data)
    n1 = L + 1
    kb, b, W, D = lay["key_bits"], lay["b"], lay["W"], lay["D"]
    none = (1 << D) - 1
    symsh = kb - lay["R"] - W * b
    wmask = (1 << (W * b)) - 1
    tile = [(cls[data[j]] if j < L else SENT) for j in range(n1 + Q + W + 1)]       # positions at or past L: sentinels
    keys = [None] * n1
    for o0 in range(0, n1, Q):
        K = SM = 0
        for t in range(W):
            c = tile[o0 + t]
            if c == SENT:
                SM |= 1 << t
                c = 0
            K = (K << b) | c
        for q in range(Q):
            ks, dist = K, none
            if SM:
                t = (SM & -SM).bit_length() - 1
                dist = t
                ks = K & ~((1 << ((W - t) * b)) - 1)
            if o0 + q < n1:
                keys[o0 + q] = (ks << symsh) | (dist << lay["dshift"])
            c = tile[o0 + q + W]
            SM >>= 1
            if c == SENT:
                SM |= 1 << (W - 1)
                c = 0
            K = ((K << b) & wmask) | c
    return keys


def key_pair_lcp(a: int, b_: int, lay) -> int:
    """sa.cuh key_pair_lcp: leading symbols two keys share, cut at the first sentinel of either window."""
    kb, b, W, D, R = lay["key_bits"], lay["b"], lay["W"], lay["D"], lay["R"]
    dmask = (1 << D) - 1
    x = a ^ b_
    common = W
    if x:
        p = kb - x.bit_length()               # identical leading bits (clz)
        if p < R:
            return 0
        common = min(common, (p - R) // b)
    da, db = (a >> lay["dshift"]) & dmask, (b_ >> lay["dshift"]) & dmask
    if da != dmask:
        common = min(common, da)
    if db != dmask:
        common = min(common, db)
    return common


def seed_lcp(skeys, order, lay):
    """The INITIAL regroup's LCP seeding (k_regroup_apply, LcpSeed): every member of a tie group is left to the Kasai kernel
    (NEED by text position, None = LCP_PENDING by slot), every other slot gets the LCP its key pair gives."""
    m = len(skeys)
    dmask = ((1 << lay["D"]) - 1) << lay["dshift"]

    def head(e):
        return e == 0 or e >= m or (skeys[e] & dmask) != dmask or skeys[e] != skeys[e - 1]

    LCP = [0] * (m + 1)
    need = [0] * m
    for e in range(m):
        if not (head(e) and head(e + 1)):
            LCP[e] = None
            need[order[e]] = 1
        else:
            LCP[e] = 0 if e == 0 else key_pair_lcp(skeys[e - 1], skeys[e], lay)
    return LCP, need


def regroup(keys, vals, slots, dist_mask, SA, RANK):
    """k_regroup_*: returns the compacted (g, s, slot) lists of still-active elements."""
    m = len(keys)

    def head(e):
        if e == 0:
            return True
        if dist_mask is not None and (keys[e] & dist_mask) != dist_mask:
            return True
        return keys[e] != keys[e - 1]

    nxt = []
    hj = 0
    for e in range(m):
        h = head(e)
        hn = head(e + 1) if e + 1 < m else True
        if h:
            hj = e
        newrank = hj if slots is None else slots[hj]
        s = vals[e]
        slot = e if slots is None else slots[e]
        RANK[s] = newrank
        SA[slot] = s
        if not (h and hn):
            nxt.append((newrank, s, slot))
    return nxt


def suffix_array(data: bytes, force_bits=None, force_w=None, seed_out=None):
    """seed_out (a dict): receives "seed" = seed_lcp(...) of the initial sort, for lcp_array(seed=...)."""
    L = len(data)
    n1 = L + 1
    cls, lay = choose_layout(data, n1, force_bits, force_w)
    keys = build_keys(data, cls, lay)
    order = sorted(range(n1), key=lambda p: keys[p])          # stable LSD radix sort
    SA = [0] * n1
    RANK = [0] * n1
    if seed_out is not None:
        seed_out["seed"] = seed_lcp([keys[p] for p in order], order, lay)
        seed_out["lay"] = lay
    act = regroup([keys[p] for p in order], order, None, ((1 << lay["D"]) - 1) << lay["dshift"], SA, RANK)
    h = lay["W"]
    rounds = 0
    while act:
        rounds += 1
        ck = [((g << 32) | RANK[s + h], s) for (g, s, _) in act]
        slots = [slot for (_, _, slot) in act]
        perm = sorted(range(len(ck)), key=lambda j: ck[j][0])
        act = regroup([ck[j][0] for j in perm], [ck[j][1] for j in perm], slots, None, SA, RANK)
        h *= 2
        assert rounds < 64
    return SA, RANK, rounds


def suffix_array_hybrid(data: bytes, gcap: int = 8, out_cap: int = 16, force_bits=None, rnd=None):
    """Prefix doubling with the HYBRID rounds of csrc/big_groups.cuh + the representative ranks of tile_sort.cuh.

    The active list is split once into S (groups of <= gcap members: k_tile_sort) and B (larger groups: one CTA of
    k_group_stream each).  Modelled: the pivot partition of a big group (pivot = most frequent of 32 evenly spaced
    samples; pivot-equal block in arbitrary member order, outliers sorted), sub-groups routed to S or B by size, ranks
    as representatives (a sub-group that still holds the slot its group was named by keeps that name; the pivot block
    otherwise takes its middle slot, every other sub-group its head slot), and the redo of a round through the radix
    path when a group has more than `out_cap` outliers (lists unified, slot list sorted, ordinary regroup; one fresh
    split is tried afterwards).  Returns (SA, RANK, rounds, stats)."""
    import random as _random
    rnd = rnd or _random.Random(1)
    L = len(data)
    n1 = L + 1
    cls, lay = choose_layout(data, n1, force_bits)
    keys = build_keys(data, cls, lay)
    order = sorted(range(n1), key=lambda p: keys[p])
    SA = [0] * n1
    RANK = [0] * n1
    act = regroup([keys[p] for p in order], order, None, ((1 << lay["D"]) - 1) << lay["dshift"], SA, RANK)
    h = lay["W"]
    stats = dict(rounds=0, hybrid_rounds=0, fallbacks=0, stream_groups=0, renamed=0, kept=0)

    def groups_of(lst):                       # contiguous runs of equal rank
        out, a = [], 0
        for j in range(1, len(lst) + 1):
            if j == len(lst) or lst[j][0] != lst[a][0]:
                out.append(lst[a:j])
                a = j
        return out

    def radix_round(lst):                     # the legacy path: global sort by (rank, rank[s+h]) + regroup
        ck = [((g << 32) | RANK[s + h], s) for (g, s, _) in lst]
        slots = sorted(slot for (_, _, slot) in lst)          # ascending slots line up with ascending ranks
        perm = sorted(range(len(ck)), key=lambda j: ck[j][0])
        return regroup([ck[j][0] for j in perm], [ck[j][1] for j in perm], slots, None, SA, RANK)

    def emit(sub, base_slot, old_rank, is_pivot_block, nextS, nextB, writes):
        """one sorted sub-group occupying slots [base_slot, base_slot + len(sub))"""
        size = len(sub)
        if base_slot <= old_rank < base_slot + size:
            name = old_rank                                    # still holds the group's name: keeps it
            stats["kept"] += size
        else:
            name = base_slot + (size >> 1) if is_pivot_block else base_slot
            stats["renamed"] += size
        for j, s in enumerate(sub):
            writes.append((s, name, base_slot + j))
        if size >= 2:
            (nextS if size <= gcap else nextB).extend((name, s, base_slot + j) for j, s in enumerate(sub))

    S, B, hybrid, fallbacks = [], [], False, 0
    while act or S or B:
        stats["rounds"] += 1
        assert stats["rounds"] < 64
        if not hybrid:
            if fallbacks < 2 and any(len(g) > gcap for g in groups_of(act)):
                gs = groups_of(act)
                S = [e for g in gs if len(g) <= gcap for e in g]
                B = [e for g in gs if len(g) > gcap for e in g]
                act, hybrid = [], True
            else:
                act = radix_round(act)
                h *= 2
                continue
        stats["hybrid_rounds"] += 1
        # consistent snapshot of the second key half for both lists (k_gather_rank before any rank is written)
        key2 = {s: RANK[s + h] for (_, s, _) in S + B}
        nextS, nextB, writes, overflow = [], [], [], False
        for grp in groups_of(S) + groups_of(B):
            old_rank = grp[0][0]
            base = min(slot for (_, _, slot) in grp)
            members = [s for (_, s, _) in grp]
            assert sorted(slot for (_, _, slot) in grp) == list(range(base, base + len(grp)))
            assert base <= old_rank < base + len(grp)          # a rank is a slot inside the group's range
            if len(grp) <= gcap:                               # k_tile_sort: full sort, sub-groups by equal key
                members.sort(key=lambda s: key2[s])
                subs = [[s for s in members if key2[s] == k] for k in sorted(set(key2[s] for s in members))]
                at = base
                for sub in subs:
                    emit(sub, at, old_rank, False, nextS, nextB, writes)
                    at += len(sub)
                continue
            stats["stream_groups"] += 1
            M = len(members)
            samples = [key2[members[(M * q) >> 5]] for q in range(32)]
            piv = max(samples, key=lambda v: (samples.count(v), -samples.index(v)))
            lt = sorted((s for s in members if key2[s] < piv), key=lambda s: (key2[s], s))
            gt = sorted((s for s in members if key2[s] > piv), key=lambda s: (key2[s], s))
            eq = [s for s in members if key2[s] == piv]
            rnd.shuffle(eq)                                    # pass B places the block in no particular order
            if len(lt) + len(gt) > out_cap:
                overflow = True
                break
            at = base
            for side in (lt, None, gt):
                if side is None:
                    emit(eq, at, old_rank, True, nextS, nextB, writes)
                    at += len(eq)
                    continue
                a = 0
                for j in range(1, len(side) + 1):
                    if j == len(side) or key2[side[j]] != key2[side[a]]:
                        emit(side[a:j], at, old_rank, False, nextS, nextB, writes)
                        at += j - a
                        a = j
        if overflow:
            # groups handled before the overflow have been written (a valid refinement); the round is redone
            for s, name, slot in writes:
                RANK[s] = name
                SA[slot] = s
            stats["fallbacks"] += 1
            fallbacks += 1
            # the unified list keeps the OLD ranks as sort keys (the keys were gathered before any write)
            lst = S + B
            ck = [((g << 32) | key2[s], s) for (g, s, _) in lst]
            slots = sorted(slot for (_, _, slot) in lst)
            perm = sorted(range(len(ck)), key=lambda j: ck[j][0])
            act = regroup([ck[j][0] for j in perm], [ck[j][1] for j in perm], slots, None, SA, RANK)
            S, B, hybrid = [], [], False
        else:
            for s, name, slot in writes:
                RANK[s] = name
                SA[slot] = s
            S, B = nextS, nextB
        h *= 2
    return SA, RANK, stats["rounds"], stats


# ------------------------------------------------------------------ stage 2: LCP (chunked Kasai)
def lcp_array(data: bytes, SA, RANK, Q=16, seed=None):
    """k_lcp_kasai: Kasai in runs of Q text positions.  seed = (LCP, need) from seed_lcp: only marked positions are
    computed (an unmarked one breaks the carry), the others keep their key-derived value."""
    L = len(data)
    n1 = L + 1
    LCP = [0] * (n1 + 1) if seed is None else list(seed[0])
    for c0 in range(0, n1, Q):
        l = 0
        for i in range(c0, min(c0 + Q, n1)):
            if seed is not None and not seed[1][i]:
                l = 0
                continue
            r = RANK[i]
            if r == 0:
                l = 0
                continue
            j = SA[r - 1]
            maxl = L - max(i, j)
            while l < maxl and data[i + l] == data[j + l]:
                l += 1
            l = min(l, maxl)
            LCP[r] = l
            if l:
                l -= 1
    return LCP


# ------------------------------------------------------------------ stage 3: trees + walk
class Trees:
    def __init__(self, LCP, SA, rc, N):
        self.rc, self.N, self.twoN = rc, N, 2 * N
        self.lcp = [LCP]
        self.f = [SA]
        self.r = [SA]
        while len(self.lcp[-1]) > 32:
            a = self.lcp[-1]
            self.lcp.append([min(a[k:k + 32]) for k in range(0, len(a), 32)])
            if len(self.f) == 1:
                fv = [self.fval(s) for s in SA]
                rv = [self.rval(s) for s in SA]
            else:
                fv, rv = self.f[-1], self.r[-1]
            self.f.append([min(fv[k:k + 32]) for k in range(0, len(fv), 32)])
            self.r.append([max(rv[k:k + 32]) for k in range(0, len(rv), 32)])
        self.nlev = len(self.lcp)

    def fval(self, s):
        if self.rc:
            return s if s < self.N else NONE_MIN
        return s

    def rval(self, s):
        return s if (self.N < s <= self.twoN) else 0

    # lpnf.cuh line_prev_less / line_next_less: a 32-entry line is probed in pieces of four entries (one 16-byte load and
    # one compare group each), with an exit after every piece
    @staticmethod
    def line_prev_less(line, hi, d):
        """highest index <= hi of `line` (a list of up to 32 entries) whose entry is < d, or -1"""
        keep = (2 << (hi & 3)) - 1
        for c in range(hi >> 2, -1, -1):
            m4 = 0
            for e in range(4):
                j = 4 * c + e
                if j < len(line) and line[j] < d:
                    m4 |= 1 << e
            m4 &= keep
            if m4:
                return 4 * c + m4.bit_length() - 1
            keep = 0xF
        return -1

    @staticmethod
    def line_next_less(line, lo, last, d):
        """lowest index in [lo, last] of `line` whose entry is < d, or -1"""
        keep = 0xF & ~((1 << (lo & 3)) - 1)
        cl = last >> 2
        for c in range(lo >> 2, cl + 1):
            m4 = 0
            for e in range(4):
                j = 4 * c + e
                if j < len(line) and line[j] < d:
                    m4 |= 1 << e
            m4 &= keep
            if c == cl:
                m4 &= (2 << (last & 3)) - 1
            if m4:
                return 4 * c + (m4 & -m4).bit_length() - 1
            keep = 0xF
        return -1

    def find_prev_less(self, p, d):
        k = p
        for _ in range(4):
            if self.lcp[0][k] < d:
                return k
            k -= 1
        lev, idx = 0, k
        while True:
            a = self.lcp[lev]
            gstart = idx & ~31
            j = self.line_prev_less(a[gstart:gstart + 32], idx - gstart, d)
            if j >= 0:
                idx = gstart + j
                break
            idx = (gstart >> 5) - 1
            lev += 1
        while lev > 0:
            lev -= 1
            a = self.lcp[lev]
            base = idx << 5
            j = self.line_prev_less(a[base:base + 32], min(31, len(a) - 1 - base), d)
            assert j >= 0                                  # the parent's minimum is one of them
            idx = base + j
        return idx

    def find_next_less(self, p, d):
        k = p
        for _ in range(4):
            if self.lcp[0][k] < d:
                return k
            k += 1
        lev, idx = 0, k
        while True:
            a = self.lcp[lev]
            gstart = idx & ~31
            j = self.line_next_less(a[gstart:gstart + 32], idx - gstart, min(31, len(a) - 1 - gstart), d)
            if j >= 0:
                idx = gstart + j
                break
            idx = (gstart >> 5) + 1
            lev += 1
        while lev > 0:
            lev -= 1
            a = self.lcp[lev]
            base = idx << 5
            j = self.line_next_less(a[base:base + 32], 0, min(31, len(a) - 1 - base), d)
            assert j >= 0
            idx = base + j
        return idx

    def find_prev_r_greater(self, p, thr):
        """largest k <= p with rval(SA[k]) > thr, or -1."""
        k = p
        for _ in range(8):
            if k < 0:
                return -1
            if self.rval(self.f[0][k]) > thr:
                return k
            k -= 1
        if k < 0:
            return -1
        lev, idx = 0, k
        while True:
            gstart = idx & ~31
            j = idx
            while j >= gstart and not self._rv(lev, j) > thr:
                j -= 1
            if j >= gstart:
                idx = j
                break
            if gstart == 0:
                return -1
            idx = (gstart >> 5) - 1
            lev += 1
        while lev > 0:
            lev -= 1
            base = idx << 5
            j = min(base + 31, self._rcount(lev) - 1)
            while j > base and not self._rv(lev, j) > thr:
                j -= 1
            idx = j
        return idx

    def find_next_r_greater(self, p, thr):
        """smallest k >= p with rval(SA[k]) > thr, or -1."""
        n1 = len(self.f[0])
        k = p
        for _ in range(8):
            if k >= n1:
                return -1
            if self.rval(self.f[0][k]) > thr:
                return k
            k += 1
        if k >= n1:
            return -1
        lev, idx = 0, k
        while True:
            cnt = self._rcount(lev)
            gend = min(idx | 31, cnt - 1)
            j = idx
            while j <= gend and not self._rv(lev, j) > thr:
                j += 1
            if j <= gend:
                idx = j
                break
            idx = (idx >> 5) + 1
            lev += 1
            if lev >= self.nlev or idx >= self._rcount(lev):
                return -1
        while lev > 0:
            lev -= 1
            base = idx << 5
            top = min(base + 31, self._rcount(lev) - 1)
            j = base
            while j < top and not self._rv(lev, j) > thr:
                j += 1
            idx = j
        return idx

    def _rv(self, lev, j):
        return self.rval(self.f[0][j]) if lev == 0 else self.r[lev][j]

    def _rcount(self, lev):
        return len(self.f[0]) if lev == 0 else len(self.r[lev])

    def lcp_range_min(self, a, b):
        """min LCP[a..b] (inclusive, a <= b)."""
        m = NONE_MIN
        l0 = self.lcp[0]
        if b - a < 96:
            for k in range(a, b + 1):
                m = min(m, l0[k])
            return m
        while a & 31:
            m = min(m, l0[a]); a += 1
        while (b + 1) & 31:
            m = min(m, l0[b]); b -= 1
        a >>= 5
        b = ((b + 1) >> 5) - 1
        lev = 1
        while a <= b:
            la = self.lcp[lev]
            if b - a < 64 or lev == self.nlev - 1:
                for k in range(a, b + 1):
                    m = min(m, la[k])
                return m
            while a & 31:
                m = min(m, la[a]); a += 1
            while (b + 1) & 31:
                m = min(m, la[b]); b -= 1
            a >>= 5
            b = ((b + 1) >> 5) - 1
            lev += 1
        return m

    def agg(self, a, b, fmin, rmax):
        if a > b:
            return fmin, rmax
        sa = self.f[0]

        def upd0(k, fmin, rmax):
            s = sa[k]
            return min(fmin, self.fval(s)), (max(rmax, self.rval(s)) if self.rc else rmax)

        if b - a < 96:
            for k in range(a, b + 1):
                fmin, rmax = upd0(k, fmin, rmax)
            return fmin, rmax
        while a & 31:
            fmin, rmax = upd0(a, fmin, rmax)
            a += 1
        while (b + 1) & 31:
            fmin, rmax = upd0(b, fmin, rmax)
            b -= 1
        a >>= 5
        b = ((b + 1) >> 5) - 1
        lev = 1
        while a <= b:
            fa, ra = self.f[lev], self.r[lev]
            if b - a < 64 or lev == self.nlev - 1:
                for k in range(a, b + 1):
                    fmin = min(fmin, fa[k])
                    rmax = max(rmax, ra[k]) if self.rc else rmax
                return fmin, rmax
            while a & 31:
                fmin = min(fmin, fa[a]); rmax = max(rmax, ra[a]) if self.rc else rmax
                a += 1
            while (b + 1) & 31:
                fmin = min(fmin, fa[b]); rmax = max(rmax, ra[b]) if self.rc else rmax
                b -= 1
            a >>= 5
            b = ((b + 1) >> 5) - 1
            lev += 1
        return fmin, rmax


K_LIN = 4  # path nodes climbed one by one before switching to a binary search over string depth


def _extend(T, st, D):
    """Node state (lo, hi, F, R) -> state of interval(D) (D <= current depth), incrementally."""
    lo, hi, F, R = st
    LCP = T.lcp[0]
    nlo = T.find_prev_less(lo - 1, D) if LCP[lo] >= D else lo
    nhi = T.find_next_less(hi + 2, D) - 1 if LCP[hi + 1] >= D else hi
    F, R = T.agg(nlo, lo - 1, F, R)
    F, R = T.agg(hi + 1, nhi, F, R)
    return (nlo, nhi, F, R)


def _search(T, cur, loD, hiD, pred, U0=None):
    """max D in [loD, hiD) with pred (loD known true or 0, hiD known false; `cur` = a failing state of
    depth >= hiD).  Returns (D*, U = interval(D*) state or None, L = interval(D*+1) state)."""
    L, U = cur, U0
    while hiD - loD > 1:
        mid = (loD + hiD) // 2
        cand = _extend(T, L, mid)
        if pred(cand, mid):
            loD, U = mid, cand
        else:
            hiD, L = mid, cand
    return loD, U, L


def _finish(rc, twoN, i, have_f, fwd_len, jF, gen, rinfo):
    """Selection rule shared by both kernels: general answer, or RC forward/RC/literal choice."""
    if not rc:
        return gen if gen else (1, i)
    have_r, dR, mR = rinfo
    rc_len = dR if have_r else 0
    use_fwd = use_lit = False
    if have_f and fwd_len >= 1:
        use_fwd = not (have_r and rc_len > fwd_len)
    elif not (have_r and rc_len > 1):
        use_lit = True
    if use_lit:
        return 1, i
    if use_fwd:
        return fwd_len, jF
    e = twoN - mR
    return rc_len, (e - rc_len + 1) | LR_RC_FLAG


def walk(T: Trees, n1, nfac, RANK, k_lin=K_LIN, Q=16, real=None):
    """Stage 3 = k_lpnf_rank (rank order, bounded climb, direct RC candidate; marks `hard` positions)
    followed by k_lpnf_hard (text order over hard positions with a Kasai-style carry).
    real = (lo, hi): only these ranks are evaluated (distributed runs: the others are virtual ranks)."""
    SA, LCP = T.f[0], T.lcp[0]
    rc, twoN = T.rc, T.twoN
    LR = [None] * nfac
    HARD = [None] * nfac
    i = 0

    def pred_f(st, D):
        return st[2] != NONE_MIN and st[2] + D <= i

    # ---- kernel 1: rank order
    for r in (range(n1) if real is None else range(real[0], real[1])):
        i = SA[r]
        if i >= nfac:
            continue
        leaf = (r, r, i, 0)
        cur = leaf
        d_node = None
        have_f = at_root = False
        dF = jF = 0
        belowF = i
        steps = 0
        while True:
            d = max(LCP[cur[0]], LCP[cur[1] + 1])
            if d == 0:
                at_root = True
                break
            if steps == k_lin:
                break
            steps += 1
            childF = cur[2]
            cur = _extend(T, cur, d)
            d_node = d
            if pred_f(cur, d):
                have_f, dF, jF, belowF = True, d, cur[2], childF
                break
        rinfo = (False, 0, 0)
        if rc:
            thr = twoN - i
            kl = T.find_prev_r_greater(r - 1, thr)
            kr = T.find_next_r_greater(r + 1, thr)
            dl = T.lcp_range_min(kl + 1, r) if kl >= 0 else 0
            dr = T.lcp_range_min(r + 1, kr) if kr >= 0 else 0
            dR = max(dl, dr)
            if dR >= 1:
                rinfo = (True, dR, _extend(T, leaf, dR)[3])
        gen, fwd_len = None, 0
        if have_f:
            part = (i - belowF) if belowF != i else 0
            gen = (part, belowF) if part > dF else (dF, jF)
            fwd_len = (i - jF) if belowF == jF else dF
        elif at_root:
            v_min = cur[2]
            gen = (i - v_min, v_min) if v_min != i else None
        else:
            # depth known to satisfy the forward predicate: D0 = i - minF(last failing ancestor) (lpnf.cuh, k_lpnf_rank)
            lb0 = i - cur[2] if (cur[2] != NONE_MIN and cur[2] < i) else 0
            HARD[i] = (cur, d_node if d_node is not None else d + 1, rinfo, lb0)
            continue
        LR[i] = _finish(rc, twoN, i, have_f, fwd_len, jF, gen, rinfo)

    # ---- kernel 2: text order over the hard positions, carrying the true forward match length
    for c0 in range(0, nfac, Q):
        prevF = 0
        for i in range(c0, min(c0 + Q, nfac)):
            if HARD[i] is None:
                prevF = 0
                continue
            cur, Dtop, rinfo, lb0 = HARD[i]
            # (the CUDA kernel re-derives `cur` as interval(Dtop-ish) from the leaf; same state)
            lb = prevF - 1 if prevF > 0 else 0
            lb = max(lb, lb0)
            assert lb0 == 0 or pred_f(_extend(T, cur, lb0), lb0)      # the parked bound holds
            U = L = None
            if lb >= 1 and lb + 1 < Dtop:
                st = _extend(T, cur, lb + 1)
                if pred_f(st, lb + 1):
                    Ds, U, L = _search(T, cur, lb + 1, Dtop, pred_f, st)
                else:
                    Ds, L = lb, st
                    U = _extend(T, st, lb)
            elif lb >= 1:
                Ds, L = lb, cur
                U = _extend(T, cur, lb)
            else:
                Ds, U, L = _search(T, cur, 0, Dtop, pred_f)
            prevF = Ds
            gen = (Ds, U[2]) if Ds >= 1 else None
            have_f, fwd_len, jF = False, 0, 0
            if rc and Ds >= 1:
                if (U[0], U[1]) != (L[0], L[1]):
                    have_f, jF = True, U[2]
                    fwd_len = (i - jF) if L[2] == U[2] else Ds
                else:
                    du = max(LCP[U[0]], LCP[U[1] + 1])
                    if du > 0:
                        P = _extend(T, U, du)
                        have_f, jF = True, P[2]
                        fwd_len = (i - jF) if U[2] == P[2] else du
            LR[i] = _finish(rc, twoN, i, have_f, fwd_len, jF, gen, rinfo)
    return LR


# ---- round-2 stage 3: node tables carry the R-max too; ONE climb settles the forward and the RC candidate ---------------
DR_UNRESOLVED = 0xFFFFFFFF


def node_tables(T: Trees, n1):
    """k_node_tables: NODE[k] = (rank naming the parent, min forward start, string depth, max rc value) of the LCP interval
    rank k names (previous / next strictly smaller LCP value; aggregates over the interval from the summary trees)."""
    LCP = T.lcp[0]
    NODE = [None] * (n1 + 1)
    for k in range(n1 + 1):
        d = 0 if k in (0, n1) else LCP[k]
        if d == 0:
            NODE[k] = (k, NONE_MIN, 0, 0)
            continue
        a = T.find_prev_less(k - 1, d)
        b1 = T.find_next_less(k + 1, d)
        fm, rm = T.agg(a, b1 - 1, NONE_MIN, 0)
        NODE[k] = (a if LCP[a] >= LCP[b1] else b1, fm, d, rm if T.rc else 0)
    return NODE


def walk_tables(T: Trees, n1, nfac, RANK, max_nodes=4, Q=16, real=None):
    """k_node_tables + k_lpnf_rank + k_lpnf_hard as of round 2: the leaf climbs its tabulated ancestors once; the first
    (deepest) ancestor whose R-max qualifies IS the RC candidate node vR (factorizer_core.hpp:269-271), the first whose
    F-min + depth <= i is vF (:264-266); nothing above vF can beat it (forward wins ties), so the climb ends there.  A
    position that exhausts the climb budget is hard; if its RC candidate was not met on the way, k_lpnf_hard finds it from
    the nearest qualifying rc rank on either side (summary-tree searches)."""
    SA, LCP = T.f[0], T.lcp[0]
    rc, twoN = T.rc, T.twoN
    NODE = node_tables(T, n1)
    LR = [None] * nfac
    HARD = [None] * nfac
    for r in (range(n1) if real is None else range(real[0], real[1])):
        i = SA[r]
        if i >= nfac:
            continue
        thr = twoN - i
        k = r if LCP[r] >= LCP[r + 1] else r + 1
        have_f = at_root = False
        dF = jF = 0
        belowF = childF = i
        dR = mR = 0
        steps = 0
        while True:
            par, m, d, rm = NODE[k]
            if d == 0:
                at_root = True
                break
            if steps == max_nodes:
                break
            steps += 1
            if rc and dR == 0 and rm > thr:
                dR, mR = d, rm
            if m != NONE_MIN and m + d <= i:
                have_f, dF, jF, belowF = True, d, m, childF
                break
            childF = m
            k = par
        rinfo = (dR >= 1, dR, mR)
        gen, fwd_len = None, 0
        if have_f:
            part = (i - belowF) if belowF != i else 0
            gen = (part, belowF) if part > dF else (dF, jF)
            fwd_len = (i - jF) if belowF == jF else dF
        elif at_root:
            gen = (i - childF, childF) if childF != i else None
        else:
            lb0 = i - childF if (childF != NONE_MIN and childF < i) else 0
            HARD[i] = (r, dR if (dR >= 1 or not rc) else DR_UNRESOLVED, lb0)
            continue
        LR[i] = _finish(rc, twoN, i, have_f, fwd_len, jF, gen, rinfo)

    def pred_f(st, D):
        return st[2] != NONE_MIN and st[2] + D <= i

    for c0 in range(0, nfac, Q):
        prevF = 0
        for i in range(c0, min(c0 + Q, nfac)):
            if HARD[i] is None:
                prevF = 0
                continue
            r, dR, lb0 = HARD[i]
            leaf = (r, r, i, 0)
            Dtop = max(LCP[r], LCP[r + 1]) + 1
            lb = max(prevF - 1 if prevF > 0 else 0, lb0)
            U = L = leaf
            if lb >= 1 and lb + 1 < Dtop:
                st = _extend(T, leaf, lb + 1)
                if pred_f(st, lb + 1):
                    Ds, U, L = _search(T, leaf, lb + 1, Dtop, pred_f, st)
                else:
                    Ds, L = lb, st
                    U = _extend(T, st, lb)
            elif lb >= 1:
                Ds, L = lb, leaf
                U = _extend(T, leaf, lb)
            else:
                Ds, U, L = _search(T, leaf, 0, Dtop, pred_f)
                if U is None:
                    U = leaf
            prevF = Ds
            gen = (Ds, U[2]) if Ds >= 1 else None
            have_f, fwd_len, jF = False, 0, 0
            if rc and Ds >= 1:
                if (U[0], U[1]) != (L[0], L[1]):
                    have_f, jF = True, U[2]
                    fwd_len = (i - jF) if L[2] == U[2] else Ds
                else:
                    du = max(LCP[U[0]], LCP[U[1] + 1])
                    if du > 0:
                        P = _extend(T, U, du)
                        have_f, jF = True, P[2]
                        fwd_len = (i - jF) if U[2] == P[2] else du
            rinfo = (False, 0, 0)
            if rc:
                if dR == DR_UNRESOLVED:
                    # an RC candidate only counts when it is deeper than the forward one: without a qualifying rc suffix
                    # inside L = interval(Ds + 1) its depth is <= fwd_len and the exact value does not matter
                    thr = twoN - i
                    dR = 0
                    if T.agg(L[0], L[1], NONE_MIN, 0)[1] > thr:
                        kl = T.find_prev_r_greater(r - 1, thr)       # rc_depth_v: nearest qualifying rc rank on either side
                        kr = T.find_next_r_greater(r + 1, thr)
                        dl = T.lcp_range_min(kl + 1, r) if kl >= 0 else 0
                        dr = T.lcp_range_min(r + 1, kr) if kr >= 0 else 0
                        dR = max(dl, dr)
                if dR >= 1:
                    rinfo = (True, dR, _extend(T, leaf, dR)[3])
            LR[i] = _finish(rc, twoN, i, have_f, fwd_len, jF, gen, rinfo)
    return LR


# ------------------------------------------------------------------ stage 4: chain
def chain(LR, nfac, start_pos, rc, chunk=1024):
    nchunks = (nfac + chunk - 1) // chunk
    EXIT = [0] * nfac
    alist = [start_pos]
    for c in range(nchunks):
        base, end = c * chunk, min((c + 1) * chunk, nfac)
        prev = None
        for e in range(base, end):
            v = e + LR[e][0]
            while v < end:
                v = v + LR[v][0]
            EXIT[e] = v
            if v < nfac and (e == base or v != prev):
                alist.append(v)
            prev = v
    REACH = [0] * nfac
    REACH[start_pos] = 1
    J = list(EXIT)
    rounds = max(1, nchunks).bit_length() + 1
    for _ in range(rounds):
        Jn = list(J)
        for x in alist:
            j = J[x]
            if j < nfac:
                if REACH[x]:
                    REACH[j] = 1
                Jn[x] = J[j]
            else:
                Jn[x] = j
        J = Jn
    out = []
    for c in range(nchunks):
        base, end = c * chunk, min((c + 1) * chunk, nfac)
        entry = next((o for o in range(base, end) if REACH[o]), None)
        o = entry
        while o is not None and o < end:
            ln, ref32 = LR[o]
            ref = ((ref32 & ~LR_RC_FLAG) | (RC_MASK if ref32 & LR_RC_FLAG else 0)) if rc else ref32
            out.append((o, ln, ref))
            o = o + ln
    return out


# ------------------------------------------------------------------ whole pipeline
def factorize_model(data: bytes, mode: str = "general", start_pos: int = 0, force_bits=None, chunk=1024,
                    k_lin=K_LIN, walk_q=16, tables=True):
    """mode: 'general' | 'rc_prepared'."""
    data = bytes(data)
    if mode == "general":
        if len(data) == 0 or start_pos >= len(data):
            return []
        rc, N, nfac = False, 0, len(data)
    else:
        if len(data) < 4:
            return []
        N = len(data) // 2 - 1
        if N == 0:
            return []
        if start_pos >= N:
            raise ValueError("start_pos must be less than the original sequence length")
        rc, nfac = True, N
    n1 = len(data) + 1
    so = {}
    SA, RANK, _ = suffix_array(data, force_bits, seed_out=so)
    LCP = lcp_array(data, SA, RANK, seed=so["seed"])          # key-derived LCP + Kasai over the marked positions
    T = Trees(LCP, SA, rc, N)
    LR = (walk_tables(T, n1, nfac, RANK, k_lin, walk_q) if tables else walk(T, n1, nfac, RANK, k_lin, walk_q))
    return chain(LR, nfac, start_pos, rc, chunk)
