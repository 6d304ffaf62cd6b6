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
def choose_layout(data: bytes, n1: int, force_bits=None):
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
        return cls, dict(key_bits=32, b=b, W=w32, D=4)
    return cls, dict(key_bits=64, b=b, W=min(29, 59 // b), D=5)


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
        keys.append(key | dist)
    return keys


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


def suffix_array(data: bytes, force_bits=None):
    L = len(data)
    n1 = L + 1
    cls, lay = choose_layout(data, n1, force_bits)
    keys = build_keys(data, cls, lay)
    order = sorted(range(n1), key=lambda p: keys[p])          # stable LSD radix sort
    SA = [0] * n1
    RANK = [0] * n1
    act = regroup([keys[p] for p in order], order, None, (1 << lay["D"]) - 1, SA, RANK)
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


# ------------------------------------------------------------------ stage 2: LCP (chunked Kasai)
def lcp_array(data: bytes, SA, RANK, Q=16):
    L = len(data)
    n1 = L + 1
    LCP = [0] * (n1 + 1)
    for c0 in range(0, n1, Q):
        l = 0
        for i in range(c0, min(c0 + Q, n1)):
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

    def find_prev_less(self, p, d):
        k = p
        for _ in range(12):
            if self.lcp[0][k] < d:
                return k
            k -= 1
        lev, idx = 0, k
        while True:
            a = self.lcp[lev]
            gstart = idx & ~31
            j = idx
            while j >= gstart and not a[j] < d:
                j -= 1
            if j >= gstart:
                idx = j
                break
            idx = (gstart >> 5) - 1
            lev += 1
        while lev > 0:
            lev -= 1
            a = self.lcp[lev]
            base = idx << 5
            j = min(base + 31, len(a) - 1)
            while j > base and not a[j] < d:
                j -= 1
            idx = j
        return idx

    def find_next_less(self, p, d):
        k = p
        for _ in range(12):
            if self.lcp[0][k] < d:
                return k
            k += 1
        lev, idx = 0, k
        while True:
            a = self.lcp[lev]
            gend = min(idx | 31, len(a) - 1)
            j = idx
            while j <= gend and not a[j] < d:
                j += 1
            if j <= gend:
                idx = j
                break
            idx = (idx >> 5) + 1
            lev += 1
        while lev > 0:
            lev -= 1
            a = self.lcp[lev]
            base = idx << 5
            top = min(base + 31, len(a) - 1)
            j = base
            while j < top and not a[j] < d:
                j += 1
            idx = j
        return idx

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


def walk(T: Trees, n1, nfac):
    """k_lpnf_walk: LR[i] = (len, ref32)."""
    SA, LCP = T.f[0], T.lcp[0]
    rc, twoN = T.rc, T.twoN
    LR = [None] * nfac
    for r in range(n1):
        i = SA[r]
        if i >= nfac:
            continue
        lo = hi = r
        curF, curR = i, 0
        have_f = have_r = False
        dF = jF = 0
        belowF = i
        dR = mR = 0
        lastF = i
        while True:
            dl, dh = LCP[lo], LCP[hi + 1]
            d = max(dl, dh)
            if d == 0:
                break
            nlo = T.find_prev_less(lo - 1, d) if dl >= d else lo
            nhi = T.find_next_less(hi + 2, d) - 1 if dh >= d else hi
            childF = curF
            curF, curR = T.agg(nlo, lo - 1, curF, curR)
            curF, curR = T.agg(hi + 1, nhi, curF, curR)
            lo, hi = nlo, nhi
            lastF = curF
            if not have_f and curF != NONE_MIN and curF + d <= i:
                have_f, dF, jF, belowF = True, d, curF, childF
                if not rc:
                    break
            if rc and not have_r and curR != 0 and (twoN - curR) < i:
                have_r, dR, mR = True, d, curR
            if rc and have_f and have_r:
                break
        if not rc:
            v_min = belowF if have_f else lastF
            if v_min == i:
                ln, ref = (1, i) if not have_f else (dF, jF)
            else:
                Lc = i - v_min
                ln, ref = (Lc, v_min) if (not have_f or Lc > dF) else (dF, jF)
        else:
            fwd_len = ((i - jF) if belowF == jF else dF) if have_f else 0
            rc_len = dR if have_r else 0
            use_fwd = use_lit = False
            if have_f and fwd_len >= 1:
                use_fwd = not (have_r and rc_len > fwd_len)
            elif not (have_r and rc_len > 1):
                use_lit = True
            if use_lit:
                ln, ref = 1, i
            elif use_fwd:
                ln, ref = fwd_len, jF
            else:
                e = twoN - mR
                ln, ref = rc_len, (e - rc_len + 1) | LR_RC_FLAG
        LR[i] = (ln, ref)
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
def factorize_model(data: bytes, mode: str = "general", start_pos: int = 0, force_bits=None, chunk=1024):
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
    SA, RANK, _ = suffix_array(data, force_bits)
    LCP = lcp_array(data, SA, RANK)
    T = Trees(LCP, SA, rc, N)
    LR = walk(T, n1, nfac)
    return chain(LR, nfac, start_pos, rc, chunk)
