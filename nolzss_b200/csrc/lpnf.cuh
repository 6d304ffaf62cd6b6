// Stage 3: per-position longest previous NON-overlapping factor, exactly as the reference resolves it.
//
// f(i) = (length, ref) is a pure function of the text, so it is evaluated
// for every text position in parallel and the greedy chain i -> i+len is extracted afterwards
// (chain.cuh).  One thread owns one suffix-array rank r (rank space keeps the SA/LCP neighbourhood
// of a warp in the same cache lines) and climbs the LCP-interval ancestors of leaf r, i.e. the
// suffix-tree path the reference walks with level_anc (factorizer_core.hpp:70-78, :256-300):
//
//   general mode (detail::nolzss, factorizer_core.hpp:51-119)
//     u = deepest path node with  min SA(u) + depth(u) <= i ;  v = path node just below u
//     -> (depth(u), minSA(u)) or (i - minSA(v), minSA(v)) or the literal (1, i)      (:82-108)
//   RC mode (detail::nolzss_multiple_dna_w_rc, factorizer_core.hpp:177-383)
//     vF = deepest node with  minF + depth <= i   (minF over suffixes starting in T, :264-266)
//     vR = deepest node with  2N - maxR < i       (maxR over suffixes starting in rc(T), :269-271)
//     fwd_len = min(lcp(i, jF), i - jF) = (minF(child of vF) == jF) ? i - jF : depth(vF)   (:322-326)
//     rc_len  = depth(vR)  (:328-330);  forward wins ties, RC needs > 1 without a forward (:338-352)
//
// Interval boundaries (previous/next smaller LCP value) and the range aggregates (min / max of SA
// over an interval) are answered from 32-ary summary trees: one 128-byte line per tree node, so a
// probe costs O(log32 n) line reads; short ranges are scanned directly.
#pragma once
#include "common.cuh"

namespace nlz {

constexpr int TREE_MAX_LEVELS = 8;
constexpr u32 NONE_MIN = 0xFFFFFFFFu;

struct Trees {
    int nlev;                          // number of levels including level 0
    const u32* lcp[TREE_MAX_LEVELS];   // lcp[0] = LCP (n1+1 entries), lcp[t] = block minima
    u32 cntL[TREE_MAX_LEVELS];
    const u32* f[TREE_MAX_LEVELS];     // f[0] = SA; f[t] = block min of F-class values
    const u32* r[TREE_MAX_LEVELS];     // r[t] = block max of R-class values (RC mode only)
    u32 cntS[TREE_MAX_LEVELS];
};

struct WalkParams {
    u32 n1;        // number of suffixes
    u32 nfac;      // positions [0, nfac) are factorized (general: L, RC: N)
    u32 N;         // RC: |S|/2 - 1
    u32 twoN;      // RC: 2N
};

template <bool RC> __device__ __forceinline__ u32 f_value(u32 s, const WalkParams& p) {
    return RC ? (s < p.N ? s : NONE_MIN) : s;
}
__device__ __forceinline__ u32 r_value(u32 s, const WalkParams& p) {
    return (s > p.N && s <= p.twoN) ? s : 0u;   // 0 = none (valid values are >= N+1 >= 1)
}

// ---- tree construction ----------------------------------------------------------------------
// level-1 nodes from level 0: LCP minima and, from SA, F minima / R maxima.
template <bool RC>
__global__ void __launch_bounds__(256)
k_tree_level1(const u32* __restrict__ LCP, u32 cntL0, const u32* __restrict__ SA, u32 cntS0,
              WalkParams p, u32* __restrict__ lcp1, u32 cntL1, u32* __restrict__ f1,
              u32* __restrict__ r1, u32 cntS1) {
    // one warp per node: coalesced 128-byte line, shuffle reduction
    const u32 node = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    const u64 e = (u64)node * 32 + lane;
    if (node < cntL1) {
        u32 v = e < cntL0 ? LCP[e] : NONE_MIN;
#pragma unroll
        for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) lcp1[node] = v;
    }
    if (node < cntS1) {
        u32 fv = NONE_MIN, rv = 0;
        if (e < cntS0) {
            u32 s = SA[e];
            fv = f_value<RC>(s, p);
            if (RC) rv = r_value(s, p);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            fv = min(fv, __shfl_xor_sync(0xffffffffu, fv, o));
            if (RC) rv = max(rv, __shfl_xor_sync(0xffffffffu, rv, o));
        }
        if (lane == 0) { f1[node] = fv; if (RC) r1[node] = rv; }
    }
}

template <bool RC>
__global__ void __launch_bounds__(256)
k_tree_level_up(const u32* __restrict__ lcpA, u32 cntLA, const u32* __restrict__ fA,
                const u32* __restrict__ rA, u32 cntSA, u32* __restrict__ lcpB, u32 cntLB,
                u32* __restrict__ fB, u32* __restrict__ rB, u32 cntSB) {
    const u32 node = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    const u64 e = (u64)node * 32 + lane;
    if (node < cntLB) {
        u32 v = e < cntLA ? lcpA[e] : NONE_MIN;
#pragma unroll
        for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) lcpB[node] = v;
    }
    if (node < cntSB) {
        u32 fv = e < cntSA ? fA[e] : NONE_MIN;
        u32 rv = (RC && e < cntSA) ? rA[e] : 0u;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            fv = min(fv, __shfl_xor_sync(0xffffffffu, fv, o));
            if (RC) rv = max(rv, __shfl_xor_sync(0xffffffffu, rv, o));
        }
        if (lane == 0) { fB[node] = fv; if (RC) rB[node] = rv; }
    }
}

// ---- queries --------------------------------------------------------------------------------
// largest k <= p with LCP[k] < d   (exists: LCP[0] = 0 < d since d >= 1)
__device__ __forceinline__ u32 find_prev_less(const Trees& T, u32 p, u32 d) {
    i64 k = p;
    const u32* l0 = T.lcp[0];
#pragma unroll 1
    for (int s = 0; s < 12; ++s) {
        if (l0[k] < d) return (u32)k;
        --k;
    }
    int lev = 0;
    i64 idx = k;
    for (;;) {
        const u32* a = T.lcp[lev];
        const i64 gstart = idx & ~31LL;
        i64 j = idx;
        for (; j >= gstart; --j)
            if (a[j] < d) break;
        if (j >= gstart) { idx = j; break; }
        idx = (gstart >> 5) - 1;
        ++lev;
    }
    while (lev > 0) {
        --lev;
        const u32* a = T.lcp[lev];
        const i64 base = idx << 5;
        i64 j = base + 31;
        if (j > (i64)T.cntL[lev] - 1) j = (i64)T.cntL[lev] - 1;
        for (; j > base; --j)
            if (a[j] < d) break;
        idx = j;
    }
    return (u32)idx;
}

// smallest k >= p with LCP[k] < d   (exists: LCP[n1] = 0)
__device__ __forceinline__ u32 find_next_less(const Trees& T, u32 p, u32 d) {
    i64 k = p;
    const u32* l0 = T.lcp[0];
#pragma unroll 1
    for (int s = 0; s < 12; ++s) {
        if (l0[k] < d) return (u32)k;
        ++k;
    }
    int lev = 0;
    i64 idx = k;
    for (;;) {
        const u32* a = T.lcp[lev];
        i64 gend = idx | 31;
        if (gend > (i64)T.cntL[lev] - 1) gend = (i64)T.cntL[lev] - 1;
        i64 j = idx;
        for (; j <= gend; ++j)
            if (a[j] < d) break;
        if (j <= gend) { idx = j; break; }
        idx = (idx >> 5) + 1;
        ++lev;
    }
    while (lev > 0) {
        --lev;
        const u32* a = T.lcp[lev];
        const i64 base = idx << 5;
        i64 top = base + 31;
        if (top > (i64)T.cntL[lev] - 1) top = (i64)T.cntL[lev] - 1;
        i64 j = base;
        for (; j < top; ++j)
            if (a[j] < d) break;
        idx = j;
    }
    return (u32)idx;
}

// aggregate F-min / R-max of SA over ranks [a, b] (inclusive; empty when a > b)
template <bool RC>
__device__ __forceinline__ void agg_range(const Trees& T, const WalkParams& p, i64 a, i64 b, u32& fmin, u32& rmax) {
    if (a > b) return;
    const u32* sa = T.f[0];
    if (b - a < 96) {
        for (i64 k = a; k <= b; ++k) {
            u32 s = sa[k];
            fmin = min(fmin, f_value<RC>(s, p));
            if (RC) rmax = max(rmax, r_value(s, p));
        }
        return;
    }
    while (a & 31) {
        u32 s = sa[a++];
        fmin = min(fmin, f_value<RC>(s, p));
        if (RC) rmax = max(rmax, r_value(s, p));
    }
    while ((b + 1) & 31) {
        u32 s = sa[b--];
        fmin = min(fmin, f_value<RC>(s, p));
        if (RC) rmax = max(rmax, r_value(s, p));
    }
    a >>= 5;
    b = ((b + 1) >> 5) - 1;
    int lev = 1;
    while (a <= b) {
        const u32* fa = T.f[lev];
        const u32* ra = T.r[lev];
        if (b - a < 64 || lev == T.nlev - 1) {
            for (i64 k = a; k <= b; ++k) {
                fmin = min(fmin, fa[k]);
                if (RC) rmax = max(rmax, ra[k]);
            }
            return;
        }
        while (a & 31) {
            fmin = min(fmin, fa[a]);
            if (RC) rmax = max(rmax, ra[a]);
            ++a;
        }
        while ((b + 1) & 31) {
            fmin = min(fmin, fa[b]);
            if (RC) rmax = max(rmax, ra[b]);
            --b;
        }
        a >>= 5;
        b = ((b + 1) >> 5) - 1;
        ++lev;
    }
}

// ---- the walk -------------------------------------------------------------------------------
// LR[i] = (ref | rc_flag<<31) << 32 | len        for every factorized position i = SA[r]
constexpr u32 LR_RC_FLAG = 0x80000000u;

// Evaluates the factor rule for suffix i at rank r; returns the number of path nodes visited.
template <bool RC>
__device__ __forceinline__ u32 lpnf_one(const Trees& T, const WalkParams& p, u32 r, u32 i, u64* __restrict__ LR) {
    const u32* LCP = T.lcp[0];
    u32 visited = 0;

    u32 lo = r, hi = r;
    u32 curF = i;        // F-min over the current node (the leaf holds suffix i, which is in T)
    u32 curR = 0;        // R-max over the current node (none)
    bool have_f = false, have_r = false;
    u32 dF = 0, jF = 0, belowF = i;   // deepest ok-forward node: depth, min start, F-min of its path child
    u32 dR = 0, mR = 0;               // deepest ok-RC node: depth, R-max
    u32 lastF = i;                    // F-min of the last node visited (child of root when the loop ends)

    for (;;) {
        u32 dl = LCP[lo], dh = LCP[hi + 1];
        u32 d = max(dl, dh);
        if (d == 0) break;                      // parent is the root
        ++visited;
        u32 nlo = (dl >= d) ? find_prev_less(T, lo - 1, d) : lo;
        u32 nhi = (dh >= d) ? find_next_less(T, hi + 2, d) - 1 : hi;
        u32 childF = curF;
        agg_range<RC>(T, p, (i64)nlo, (i64)lo - 1, curF, curR);
        agg_range<RC>(T, p, (i64)hi + 1, (i64)nhi, curF, curR);
        lo = nlo; hi = nhi;
        lastF = curF;
        if (!have_f && curF != NONE_MIN && (u64)curF + d <= (u64)i) {
            have_f = true; dF = d; jF = curF; belowF = childF;
            if (!RC) break;
        }
        if (RC && !have_r && curR != 0 && (p.twoN - curR) < i) {
            have_r = true; dR = d; mR = curR;
        }
        if (RC && have_f && have_r) break;
    }

    u32 len, ref;
    if (!RC) {
        // v = node just below u (or the child of the root when no u exists)
        u32 v_min = have_f ? belowF : lastF;
        if (v_min == i) {                                   // factorizer_core.hpp:82
            if (!have_f) { len = 1; ref = i; }              // :83-87
            else { len = dF; ref = jF; }                    // :89-94
        } else {
            u32 Lc = i - v_min;                             // :96-97 (lcp(i, v_min) >= depth(v) > i - v_min)
            if (!have_f || Lc > dF) { len = Lc; ref = v_min; }   // :104-107
            else { len = dF; ref = jF; }                    // :98-102
        }
    } else {
        u32 fwd_len = 0;
        if (have_f) fwd_len = (belowF == jF) ? (i - jF) : dF;    // :322-326
        u32 rc_len = have_r ? dR : 0;                            // :328-330
        bool use_fwd = false, use_lit = false;
        if (have_f && fwd_len >= 1) use_fwd = !(have_r && rc_len > fwd_len);   // :338-344
        else if (!(have_r && rc_len > 1)) use_lit = true;                      // :346-351
        if (use_lit) { len = 1; ref = i; }
        else if (use_fwd) { len = fwd_len; ref = jF; }
        else {
            u32 e = p.twoN - mR;                                 // smallest RC end in T coordinates
            len = rc_len;
            ref = (e - rc_len + 1) | LR_RC_FLAG;                 // :362-364 (start-anchored + RC flag)
        }
    }
    LR[i] = ((u64)ref << 32) | (u64)len;
    return visited;
}

template <bool RC>
__global__ void __launch_bounds__(256)
k_lpnf_walk(Trees T, WalkParams p, u64* __restrict__ LR, unsigned long long* __restrict__ visit_counter) {
    const u32 r = blockIdx.x * 256 + threadIdx.x;
    u32 visited = 0;
    if (r < p.n1) {
        const u32 i = T.f[0][r];
        if (i < p.nfac) visited = lpnf_one<RC>(T, p, r, i, LR);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) visited += __shfl_xor_sync(0xffffffffu, visited, o);
    if ((threadIdx.x & 31) == 0 && visited) atomicAdd(visit_counter, (unsigned long long)visited);
}

}  // namespace nlz
