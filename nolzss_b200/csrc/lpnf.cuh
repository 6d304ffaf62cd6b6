// Stage 3: per-position longest previous NON-overlapping factor, exactly as the reference resolves it.
//
// f(i) = (length, ref) is a pure function of the text, so it is evaluated
// for every text position in parallel and the greedy chain i -> i+len is extracted afterwards
// (chain.cuh).  For each position the kernels search the LCP-interval ancestors of leaf r = ISA[i],
// i.e. the suffix-tree path the reference walks with level_anc (factorizer_core.hpp:70-78, :256-300):
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
// over an interval) are answered from 32-ary summary trees (one 128-byte line per tree node).
//   k_node_tables for every rank k, the LCP interval it names (previous / next smaller LCP value),
//                the minimum forward start AND the maximum rc value inside it: the suffix tree's
//                internal nodes, tabulated once so that a climb step is one 16-byte load.
//   k_lpnf_rank  one thread per suffix-array RANK (a warp shares the cache lines around its ranks):
//                ONE climb over the ancestors of its leaf meets vR (first ancestor whose R-max
//                qualifies; that R-max is the RC source) and vF (1-2 nodes on random DNA, up to the
//                copy count inside tandem arrays), then applies the selection rule.  Positions that
//                exhaust the climb budget (low-complexity text) are flagged.
//   k_lpnf_hard  text order over the flagged positions: the forward predicate is monotone in the
//                string depth D, so the answer is a binary search over D that grows interval(D)
//                incrementally from the deepest failing node; a Kasai-style carry (the match at i
//                is at least the match at i-1 minus one) makes consecutive positions O(1) probes.
//                Tree lines are read with eight independent 16-byte loads (one latency per level).
//                Also settles the RC candidate of flagged positions whose climb met no vR.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

#include "common.cuh"

namespace nlz {

namespace cg = cooperative_groups;

constexpr int TREE_MAX_LEVELS = 8;
constexpr u32 NONE_MIN = 0xFFFFFFFFu;

struct Trees {
    int nlev;                          // number of levels including level 0
    const u32* lcp[TREE_MAX_LEVELS];   // lcp[0] = LCP (n1+1 entries), lcp[t] = block minima
    u32 cntL[TREE_MAX_LEVELS];
    const u32* f[TREE_MAX_LEVELS];     // f[0] = F0 (leaf values, see k_leaf_values); f[t] = block min of F-class values
    const u32* r[TREE_MAX_LEVELS];     // r[0] = R0; r[t] = block max of R-class values (RC mode only)
    u32 cntS[TREE_MAX_LEVELS];
};

struct WalkParams {
    u32 n1;        // number of suffixes
    u32 nfac;      // positions [0, nfac) are factorized (general: L, RC: N)
    u32 N;         // RC: |S|/2 - 1
    u32 twoN;      // RC: 2N
    // distributed runs (dist.cuh): the arrays hold virtual ranks around the real ones; only ranks in
    // [real_lo, real_hi) are evaluated, and a global rank r lives at index r + rank_add (mod 2^32).
    // Single GPU: real_lo = 0, real_hi = n1, rank_add = 0.
    u32 real_lo, real_hi, rank_add;
};

// Leaf values.  Every rank k carries two 32-bit values derived from its suffix start s = SA[k] (an S-position, up
// to 33 bits on the wide distributed path -- which is why stage 3 never reads SA itself):
//   F0[k]  forward class: the T-coordinate s          (general mode: every s; RC mode: s < N), else NONE_MIN
//   R0[k]  rc class (RC mode, N < s <= 2N):  s - N in [1, N]  -- larger = smaller T-end e = 2N - s = N - R0;  0 = none
// T-coordinates stay below 2^32 for any text this build takes (nfac <= 0xFFFFFFF0), so the node tables, the summary
// trees and the per-position results are 32-bit on every path.
template <bool RC> __device__ __forceinline__ u32 f_value(u64 s, const WalkParams& p) {
    return RC ? (s < (u64)p.N ? (u32)s : NONE_MIN) : (u32)s;
}
__device__ __forceinline__ u32 r_value(u64 s, const WalkParams& p) {
    return (s > (u64)p.N && s <= 2ull * p.N) ? (u32)(s - p.N) : 0u;
}
// SAT = u32 (one GPU) or u64 (distributed, S-positions).  F0 may alias SA when SAT is u32 (in-place conversion).
template <bool RC, typename SAT>
__global__ void __launch_bounds__(256)
k_leaf_values(const SAT* SA, u32 cnt, WalkParams p, u32* F0, u32* __restrict__ R0) {
    const u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k >= cnt) return;
    const u64 s = (u64)SA[k];
    F0[k] = f_value<RC>(s, p);
    if (RC) R0[k] = r_value(s, p);
}

// ---- tree construction ----------------------------------------------------------------------
// level-1 nodes from level 0: LCP minima and, from SA, F minima / R maxima.
template <bool RC>
__global__ void __launch_bounds__(256)
k_tree_level1(const u32* __restrict__ LCP, u32 cntL0, const u32* __restrict__ F0, const u32* __restrict__ R0, u32 cntS0,
              u32* __restrict__ lcp1, u32 cntL1, u32* __restrict__ f1,
              u32* __restrict__ r1, u32 cntS1) {
    // one warp per node: coalesced 128-byte line, shuffle reduction
    const u32 node = (blockIdx.x * 256 + threadIdx.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    const u64 e = (u64)node * 32 + lane;
    if (node < cntL1) {
        u32 v = e < cntL0 ? LCP[e] : NONE_MIN;
        v = __reduce_min_sync(0xffffffffu, v);
        if (lane == 0) lcp1[node] = v;
    }
    if (node < cntS1) {
        u32 fv = NONE_MIN, rv = 0;
        if (e < cntS0) {
            fv = F0[e];
            if (RC) rv = R0[e];
        }
        fv = __reduce_min_sync(0xffffffffu, fv);
        if (RC) rv = __reduce_max_sync(0xffffffffu, rv);
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
        v = __reduce_min_sync(0xffffffffu, v);
        if (lane == 0) lcpB[node] = v;
    }
    if (node < cntSB) {
        u32 fv = e < cntSA ? fA[e] : NONE_MIN;
        u32 rv = (RC && e < cntSA) ? rA[e] : 0u;
        fv = __reduce_min_sync(0xffffffffu, fv);
        if (RC) rv = __reduce_max_sync(0xffffffffu, rv);
        if (lane == 0) { fB[node] = fv; if (RC) rB[node] = rv; }
    }
}

// ---- queries --------------------------------------------------------------------------------
// Two flavours of every query: scalar (one thread, early-exit word probes; cheap when the answer is
// a few entries away, the common case in rank order where neighbouring lanes share the lines) and
// cooperative (VEC = true: the 8 lanes of a thread_block_tile<8> call with identical arguments, each
// lane owns 16 bytes of the 128-byte node line, so a node costs one 16-byte load per lane plus a
// redux -- one memory latency per tree level at full coalescing; used where searches are long).
// All arrays are padded so a whole line can be read at the end of an array.
using Tile8 = cg::thread_block_tile<8>;
__device__ __forceinline__ Tile8 tile8() { return cg::tiled_partition<8>(cg::this_thread_block()); }

__device__ __forceinline__ u32 valid_mask(i64 count) {   // count >= 1
    return count >= 32 ? 0xFFFFFFFFu : ((1u << (u32)count) - 1u);
}
__device__ __forceinline__ u32 bits_upto(u32 hi_bit) {   // bits [0, hi_bit]
    return hi_bit >= 31 ? 0xFFFFFFFFu : ((2u << hi_bit) - 1u);
}

// bit k set iff LCP-tree node (lev, gstart + k) < d   (whole line, cooperative)
__device__ __forceinline__ u32 mask_lcp_less_v(const Trees& T, int lev, i64 gstart, u32 d) {
    const Tile8 t = tile8();
    const u32 q = t.thread_rank();
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(T.lcp[lev] + gstart) + q);
    const u32 m4 = (x.x < d ? 1u : 0u) | (x.y < d ? 2u : 0u) | (x.z < d ? 4u : 0u) | (x.w < d ? 8u : 0u);
    const u32 m = cg::reduce(t, m4 << (4 * q), cg::bit_or<u32>());
    return m & valid_mask((i64)T.cntL[lev] - gstart);
}

// Per-thread line probes (scalar flavour): a 128-byte line is read in 16-byte pieces, four entries per load and compare
// group, with an exit after every piece -- the entry-by-entry loops this replaces spent ~6 instructions per entry and up
// to 31 entries per line and level, and in k_node_tables a warp waits for its slowest lane.
// highest index <= hi (0..31) of the line whose entry is < d, or -1
__device__ __forceinline__ int line_prev_less(const u32* __restrict__ line, int hi, u32 d) {
    const uint4* q = reinterpret_cast<const uint4*>(line);
    u32 keep = (2u << (hi & 3)) - 1u;
#pragma unroll 1
    for (int c = hi >> 2; c >= 0; --c) {
        const uint4 x = __ldg(q + c);
        const u32 m4 = ((x.x < d ? 1u : 0u) | (x.y < d ? 2u : 0u) | (x.z < d ? 4u : 0u) | (x.w < d ? 8u : 0u)) & keep;
        if (m4) return 4 * c + (31 - __clz(m4));
        keep = 0xFu;
    }
    return -1;
}
// lowest index in [lo, last] (0..31) of the line whose entry is < d, or -1
__device__ __forceinline__ int line_next_less(const u32* __restrict__ line, int lo, int last, u32 d) {
    const uint4* q = reinterpret_cast<const uint4*>(line);
    u32 keep = 0xFu & ~((1u << (lo & 3)) - 1u);
    const int cl = last >> 2;
#pragma unroll 1
    for (int c = lo >> 2; c <= cl; ++c) {
        const uint4 x = __ldg(q + c);
        u32 m4 = ((x.x < d ? 1u : 0u) | (x.y < d ? 2u : 0u) | (x.z < d ? 4u : 0u) | (x.w < d ? 8u : 0u)) & keep;
        if (c == cl) m4 &= (2u << (last & 3)) - 1u;
        if (m4) return 4 * c + (__ffs(m4) - 1);
        keep = 0xFu;
    }
    return -1;
}
__device__ __forceinline__ int line_last(const Trees& T, int lev, i64 gstart) {   // last valid entry of the line
    const i64 rest = (i64)T.cntL[lev] - 1 - gstart;
    return rest > 31 ? 31 : (int)rest;
}

// largest k <= pos with LCP[k] < d   (exists: LCP[0] = 0 < d since d >= 1)
template <bool VEC>
__device__ __forceinline__ u32 find_prev_less(const Trees& T, u32 pos, u32 d) {
    i64 idx = pos;
    int lev = 0;
    if (VEC) {
        for (;;) {
            const i64 gstart = idx & ~31LL;
            const u32 m = mask_lcp_less_v(T, lev, gstart, d) & bits_upto((u32)(idx - gstart));
            if (m) { idx = gstart + (31 - __clz(m)); break; }
            idx = (gstart >> 5) - 1;      // >= 0: the group holding LCP[0] always matches
            ++lev;
        }
        while (lev > 0) {
            --lev;
            const i64 gstart = idx << 5;
            idx = gstart + (31 - __clz(mask_lcp_less_v(T, lev, gstart, d)));
        }
        return (u32)idx;
    }
    const u32* l0 = T.lcp[0];
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {        // the common case in rank order: the boundary is a neighbour
        if (l0[idx] < d) return (u32)idx;
        --idx;
    }
    for (;;) {
        const i64 gstart = idx & ~31LL;
        const int j = line_prev_less(T.lcp[lev] + gstart, (int)(idx - gstart), d);
        if (j >= 0) { idx = gstart + j; break; }
        idx = (gstart >> 5) - 1;
        ++lev;
    }
    while (lev > 0) {
        --lev;
        const i64 base = idx << 5;
        idx = base + line_prev_less(T.lcp[lev] + base, line_last(T, lev, base), d);   // exists: the parent's minimum
    }
    return (u32)idx;
}

// smallest k >= pos with LCP[k] < d   (exists: LCP[n1] = 0)
template <bool VEC>
__device__ __forceinline__ u32 find_next_less(const Trees& T, u32 pos, u32 d) {
    i64 idx = pos;
    int lev = 0;
    if (VEC) {
        for (;;) {
            const i64 gstart = idx & ~31LL;
            const u32 m = mask_lcp_less_v(T, lev, gstart, d) & ~(bits_upto((u32)(idx - gstart)) >> 1);
            if (m) { idx = gstart + (__ffs(m) - 1); break; }
            idx = (gstart >> 5) + 1;      // exists: the last group of every level holds LCP[n1] = 0
            ++lev;
        }
        while (lev > 0) {
            --lev;
            const i64 gstart = idx << 5;
            idx = gstart + (__ffs(mask_lcp_less_v(T, lev, gstart, d)) - 1);
        }
        return (u32)idx;
    }
    const u32* l0 = T.lcp[0];
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        if (l0[idx] < d) return (u32)idx;
        ++idx;
    }
    for (;;) {
        const i64 gstart = idx & ~31LL;
        const int j = line_next_less(T.lcp[lev] + gstart, (int)(idx - gstart), line_last(T, lev, gstart), d);
        if (j >= 0) { idx = gstart + j; break; }
        idx = (gstart >> 5) + 1;
        ++lev;
    }
    while (lev > 0) {
        --lev;
        const i64 base = idx << 5;
        idx = base + line_next_less(T.lcp[lev] + base, 0, line_last(T, lev, base), d);
    }
    return (u32)idx;
}

// fold entries [lob, hib] of node line (lev, gstart) into the F-min / R-max aggregates
template <bool RC, bool WANT_R, bool VEC>
__device__ __forceinline__ void agg_line(const Trees& T, const WalkParams& p, int lev, i64 gstart, u32 lob, u32 hib,
                                         u32& fmin, u32& rmax) {
    if (VEC) {
        const Tile8 t = tile8();
        const u32 q = t.thread_rank();
        const uint4 xf = __ldg(reinterpret_cast<const uint4*>(T.f[lev] + gstart) + q);
        const u32 fv4[4] = {xf.x, xf.y, xf.z, xf.w};
        u32 fv = NONE_MIN, rv = 0;
        {
            uint4 xr = make_uint4(0, 0, 0, 0);
            if (WANT_R) xr = __ldg(reinterpret_cast<const uint4*>(T.r[lev] + gstart) + q);
            const u32 rv4[4] = {xr.x, xr.y, xr.z, xr.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const u32 k = 4 * q + j;
                if (k >= lob && k <= hib) {
                    fv = min(fv, fv4[j]);
                    if (WANT_R) rv = max(rv, rv4[j]);
                }
            }
        }
        fmin = min(fmin, cg::reduce(t, fv, cg::less<u32>()));
        if (WANT_R) rmax = max(rmax, cg::reduce(t, rv, cg::greater<u32>()));
        return;
    }
    {
        // one thread: 16-byte pieces, whole pieces folded with three min / max each
        const uint4* fa = reinterpret_cast<const uint4*>(T.f[lev] + gstart);
        const uint4* ra = reinterpret_cast<const uint4*>(T.r[lev] + gstart);
        const u32 c0 = lob >> 2, c1 = hib >> 2;
#pragma unroll 1
        for (u32 c = c0; c <= c1; ++c) {
            const uint4 xf = __ldg(fa + c);
            uint4 xr = make_uint4(0, 0, 0, 0);
            if (WANT_R) xr = __ldg(ra + c);
            const u32 lo = c == c0 ? (lob & 3u) : 0u, hi = c == c1 ? (hib & 3u) : 3u;
            if (lo == 0 && hi == 3) {
                fmin = min(fmin, min(min(xf.x, xf.y), min(xf.z, xf.w)));
                if (WANT_R) rmax = max(rmax, max(max(xr.x, xr.y), max(xr.z, xr.w)));
            } else {
                const u32 fv[4] = {xf.x, xf.y, xf.z, xf.w};
                const u32 rv[4] = {xr.x, xr.y, xr.z, xr.w};
#pragma unroll
                for (u32 e = 0; e < 4; ++e) {
                    if (e >= lo && e <= hi) {
                        fmin = min(fmin, fv[e]);
                        if (WANT_R) rmax = max(rmax, rv[e]);
                    }
                }
            }
        }
    }
}

// aggregate F-min / R-max of SA over ranks [a, b] (inclusive; empty when a > b): at most two
// partial lines per level
template <bool RC, bool WANT_R, bool VEC>
__device__ __forceinline__ void agg_range(const Trees& T, const WalkParams& p, i64 a, i64 b, u32& fmin, u32& rmax) {
    int lev = 0;
    while (a <= b) {
        const i64 ga = a & ~31LL, gb = b & ~31LL;
        if (ga == gb) { agg_line<RC, WANT_R, VEC>(T, p, lev, ga, (u32)(a - ga), (u32)(b - ga), fmin, rmax); return; }
        if (a != ga) { agg_line<RC, WANT_R, VEC>(T, p, lev, ga, (u32)(a - ga), 31u, fmin, rmax); a = ga + 32; }
        if (((b + 1) & 31) != 0) { agg_line<RC, WANT_R, VEC>(T, p, lev, gb, 0u, (u32)(b - gb), fmin, rmax); b = gb - 1; }
        if (a > b) return;
        a >>= 5;
        b = ((b + 1) >> 5) - 1;
        ++lev;
    }
}

// ---- cooperative (8-lane) versions: RC candidate depth of one leaf from scratch ---------------------------------------
// bit k set iff R-tree node (lev, gstart + k) > thr   (whole line; entries past the end of the level are masked)
__device__ __forceinline__ u32 mask_r_greater_v(const Trees& T, int lev, i64 gstart, u32 thr) {
    const Tile8 t = tile8();
    const u32 q = t.thread_rank();
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(T.r[lev] + gstart) + q);
    const u32 m4 = (x.x > thr ? 1u : 0u) | (x.y > thr ? 2u : 0u) | (x.z > thr ? 4u : 0u) | (x.w > thr ? 8u : 0u);
    const u32 m = cg::reduce(t, m4 << (4 * q), cg::bit_or<u32>());
    return m & valid_mask((i64)T.cntS[lev] - gstart);
}
// largest k <= q with R0[k] > thr, or -1
__device__ __forceinline__ i64 find_prev_r_greater_v(const Trees& T, i64 q, u32 thr) {
    if (q < 0) return -1;
    int lev = 0;
    i64 idx = q;
    for (;;) {
        const i64 gstart = idx & ~31LL;
        const u32 m = mask_r_greater_v(T, lev, gstart, thr) & bits_upto((u32)(idx - gstart));
        if (m) { idx = gstart + (31 - __clz(m)); break; }
        if (gstart == 0) return -1;
        idx = (gstart >> 5) - 1;
        ++lev;
    }
    while (lev > 0) {
        --lev;
        const i64 gstart = idx << 5;
        idx = gstart + (31 - __clz(mask_r_greater_v(T, lev, gstart, thr)));   // non-empty: the parent's maximum is one of them
    }
    return idx;
}
// smallest k >= q with R0[k] > thr, or -1
__device__ __forceinline__ i64 find_next_r_greater_v(const Trees& T, i64 q, u32 thr) {
    if (q >= (i64)T.cntS[0]) return -1;
    int lev = 0;
    i64 idx = q;
    for (;;) {
        const i64 gstart = idx & ~31LL;
        const u32 m = mask_r_greater_v(T, lev, gstart, thr) & ~(bits_upto((u32)(idx - gstart)) >> 1);
        if (m) { idx = gstart + (__ffs(m) - 1); break; }
        idx = (gstart >> 5) + 1;
        ++lev;
        if (lev >= T.nlev || idx >= (i64)T.cntS[lev]) return -1;
    }
    while (lev > 0) {
        --lev;
        const i64 gstart = idx << 5;
        idx = gstart + (__ffs(mask_r_greater_v(T, lev, gstart, thr)) - 1);
    }
    return idx;
}
// minimum of entries [lob, hib] of LCP-tree line (lev, gstart)
__device__ __forceinline__ u32 lcp_line_min_v(const Trees& T, int lev, i64 gstart, u32 lob, u32 hib) {
    const Tile8 t = tile8();
    const u32 q = t.thread_rank();
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(T.lcp[lev] + gstart) + q);
    const u32 v4[4] = {x.x, x.y, x.z, x.w};
    u32 v = NONE_MIN;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const u32 k = 4 * q + j;
        if (k >= lob && k <= hib) v = min(v, v4[j]);
    }
    return cg::reduce(t, v, cg::less<u32>());
}
// min LCP[a..b] (inclusive, a <= b)
__device__ __forceinline__ u32 lcp_range_min_v(const Trees& T, i64 a, i64 b) {
    u32 m = NONE_MIN;
    int lev = 0;
    while (a <= b) {
        const i64 ga = a & ~31LL, gb = b & ~31LL;
        if (ga == gb) return min(m, lcp_line_min_v(T, lev, ga, (u32)(a - ga), (u32)(b - ga)));
        if (a != ga) { m = min(m, lcp_line_min_v(T, lev, ga, (u32)(a - ga), 31u)); a = ga + 32; }
        if (((b + 1) & 31) != 0) { m = min(m, lcp_line_min_v(T, lev, gb, 0u, (u32)(b - gb))); b = gb - 1; }
        if (a > b) return m;
        a >>= 5;
        b = ((b + 1) >> 5) - 1;
        ++lev;
    }
    return m;
}
// Depth of the deepest ancestor of leaf r that holds an rc(T) suffix whose value exceeds thr (0: none but the root) = the
// LCA of r with the nearest such rank on either side.  k_lpnf_rank finds it on its climb; this is for the positions whose
// climb ran out of budget first (inside long tandem arrays the qualifying rank can be millions of ranks away).
__device__ __forceinline__ u32 rc_depth_v(const Trees& T, u32 r, u32 thr) {
    u32 dl = 0, dr = 0;
    const i64 kl = find_prev_r_greater_v(T, (i64)r - 1, thr);
    if (kl >= 0) dl = lcp_range_min_v(T, kl + 1, (i64)r);
    const i64 kr = find_next_r_greater_v(T, (i64)r + 1, thr);
    if (kr >= 0) dr = lcp_range_min_v(T, (i64)r + 1, kr);
    return max(dl, dr);
}

// ---- the factor rule ------------------------------------------------------------------------
// LR[i] = ref << 32 | len for every factorized position i; FLAGS[i] (one byte): bit 0 = "hard" (left to k_lpnf_hard),
// bit 1 = the factor is a reverse-complement one.  (ref and len both need 32 bits at genome scale -- refs of a 3.1 Gbp
// text exceed 2^31 -- so the RC flag cannot ride in either.)
constexpr u8 FLAG_HARD = 1, FLAG_RC = 2;
// Ancestors climbed in rank order before a position is "hard" (left to the depth search of k_lpnf_hard).  Chromosome-scale
// texts: nearly every position that needs more than 64 climbs needs thousands (megabase tandem arrays) -- 250 Mbp text,
// budget 512 -> 64: lpnf_rank 38 -> 32 ms with 1.5 % more hard positions; below 24 the depth search costs more than the
// climbs save.  Small texts: their arrays have a few hundred copies, which the climb finishes cheaper than the depth
// search (5 Mbp text: budget 64 costs 0.9 ms of k_lpnf_hard to save 0.1 ms of climbs).
constexpr int WALK_MAX_NODES = 512;
constexpr int WALK_MAX_NODES_LARGE = 64;
constexpr u32 WALK_LARGE_TEXT = 1u << 26;   // suffixes
constexpr int WALK_Q = 32;             // consecutive text positions per 8-lane tile in k_lpnf_hard (the carried bound
                                       // links them: the first one pays a full bisection, the others 2-3 probes)

struct NodeState {
    u32 lo, hi;   // rank interval
    u32 F;        // min of F-class suffix starts in the interval
    u32 R;        // max of R-class suffix starts in the interval (only when asked for)
};

// State of interval(D) = maximal rank interval around `s` whose LCP values are all >= D, for any
// D <= depth(s): grows the interval and folds only the newly covered ranks into the aggregates.
template <bool RC, bool WANT_R, bool VEC>
__device__ __forceinline__ NodeState extend_to(const Trees& T, const WalkParams& p, NodeState s, u32 D) {
    const u32* LCP = T.lcp[0];
    const u32 nlo = (LCP[s.lo] >= D) ? find_prev_less<VEC>(T, s.lo - 1, D) : s.lo;
    const u32 nhi = (LCP[s.hi + 1] >= D) ? find_next_less<VEC>(T, s.hi + 2, D) - 1 : s.hi;
    agg_range<RC, WANT_R, VEC>(T, p, (i64)nlo, (i64)s.lo - 1, s.F, s.R);
    agg_range<RC, WANT_R, VEC>(T, p, (i64)s.hi + 1, (i64)nhi, s.F, s.R);
    s.lo = nlo;
    s.hi = nhi;
    return s;
}

// forward predicate of the reference (factorizer_core.hpp:75 and :266): min start + depth <= i
__device__ __forceinline__ bool pred_f(const NodeState& s, u32 D, u32 i) {
    return s.F != NONE_MIN && (u64)s.F + D <= (u64)i;
}

// Selection between forward candidate, RC candidate and literal (factorizer_core.hpp:335-365), or the
// general-mode answer (:82-108).  The RC source (largest rc start = smallest T-end inside the chosen
// node, :290-299) is a range max over interval(dR) and is only evaluated when the RC candidate wins.
template <bool RC, bool VEC>
__device__ __forceinline__ u64 select_factor(const Trees& T, const WalkParams& p, u32 r, u32 i, bool have_f,
                                             u32 fwd_len, u32 jF, u32 gen_len, u32 gen_ref, u32 dR, bool& is_rc) {
    u32 len, ref;
    is_rc = false;
    if (!RC) {
        if (gen_len >= 1) { len = gen_len; ref = gen_ref; }
        else { len = 1; ref = i; }
    } else {
        const bool have_r = dR >= 1;
        const u32 rc_len = dR;                                                  // :328-330
        bool use_fwd = false, use_lit = false;
        if (have_f && fwd_len >= 1) use_fwd = !(have_r && rc_len > fwd_len);    // :338-344
        else if (!(have_r && rc_len > 1)) use_lit = true;                       // :346-351
        if (use_lit) { len = 1; ref = i; }
        else if (use_fwd) { len = fwd_len; ref = jF; }
        else {
            NodeState leaf;
            leaf.lo = r; leaf.hi = r; leaf.F = i; leaf.R = 0;
            const u32 mR = extend_to<RC, true, VEC>(T, p, leaf, dR).R;
            const u32 e = p.N - mR;                              // smallest RC end in T coordinates (R0 = s - N, e = 2N - s)
            len = rc_len;
            ref = e - rc_len + 1;                                // :362-364 (start-anchored; the RC flag travels in the flag plane)
            is_rc = true;
        }
    }
    return ((u64)ref << 32) | (u64)len;
}

// ---- per-node table ----------------------------------------------------------------------------
// Every rank k with LCP[k] > 0 names the LCP interval ("node") [a, b1-1] of string depth LCP[k] (a / b1 =
// nearest strictly smaller LCP value to the left / right).  NODE[k] = {rank that names the parent node, minimum
// forward start inside the node, string depth, maximum rc value inside the node (RC mode)}: the suffix tree's internal
// nodes, tabulated once, so that a leaf climbs to the root with ONE 16-byte load per ancestor and meets both candidates
// on the way.  The parent of [a, b1-1] is named by a if LCP[a] >= LCP[b1], else by b1.  Ranks with LCP 0 (and the guard
// entry n1) name the root: depth 0.
// Rank order, one thread per rank: neighbouring lanes name nested or identical nodes and share the cache lines.  Interval
// sizes follow a 1/s law -- a third of the ranks name a node that reaches more than 16 entries to one side, every warp
// holds several, and the warp waits for them: the searches and the range aggregate read the lines in 16-byte pieces
// (line_prev_less / line_next_less / agg_line).  Tried and dropped (r2, 250 Mbp text, 33 ms for this kernel with scalar
// entry-by-entry loops): short scans for every rank + a shared-memory queue of the rest served by 8-lane tiles (43 ms: the
// four tiles of a warp diverge and serialise) or by whole warps (77 ms: one dependent chain per warp instead of 32), and
// the same scans + a compacted list for a second kernel (33 ms: the list's entries cost ~3700 instructions each).
template <bool RC, bool WANT_R>
__global__ void __launch_bounds__(256)
k_node_tables(Trees T, WalkParams p, uint4* __restrict__ NODE) {
    const u32 k = blockIdx.x * 256 + threadIdx.x;
    if (k > p.n1) return;
    const u32 d = (k == 0 || k == p.n1) ? 0u : T.lcp[0][k];
    if (d == 0) { NODE[k] = make_uint4(k, NONE_MIN, 0u, 0u); return; }
    const u32 a = find_prev_less<false>(T, k - 1, d);
    const u32 b1 = find_next_less<false>(T, k + 1, d);
    u32 fmin = NONE_MIN, rmax = 0;
    agg_range<RC, WANT_R, false>(T, p, (i64)a, (i64)b1 - 1, fmin, rmax);
    const u32 la = T.lcp[0][a], lb = T.lcp[0][b1];
    NODE[k] = make_uint4(la >= lb ? a : b1, fmin, d, rmax);
}

// ---- ranks to evaluate ---------------------------------------------------------------------------
// RC mode: half of the ranks hold rc(T) suffixes and have no factor to compute.  The ranks that do
// (real rank, suffix start < nfac) are compacted, in rank order, so that every lane of k_lpnf_rank works.
// MODE 1: per-CTA counts; MODE 2: ordered write (cta_off = exclusive scan of the counts).
constexpr int FR_TILE = 2048;
template <int MODE>
__global__ void __launch_bounds__(256)
k_forward_ranks(const u32* __restrict__ F0, WalkParams p, u32* __restrict__ cta_off, u32* __restrict__ list) {
    __shared__ u32 wcnt[8];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 run = MODE == 2 ? cta_off[blockIdx.x] : 0u;
#pragma unroll 1
    for (int t = 0; t < FR_TILE / 256; ++t) {
        const u32 r = blockIdx.x * FR_TILE + t * 256 + threadIdx.x;
        const bool keep = r >= p.real_lo && r < p.real_hi && F0[r] < p.nfac;
        const u32 bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcnt[w] = __popc(bal);
        __syncthreads();
        u32 before = 0, tot = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const u32 c = wcnt[i]; if (i < (int)w) before += c; tot += c; }
        if (MODE == 2 && keep) list[run + before + __popc(bal & lanemask_lt())] = r;
        run += tot;
        __syncthreads();
    }
    if (MODE == 1 && threadIdx.x == 0) cta_off[blockIdx.x] = run;
}

// ---- kernel 1: rank order, one climb settles both candidates -------------------------------------------------------
// One thread per suffix-array rank that holds a factorized position (RC mode: the compacted forward ranks).  The leaf
// climbs its tabulated ancestors: the first (deepest) one whose R-max exceeds N - i is the RC candidate node vR (deepest
// ancestor holding an rc(T) suffix whose T-end lies before i, factorizer_core.hpp:269-271) and that R-max is the RC source
// (:290-299); the first one with F-min + depth <= i is vF (:264-266).  Nothing above vF can beat the forward candidate
// (fwd_len >= depth(vF), forward wins ties), so the climb ends at vF.  (Round 1 found the RC candidate by hopping over the
// neighbouring rc(T) ranks -- two segmented scans to prepare, up to 24 hops and a scalar tree search per lane, 4200 issue
// slots per warp against a few hundred now.)  A position whose climb runs out of budget is hard; if no qualifying ancestor
// was met by then its RC depth is left to k_lpnf_hard (DR_UNRESOLVED).
// BYLIST (distributed runs): results are indexed by the thread's work index t (the position in the compacted list of
// forward ranks, or r - real_lo without a list) instead of by the text position -- the text positions of a rank range
// are scattered over the whole text, and the results travel to their position owners as (position, value) records.
constexpr u32 DR_UNRESOLVED = 0xFFFFFFFFu;
// (The kernel waits on dependent loads -- list -> leaf -> ancestor -> ancestor ...; r2 profile: 50 stall cycles per issue
// slot, a quarter of the issue slots used at full occupancy.  Two leaves per thread climbing in lockstep, their ancestor
// loads in flight together, gave 17.8 -> 17.3 ms on the 250 Mbp text and 0.32 -> 0.53 ms on the 5 Mbp one: the bound is the
// random-access rate of the table, and the longer loop body costs more than the overlap gains.  One leaf per thread.)
template <bool RC, bool BYLIST>
__global__ void __launch_bounds__(256, 8)
k_lpnf_rank(Trees T, WalkParams p, const uint4* __restrict__ NODE, const u32* __restrict__ list,
             const u32* __restrict__ nlist, int max_nodes, u64* __restrict__ LR, u8* __restrict__ HARD,
             unsigned long long* __restrict__ counters) {
    const u32 t = blockIdx.x * 256 + threadIdx.x;
    u32 r = list ? 0xFFFFFFFFu : t + (BYLIST ? p.real_lo : 0u);
    if (list && t < *nlist) r = list[t];                    // compacted forward ranks (RC mode)
    u32 visited = 0, hard = 0;
    const u32* LCP = T.lcp[0];
    const u32* F0 = T.f[0];
    u32 i = 0xFFFFFFFFu;
    if (r >= p.real_lo && r < p.real_hi) i = F0[r];         // r = 0xFFFFFFFF: no work
    if (i < p.nfac) {
        const u32 o = BYLIST ? t : i;                       // where this position's results go
        bool have_f = false, at_root = false;
        u32 dF = 0, jF = 0, belowF = i;  // deepest ok-forward node: depth, min start, F-min of its path child
        u32 childF = i;                  // F-min of the last node that failed (starts at the leaf)
        u32 dR = 0, mR = 0;              // RC candidate node: depth and R-max (0: not met yet)
        const u32 thr = p.N - i;         // an rc suffix qualifies when its T-end N - R0 is < i
        {
            const u32 dl = LCP[r], dh = LCP[r + 1];
            u32 k = dl >= dh ? r : r + 1;            // names the parent of the leaf
#pragma unroll 1
            for (int step = 0;; ++step) {
                const uint4 nd = __ldg(NODE + k);    // {parent, F-min, depth, R-max}
                const u32 d = nd.z;
                if (d == 0) { at_root = true; break; }
                if (step == max_nodes) break;
                ++visited;
                if (RC && dR == 0 && nd.w > thr) { dR = d; mR = nd.w; }          // factorizer_core.hpp:269-271
                const u32 m = nd.y;
                if (m != NONE_MIN && (u64)m + d <= (u64)i) {                      // :75 / :264-266
                    have_f = true; dF = d; jF = m; belowF = childF;
                    break;
                }
                childF = m;
                k = nd.x;
            }
        }
        if (have_f || at_root) {
            u32 len, ref;
            bool is_rc = false;
            u32 gen_len, gen_ref, fwd_len = 0;
            if (have_f) {
                const u32 part = (belowF != i) ? i - belowF : 0;
                if (part > dF) { gen_len = part; gen_ref = belowF; }    // :104-107
                else { gen_len = dF; gen_ref = jF; }                    // :89-94, :98-102
                fwd_len = (belowF == jF) ? (i - jF) : dF;               // :322-326
            } else {
                gen_len = (childF != i) ? i - childF : 0;               // :96-107 with u = root, or literal
                gen_ref = childF;
            }
            if (!RC) {
                if (gen_len >= 1) { len = gen_len; ref = gen_ref; }
                else { len = 1; ref = i; }
            } else {
                const bool have_r = dR >= 1;
                bool use_fwd = false, use_lit = false;
                if (have_f && fwd_len >= 1) use_fwd = !(have_r && dR > fwd_len);   // :338-344
                else if (!(have_r && dR > 1)) use_lit = true;                      // :346-351
                if (use_lit) { len = 1; ref = i; }
                else if (use_fwd) { len = fwd_len; ref = jF; }
                else {
                    const u32 e = p.N - mR;                  // smallest RC end in T coordinates (R0 = s - N, e = 2N - s)
                    len = dR;
                    ref = e - dR + 1;                        // :362-364
                    is_rc = true;
                }
            }
            LR[o] = ((u64)ref << 32) | (u64)len;
            if (is_rc) HARD[o] = FLAG_RC;           // the plane is zeroed beforehand: most positions need no (scattered) store
        } else {
            // parked for k_lpnf_hard: a depth known to satisfy the forward predicate (see k_lpnf_rank) and the RC depth
            const u32 lb0 = (childF != NONE_MIN && childF < i) ? i - childF : 0u;
            LR[o] = ((u64)lb0 << 32) | (u64)((RC && dR == 0) ? DR_UNRESOLVED : dR);
            HARD[o] = FLAG_HARD;
            hard = 1;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        visited += __shfl_xor_sync(0xffffffffu, visited, o);
        hard += __shfl_xor_sync(0xffffffffu, hard, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (visited) atomicAdd(counters, (unsigned long long)visited);
        if (hard) atomicAdd(counters + 1, (unsigned long long)hard);
    }
}

// ---- kernel 2: text order over the hard positions, one 8-lane tile per run ----------------------
// Largest D in [loD, hiD) that satisfies the forward predicate (loD: known true, or 0; hiD: known
// false; `cur`: a failing state of depth >= hiD).  U = interval(D*) when D* > initial loD,
// L = interval(D*+1).  Every probe extends the deepest known-failing node.
template <bool RC>
__device__ __forceinline__ u32 depth_search(const Trees& T, const WalkParams& p, u32 i, const NodeState& cur,
                                            u32 loD, u32 hiD, NodeState& U, NodeState& L, u32& probes) {
    L = cur;
    // Gallop upwards from the known-true bound first: along consecutive positions the carried bound is within a few
    // symbols of the answer (inside a tandem array the answer is constant over a period and the upper bound is the
    // rest of the array: ~20 bisection probes, each a fresh interval search, against 2-3 galloping ones).
    for (u32 step = 1; loD >= 1 && hiD - loD > 1; step <<= 1) {     // no carried bound (loD = 0): bisect at once
        const u32 mid = loD + step;
        if (mid >= hiD || step >= (1u << 30)) break;
        const NodeState cand = extend_to<RC, false, true>(T, p, L, mid);
        ++probes;
        if (pred_f(cand, mid, i)) { loD = mid; U = cand; }
        else { hiD = mid; L = cand; break; }
    }
    while (hiD - loD > 1) {
        const u32 mid = loD + ((hiD - loD) >> 1);
        const NodeState cand = extend_to<RC, false, true>(T, p, L, mid);
        if (pred_f(cand, mid, i)) { loD = mid; U = cand; }
        else { hiD = mid; L = cand; }
        ++probes;
    }
    return loD;
}

// BYLIST (distributed runs): the work items are the entries of k_lpnf_rank's list (rank order; `nwork` of them, or
// *nlist), the leaf is list[t] and no RANK lookup is needed; consecutive items are unrelated text positions, so there is
// no carried bound -- the search starts from the depth k_lpnf_rank parked.
template <bool RC, bool BYLIST>
__global__ void __launch_bounds__(256)
k_lpnf_hard(Trees T, WalkParams p, const u32* __restrict__ RANK, const u32* __restrict__ list, const u32* __restrict__ nlist,
            u32 nwork, u64* __restrict__ LR, u8* __restrict__ HARD, unsigned long long* __restrict__ counters) {
    const Tile8 t8 = tile8();
    const u64 tile_id = ((u64)blockIdx.x * 256 + threadIdx.x) >> 3;
    const u64 i0 = tile_id * WALK_Q;
    if (BYLIST && nlist) nwork = *nlist;
    const u64 bound = BYLIST ? (u64)nwork : (u64)p.nfac;
    if (i0 >= bound) return;                                    // tile-uniform
    const u32 q = t8.thread_rank();
    u32 todo = 0;                                               // bit k: item i0 + k is hard
#pragma unroll
    for (int j = 0; j < WALK_Q / 8; ++j) {
        const u64 ii = i0 + (u64)(j * 8) + q;
        todo |= t8.ballot(ii < bound && (HARD[ii] & FLAG_HARD) != 0) << (8 * j);
    }
    if (!todo) return;
    const u32* LCP = T.lcp[0];
    u32 visited = 0;
    u32 prevF = 0;               // true longest non-overlapping forward match of the previous position
    int prev_k = -2;
    while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        if (BYLIST || k != prev_k + 1) prevF = 0;                // the carry only links consecutive positions
        prev_k = k;
        const u32 o = (u32)(i0 + k);                             // where the item's results live
        const u32 r = BYLIST ? (list ? __ldg(list + o) : o + p.real_lo) : __ldg(RANK + o) + p.rank_add;
        const u32 i = BYLIST ? __ldg(T.f[0] + r) : o;
        NodeState leaf;
        leaf.lo = r; leaf.hi = r; leaf.F = i; leaf.R = 0;
        const u32 Dtop = max(__ldg(LCP + r), __ldg(LCP + r + 1)) + 1;   // deeper than the leaf's parent nothing matches
        const u64 parked = LR[o];                                // {depth known to hold, RC candidate depth} from k_lpnf_rank
        u32 lb = prevF > 0 ? prevF - 1 : 0;                      // Kasai-style lower bound (known to hold)
        lb = max(lb, (u32)(parked >> 32));
        NodeState U = leaf, L = leaf;
        u32 Ds;
        if (lb >= 1 && lb + 1 < Dtop) {
            const NodeState st = extend_to<RC, false, true>(T, p, leaf, lb + 1);
            ++visited;
            if (pred_f(st, lb + 1, i)) {
                U = st;
                Ds = depth_search<RC>(T, p, i, leaf, lb + 1, Dtop, U, L, visited);
            } else {
                Ds = lb; L = st;
                U = extend_to<RC, false, true>(T, p, st, lb);
                ++visited;
            }
        } else if (lb >= 1) {                                    // lb + 1 == Dtop
            Ds = lb; L = leaf;
            U = extend_to<RC, false, true>(T, p, leaf, lb);
            ++visited;
        } else {
            Ds = depth_search<RC>(T, p, i, leaf, 0, Dtop, U, L, visited);
        }
        prevF = Ds;
        bool have_f = false;
        u32 fwd_len = 0, jF = 0, gen_len = 0, gen_ref = i;
        if (Ds >= 1) {
            gen_len = Ds; gen_ref = U.F;
            if (RC) {
                if (U.lo != L.lo || U.hi != L.hi) {              // U is a node of depth Ds, L its path child
                    have_f = true; jF = U.F;
                    fwd_len = (L.F == U.F) ? (i - jF) : Ds;
                } else {                                         // Ds lies inside U's edge: vF = parent(U)
                    const u32 du = max(__ldg(LCP + U.lo), __ldg(LCP + U.hi + 1));
                    if (du > 0) {
                        const NodeState P = extend_to<RC, false, true>(T, p, U, du);
                        ++visited;
                        have_f = true; jF = P.F;
                        fwd_len = (U.F == P.F) ? (i - jF) : du;
                    }
                }
            }
        }
        u32 dR = (u32)parked;
        if (RC && dR == DR_UNRESOLVED) {
            // The climb of k_lpnf_rank met no qualifying ancestor before its budget ran out.  An RC candidate only counts
            // when it is deeper than the forward one, and every ancestor deeper than Ds lies inside L = interval(Ds + 1):
            // without a qualifying rc suffix in L the candidate depth is at most Ds (at most depth(vF) when Ds lies inside an
            // edge) <= fwd_len, forward or literal wins whatever the exact value -- the common case inside tandem arrays, where
            // the nearest qualifying rank is millions of ranks away.  Otherwise the exact depth comes from the nearest
            // qualifying rank on either side.
            const u32 thr = p.N - i;
            u32 fL = NONE_MIN, rL = 0;
            agg_range<RC, true, true>(T, p, (i64)L.lo, (i64)L.hi, fL, rL);
            ++visited;
            dR = 0;
            if (rL > thr) { dR = rc_depth_v(T, r, thr); ++visited; }
        }
        bool is_rc;
        const u64 lr = select_factor<RC, true>(T, p, r, i, have_f, fwd_len, jF, gen_len, gen_ref, dR, is_rc);
        t8.sync();                                               // every lane has read the parked value
        if (q == 0) { LR[o] = lr; HARD[o] = is_rc ? FLAG_RC : 0; }
    }
    if (q == 0 && visited) atomicAdd(counters, (unsigned long long)visited);
}

}  // namespace nlz
