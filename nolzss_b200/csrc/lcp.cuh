// Stage 2: LCP array, LCP[r] = lcp(SA[r-1], SA[r]), in rank order.
//
// Replaces the LCP / tree-depth support of the reference's sdsl::cst_sada (queries cst.depth(),
// cst.lca() at /root/reference/src/cpp/factorizer_core.hpp:73,258 and factorizer_helpers.hpp:20-24).
// Kasai's invariant (PLCP[i] >= PLCP[i-1] - 1) is kept per thread over a short run of consecutive
// text positions, so the worst case stays O(n + n/Q * maxLCP/8) instead of the O(n * maxLCP) of a
// per-rank direct compare; RANK (= ISA, produced by stage 1) replaces the Phi array, and the result
// is scattered straight into rank order, so neither Phi nor PLCP is ever materialised.
// Symbols are compared 8 bytes at a time on the raw text: sentinel-class bytes occur once, so a
// raw-byte match can never run across one.
#pragma once
#include "common.cuh"

namespace nlz {

__device__ __forceinline__ u64 load8_unaligned(const u64* __restrict__ xw, u64 p) {
    const u64 q = p >> 3;
    const u32 sh = (u32)(p & 7) * 8;
    u64 lo = __ldg(xw + q);
    if (sh == 0) return lo;
    u64 hi = __ldg(xw + q + 1);
    return (lo >> sh) | (hi << (64 - sh));
}

// number of equal leading bytes of x[a..L) and x[b..L), given that the first l0 are known equal
__device__ __forceinline__ u32 extend_match(const u64* __restrict__ xw, u64 L, u64 a, u64 b, u32 l0) {
    const u64 hi = a > b ? a : b;
    const u32 maxl = (u32)(L - hi);
    u32 l = l0;
    while (l < maxl) {
        u64 x = load8_unaligned(xw, a + l) ^ load8_unaligned(xw, b + l);
        if (x) { l += (u32)(__ffsll((long long)x) - 1) >> 3; break; }
        l += 8;
    }
    return l < maxl ? l : maxl;
}

constexpr int LCP_Q = 16;  // consecutive text positions per thread

__global__ void __launch_bounds__(256)
k_lcp_kasai(const u8* __restrict__ x, u64 L, u32 n1, const u32* __restrict__ SA,
            const u32* __restrict__ RANK, u32* __restrict__ LCP) {
    const u64* xw = reinterpret_cast<const u64*>(x);
    const u64 c = (u64)blockIdx.x * 256 + threadIdx.x;
    u64 i = c * LCP_Q;
    if (i >= n1) return;
    u64 iend = i + LCP_Q;
    if (iend > n1) iend = n1;
    u32 l = 0;
    for (; i < iend; ++i) {
        u32 r = RANK[i];
        if (r == 0) { LCP[0] = 0; l = 0; continue; }
        u32 j = SA[r - 1];
        l = extend_match(xw, L, i, (u64)j, l);
        LCP[r] = l;
        if (l) --l;
    }
    if (c == 0) LCP[n1] = 0;  // right guard used by the interval walks
}

}  // namespace nlz
