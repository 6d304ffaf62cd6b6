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
#include "sa.cuh"

namespace nlz {

__device__ __forceinline__ u64 load8_unaligned(const u64* __restrict__ xw, u64 p) {
    const u64 q = p >> 3;
    const u32 sh = (u32)(p & 7) * 8;
    u64 lo = __ldg(xw + q);
    if (sh == 0) return lo;
    u64 hi = __ldg(xw + q + 1);
    return (lo >> sh) | (hi << (64 - sh));
}

constexpr int LCP_Q = 32;        // consecutive text positions per thread (Kasai run)
constexpr int LCP_LOCAL_WORDS = 2;  // 8-byte words a lane compares alone before asking the warp for help

// 32 lanes compare x[a+l0 ..) with x[b+l0 ..), 32 consecutive 8-byte words per step (two coalesced
// 256-byte reads), a ballot finds the first mismatch.  All lanes pass the same arguments.
__device__ __forceinline__ u32 warp_extend_match(const u64* __restrict__ xw, u64 a, u64 b, u32 l0, u32 ml, u32 lane) {
    u32 res = ml;
    for (u32 off = l0; off < ml; off += 256) {
        const u32 my = off + 8 * lane;
        u64 d = 0;
        if (my < ml) d = load8_unaligned(xw, a + my) ^ load8_unaligned(xw, b + my);
        const u32 hb = __ballot_sync(0xffffffffu, d != 0);
        if (hb) {
            const int f = __ffs(hb) - 1;
            const u64 df = __shfl_sync(0xffffffffu, d, f);
            res = off + 8 * f + ((u32)(__ffsll((long long)df) - 1) >> 3);
            if (res > ml) res = ml;
            break;
        }
    }
    return res;
}

// Each lane runs Kasai over LCP_Q consecutive text positions; a warp therefore owns 32*LCP_Q
// consecutive positions.
//  * run starts (k = 0): Kasai's inequality PLCP[i+d] >= PLCP[i] - d also links the STARTS of
//    neighbouring lanes (d = LCP_Q), so the 32 start comparisons are done by the whole warp in lane
//    order, each beginning LCP_Q symbols short of its predecessor's result: a 100-kbp tandem repeat
//    costs the warp one long comparison instead of 32.
//  * inside a run (k > 0): the lane continues from l-1 on its own; a lane whose match still outlasts
//    LCP_LOCAL_WORDS words hands the comparison to its warp.
// BATCH: suffixes of different records share nothing, and a match stops at the sentinel that ends
// either segment (batch sentinels all carry the same byte, so the raw compare alone would run on).
// DIST (one text across GPUs, dist2.cuh): this GPU owns the text positions [pos0, pos1) and the slice of RANK that
// belongs to them (global ranks, PT = u64: up to 33 bits).  PHI[i - pos0] = SA[RANK[i] - 1] (an S-position, PT) arrives
// from the owners of the ranks; the result goes to PLCP (text order) and is sent to the owners of the ranks
// afterwards (bucketed exchange).
template <typename PT>
struct LcpDistT {
    const PT* PHI;
    u32* PLCP;
    u64 pos0, pos1;
    PT rank0;      // value of global rank 0 (0, or the test hook's rank bias)
    __device__ __forceinline__ void store(u64 i, u32 l) const { PLCP[i - pos0] = l; }
};
using LcpDist = LcpDistT<u32>;

template <bool BATCH, bool DIST, typename PT = u32>
__global__ void __launch_bounds__(256)
k_lcp_kasai(const u8* __restrict__ x, u64 L, u64 n1, const u32* __restrict__ SA,
            const PT* __restrict__ RANK, u32* __restrict__ LCP, BatchView bv, LcpDistT<PT> ld) {
    const u64* xw = reinterpret_cast<const u64*>(x);
    const u32 lane = threadIdx.x & 31;
    const u64 c = (u64)blockIdx.x * 256 + threadIdx.x;
    const u64 i0 = (DIST ? (u64)ld.pos0 : 0ull) + c * LCP_Q;
    if (DIST) n1 = ld.pos1;
    if (i0 - (u64)lane * LCP_Q >= n1) return;      // whole warp out of range (warp-uniform)
    if (!DIST && c == 0) LCP[n1] = 0;              // right guard used by the interval walks
    u32 l = 0;
#pragma unroll 1
    for (int k = 0; k < LCP_Q; ++k) {
        const u64 i = i0 + k;
        PT r = 0, j = 0;
        u32 maxl = 0;
        bool need = false;
        if (i < n1) {
            r = RANK[DIST ? i - ld.pos0 : i];
            if (DIST ? r == ld.rank0 : r == 0) { if (DIST) ld.store(i, 0); else LCP[0] = 0; l = 0; }
            else {
                j = DIST ? ld.PHI[i - ld.pos0] : (PT)SA[r - 1];
                const u64 room = L - (i > (u64)j ? i : (u64)j);
                maxl = room > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (u32)room;     // a match ends at a unique sentinel: < 2^32 anyway
                if (BATCH) {
                    if (bv.REC[i] != bv.REC[j]) maxl = 0;
                    else maxl = min(maxl, min(batch_cap(bv, (u32)i), batch_cap(bv, (u32)j)));
                }
                if (l > maxl) l = maxl;
                need = true;
            }
        }
        if (k == 0) {
            u32 carry = 0;                          // result at the previous lane's run start
#pragma unroll 1
            for (int src = 0; src < 32; ++src) {
                const u64 a = __shfl_sync(0xffffffffu, i, src);
                const u64 b = __shfl_sync(0xffffffffu, (u64)j, src);
                const u32 ml = __shfl_sync(0xffffffffu, maxl, src);
                const bool nd = __shfl_sync(0xffffffffu, need ? 1u : 0u, src) != 0;
                u32 res = 0;
                if (nd) {
                    u32 l0 = carry > (u32)LCP_Q ? carry - LCP_Q : 0u;
                    if (l0 > ml) l0 = ml;
                    l0 &= ~7u;                      // keep the word grid of the first lane's offset
                    res = warp_extend_match(xw, a, b, l0, ml, lane);
                }
                if ((int)lane == src) l = res;
                carry = res;
            }
        } else {
            bool pending_me = false;
            if (need) {
                pending_me = true;
#pragma unroll
                for (int t = 0; t < LCP_LOCAL_WORDS; ++t) {
                    if (pending_me) {
                        if (l >= maxl) { l = maxl; pending_me = false; }
                        else {
                            u64 d = load8_unaligned(xw, i + l) ^ load8_unaligned(xw, (u64)j + l);
                            if (d) {
                                l += (u32)(__ffsll((long long)d) - 1) >> 3;
                                if (l > maxl) l = maxl;
                                pending_me = false;
                            } else {
                                l += 8;
                            }
                        }
                    }
                }
                if (pending_me && l >= maxl) { l = maxl; pending_me = false; }
            }
            u32 pending = __ballot_sync(0xffffffffu, pending_me);
            while (pending) {
                const int src = __ffs(pending) - 1;
                const u64 a = __shfl_sync(0xffffffffu, i, src);
                const u64 b = __shfl_sync(0xffffffffu, (u64)j, src);
                const u32 l0 = __shfl_sync(0xffffffffu, l, src);
                const u32 ml = __shfl_sync(0xffffffffu, maxl, src);
                const u32 res = warp_extend_match(xw, a, b, l0, ml, lane);
                if ((int)lane == src) l = res;
                pending &= pending - 1;
            }
        }
        if (need) {
            if (DIST) ld.store(i, l); else LCP[r] = l;
            if (l) --l;
        }
    }
}

}  // namespace nlz
