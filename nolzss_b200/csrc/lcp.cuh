// Stage 2: LCP array, LCP[r] = lcp(SA[r-1], SA[r]), in rank order.
//
// Replaces the LCP / tree-depth support of the reference's sdsl::cst_sada (queries cst.depth(),
// cst.lca() at /root/reference/src/cpp/factorizer_core.hpp:73,258 and factorizer_helpers.hpp:20-24).
//
// Three kernels, so that every random memory access sits in a kernel without a dependent chain behind it:
//   k_phi_gather   PHI[i] = SA[RANK[i] - 1]            (coalesced over i, one random read; fully parallel)
//   k_lcp_kasai    PLCP[i] = lcp(i, PHI[i])            (text order: Kasai's invariant PLCP[i] >= PLCP[i-1] - 1 is kept
//                                                       per lane over a run of consecutive positions, so the worst case
//                                                       stays O(n + n/Q * maxLCP/8); one random read -- the partner's
//                                                       text -- per position, coalesced PHI reads and PLCP writes)
//   k_lcp_scatter  LCP[RANK[i]] = PLCP[i]              (coalesced reads, one random write; fully parallel)
// Round 1 did all of it in the Kasai kernel: RANK[i] -> SA[r-1] -> text -> LCP[r] is a chain of three dependent random
// accesses per position behind a per-lane serial loop, and the kernel was bound by that latency (21 % of DRAM
// throughput, 250 Mbp text: 55 ms).  The distributed path (dist2.cuh) has the same three phases with an exchange
// where the gather and the scatter are.
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

// "no predecessor" (global rank 0) in a PHI array
template <typename PT> struct PhiNone;
template <> struct PhiNone<u32> { static constexpr u32 value = 0xFFFFFFFFu; };
template <> struct PhiNone<u64> { static constexpr u64 value = (1ull << 34) - 1; };

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

__global__ void __launch_bounds__(256)
k_phi_gather(const u32* __restrict__ SA, const u32* __restrict__ RANK, u32 n1, u32* __restrict__ PHI) {
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n1) return;
    const u32 r = RANK[i];
    PHI[i] = r ? __ldg(SA + (r - 1)) : PhiNone<u32>::value;
}
__global__ void __launch_bounds__(256)
k_lcp_scatter(const u32* __restrict__ PLCP, const u32* __restrict__ RANK, u32 n1, u32* __restrict__ LCP) {
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i == 0) LCP[n1] = 0;                       // right guard used by the interval walks
    if (i >= n1) return;
    LCP[RANK[i]] = PLCP[i];
}

// Each lane runs Kasai over LCP_Q consecutive text positions; a warp therefore owns 32*LCP_Q consecutive positions
// of the slice [pos0, pos1) (one GPU: the whole text).  Per position the lane continues from l-1 on its own for up to
// LCP_LOCAL_WORDS words; a lane whose match still goes on hands the comparison to its warp (32 words per step).  At the
// run starts (k = 0) the lanes that need help are served in lane order, and Kasai's inequality PLCP[i+d] >= PLCP[i] - d
// also links the starts of neighbouring lanes (d = LCP_Q): each starts LCP_Q symbols short of its predecessor's result,
// so a 100-kbp tandem repeat costs the warp one long comparison instead of 32.
// BATCH: suffixes of different records share nothing, and a match stops at the sentinel that ends either segment
// (batch sentinels all carry the same byte, so the raw compare alone would run on).
template <typename PT>
struct LcpSlice {
    const PT* PHI;     // PHI[i - pos0] = S-position of the suffix that precedes suffix i in rank order
    u32* PLCP;         // out, same indexing
    u64 pos0, pos1;
};

template <bool BATCH, typename PT>
__global__ void __launch_bounds__(256)
k_lcp_kasai(const u8* __restrict__ x, u64 L, BatchView bv, LcpSlice<PT> ld) {
    const u64* xw = reinterpret_cast<const u64*>(x);
    const u32 lane = threadIdx.x & 31;
    const u64 c = (u64)blockIdx.x * 256 + threadIdx.x;
    const u64 i0 = ld.pos0 + c * LCP_Q;
    const u64 n1 = ld.pos1;
    if (i0 - (u64)lane * LCP_Q >= n1) return;      // whole warp out of range (warp-uniform)
    u32 l = 0;
#pragma unroll 1
    for (int k = 0; k < LCP_Q; ++k) {
        const u64 i = i0 + k;
        PT j = 0;
        u32 maxl = 0;
        bool need = false;
        if (i < n1) {
            j = ld.PHI[i - ld.pos0];
            if (j == PhiNone<PT>::value) { ld.PLCP[i - ld.pos0] = 0; l = 0; }
            else {
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
            u32 l0 = __shfl_sync(0xffffffffu, l, src);
            const u32 ml = __shfl_sync(0xffffffffu, maxl, src);
            if (k == 0 && src > 0) {                // run starts: the previous lane's (final) result bounds this one from below
                const u32 carry = __shfl_sync(0xffffffffu, l, src - 1);
                const u32 lb = carry > (u32)LCP_Q ? carry - LCP_Q : 0u;
                if (lb > l0) l0 = lb < ml ? lb : ml;
            }
            const u32 res = warp_extend_match(xw, a, b, l0, ml, lane);
            if ((int)lane == src) l = res;
            pending &= pending - 1;
        }
        if (need) {
            ld.PLCP[i - ld.pos0] = l;
            if (l) --l;
        }
    }
}

}  // namespace nlz
