// Stage 2: LCP array, LCP[r] = lcp(SA[r-1], SA[r]), in rank order.
//
// Replaces the LCP / tree-depth support of the reference's sdsl::cst_sada (queries cst.depth(),
// cst.lca() at /root/reference/src/cpp/factorizer_core.hpp:73,258 and factorizer_helpers.hpp:20-24).
//
// Most LCP values never touch the text: after the initial radix sort the packed-prefix keys of adjacent suffixes are at
// hand, and whenever they differ (or hold a sentinel) their common prefix IS the LCP (key_pair_lcp in sa.cuh, written
// by the first regroup, coalesced).  Only the members of tie groups -- the suffixes that go through prefix doubling,
// i.e. the repeats: 22 % of the 250 Mbp text -- share their whole window with a neighbour; they are marked (NEED, one byte per text position) and resolved here by a Kasai pass in text order:
// Kasai's invariant PLCP[i] >= PLCP[i-1] - 1 is kept per lane over a run of consecutive positions, so the worst case
// stays O(n + n/Q * maxLCP/8); marked positions cluster (repeats are contiguous in the text), unmarked stretches cost
// one 32-byte flag load per lane and a ballot per warp.  Every marked position costs three random accesses
// (RANK[i] -> SA[r-1] -> the partner's text; the LCP[r] store) -- at chromosome scale the stage is bound by the
// GPU's random-access rate (~27 G sectors/s over multi-GB arrays), which is why the count matters: round 1 paid them
// for EVERY position (250 Mbp text: 55 ms).  The distributed path (dist2.cuh) runs the same kernel on a position slice
// with Phi delivered by the rank owners (PHIIN) and PLCP sent back, again for the marked positions only.
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

// 32 lanes compare x[a+l0 ..) with x[b+l0 ..): the first step covers 32 consecutive 8-byte words (two coalesced 256-byte
// reads) -- most calls end there --, every further step four such blocks with all eight loads of a lane in flight before
// the first ballot: a match of megabytes (tandem arrays) is a chain of dependent steps, and it is the load latency per
// step, not the bandwidth, that a warp waits for.  All lanes pass the same arguments.
__device__ __forceinline__ u32 warp_extend_match(const u64* __restrict__ xw, u64 a, u64 b, u32 l0, u32 ml, u32 lane) {
    if (l0 >= ml) return ml;
    {
        const u32 my = l0 + 8 * lane;
        u64 d = 0;
        if (my < ml) d = load8_unaligned(xw, a + my) ^ load8_unaligned(xw, b + my);
        const u32 hb = __ballot_sync(0xffffffffu, d != 0);
        if (hb) {
            const int f = __ffs(hb) - 1;
            const u64 df = __shfl_sync(0xffffffffu, d, f);
            const u32 res = l0 + 8 * f + ((u32)(__ffsll((long long)df) - 1) >> 3);
            return res > ml ? ml : res;
        }
    }
    for (u32 off = l0 + 256; off < ml; off += 1024) {
        u64 d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u32 my = off + 256 * q + 8 * lane;
            d[q] = (my < ml && my >= off) ? (load8_unaligned(xw, a + my) ^ load8_unaligned(xw, b + my)) : 0ull;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const u32 hb = __ballot_sync(0xffffffffu, d[q] != 0);
            if (hb) {
                const int f = __ffs(hb) - 1;
                const u64 df = __shfl_sync(0xffffffffu, d[q], f);
                const u32 res = off + 256 * q + 8 * f + ((u32)(__ffsll((long long)df) - 1) >> 3);
                return res > ml ? ml : res;
            }
        }
        if (off + 1024 < off) break;               // (u32 wrap guard; ml < 2^32 - 16)
    }
    return ml;
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
    const u8* NEED;    // NEED[i - pos0] != 0: position i is resolved here
    // PHIIN (distributed): PHI[i - pos0] = S-position of the suffix that precedes suffix i in rank order; out: PLCP[i - pos0]
    const PT* PHI;
    u32* PLCP;
    // !PHIIN (one GPU): the predecessor is SA[RANK[i] - 1]; out: LCP[RANK[i]]
    const u32* SA;
    const u32* RANK;
    u32* LCP;
    u64 pos0, pos1;
    u32 blocks_per_warp;   // consecutive 1024-position blocks a warp works through, carrying Kasai's bound across them
};
// A warp inside a megabase tandem array that starts from nothing re-scans the rest of the array for its first position
// (milliseconds of dependent steps); carried across consecutive blocks that scan happens once per blocks_per_warp * 1024
// positions.  Small texts keep one block per warp (parallelism first).
static inline u32 lcp_blocks_per_warp(u64 positions) {
    u64 b = positions / (1024ull * 6000ull);       // keep >= ~6000 warps (most exit at once: unmarked blocks)
    return (u32)(b < 1 ? 1 : (b > 32 ? 32 : b));
}

template <bool BATCH, bool PHIIN, typename PT>
__global__ void __launch_bounds__(256)
k_lcp_kasai(const u8* __restrict__ x, u64 L, BatchView bv, LcpSlice<PT> ld) {
    const u64* xw = reinterpret_cast<const u64*>(x);
    const u32 lane = threadIdx.x & 31;
    const u64 warp_id = ((u64)blockIdx.x * 256 + threadIdx.x) >> 5;
    const u64 n1 = ld.pos1;
    u32 carry_in = 0;                               // Kasai's bound for the first position of the next block
#pragma unroll 1
    for (u32 blk = 0; blk < ld.blocks_per_warp; ++blk) {
    const u64 i0 = ld.pos0 + (warp_id * ld.blocks_per_warp + blk) * (32ull * LCP_Q) + (u64)lane * LCP_Q;
    if (i0 - (u64)lane * LCP_Q >= n1) return;      // whole warp out of range (warp-uniform)
    // the marks of this lane's run: 32 bytes (pos0 and the runs are 32-aligned; the array is padded)
    u32 todo = 0;
    if (i0 < n1) {
        const uint4* f4 = reinterpret_cast<const uint4*>(ld.NEED + (i0 - ld.pos0));
        const uint4 fa = __ldg(f4), fb = __ldg(f4 + 1);
        const u32 wds[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int bq = 0; bq < 4; ++bq) todo |= ((wds[q] >> (8 * bq)) & 0xFFu) ? (1u << (4 * q + bq)) : 0u;
    }
    if (!__any_sync(0xffffffffu, todo != 0)) { carry_in = 0; continue; }
    u32 l = lane == 0 ? carry_in : 0u;
    u32 r = 0;
#pragma unroll 1
    for (int k = 0; k < LCP_Q; ++k) {
        const u64 i = i0 + k;
        PT j = 0;
        u32 maxl = 0;
        bool need = ((todo >> k) & 1u) != 0 && i < n1;
        if (!__any_sync(0xffffffffu, need)) { l = 0; continue; }
        if (need) {
            if (PHIIN) j = ld.PHI[i - ld.pos0];
            else { r = ld.RANK[i]; j = r ? (PT)__ldg(ld.SA + (r - 1)) : PhiNone<PT>::value; }
            if (j == PhiNone<PT>::value) {
                if (PHIIN) ld.PLCP[i - ld.pos0] = 0; else ld.LCP[0] = 0;
                l = 0; need = false;
            } else {
                const u64 room = L - (i > (u64)j ? i : (u64)j);
                maxl = room > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (u32)room;     // a match ends at a unique sentinel: < 2^32 anyway
                if (BATCH) {
                    if (bv.REC[i] != bv.REC[j]) maxl = 0;
                    else maxl = min(maxl, min(batch_cap(bv, (u32)i), batch_cap(bv, (u32)j)));
                }
                if (l > maxl) l = maxl;
            }
        } else {
            l = 0;                                  // an unmarked position breaks the carry
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
            if (PHIIN) ld.PLCP[i - ld.pos0] = l; else ld.LCP[r] = l;
            if (l) --l;
        }
    }
    carry_in = __shfl_sync(0xffffffffu, l, 31);    // bound for position (last of this block) + 1
    }
}

}  // namespace nlz
