// One text across G GPUs, second generation: EVERYTHING is partitioned, nothing but the text is replicated, and the
// index path is 64-bit clean (configs[3] and configs[4] of BASELINE.json: a 3.1 Gbp genome in RC mode indexes
// 6.2 * 10^9 suffixes -- S-positions and global ranks need 33 bits).
//
// The reference builds ONE suffix tree serially even in its parallel mode
// (/root/reference/src/cpp/parallel_factorizer.cpp:78-84) with 64-bit SDSL vectors
// (/root/reference/src/cpp/factorizer_core.hpp:195-232); there is nothing to port -- this is the north-star design:
//
//   * positions:  GPU g owns the S-positions [g*chunk, (g+1)*chunk) -- the slice of RANK (= ISA, global ranks, u64)
//                 that belongs to them, their Phi / PLCP values, and (for T-positions, chunkT) the per-position
//                 results and the chain;
//   * ranks:      GPU g owns the global ranks [base[g], base[g+1]) -- a histogram of the leading key bits (first 12
//                 bases) over every GPU's slice, summed through peer memory, gives G equal bucket ranges (sample-sort
//                 style splitters); the (key, position) pairs travel to the owner of their bucket in position order
//                 (stable), are radix-sorted there and refined by prefix doubling.  A tie group is contiguous in rank
//                 order and never leaves its GPU.
//   * exchanges:  every step that crosses the two partitions is a BUCKETED BULK EXCHANGE: items are bucketed by
//                 destination GPU into a contiguous staging list (one counting pass, one scatter pass), the G x G count
//                 matrix travels with a barrier, each bucket moves with ONE bulk copy over NVLink into the receiver's
//                 inbox, and the receiver applies the items locally.  (Scattered fine-grained peer stores collapse on
//                 this machine beyond ~1 GB of span; bulk copies run at link speed.)  Used for: the initial
//                 (key, position) pairs; refined ranks -> position owners; rank REQUESTS (position s+h -> its owner)
//                 and RESPONSES (the owner answers and pushes every answer into its requester's answer region); Phi -> position
//                 owners; PLCP -> rank owners; per-position results -> T-position owners.
//   * widths:     inside a GPU everything is 32-bit: local slots and list indices (< 2^30 per GPU), suffix HANDLES
//                 (the arrival index of a suffix at its rank owner; its S-position is pos0[sender] + OFF[handle]),
//                 T-coordinates (< 2^32).  Only the second half of a doubling key (a global rank, 33 bits: keys are
//                 group << 33 | rank, tile_sort.cuh) and the exchanged records are wider; records carry
//                 destination-LOCAL indices, so each fits one u64.
//   * stage 3:    rank space, local range, with the neighbours' edge staircases appended as virtual ranks (boundary
//                 exchange, k_dist_edges in dist.cuh); results are indexed by work item and travel to the T-position
//                 owners.
//   * chain:      position slices; the exit-node doubling follows pointers across slices through peer memory
//                 (chain.cuh, ChainDom); every GPU emits the factors of its slice, rank 0 gathers them.
#pragma once
#include "common.cuh"
#include "dist.cuh"
#include "sa.cuh"

namespace nlz {

constexpr u64 D2_MASK34 = (1ull << 34) - 1;
constexpr u64 D2_MASK33 = (1ull << 33) - 1;      // a global rank (n' < 2^33)
constexpr u64 D2_NO_PHI = D2_MASK34;            // Phi of global rank 0 (= PhiNone<u64>::value in lcp.cuh: Kasai stores 0 for it)
constexpr u32 D2_MAX_LOCAL = 0x3FFFFF00u;       // suffixes / positions one GPU can own (30-bit local indices)

// suffix handle -> S-position.  A handle is the index of the suffix in the INITIALLY SORTED list of its rank owner
// (so the active lists, which stay in slot order, look their positions up nearly sequentially -- with arrival-order
// handles every lookup was a scattered read, four per member and round).  The position is kept as the sender's GPU
// number and the offset inside that sender's position slice: 5 bytes instead of a 33-bit value in 8.
struct HandleMap {
    const u32* off;
    const u8* snd;
    u64 chunk;
    __device__ __forceinline__ u64 pos(u32 h) const { return (u64)snd[h] * chunk + off[h]; }
};
// after the initial sort: arrival index (pairs of sender g occupy [seg[g], seg[g+1])) of the e-th sorted suffix ->
// handle tables in sorted order
struct ArrivalSegs { u32 seg[MAX_PEERS + 1]; int G; };
__global__ void __launch_bounds__(256)
k_d2_sorted_handles(const u32* __restrict__ arrival_sorted, u32 cnt, const u32* __restrict__ off_arrival, ArrivalSegs sg,
                    u32* __restrict__ off_sorted, u8* __restrict__ snd_sorted) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u32 a = arrival_sorted[e];
    int g = 0;
#pragma unroll
    for (int q = 1; q < MAX_PEERS; ++q) g += (q < sg.G && a >= sg.seg[q]) ? 1 : 0;
    off_sorted[e] = off_arrival[a];
    snd_sorted[e] = (u8)g;
}

struct RankBases {
    u64 base[MAX_PEERS + 1];
    int G;
    __device__ __forceinline__ int owner(u64 r) const {
        int g = 0;
#pragma unroll
        for (int q = 1; q < MAX_PEERS; ++q) g += (q < G && r >= base[q]) ? 1 : 0;
        return g;
    }
};

// ---- partition ---------------------------------------------------------------------------------------------------
// GH[b] = sum over GPUs of their slice histograms (coalesced peer loads); flags a bucket beyond 2^32 - 1.
struct HistPeers { const u32* h[MAX_PEERS]; int G; };
__global__ void __launch_bounds__(256)
k_d2_hist_sum(HistPeers hp, u32 nb, u32* __restrict__ GH, u32* __restrict__ overflow) {
    const u32 b = blockIdx.x * 256 + threadIdx.x;
    if (b >= nb) return;
    u64 s = 0;
    for (int g = 0; g < hp.G; ++g) s += hp.h[g][b];
    if (s > 0xFFFFFFFFull) { *overflow = 1; s = 0xFFFFFFFFull; }
    GH[b] = (u32)s;
}
// tile sums (u64) of GH, 4096 buckets per tile
__global__ void __launch_bounds__(256)
k_d2_hist_tiles(const u32* __restrict__ GH, u32 nb, u64* __restrict__ TS) {
    __shared__ u64 ws[8];
    const u32 base = blockIdx.x * 4096;
    u64 s = 0;
    for (u32 i = threadIdx.x; i < 4096; i += 256) if (base + i < nb) s += GH[base + i];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { u64 t = 0; for (int i = 0; i < 8; ++i) t += ws[i]; TS[blockIdx.x] = t; }
}
// splitters: rank g owns buckets [split[g], split[g+1]) = global ranks [base[g], base[g+1]); split[g] = the largest
// bucket whose exclusive prefix is <= n1*g/G.  One warp per splitter (G <= 8 warps), two-level search.
__global__ void __launch_bounds__(288)
k_d2_splitters(const u32* __restrict__ GH, const u64* __restrict__ TS, u32 nb, u64 n1, int G, u32* __restrict__ split,
               u64* __restrict__ base) {
    const int g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (g > G) return;
    if (g == 0) { if (lane == 0) { split[0] = 0; base[0] = 0; } return; }
    if (g == G) { if (lane == 0) { split[G] = nb; base[G] = n1; } return; }
    const u64 target = n1 / (u64)G * (u64)g + (n1 % (u64)G) * (u64)g / (u64)G;
    const u32 ntiles = (nb + 4095) / 4096;
    // tile: the last tile whose exclusive prefix is <= target
    u64 run = 0;            // exclusive prefix of the current 32-tile batch
    u32 tile = 0;
    u64 tile_prefix = 0;
    for (u32 t0 = 0; t0 < ntiles; t0 += 32) {
        const u32 t = t0 + lane;
        const u64 v = t < ntiles ? TS[t] : 0;
        u64 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u64 x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        const u64 excl = run + inc - v;
        const u32 ok = __ballot_sync(0xffffffffu, t < ntiles && excl <= target);
        if (ok) {
            const int last = 31 - __clz(ok);
            tile = t0 + last;
            tile_prefix = __shfl_sync(0xffffffffu, excl, last);
        }
        run += __shfl_sync(0xffffffffu, inc, 31);
        if (ok != 0xFFFFFFFFu) break;
    }
    // bucket inside the tile
    u64 run2 = tile_prefix;
    u32 best = tile * 4096;
    u64 best_prefix = tile_prefix;
    for (u32 b0 = tile * 4096; b0 < tile * 4096 + 4096 && b0 < nb; b0 += 32) {
        const u32 b = b0 + lane;
        const u64 v = b < nb ? GH[b] : 0;
        u64 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u64 x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += x; }
        const u64 excl = run2 + inc - v;
        const u32 ok = __ballot_sync(0xffffffffu, b < nb && excl <= target);
        if (ok) {
            const int last = 31 - __clz(ok);
            best = b0 + last;
            best_prefix = __shfl_sync(0xffffffffu, excl, last);
        }
        run2 += __shfl_sync(0xffffffffu, inc, 31);
        if (ok != 0xFFFFFFFFu) break;
    }
    if (lane == 0) { split[g] = best; base[g] = best_prefix; }
}

// Keys of this GPU's position slice [pos_lo, pos_hi), one CTA per KB_TP positions (kb_stage / kb_key of sa.cuh).
//   MODE 0: histogram of the key prefixes (top pbits bits) into hist[2^pbits];
//   MODE 1: per-CTA, per-destination counts  ->  counts[dest * nctas + cta];
//   MODE 2: stable (position order) scatter of (key, offset in the slice) into the staging buckets; counts[] now holds
//           the exclusive scan of MODE 1's counts (bucket d starts at counts[d * nctas]).
struct Splits { u32 split[MAX_PEERS + 1]; int G; };
template <int MODE>
__global__ void __launch_bounds__(256)
k_d2_keys(const u8* __restrict__ x, u64 L, u64 pos_lo, u64 pos_hi, ClassTable tab, KeyLayout lay, int pbits, Splits sp,
          u32* __restrict__ hist_or_counts, u32 nctas, u64* __restrict__ keys, u32* __restrict__ offs) {
    __shared__ u8 cls[256];
    __shared__ __align__(16) u8 tile[KB_TP + KB_HALO];
    __shared__ u32 wcnt[8][MAX_PEERS];
    __shared__ u32 s_run[MAX_PEERS];
    __shared__ u64 skey[KB_SKEY];
    // kb_stage with an explicit base
    cls[threadIdx.x] = tab.cls[threadIdx.x];
    const u64 base = pos_lo + (u64)blockIdx.x * KB_TP;          // pos_lo is a multiple of KB_TP: aligned word loads
    {
        const u32* xw = reinterpret_cast<const u32*>(x + base);
        u32* tw = reinterpret_cast<u32*>(tile);
        for (int i = threadIdx.x; i < (KB_TP + KB_HALO) / 4; i += 256) {
            const u64 byte0 = base + (u64)i * 4;
            tw[i] = (byte0 < L + 64) ? xw[i] : 0u;
        }
    }
    if (MODE != 0 && threadIdx.x < MAX_PEERS) s_run[threadIdx.x] = MODE == 2 ? hist_or_counts[(size_t)threadIdx.x * nctas + blockIdx.x] : 0u;
    for (int i = threadIdx.x; i < 8 * MAX_PEERS; i += 256) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    kb_tile_keys<u64>(cls, tile, base, L, lay, skey);            // rolling windows (sa.cuh); lay.R = 0 here
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 mine[MAX_PEERS];
#pragma unroll
    for (int d = 0; d < MAX_PEERS; ++d) mine[d] = 0;
#pragma unroll 1
    for (int r = 0; r < KB_TP / 256; ++r) {
        const int o = r * 256 + threadIdx.x;
        const u64 p = base + o;
        u64 key = 0;
        int dest = -1;
        if (p < pos_hi) {
            key = skey[o + (o >> 3)];
            const u32 pre = (u32)(key >> (lay.key_bits - pbits));
            if (MODE == 0) atomicAdd(&hist_or_counts[pre], 1u);
            else {
                dest = 0;
#pragma unroll
                for (int q = 1; q < MAX_PEERS; ++q) dest += (q < sp.G && pre >= sp.split[q]) ? 1 : 0;
            }
        }
        if (MODE == 1) {
#pragma unroll
            for (int d = 0; d < MAX_PEERS; ++d) mine[d] += dest == d ? 1u : 0u;
        }
        if (MODE == 2) {
            // invalid lanes carry dest = -1: they match each other and write nothing
            const u32 same = __match_any_sync(0xffffffffu, dest);
            if (dest >= 0 && lane == (u32)(__ffs(same) - 1)) wcnt[w][dest] = __popc(same);
            __syncthreads();
            if (dest >= 0) {
                u32 before = 0;
                for (u32 i = 0; i < w; ++i) before += wcnt[i][dest];
                const u32 dst = s_run[dest] + before + __popc(same & lanemask_lt());
                keys[dst] = key;
                offs[dst] = (u32)(p - pos_lo);
            }
            __syncthreads();
            if (threadIdx.x < MAX_PEERS) {
                u32 tot = 0;
                for (int i = 0; i < 8; ++i) tot += wcnt[i][threadIdx.x];
                s_run[threadIdx.x] += tot;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < 8 * MAX_PEERS; i += 256) (&wcnt[0][0])[i] = 0;
            __syncthreads();
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int d = 0; d < MAX_PEERS; ++d) {
            u32 v = mine[d];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) wcnt[w][d] = v;
        }
        __syncthreads();
        if (threadIdx.x < MAX_PEERS) {
            u32 tot = 0;
            for (int i = 0; i < 8; ++i) tot += wcnt[i][threadIdx.x];
            hist_or_counts[(size_t)threadIdx.x * nctas + blockIdx.x] = tot;
        }
    }
}

// per-destination totals from the scanned [dest][cta] count table (exclusive scan over MAX_PEERS * nctas entries;
// `total` = the scan's grand total)
__global__ void k_d2_bucket_totals(const u32* __restrict__ scanned, u32 nctas, u32 total, int G, u32* __restrict__ cnts) {
    const int d = threadIdx.x;
    if (d >= MAX_PEERS) return;
    const u32 lo = scanned[(size_t)d * nctas];
    const u32 hi = d + 1 < MAX_PEERS ? scanned[(size_t)(d + 1) * nctas] : total;
    cnts[d] = d < G ? hi - lo : 0u;
}

// arrival -> sort input: KEY[a] = inbox key, VAL[a] = a (the suffix handle), OFF[a] = inbox offset
__global__ void __launch_bounds__(256)
k_d2_unpack_keys(const u64* __restrict__ in_keys, const u32* __restrict__ in_offs, u32 cnt, u64* __restrict__ key,
                 u32* __restrict__ val, u32* __restrict__ off) {
    const u32 a = blockIdx.x * 256 + threadIdx.x;
    if (a >= cnt) return;
    key[a] = in_keys[a];
    val[a] = a;
    off[a] = in_offs[a];
}

// ---- generic bucketing of items by destination ------------------------------------------------------------------------
// F::item(t, dest, a, b) -> valid.  PASS 0: counts[dest] += 1;  PASS 1: staging[cursor[dest]++] = item (cursor starts at
// the bucket starts; order inside a bucket is irrelevant).  One atomic range reservation per CTA and destination.
template <typename F, int PASS, bool TWO>
__global__ void __launch_bounds__(256)
k_d2_bucket(F f, u32 count, const u32* __restrict__ count_dev, u32* __restrict__ counts_or_cursor, u64* __restrict__ stA,
            u64* __restrict__ stB) {
    __shared__ u32 cnt[MAX_PEERS], basev[MAX_PEERS];
    if (count_dev) count = min(count, *count_dev);
    if (threadIdx.x < MAX_PEERS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 t = blockIdx.x * 256 + threadIdx.x;
    u32 dest = 0xFFu;
    u64 a = 0, b = 0;
    bool valid = t < count;
    if (valid) valid = f.item(t, dest, a, b);
    if (!valid) dest = 0xFFu;
    const u32 same = __match_any_sync(0xffffffffu, dest);
    const u32 leader = __ffs(same) - 1, lane = threadIdx.x & 31;
    u32 off = 0;
    if (valid && lane == leader) off = atomicAdd(&cnt[dest], __popc(same));
    off = __shfl_sync(0xffffffffu, off, leader) + __popc(same & lanemask_lt());
    __syncthreads();
    if (threadIdx.x < MAX_PEERS && cnt[threadIdx.x]) basev[threadIdx.x] = atomicAdd(&counts_or_cursor[threadIdx.x], cnt[threadIdx.x]);
    if (PASS == 0) return;
    __syncthreads();
    if (valid) {
        stA[basev[dest] + off] = a;
        if (TWO) stB[basev[dest] + off] = b;
    }
}
__global__ void k_d2_bucket_starts(const u32* __restrict__ counts, u32* __restrict__ cursor, int G) {
    if (threadIdx.x == 0) { u32 run = 0; for (int g = 0; g < MAX_PEERS; ++g) { cursor[g] = run; run += g < G ? counts[g] : 0u; } }
}

// refined ranks: records (local rank << 32 | handle) of this round -> (position offset at its owner << 34 |
// "LCP is pending" mark << 33 | global rank).  The mark (first regroup only, UPD_NEED_BIT) tells the position owner
// that this suffix's LCP has to be computed from the text (lcp.cuh).
struct UpdItem {
    const u64* upd;
    HandleMap hm;
    u64 rbase;              // first global rank of this GPU
    __device__ __forceinline__ bool item(u32 t, u32& dest, u64& a, u64& b) const {
        const u64 u = upd[t];
        const u64 p = hm.pos((u32)u);
        const u64 d = p / hm.chunk;
        dest = (u32)d;
        a = ((p - d * hm.chunk) << 34) | ((u >> 63) << 33) | (rbase + (u64)((u32)(u >> 32) & 0x7FFFFFFFu));
        return true;
    }
};
__global__ void __launch_bounds__(256)
k_d2_apply_ranks(const u64* __restrict__ inbox, u32 cnt, u64* __restrict__ RANKL, u8* __restrict__ NEED) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u64 u = inbox[e];
    RANKL[u >> 34] = u & D2_MASK33;
    if ((u >> 33) & 1ull) NEED[u >> 34] = 1;
}

// rank requests of a doubling round: active list entry j (S list [0, mS), B list [b0, b0 + mB)) asks the owner of
// position pos(val[j]) + h:  (position offset at the owner << 32 | j)
struct ReqItem {
    const u32* val;
    HandleMap hm;
    u64 h;
    u32 mS, b0;
    __device__ __forceinline__ bool item(u32 t, u32& dest, u64& a, u64& b) const {
        const u32 j = t < mS ? t : b0 + (t - mS);
        const u64 p = hm.pos(val[j]) + h;
        const u64 d = p / hm.chunk;
        dest = (u32)d;
        a = ((p - d * hm.chunk) << 32) | (u64)j;
        return true;
    }
};
// the owner answers the requests of its inbox ...
// ... and pushes every answer straight into the requester's answer region (coalesced peer stores, in the requester's
// staging order): request e of source g's bucket [in_off[g], in_off[g+1]) -> dst[g][e - in_off[g]]
struct ServeGeom { u64* dst[MAX_PEERS]; u32 in_off[MAX_PEERS + 1]; int G; };
__global__ void __launch_bounds__(256)
k_d2_serve_push(const u64* __restrict__ inbox, u32 cnt, const u64* __restrict__ RANKL, u64 nloc, ServeGeom sg) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u64 q = inbox[e] >> 32;
    int g = 0;
#pragma unroll
    for (int s = 1; s < MAX_PEERS; ++s) g += (s < sg.G && e >= sg.in_off[s]) ? 1 : 0;
    sg.dst[g][e - sg.in_off[g]] = q < nloc ? RANKL[q] : 0ull;
}
// the requester completes its keys: key[j] |= response (requests and responses share their staging order)
__global__ void __launch_bounds__(256)
k_d2_apply_resp(const u64* __restrict__ req, const u64* __restrict__ resp, u32 cnt, u64* __restrict__ key) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    key[(u32)req[e]] |= resp[e];
}

// suffix array of the local rank range as S-positions
__global__ void __launch_bounds__(256)
k_d2_sa_positions(const u32* __restrict__ SAh, u32 cnt, HandleMap hm, u64* __restrict__ SA64) {
    const u32 r = blockIdx.x * 256 + threadIdx.x;
    if (r >= cnt) return;
    SA64[r] = hm.pos(SAh[r]);
}

// Phi: rank owner -> position owner.  item = local rank r:  (offset of SA[r] at its owner << 34 | SA[r-1])
struct PhiItem {
    const u64* SA64;
    const u32* LCP;         // local rank order: only the ranks whose LCP is still pending send their Phi
    u64 left_sa;            // SA of the global rank just before this GPU's range (D2_NO_PHI for global rank 0)
    u64 chunk;
    __device__ __forceinline__ bool item(u32 t, u32& dest, u64& a, u64& b) const {
        if (LCP[t] != LCP_PENDING) return false;
        const u64 s = SA64[t];
        const u64 prev = t ? SA64[t - 1] : left_sa;
        const u64 d = s / chunk;
        dest = (u32)d;
        a = ((s - d * chunk) << 34) | prev;
        return true;
    }
};
__global__ void __launch_bounds__(256)
k_d2_apply_phi(const u64* __restrict__ inbox, u32 cnt, u64* __restrict__ PHI) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u64 u = inbox[e];
    PHI[u >> 34] = u & D2_MASK34;
}
// PLCP: position owner -> rank owner.  item = local position t:  (local rank at its owner << 32 | PLCP[t])
struct LcpItem {
    const u64* RANKL;
    const u32* PLCP;
    const u8* NEED;         // only the marked positions were resolved
    RankBases rb;
    __device__ __forceinline__ bool item(u32 t, u32& dest, u64& a, u64& b) const {
        if (!NEED[t]) return false;
        const u64 r = RANKL[t];
        const int g = rb.owner(r);
        dest = (u32)g;
        a = ((r - rb.base[g]) << 32) | (u64)PLCP[t];
        return true;
    }
};
__global__ void __launch_bounds__(256)
k_d2_apply_lcp(const u64* __restrict__ inbox, u32 cnt, u32* __restrict__ LCP) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u64 u = inbox[e];
    LCP[u >> 32] = (u32)u;
}
// per-position results: work item t of stage 3 (rank r = list[t] or real_lo + t, text position i = F0[r]) ->
// owner of T-position i:  A = (offset at the owner << 32 | len),  B = (rc flag << 32 | ref)
struct LrItem {
    const u64* LR;
    const u8* FLAGS;
    const u32* list;        // compacted forward ranks (RC mode) or nullptr
    const u32* F0;
    u32 real_lo, nfac, chunkT;
    __device__ __forceinline__ bool item(u32 t, u32& dest, u64& a, u64& b) const {
        const u32 r = list ? list[t] : real_lo + t;
        const u32 i = F0[r];
        if (i >= nfac) return false;
        const u64 lr = LR[t];
        const u32 d = i / chunkT;
        dest = d;
        a = ((u64)(i - d * chunkT) << 32) | (u64)(u32)lr;
        b = ((u64)((FLAGS[t] & 2) ? 1u : 0u) << 32) | (lr >> 32);
        return true;
    }
};
__global__ void __launch_bounds__(256)
k_d2_apply_lr(const u64* __restrict__ inA, const u64* __restrict__ inB, u32 cnt, u64* __restrict__ LRT, u8* __restrict__ FL) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u64 a = inA[e], b = inB[e];
    const u32 il = (u32)(a >> 32);
    LRT[il] = ((b & 0xFFFFFFFFull) << 32) | (a & 0xFFFFFFFFull);
    FL[il] = (b >> 32) ? 2 : 0;
}

// DNA_RC text of a distributed run: every GPU prepares its slice [lo, hi) of T (staged in its own X) and stores both
// strands into the X of EVERY GPU (coalesced peer stores): S = T s0 rc(T) s1, rc base i at 2n - i.
struct XPeers { u8* p[MAX_PEERS]; int n; };
__global__ void __launch_bounds__(256)
k_d2_prepare_dna_rc_slice(const u8* T, u64 n, u64 lo, u64 hi, XPeers X, bool write_sentinels, unsigned long long* __restrict__ first_bad) {
    for (u64 i = lo + (u64)blockIdx.x * 256 + threadIdx.x; i < hi; i += (u64)gridDim.x * 256) {
        const u8 c = T[i];
        u8 u = c, k = 0;
        switch (c) {
            case 'A': case 'a': u = 'A'; k = 'T'; break;
            case 'C': case 'c': u = 'C'; k = 'G'; break;
            case 'G': case 'g': u = 'G'; k = 'C'; break;
            case 'T': case 't': u = 'T'; k = 'A'; break;
            default: atomicMin(first_bad, (unsigned long long)i); k = c; break;
        }
        for (int g = 0; g < X.n; ++g) { X.p[g][i] = u; X.p[g][2 * n - i] = k; }
    }
    if (write_sentinels && blockIdx.x == 0 && threadIdx.x == 0)
        for (int g = 0; g < X.n; ++g) { X.p[g][n] = 1; X.p[g][2 * n + 1] = 2; }
}

__global__ void k_d2_set_u64(u64* p, u64 v) { *p = v; }

// One launch moves every bucket of an 8-byte-item exchange: element e of the staging list belongs to the bucket g with
// off[g] <= e < off[g+1] and goes to dst[g][e - off[g]] (coalesced peer stores).
constexpr u32 D2_KERNEL_PUSH_MAX = 8u << 20;       // items (64 MB); larger exchanges use the copy engines
struct PushGeom { u64* dst[MAX_PEERS]; u32 off[MAX_PEERS + 1]; int G; };
__device__ __forceinline__ int pg_bucket(const PushGeom& pg, u32 e) {
    int g = 0;
#pragma unroll
    for (int q = 1; q < MAX_PEERS; ++q) g += (q < pg.G && e >= pg.off[q]) ? 1 : 0;
    return g;
}
__global__ void __launch_bounds__(256)
k_d2_push(const u64* __restrict__ staging, u32 total, PushGeom pg) {
    for (u32 e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        const int g = pg_bucket(pg, e);
        pg.dst[g][e - pg.off[g]] = staging[e];
    }
}

}  // namespace nlz
