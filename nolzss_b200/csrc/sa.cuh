// Stage 0/1: symbol classes, packed-prefix keys and suffix-array construction by prefix doubling.
//
// Replaces the reference's construct_im(cst, text, 1) (SDSL; call sites
// /root/reference/src/cpp/factorizer.cpp:381 and factorizer_core.hpp:208): the suffix array of
// text·$ and its inverse.  Design (B200-first, not the CSA of the reference):
//   * bytes that occur exactly once in the text (record sentinels, and the virtual terminator at
//     position L) are "sentinel class": unique symbols, ordered by text position, smaller than
//     every repeated byte.  The factorization only depends on the suffix TREE, which is invariant
//     under any order of the alphabet, so this order is free to choose.
//   * repeated bytes get dense b-bit codes (b=2 for DNA); the first W symbols of every suffix are
//     packed MSB-first into one 32- or 64-bit key whose low D bits hold the offset of the first
//     sentinel inside the window (all ones = none).  One stable LSD radix sort of those keys orders
//     all suffixes by their first W symbols; suffixes whose window holds a sentinel are already
//     unique and correctly ordered (stable sort == position order).
//   * the remaining tie groups are refined by prefix doubling (Manber-Myers / Larsson-Sadakane)
//     with discarding: only suffixes still in a group of size >= 2 are re-sorted, by the 64-bit
//     key (group head slot, rank[s+h]).  RANK doubles as the inverse suffix array at the end.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace nlz {

constexpr u32 SENT_CLASS = 0xFFu;

struct ClassTable {
    u8 cls[256];  // dense code, or SENT_CLASS
};

struct KeyLayout {
    int key_bits;   // 32 or 64
    int b;          // bits per symbol
    int W;          // symbols per key
    int D;          // bits of the sentinel-offset field
    int R;          // bits of the record-id prefix (batch mode: suffixes are grouped by record first)
    // Bit layout, MSB first: [R record id][W*b symbols][D sentinel offset][unused zeros].  The offset field sits directly
    // below the symbols, so the initial sort covers ONE bit range [dshift(), key_bits) -- and W is chosen as small as
    // the text allows (layout_symbols in api.cu): a 250 Mbp text sorts 47 bits in 6 passes instead of 64 in 8.
    __host__ __device__ int dshift() const { return key_bits - R - W * b - D; }
};

// Batch mode (many independent records in one text): record id of every text position and the
// geometry needed to stop matches at record boundaries (all sentinels carry the same byte there).
struct BatchView {
    const u32* REC;      // record id per position (forward half, rc half and their sentinels); k for the terminator
    const u32* fstart;   // start of record b in the forward half
    const u32* flen;     // its length
    u32 k;               // number of records
    u32 N;               // RC layout: |S|/2 - 1 (positions > N are the rc half); 0xFFFFFFFF when there is no rc half
};
// symbols of position p up to (excluding) the sentinel that ends its segment
__device__ __forceinline__ u32 batch_cap(const BatchView& bv, u32 p) {
    const u32 b = bv.REC[p];
    if (b >= bv.k) return 0;
    const u32 fs = bv.fstart[b];
    if (p <= bv.N) return fs + bv.flen[b] - p;          // forward half (0 at the sentinel itself)
    return 2 * bv.N - fs + 1 - p;                         // rc half: its segment ends at the mirror of the sentinel before fs
}

// Where refined ranks are written.  One GPU: its RANK array (`rank`).  Distributed (one rank range of the suffix
// array per GPU, dist2.cuh): `rank` is null and every CHANGED rank leaves as a record (local rank << 32 | suffix
// handle) in `upd` -- a contiguous list (appended with one atomic range reservation per CTA, *upd_count entries) that
// is bucketed by position owner, bulk-copied over NVLink and applied to the owners' slices of RANK (scattered peer
// stores collapse beyond ~1 GB of span; bulk copies do not).  A member that stays in the sub-group that keeps its
// group's name keeps its rank: neither a store nor a record.
// `base`: 0 on both paths today (slots and group names are local to the GPU; dist2 adds the GPU's first global rank
// when it turns a record into an exchange item).
struct RankDst {
    u32* rank;
    u64* upd;
    u32* upd_count;
    u32 base;
};
// CTA-wide exclusive prefix of one count per thread (blockDim.x <= 1024); `ws` holds 33 words.
__device__ __forceinline__ u32 cta_excl_scan(u32 v, u32* ws, u32& total) {
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (u32)o) inc += t;
    }
    __syncthreads();
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    u32 carry = 0, tot = 0;
    for (u32 i = 0; i < nw; ++i) { const u32 x = ws[i]; if (i < w) carry += x; tot += x; }
    total = tot;
    return carry + inc - v;
}
constexpr int MAX_PEERS = 8;

// ---------------------------------------------------------------- byte histogram
__global__ void __launch_bounds__(256) k_byte_hist(const u8* __restrict__ x, u64 L, u32* __restrict__ hist) {
    __shared__ u32 h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    const u32 w = threadIdx.x >> 5;
    const u64 nwords = L >> 2;
    const u32* xw = reinterpret_cast<const u32*>(x);
    for (u64 i = (u64)blockIdx.x * 256 + threadIdx.x; i < nwords; i += (u64)gridDim.x * 256) {
        u32 v = xw[i];
        atomicAdd(&h[w][v & 255], 1u);
        atomicAdd(&h[w][(v >> 8) & 255], 1u);
        atomicAdd(&h[w][(v >> 16) & 255], 1u);
        atomicAdd(&h[w][v >> 24], 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x < (L & 3)) atomicAdd(&h[0][x[(nwords << 2) + threadIdx.x]], 1u);
    __syncthreads();
    u32 s = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) s += h[ww][threadIdx.x];
    if (s) atomicAdd(&hist[threadIdx.x], s);
}

// ---------------------------------------------------------------- key construction
// One thread per suffix; the CTA stages its text window in shared memory (coalesced 4-byte loads).
constexpr int KB_TP = 2048;            // suffixes per CTA
constexpr int KB_HALO = 32;            // >= max W

__device__ __forceinline__ void kb_stage(const u8* __restrict__ x, u64 L, const ClassTable& tab, u8* cls, u8* tile) {
    cls[threadIdx.x] = tab.cls[threadIdx.x];
    const u64 base = (u64)blockIdx.x * KB_TP;
    // x is padded with >= 64 readable bytes past L (workspace contract)
    const u32* xw = reinterpret_cast<const u32*>(x + base);
    u32* tw = reinterpret_cast<u32*>(tile);
    for (int i = threadIdx.x; i < (KB_TP + KB_HALO) / 4; i += 256) {
        u64 byte0 = base + (u64)i * 4;
        tw[i] = (byte0 < L + 64) ? xw[i] : 0u;
    }
    __syncthreads();
}

template <typename KeyT>
__device__ __forceinline__ KeyT kb_key(const u8* cls, const u8* tile, int o, u64 p, u64 L, const KeyLayout& lay,
                                       const u32* __restrict__ REC) {
    const int kb = lay.key_bits, b = lay.b, W = lay.W;
    KeyT key = 0;
    KeyT dist = ((KeyT)1 << lay.D) - 1;
    const int dsh = lay.dshift();
    int sh = kb - lay.R - b;
    if (lay.R) key = (KeyT)REC[p] << (kb - lay.R);
    for (int t = 0; t < W; ++t, sh -= b) {
        u32 c = (p + t < L) ? (u32)cls[tile[o + t]] : SENT_CLASS;
        if (c == SENT_CLASS) { dist = (KeyT)t; break; }
        key |= (KeyT)c << sh;
    }
    return key | (dist << dsh);
}

// Every thread builds the keys of KB_Q CONSECUTIVE suffixes: the window of the first one symbol by symbol, every further
// one by shifting a symbol out and one in (one key per ~25 instructions instead of one per 10 W), with the sentinel flags
// of the window rolling along in a bit mask (symbols from the first sentinel on are zero in the key, its offset goes into
// the offset field -- kb_key above is the definition).  The keys are left in shared memory (index o + (o >> 3): one pad
// element per 8, so a thread's run of 8 starts in its own banks) for coalesced stores.  250 Mbp text: 7.9 -> 1.9 ms.
// On entry `tile` holds the CTA's text window (kb_stage), on exit the symbol classes; `base` = text position of tile[0].
constexpr int KB_Q = KB_TP / 256;
constexpr int KB_SKEY = KB_TP + KB_TP / 8;
template <typename KeyT>
__device__ __forceinline__ void kb_tile_keys(const u8* cls, u8* tile, u64 base, u64 L, const KeyLayout& lay, KeyT* skey) {
    // bytes -> classes, in place; positions at or past L count as sentinels
    for (int j = threadIdx.x; j < KB_TP + KB_HALO; j += 256) tile[j] = (base + j < L) ? cls[tile[j]] : (u8)SENT_CLASS;
    __syncthreads();
    const int b = lay.b, W = lay.W;
    const int o0 = threadIdx.x * KB_Q;
    u64 K = 0;                                        // the W symbols of the window, sentinels as 0
    u32 SM = 0;                                       // bit t: symbol t of the window is a sentinel
    for (int t = 0; t < W; ++t) {
        u32 c = tile[o0 + t];
        if (c == SENT_CLASS) { SM |= 1u << t; c = 0; }
        K = (K << b) | c;
    }
    const u64 wmask = (1ull << (W * b)) - 1;          // W * b <= 59
    const int symsh = lay.key_bits - lay.R - W * b, dsh = lay.dshift();
    const KeyT none = ((KeyT)1 << lay.D) - 1;
#pragma unroll 1
    for (int q = 0; q < KB_Q; ++q) {
        u64 ks = K;
        KeyT dist = none;
        if (SM) {
            const int t = __ffs(SM) - 1;
            dist = (KeyT)t;
            ks = K & ~((1ull << ((W - t) * b)) - 1);  // symbols t .. W-1 dropped
        }
        const int o = o0 + q;
        skey[o + (o >> 3)] = ((KeyT)ks << symsh) | (dist << dsh);
        u32 c = tile[o + W];                          // <= KB_TP - 1 + W < KB_TP + KB_HALO
        SM >>= 1;
        if (c == SENT_CLASS) { SM |= 1u << (W - 1); c = 0; }
        K = ((K << b) & wmask) | c;
    }
    __syncthreads();
}

template <typename KeyT>
__global__ void __launch_bounds__(256)
k_build_keys(const u8* __restrict__ x, u64 L, u32 n1, ClassTable tab, KeyLayout lay, const u32* __restrict__ REC,
             KeyT* __restrict__ keys, u32* __restrict__ vals) {
    __shared__ u8 cls[256];
    __shared__ __align__(16) u8 tile[KB_TP + KB_HALO];
    __shared__ KeyT skey[KB_SKEY];
    kb_stage(x, L, tab, cls, tile);
    const u64 base = (u64)blockIdx.x * KB_TP;
    kb_tile_keys<KeyT>(cls, tile, base, L, lay, skey);
#pragma unroll 1
    for (int r = 0; r < KB_Q; ++r) {
        const int o = r * 256 + threadIdx.x;
        const u64 p = base + o;
        if (p >= n1) break;
        KeyT key = skey[o + (o >> 3)];
        if (lay.R) key |= (KeyT)REC[p] << (lay.key_bits - lay.R);
        keys[p] = key;
        vals[p] = (u32)p;
    }
}

// ---------------------------------------------------------------- regroup (head flags, ranks, compaction)
constexpr int RG_THREADS = 256;
constexpr int RG_ITEMS = 8;
constexpr int RG_TILE = RG_THREADS * RG_ITEMS;

template <typename KeyT, bool INITIAL>
__device__ __forceinline__ bool rg_is_head(const KeyT* __restrict__ keys, u32 e, KeyT dist_mask) {
    if (e == 0) return true;
    KeyT k = keys[e];
    if (INITIAL && (k & dist_mask) != dist_mask) return true;  // window holds a sentinel: unique
    return k != keys[e - 1];
}

// phase A: per-tile (max head index + 1, number of still-active elements)
template <typename KeyT, bool INITIAL>
__global__ void __launch_bounds__(RG_THREADS)
k_regroup_reduce(const KeyT* __restrict__ keys, u32 m, KeyT dist_mask, u32* __restrict__ pmax,
                 u32* __restrict__ psum) {
    __shared__ u32 smax[RG_THREADS / 32], ssum[RG_THREADS / 32];
    const u64 tile_start = (u64)blockIdx.x * RG_TILE;
    u32 lmax = 0, lsum = 0;
#pragma unroll
    for (int t = 0; t < RG_ITEMS; ++t) {
        u64 e = tile_start + (u64)t * RG_THREADS + threadIdx.x;
        if (e < m) {
            bool h = rg_is_head<KeyT, INITIAL>(keys, (u32)e, dist_mask);
            bool hn = (e + 1 < m) ? rg_is_head<KeyT, INITIAL>(keys, (u32)e + 1, dist_mask) : true;
            if (h) lmax = max(lmax, (u32)e + 1);
            lsum += (h && hn) ? 0u : 1u;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        lmax = max(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    }
    if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = lmax; ssum[threadIdx.x >> 5] = lsum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 a = 0, b = 0;
        for (int i = 0; i < RG_THREADS / 32; ++i) { a = max(a, smax[i]); b += ssum[i]; }
        pmax[blockIdx.x] = a;
        psum[blockIdx.x] = b;
    }
}

// phase B: exclusive (max, sum) scan of the tile partials by one CTA; writes the new active count.
__global__ void __launch_bounds__(1024)
k_regroup_scan_partials(u32* __restrict__ pmax, u32* __restrict__ psum, u32 ntiles, u32* __restrict__ m_out) {
    // m_out[0] = new active count; m_out[3] = largest new group (atomicMax'ed by k_regroup_apply)
    if (threadIdx.x == 0) m_out[3] = 0;
    __shared__ u32 wmax[32], wsum[32];
    const u32 per = (ntiles + 1023) / 1024;
    u32 b = threadIdx.x * per, e = b + per;
    if (b > ntiles) b = ntiles;
    if (e > ntiles) e = ntiles;
    u32 lm = 0, ls = 0;
    for (u32 i = b; i < e; ++i) { lm = max(lm, pmax[i]); ls += psum[i]; }
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 im = lm, is = ls;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 tm = __shfl_up_sync(0xffffffffu, im, o), ts = __shfl_up_sync(0xffffffffu, is, o);
        if (lane >= o) { im = max(im, tm); is += ts; }
    }
    if (lane == 31) { wmax[w] = im; wsum[w] = is; }
    __syncthreads();
    u32 cm = 0, cs = 0, tot = 0;
    for (u32 i = 0; i < 32; ++i) {
        if (i < w) { cm = max(cm, wmax[i]); cs += wsum[i]; }
        tot += wsum[i];
    }
    // exclusive prefix for this thread = carries of previous warps + previous lanes
    u32 pm = __shfl_up_sync(0xffffffffu, im, 1), ps = __shfl_up_sync(0xffffffffu, is, 1);
    if (lane == 0) { pm = 0; ps = 0; }
    u32 rm = max(cm, pm), rs = cs + ps;
    for (u32 i = b; i < e; ++i) {
        u32 tm = pmax[i], ts = psum[i];
        pmax[i] = rm;
        psum[i] = rs;
        rm = max(rm, tm);
        rs += ts;
    }
    if (threadIdx.x == 0) *m_out = tot;
}

// LCP of two adjacent suffixes of the initially sorted list, from their packed-prefix keys alone: the number of
// leading symbols the keys share, cut at the first sentinel of either window (sentinels are unique symbols; batch
// sentinels end their record).  Exact whenever the keys differ or a window holds a sentinel; W (the window) otherwise --
// those pairs (members of a tie group behind its head) are the only ones the Kasai kernel has to look at.
template <typename KeyT>
__device__ __forceinline__ u32 key_pair_lcp(KeyT a, KeyT b, const KeyLayout& lay) {
    const KeyT dmask = ((KeyT)1 << lay.D) - 1;
    const int dsh = lay.dshift();
    const KeyT x = a ^ b;
    u32 common = (u32)lay.W;
    if (x) {
        const int p = sizeof(KeyT) == 8 ? __clzll((long long)x) : __clz((int)x);     // identical leading bits
        if (p < lay.R) return 0;                                                      // different records (batch mode)
        const u32 cs = (u32)(p - lay.R) / (u32)lay.b;
        if (cs < common) common = cs;
    }
    const KeyT da = (a >> dsh) & dmask, db = (b >> dsh) & dmask;
    if (da != dmask && (u32)da < common) common = (u32)da;
    if (db != dmask && (u32)db < common) common = (u32)db;
    return common;
}
constexpr u32 LCP_PENDING = 0xFFFFFFFFu;      // LCP entry left to the Kasai kernel
// INITIAL regroup only: where the key-derived LCP values and the "Kasai has to compute this one" marks go
struct LcpSeed {
    u32* LCP;            // rank order (local slots); nullptr: no seeding
    u8* NEED;            // by text position (one GPU); nullptr on the distributed path (the mark travels in the record)
    KeyLayout lay;
    bool first_pending;  // element 0 follows a suffix of ANOTHER GPU: its LCP cannot come from a key pair
    // Large texts, one GPU: the inverse suffix array RANK[s] = rank is a scatter of n' 4-byte stores over gigabytes --
    // bound by the GPU's random-access rate.  Instead the new ranks are written in rank order here (coalesced), the
    // (suffix, rank) pairs are partitioned by the top 8 bits of the suffix position (one radix pass), and a last
    // kernel scatters them window by window (8 MB of RANK at a time: L2-resident).  nullptr: direct scatter.
    u32* RANKOUT;
    // With RANKOUT and n' < 2^31 the mark travels in bit 31 of the rank and k_scatter_pairs sets NEED inside the same
    // L2-resident window (110 M scattered byte stores of the 250 Mbp text otherwise: ~64 B of DRAM traffic each).
    bool need_in_rankout;
};
constexpr u32 RANKOUT_NEED_BIT = 0x80000000u;
template <bool NEEDBIT>
__global__ void __launch_bounds__(256)
k_scatter_pairs(const u32* __restrict__ pos, const u32* __restrict__ val, u32 m, u32* __restrict__ dst, u8* __restrict__ need) {
    for (u32 e = blockIdx.x * 256 + threadIdx.x; e < m; e += gridDim.x * 256) {
        const u32 s = pos[e], v = val[e];
        if (NEEDBIT) {
            dst[s] = v & ~RANKOUT_NEED_BIT;
            if (v & RANKOUT_NEED_BIT) need[s] = 1;
        } else {
            dst[s] = v;
        }
    }
}
constexpr u64 UPD_NEED_BIT = 1ull << 63;      // INITIAL records: the suffix is a member of a tie group (its LCP is pending)

// phase C: ranks, SA write-back and compaction of the still-active elements.  GS: the doubling keys are
// group << GS | rank (tile_sort.cuh); 33 on the distributed path, where RANK.rank is null (records only).
template <typename KeyT, bool INITIAL, int GS = 32>
__global__ void __launch_bounds__(RG_THREADS)
k_regroup_apply(const KeyT* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ slots,
                u32 m, KeyT dist_mask, const u32* __restrict__ pmax, const u32* __restrict__ psum,
                u32* __restrict__ SA, RankDst RANK, u64* __restrict__ key_next,
                u32* __restrict__ val_next, u32* __restrict__ slot_next, u32* __restrict__ maxg_out,
                LcpSeed seed = LcpSeed{nullptr, nullptr, KeyLayout{64, 1, 1, 1, 0}, false, nullptr, false}) {
    __shared__ u8 sh_head[RG_TILE + 8];
    __shared__ u32 wmax[RG_THREADS / 32], wsum[RG_THREADS / 32];
    const u64 tile_start = (u64)blockIdx.x * RG_TILE;
    const u32 valid = (u32)((m - tile_start) < (u64)RG_TILE ? (m - tile_start) : (u64)RG_TILE);
#pragma unroll
    for (int t = 0; t < RG_ITEMS; ++t) {
        u32 idx = t * RG_THREADS + threadIdx.x;
        if (idx < valid) sh_head[idx] = rg_is_head<KeyT, INITIAL>(keys, (u32)(tile_start + idx), dist_mask) ? 1 : 0;
    }
    if (threadIdx.x == 0) {
        u64 e = tile_start + valid;
        sh_head[valid] = (e < m) ? (rg_is_head<KeyT, INITIAL>(keys, (u32)e, dist_mask) ? 1 : 0) : 1;
    }
    __syncthreads();
    const u32 i0 = threadIdx.x * RG_ITEMS;
    u32 hm[RG_ITEMS], ps[RG_ITEMS];
    bool act[RG_ITEMS];
    u32 gmax = 0;
    u32 run_max = 0, run_sum = 0;
    u64 upd_rec[RG_ITEMS];
    u32 nupd = 0;
#pragma unroll
    for (int q = 0; q < RG_ITEMS; ++q) {
        u32 idx = i0 + q;
        hm[q] = 0; ps[q] = 0; act[q] = false;
        if (idx < valid) {
            bool h = sh_head[idx] != 0;
            bool a = !(h && sh_head[idx + 1] != 0);
            if (h) run_max = (u32)(tile_start + idx) + 1;
            hm[q] = run_max;
            ps[q] = run_sum;
            run_sum += a ? 1u : 0u;
            act[q] = a;
        }
    }
    // CTA-wide exclusive (max, sum) scan of the per-thread aggregates
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 im = run_max, is = run_sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 tm = __shfl_up_sync(0xffffffffu, im, o), ts = __shfl_up_sync(0xffffffffu, is, o);
        if (lane >= o) { im = max(im, tm); is += ts; }
    }
    if (lane == 31) { wmax[w] = im; wsum[w] = is; }
    __syncthreads();
    u32 cm = pmax[blockIdx.x], cs = psum[blockIdx.x];
    for (u32 i = 0; i < w; ++i) { cm = max(cm, wmax[i]); cs += wsum[i]; }
    u32 pm = __shfl_up_sync(0xffffffffu, im, 1), pps = __shfl_up_sync(0xffffffffu, is, 1);
    if (lane == 0) { pm = 0; pps = 0; }
    cm = max(cm, pm);
    cs += pps;
#pragma unroll
    for (int q = 0; q < RG_ITEMS; ++q) {
        u32 idx = i0 + q;
        if (idx < valid) {
            u32 e = (u32)(tile_start + idx);
            u32 hj = max(hm[q], cm) - 1;          // index (in this sorted list) of the group head
            u32 newrank = INITIAL ? hj + RANK.base : slots[hj];
            u32 s = vals ? vals[e] : e;            // (distributed path, first regroup: the sorted index IS the suffix handle)
            u32 slot = INITIAL ? e + RANK.base : slots[e];
            // the old rank of a member is its group's head slot = the high half of its doubling key
            const bool changed = INITIAL || newrank != (u32)((u64)keys[e] >> GS);
            bool lcp_pending = false;
            if (INITIAL && seed.LCP) {
                // A singleton's LCP with its predecessor follows from the two keys.  The members of a tie group share their
                // whole window (LCP >= W) and the doubling rounds still permute them inside the group's slot range -- which
                // suffix ends up in which slot is not known yet -- so EVERY member is left to the Kasai kernel (for the one
                // that lands in the head slot it recomputes what the key pair says).
                lcp_pending = act[q] || (e == 0 && seed.first_pending);
                seed.LCP[slot - RANK.base] = lcp_pending ? LCP_PENDING : (e == 0 ? 0u : key_pair_lcp<KeyT>(keys[e - 1], keys[e], seed.lay));
                if (lcp_pending && seed.NEED && !seed.need_in_rankout) seed.NEED[s] = 1;
            }
            if (INITIAL && seed.RANKOUT) seed.RANKOUT[e] = newrank | ((seed.need_in_rankout && lcp_pending) ? RANKOUT_NEED_BIT : 0u);
            if (changed) {
                if (RANK.rank) RANK.rank[s] = newrank;
                if (RANK.upd) {
                    if (INITIAL) RANK.upd[e] = ((u64)newrank << 32) | (u64)s | (lcp_pending ? UPD_NEED_BIT : 0ull);      // every rank is new: dense list
                    else { upd_rec[nupd] = ((u64)newrank << 32) | (u64)s; ++nupd; }
                }
            }
            SA[slot - RANK.base] = s;
            if (act[q]) {
                u32 pos = cs + ps[q];
                key_next[pos] = (u64)newrank << GS;
                val_next[pos] = s;
                slot_next[pos] = slot;
                if (sh_head[idx + 1] != 0) gmax = max(gmax, e - hj + 1);   // last element of its group
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && gmax > 1) atomicMax(maxg_out, gmax);
    if (!INITIAL && RANK.upd) {                      // append this CTA's records of changed ranks
        __shared__ u32 ws[33];
        __shared__ u32 s_ubase;
        u32 total;
        const u32 off = cta_excl_scan(nupd, ws, total);
        if (threadIdx.x == 0) s_ubase = total ? atomicAdd(RANK.upd_count, total) : 0u;
        __syncthreads();
        for (u32 j = 0; j < nupd; ++j) RANK.upd[s_ubase + off + j] = upd_rec[j];
    }
}

// KEY[j] |= RANK[VAL[j] + h]   (second half of the doubling key).  The list length is `m`, or *m_dev when the
// host runs ahead of the device (pipelined rounds: the grid is sized from an upper bound).
__global__ void __launch_bounds__(256)
k_gather_rank(u64* __restrict__ key, const u32* __restrict__ val, u32 m, const u32* __restrict__ m_dev,
              const u32* __restrict__ RANK, u64 h, u32 n1, u32* __restrict__ ctr) {
    if (m_dev) m = *m_dev;
    u32 j = blockIdx.x * 256 + threadIdx.x;
    if (j == 0 && ctr) { ctr[0] = 0; ctr[3] = 0; }   // next round's active count / largest group (fused regroup)
    if (j >= m) return;
    u64 p = (u64)val[j] + h;
    u32 r = p < n1 ? RANK[p] : 0u;   // p < n1 always holds for active suffixes; guard keeps reads in range
    key[j] |= (u64)r;
}

}  // namespace nlz
