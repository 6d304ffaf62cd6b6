// Doubling round, fused: gather the second key half RANK[s+h] and sort every tie group by it, in
// shared memory, in ONE pass over the active set (instead of 6 LSD radix passes over HBM).
//
// The active list is a concatenation of tie groups (equal high key half g = group head slot), each
// contiguous.  CTA t owns the groups whose head lies in [t*tile, (t+1)*tile); because no group is
// larger than `maxg` (checked on the host from the previous regroup), its elements fit in
// tile + maxg <= 4096 shared-memory slots.  Rounds whose largest group exceeds that capacity fall
// back to the global radix sort.
//
// Sorting inside a tile exploits what prefix doubling does to repeats: in a round, most members of a
// large group (a tandem array, a repeat family) receive the SAME second key -- they are still tied
// -- and only a few "outliers" separate.  Per group a pivot (majority of three samples) splits the
// members into  [smaller outliers | pivot-equal block | larger outliers]:
//   * pivot-equal members keep their order: position = #smaller + (segmented prefix count);
//   * every other member ranks itself by counting smaller keys in its group (all-pairs);
//   * if the all-pairs work of a tile would be excessive (large groups without a majority), the
//     tile falls back to a bitonic network on the composite key (group, rank).
// Work per member is O(1) for the tied majority instead of O(log^2) compare-exchanges.
//
// The same CTA then regroups its (now sorted) groups -- they never straddle tiles -- so a doubling
// round is two launches: k_gather_rank (reads a consistent RANK snapshot) and this kernel, which also
// detects the new sub-groups, writes the refined ranks (RANK = ISA in the end) and the suffix array
// slots, and appends the members that are still tied to the next round's list (one atomic range
// reservation per tile; the order of groups in that list is irrelevant).
//
// Key layout: key = group << GS | rank.  GS = 32 on one GPU (global slots and ranks below 2^32).  The distributed
// path (dist2.cuh) instantiates GS = 33: groups are named by LOCAL slots (a rank range holds fewer than 2^30
// suffixes) while the second half is a GLOBAL rank of up to 33 bits (6.2 * 10^9 suffixes for a 3.1 Gbp text in RC
// mode); there the pivot and the compacted outliers are kept as tile indices and compared through the full keys.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "sa.cuh"

namespace nlz {

template <int GS> struct KeyHalves {
    using RT = typename std::conditional<GS == 32, u32, u64>::type;     // rank type
    static __device__ __forceinline__ u32 grp(u64 k) { return (u32)(k >> GS); }
    static __device__ __forceinline__ RT rank(u64 k) { return GS == 32 ? (RT)k : (RT)(k & ((1ull << GS) - 1)); }
};

constexpr int TSORT_THREADS = 1024;
constexpr int TSORT_SLOTS = 4096;
constexpr int TSORT_PER_THREAD = TSORT_SLOTS / TSORT_THREADS;   // 4
constexpr u32 TSORT_ALLPAIRS_BUDGET = 320u * 1024u;             // compare steps a tile may spend on counting (beyond: bitonic network)
// shared memory layout (dynamic).  Groups of the active list have >= 2 members, so group start / 2 is
// a unique per-group index: the per-group tables need SLOTS/2 entries.
constexpr size_t TSORT_OFF_KEY = 0;                                          // u64[SLOTS]
constexpr size_t TSORT_OFF_VAL = TSORT_OFF_KEY + (size_t)TSORT_SLOTS * 8;    // u32[SLOTS]
constexpr size_t TSORT_OFF_SLOT = TSORT_OFF_VAL + (size_t)TSORT_SLOTS * 4;   // u32[SLOTS]   suffix-array slots of the owned range
constexpr size_t TSORT_OFF_PIV = TSORT_OFF_SLOT + (size_t)TSORT_SLOTS * 4;   // u32[SLOTS/2] pivot rank of the group
constexpr size_t TSORT_OFF_GLE = TSORT_OFF_PIV + (size_t)TSORT_SLOTS * 2;    // u32[SLOTS/2] group end << 16 | #members < pivot
constexpr size_t TSORT_OFF_GS = TSORT_OFF_GLE + (size_t)TSORT_SLOTS * 2;     // u16[SLOTS]   group start of every member
constexpr size_t TSORT_OFF_EQ = TSORT_OFF_GS + (size_t)TSORT_SLOTS * 2;      // u16[SLOTS]   exclusive prefix counts
constexpr size_t TSORT_OFF_FLAG = TSORT_OFF_EQ + (size_t)TSORT_SLOTS * 2;    // u8 [SLOTS]   head / pivot-equal flags
constexpr size_t TSORT_OFF_ACT = TSORT_OFF_FLAG + (size_t)TSORT_SLOTS;       // u8 [SLOTS]   still-tied flags
constexpr size_t TSORT_OFF_POS = TSORT_OFF_ACT + (size_t)TSORT_SLOTS;        // u16[SLOTS]   sorted position of every member
constexpr size_t TSORT_OFF_MISC = TSORT_OFF_POS + (size_t)TSORT_SLOTS * 2;   // u32[128]
constexpr size_t TSORT_SMEM = TSORT_OFF_MISC + 512;
constexpr u32 TSORT_BIG_OUTLIERS = 64;     // groups with more outliers are ranked by whole warps (one warp per outlier)

// Scans over the SLOTS one-byte flags, 4 consecutive flags per thread.
// MAXPOS: out[o] = index of the last set flag at or before o (flag[0] must be set).
// otherwise: out[o] = number of set flags before o (exclusive prefix count).
template <bool MAXPOS>
__device__ __forceinline__ void tsort_scan_flags(const u8* __restrict__ flag, unsigned short* __restrict__ out,
                                                 u32* __restrict__ wscratch) {
    const u32 t = threadIdx.x, lane = t & 31, w = t >> 5;
    const u32 base = t * TSORT_PER_THREAD;
    static_assert(TSORT_PER_THREAD == 4, "flag scan reads one 4-byte word per thread");
    const u32 words[1] = {*reinterpret_cast<const u32*>(flag + base)};
    u32 loc[TSORT_PER_THREAD];
    u32 run = 0;
#pragma unroll
    for (int j = 0; j < TSORT_PER_THREAD; ++j) {
        const u32 f = (words[j >> 2] >> (8 * (j & 3))) & 0xFFu;
        if (MAXPOS) { if (f) run = base + j + 1; loc[j] = run; }
        else { loc[j] = run; run += f; }
    }
    u32 inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (u32)o) inc = MAXPOS ? max(inc, x) : inc + x;
    }
    if (lane == 31) wscratch[w] = inc;
    __syncthreads();
    u32 carry = 0;
    for (u32 i = 0; i < w; ++i) carry = MAXPOS ? max(carry, wscratch[i]) : carry + wscratch[i];
    u32 prevl = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prevl = 0;
    carry = MAXPOS ? max(carry, prevl) : carry + prevl;
#pragma unroll
    for (int j = 0; j < TSORT_PER_THREAD; ++j)
        out[base + j] = (unsigned short)(MAXPOS ? (max(loc[j], carry) - 1) : (loc[j] + carry));
    __syncthreads();
}

template <int GS>
__global__ void __launch_bounds__(TSORT_THREADS, 2)
k_tile_sort(const u64* __restrict__ key_in, const u32* __restrict__ val_in, const u32* __restrict__ slot_in, u32 m,
            const u32* __restrict__ m_dev /* overrides m when the host runs ahead (pipelined rounds) */,
            u32 tile, u32 maxg, u32* __restrict__ SA, RankDst RANK, u64* __restrict__ key_next,
            u32* __restrict__ val_next, u32* __restrict__ slot_next, u32* __restrict__ out_m /* m' */,
            u32* __restrict__ out_maxg /* largest new group */, int dbg) {
    extern __shared__ __align__(16) unsigned char tsort_smem[];
    u64* skey = reinterpret_cast<u64*>(tsort_smem + TSORT_OFF_KEY);
    u32* sval = reinterpret_cast<u32*>(tsort_smem + TSORT_OFF_VAL);
    u32* gpiv = reinterpret_cast<u32*>(tsort_smem + TSORT_OFF_PIV);
    u32* gle = reinterpret_cast<u32*>(tsort_smem + TSORT_OFF_GLE);
    unsigned short* sgs = reinterpret_cast<unsigned short*>(tsort_smem + TSORT_OFF_GS);
    unsigned short* seq = reinterpret_cast<unsigned short*>(tsort_smem + TSORT_OFF_EQ);
    u8* flag = reinterpret_cast<u8*>(tsort_smem + TSORT_OFF_FLAG);
    u8* act = reinterpret_cast<u8*>(tsort_smem + TSORT_OFF_ACT);
    u32* sslot = reinterpret_cast<u32*>(tsort_smem + TSORT_OFF_SLOT);
    u32* misc = reinterpret_cast<u32*>(tsort_smem + TSORT_OFF_MISC);
    u32& s_first = misc[0];
    u32& s_end = misc[1];
    u32& s_cost = misc[2];
    u32& s_base = misc[3];
    u32& s_nbig = misc[4];
    u32* wscratch = misc + 8;   // one word per warp
    unsigned short* biglist = reinterpret_cast<unsigned short*>(misc + 64);   // starts of the groups with many outliers
    unsigned short* spos = reinterpret_cast<unsigned short*>(tsort_smem + TSORT_OFF_POS);

    using KH = KeyHalves<GS>;
    using RT = typename KH::RT;
    constexpr bool WIDE = GS != 32;
    const u32 tid = threadIdx.x, lane = tid & 31;
    if (m_dev) m = *m_dev;
    const u32 a = blockIdx.x * tile;
    if (a >= m) return;                // grid sized from an upper bound of m
    u32 b = a + tile;
    if (b > m) b = m;
    u32 load_end = b + maxg;          // a group headed before b ends before b + maxg
    if (load_end > m) load_end = m;
    const u32 nload = load_end - a;   // <= TSORT_SLOTS
    if (tid == 0) { s_first = 0xFFFFFFFFu; s_end = load_end; s_cost = 0; s_nbig = 0; }
    __syncthreads();
    // pass 1: find the first owned head (>= a) and the first foreign head (>= b)
    // (one shared-memory atomic per warp: thousands of heads hitting one address would serialise the CTA)
    for (u32 ob = 0; ob < nload; ob += TSORT_THREADS) {
        const u32 o = ob + tid;
        u32 fo = 0xFFFFFFFFu, fe = 0xFFFFFFFFu;
        if (o < nload) {
            const u32 j = a + o;
            const u32 g = KH::grp(key_in[j]);
            const bool head = (j == 0) || (KH::grp(key_in[j - 1]) != g);
            if (head) {
                if (j < b) fo = j;
                else fe = j;
            }
        }
        fo = __reduce_min_sync(0xffffffffu, fo);
        fe = __reduce_min_sync(0xffffffffu, fe);
        if (lane == 0) {
            if (fo != 0xFFFFFFFFu) atomicMin(&s_first, fo);
            if (fe != 0xFFFFFFFFu) atomicMin(&s_end, fe);
        }
    }
    __syncthreads();
    const u32 first = s_first, end = s_end;
    if (first == 0xFFFFFFFFu) return;             // no group starts in this tile
    const u32 cnt = end - first;                  // owned elements [first, end)
    // pass 2: composite keys (group head slot, RANK[s+h]) of the owned elements
    for (u32 o = tid; o < TSORT_SLOTS; o += TSORT_THREADS) {
        u64 k = ~0ull;
        u32 v = 0;
        if (o < cnt) { k = key_in[first + o]; v = val_in[first + o]; }
        skey[o] = k;
        sval[o] = v;
    }
    __syncthreads();
    for (u32 o = tid; o < TSORT_SLOTS; o += TSORT_THREADS)
        flag[o] = (o < cnt && (o == 0 || KH::grp(skey[o]) != KH::grp(skey[o - 1]))) ? 1 : 0;
    __syncthreads();
    tsort_scan_flags<true>(flag, sgs, wscratch);                  // sgs[o] = start of o's group
    // group ends (high half of gle), by group
    for (u32 o = tid; o < cnt; o += TSORT_THREADS) {
        if (o > 0 && flag[o]) gle[sgs[o - 1] >> 1] = o << 16;
        if (o == cnt - 1) gle[sgs[o] >> 1] = cnt << 16;
    }
    __syncthreads();
    // pivot per group: majority of the ranks of its first, middle and last member
    for (u32 o = tid; o < cnt; o += TSORT_THREADS) {
        if (flag[o]) {
            const u32 e = gle[o >> 1] >> 16;
            const RT ka = KH::rank(skey[o]), kb = KH::rank(skey[(o + e) >> 1]), kc = KH::rank(skey[e - 1]);
            // 32-bit ranks: the pivot rank itself; wide ranks: the tile index of a member that carries it
            if (WIDE) gpiv[o >> 1] = (ka == kb || ka == kc) ? o : ((o + e) >> 1);
            else gpiv[o >> 1] = (u32)((ka == kb || ka == kc) ? ka : kb);
        }
    }
    __syncthreads();
    // members below the pivot (warp-aggregated counting) and pivot-equal flags
    for (u32 ob = 0; ob < cnt; ob += TSORT_THREADS) {
        const u32 o = ob + tid;
        const bool valid = o < cnt;
        u32 gs = 0xFFFFFFFFu;
        bool less = false, eq = false;
        if (valid) {
            gs = sgs[o];
            const RT k2 = KH::rank(skey[o]);
            const RT pv = WIDE ? KH::rank(skey[gpiv[gs >> 1]]) : (RT)gpiv[gs >> 1];
            less = k2 < pv;
            eq = k2 == pv;
            if (dbg & 2) eq = false;
            flag[o] = eq ? 1 : 0;
        }
        const u32 same = __match_any_sync(0xffffffffu, gs);
        const u32 lessm = __ballot_sync(0xffffffffu, less);
        const u32 c = __popc(same & lessm);
        if (valid && c && lane == (u32)(__ffs(same) - 1)) atomicAdd(&gle[gs >> 1], c);   // low half: < 2^16
    }
    __syncthreads();
    tsort_scan_flags<false>(flag, seq, wscratch);                 // seq[o] = pivot-equal members before o (tile-wide)
    // Outliers (members that differ from their group's pivot) are compacted, 32-bit ranks only, into `outk` (the
    // slot table's shared memory, which is loaded after the sort): outlier o lands at o - seq[o], so the outliers of a
    // group are contiguous and in member order.  An outlier then ranks itself among its group's OUTLIERS only --
    // the pivot block is accounted for in one step -- so a group costs outliers^2 compares, not outliers x size.
    // Groups with many outliers (the far end of a tandem array late in the doubling) would make a few threads loop
    // long while the CTA waits at the barrier: their outliers are ranked afterwards by whole warps.
    u32* outk = sslot;
    for (u32 ob = 0; ob < cnt; ob += TSORT_THREADS) {
        const u32 o = ob + tid;
        u32 cost = 0;
        if (o < cnt) {
            if (!flag[o]) outk[o - (u32)seq[o]] = WIDE ? o : (u32)skey[o];   // wide ranks: the member's tile index
            if (o == 0 || sgs[o] != sgs[o - 1]) {
                const u32 e = gle[o >> 1] >> 16, size = e - o;
                const u32 eqc = (u32)seq[e - 1] + flag[e - 1] - (u32)seq[o];
                const u32 outl = size - eqc;
                cost = outl * outl;
                if (outl > TSORT_BIG_OUTLIERS) biglist[atomicAdd(&s_nbig, 1u)] = (unsigned short)o;   // at most SLOTS / 65 groups
            }
        }
        cost = __reduce_add_sync(0xffffffffu, cost);
        if (lane == 0 && cost) atomicAdd(&s_cost, cost);
    }
    __syncthreads();
    const bool used_bitonic = !((s_cost <= TSORT_ALLPAIRS_BUDGET || (dbg & 4)) && !(dbg & 1));
    if (!used_bitonic) {
        // sorted positions: pivot-equal members by prefix counts, outliers of small groups by their own thread
#pragma unroll 1
        for (int q = 0; q < TSORT_PER_THREAD; ++q) {
            const u32 o = q * TSORT_THREADS + tid;
            if (o < cnt) {
                const u32 gs = sgs[o];
                const u32 ge = gle[gs >> 1];
                if (flag[o]) {
                    spos[o] = (unsigned short)(gs + (ge & 0xFFFFu) + ((u32)seq[o] - (u32)seq[gs]));
                } else {
                    const u32 e = ge >> 16;
                    const u32 xb = gs - (u32)seq[gs];                               // outliers of the group: outk[xb, xe)
                    const u32 xe = e - ((u32)seq[e - 1] + flag[e - 1]);
                    if (xe - xb <= TSORT_BIG_OUTLIERS) {
                        const u32 me = o - (u32)seq[o];
                        const RT k32 = KH::rank(skey[o]);
                        u32 smaller = 0;
                        for (u32 x = xb; x < me; ++x) smaller += (WIDE ? KH::rank(skey[outk[x]]) : (RT)outk[x]) <= k32 ? 1u : 0u;       // earlier members win ties
                        for (u32 x = me + 1; x < xe; ++x) smaller += (WIDE ? KH::rank(skey[outk[x]]) : (RT)outk[x]) < k32 ? 1u : 0u;
                        const u32 eqc = (e - gs) - (xe - xb);
                        const RT pv = WIDE ? KH::rank(skey[gpiv[gs >> 1]]) : (RT)gpiv[gs >> 1];
                        spos[o] = (unsigned short)(gs + smaller + (k32 > pv ? eqc : 0u));
                    }
                }
            }
        }
        // outliers of the big groups: one warp per outlier, lanes stride over the group's outliers
        const u32 nbig = s_nbig, warp = tid >> 5;
        for (u32 gi = 0; gi < nbig; ++gi) {
            const u32 gs = biglist[gi];
            const u32 ge = gle[gs >> 1], e = ge >> 16;
            const u32 xb = gs - (u32)seq[gs];
            const u32 xe = e - ((u32)seq[e - 1] + flag[e - 1]);
            const u32 eqc = (e - gs) - (xe - xb);
            const RT piv = WIDE ? KH::rank(skey[gpiv[gs >> 1]]) : (RT)gpiv[gs >> 1];
            for (u32 o = gs + warp; o < e; o += TSORT_THREADS / 32) {
                if (flag[o]) continue;                                              // warp-uniform
                const u32 me = o - (u32)seq[o];
                const RT k32 = KH::rank(skey[o]);
                u32 smaller = 0;
                for (u32 x = xb + lane; x < xe; x += 32) {
                    const RT kx = WIDE ? KH::rank(skey[outk[x]]) : (RT)outk[x];
                    smaller += (kx < k32 || (kx == k32 && x < me)) ? 1u : 0u;
                }
                smaller = __reduce_add_sync(0xffffffffu, smaller);
                if (lane == 0) spos[o] = (unsigned short)(gs + smaller + (k32 > piv ? eqc : 0u));
            }
        }
        __syncthreads();
        // permute the tile in place
        u64 rk[TSORT_PER_THREAD];
        u32 rv[TSORT_PER_THREAD];
        unsigned short rp[TSORT_PER_THREAD];
#pragma unroll
        for (int q = 0; q < TSORT_PER_THREAD; ++q) {
            const u32 o = q * TSORT_THREADS + tid;
            rk[q] = 0; rv[q] = 0; rp[q] = 0xFFFF;
            if (o < cnt) { rk[q] = skey[o]; rv[q] = sval[o]; rp[q] = spos[o]; }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < TSORT_PER_THREAD; ++q)
            if (rp[q] != 0xFFFF) { skey[rp[q]] = rk[q]; sval[rp[q]] = rv[q]; }
        __syncthreads();
    } else {
        // fallback: bitonic network over the padded tile.  The groups of the list are in no particular
        // order of their head slots, so the network sorts by (position of the group in the tile, rank):
        // groups stay where they are, members are ordered inside them.
        for (u32 o = tid; o < cnt; o += TSORT_THREADS) skey[o] = ((u64)sgs[o] << GS) | (u64)KH::rank(skey[o]);
        __syncthreads();
        u32 ns = 32;
        while (ns < cnt) ns <<= 1;
        for (u32 k = 2; k <= ns; k <<= 1) {
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                for (u32 t = tid; t < (ns >> 1); t += TSORT_THREADS) {
                    const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const u32 l = i + j;
                    const bool asc = (i & k) == 0;
                    const u64 ki = skey[i], kl = skey[l];
                    if ((ki > kl) == asc) {
                        skey[i] = kl; skey[l] = ki;
                        const u32 vi = sval[i];
                        sval[i] = sval[l]; sval[l] = vi;
                    }
                }
                __syncthreads();
            }
        }
    }
    // ---- regroup: the tile's members are sorted by (group, rank); sub-groups = runs of equal keys
    for (u32 o = tid; o < TSORT_SLOTS; o += TSORT_THREADS) {
        flag[o] = (o < cnt && (o == 0 || skey[o] != skey[o - 1])) ? 1 : 0;
        sslot[o] = o < cnt ? slot_in[first + o] : 0u;      // suffix-array slots of the owned range (outk is dead now)
    }
    __syncthreads();
    tsort_scan_flags<true>(flag, sgs, wscratch);                  // sgs[o] = start of o's new sub-group
    // still-tied members: not (head and next is head); flags reused: bit0 head, bit1 active
    u32 gmax = 0;
    for (u32 o = tid; o < TSORT_SLOTS; o += TSORT_THREADS) {
        u8 f = 0;
        if (o < cnt) {
            const bool hd = flag[o] != 0;
            const bool nh = (o + 1 >= cnt) || flag[o + 1] != 0;
            f = (hd && nh) ? 0 : 1;
            if (nh) gmax = max(gmax, o - (u32)sgs[o] + 1);          // last member of its sub-group
        }
        act[o] = f;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if (lane == 0 && gmax > 1) atomicMax(out_maxg, gmax);
    __syncthreads();
    tsort_scan_flags<false>(act, seq, wscratch);                  // seq[o] = still-tied members before o
    if (tid == 0) {
        const u32 total = (u32)seq[cnt - 1] + act[cnt - 1];
        s_base = total ? atomicAdd(out_m, total) : 0u;
    }
    __syncthreads();
    const u32 base = s_base;
    u64 upd_rec[TSORT_PER_THREAD];
    u32 nupd = 0;
#pragma unroll
    for (int q = 0; q < TSORT_PER_THREAD; ++q) {
        const u32 o = q * TSORT_THREADS + tid;
        if (o < cnt) {
            const u32 s = sval[o];
            const u32 slot = sslot[o];
            const u32 gid0 = KH::grp(skey[o]);
            // New rank of the member's sub-group: any slot inside the sub-group's slot range is a valid name (the
            // ranges of different groups are disjoint; a singleton's range is its slot).  The sub-group that still
            // holds the slot the old group was named by KEEPS that name -- a tandem-array group that sheds a few
            // members per round is then renamed (one scattered store per member, one record per replica on several
            // GPUs) only when its name falls out of its range -- every other sub-group takes its head slot.
            // The old name's tile position follows from the slots being consecutive inside a group.
            u32 newrank = sslot[sgs[o]];
            if (!used_bitonic) {
                const u32 pg = o + (gid0 - slot);                  // tile position of the member that holds slot gid0
                if (pg < cnt && sgs[pg] == sgs[o]) newrank = gid0;
            }
            // old rank = the high key half.  The bitonic path keyed the members by their group's tile position and has
            // lost it; it cannot assume the head slot either (a group that comes from k_group_stream may be named by
            // any slot of its range), so there every rank is stored.
            const u32 gid = KH::grp(skey[o]);
            const u32 oldrank = used_bitonic ? 0xFFFFFFFFu : gid;
            if (newrank != oldrank) {
                if (RANK.rank) RANK.rank[s] = newrank;
                if (RANK.upd) { upd_rec[nupd] = ((u64)newrank << 32) | (u64)s; ++nupd; }
            }
            SA[slot - RANK.base] = s;
            if (act[o]) {
                const u32 pos = base + seq[o];
                key_next[pos] = (u64)newrank << GS;
                val_next[pos] = s;
                slot_next[pos] = slot;
            }
        }
    }
    if (RANK.upd) {                                   // append this tile's records of changed ranks
        u32 total;
        const u32 off = cta_excl_scan(nupd, wscratch, total);
        if (tid == 0) s_base = total ? atomicAdd(RANK.upd_count, total) : 0u;
        __syncthreads();
        for (u32 j = 0; j < nupd; ++j) RANK.upd[s_base + off + j] = upd_rec[j];
    }
}

}  // namespace nlz
