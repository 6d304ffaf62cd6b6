// Doubling rounds when some tie groups exceed the shared-memory tile of tile_sort.cuh ("hybrid rounds").
//
// A text with long tandem arrays keeps a few thousand tie groups of 10^3..10^5 members alive for
// log2(longest repeat) rounds (a 250 Mbp text with Mbp-long arrays: 10^8 active suffixes for 13 rounds).
// Re-sorting the whole active list with the global radix sort costs 8 passes + a regroup per round although
// nearly every member of such a group receives the SAME second key (it is still tied) and only the members near
// the end of the array separate.  Hybrid rounds keep two lists in the same buffers:
//   S  (left-aligned,  [0, mS))        groups of <= gcap members  -> k_tile_sort, unchanged;
//   B  (right-aligned, [end - mB, end)) groups of  > gcap members  -> k_group_stream, one CTA per group.
// k_group_stream streams its group twice: pass A classifies every member against a pivot (the most frequent of 32
// samples), counts the members below it and collects the OUTLIERS (members that differ from the pivot) in shared
// memory; the outliers are sorted there (bitonic network); pass B places the pivot-equal block behind the smaller
// outliers.  Sub-groups (runs of equal keys) get their refined ranks, the suffix array slots are rewritten, and
// the members that are still tied go to the next round's lists: sub-groups of <= gcap members are appended to S
// (one atomic range reservation), larger ones to B (reserved downwards from `end`, so both lists stay compact
// without knowing each other's size).  A group with more outliers than fit in shared memory raises a flag and
// leaves everything untouched: the host redoes the round with the radix path (api.cu).
//
// Invariants used: a group's members are contiguous in the list, in suffix-array slot order; the group's id (high
// key half = its current rank) is a slot inside the group's slot range -- the head slot after a regroup or a tile
// sort, possibly a middle slot after k_group_stream (see "Ranks" in the kernel); every group of B has more than
// gcap members, so a chunk of gcap list positions holds at most one group head and CTA c simply looks for a head
// in chunk c.
#pragma once
#include "common.cuh"
#include "sa.cuh"
#include "tile_sort.cuh"

namespace nlz {

constexpr int GS_THREADS = 1024;
// Two sizes of the outlier buffer.  A tandem-array group sheds only a few members per round, so nearly every group fits
// the SMALL buffer (48 KB: two 1024-thread CTAs per SM -- with the one 193 KB buffer of round 1 the kernel ran one CTA
// per SM at 48 % active warps); a group whose outliers do not fit is appended to an overflow list and redone by a
// second, persistent launch with the BIG buffer.
constexpr u32 GS_CAP_SMALL = 4096;
constexpr u32 GS_CAP_BIG = 16384;                                 // outliers a CTA can sort in shared memory
// shared memory layout for a buffer of CAP outliers: u64[CAP] (key2 << shift | suffix), sorted | u16[CAP] head index of
// the outlier's sub-group | u16[CAP] sub-group size, stored at its head (CAP < 65536) | u32[128] misc | big sub-group
// list: u32 head[64], u32 base[64]
__host__ __device__ constexpr size_t gs_off_head(u32 cap) { return (size_t)cap * 8; }
__host__ __device__ constexpr size_t gs_off_size(u32 cap) { return gs_off_head(cap) + (size_t)cap * 2; }
__host__ __device__ constexpr size_t gs_off_misc(u32 cap) { return gs_off_size(cap) + (size_t)cap * 2; }
__host__ __device__ constexpr size_t gs_smem(u32 cap) { return gs_off_misc(cap) + 512 + 2 * 64 * 4; }
constexpr u32 GS_MAX_BIGSUB = 64;                                 // outlier sub-groups larger than gcap per group (more: fallback)

template <int GS> __device__ __forceinline__ u32 gs_grp(u64 k) { return KeyHalves<GS>::grp(k); }
// sorted outliers in shared memory: rank << OKV_SHIFT | suffix.  GS = 33 (distributed path): ranks of up to 34 bits,
// suffixes are local handles below 2^30.
template <int GS> struct OkvLayout { static constexpr int SHIFT = GS == 32 ? 32 : 30; };

// ---- split of a unified list into S and B (once, when the hybrid rounds start) -------------------------------
// member j belongs to a group of more than gcap members iff the list position gcap behind its group's head still
// holds the same group.
template <int GS>
__device__ __forceinline__ bool bg_is_big(const u64* __restrict__ key, const u32* __restrict__ slot, u32 m, u32 j, u32 gcap) {
    const u32 g = gs_grp<GS>(key[j]);
    const u32 hj = j - (slot[j] - g);
    const u32 p = hj + gcap;
    return p < m && gs_grp<GS>(key[p]) == g;
}

template <int GS>
__global__ void __launch_bounds__(RG_THREADS)
k_split_count(const u64* __restrict__ key, const u32* __restrict__ slot, u32 m, u32 gcap, u32* __restrict__ psum) {
    __shared__ u32 ws[33];
    const u64 tile_start = (u64)blockIdx.x * RG_TILE;
    u32 c = 0;
#pragma unroll
    for (int t = 0; t < RG_ITEMS; ++t) {
        const u64 e = tile_start + (u64)t * RG_THREADS + threadIdx.x;
        if (e < m && bg_is_big<GS>(key, slot, m, (u32)e, gcap)) ++c;
    }
    u32 total;
    cta_excl_scan(c, ws, total);
    if (threadIdx.x == 0 && psum) psum[blockIdx.x] = total;
}

// stable partition: small-group members to [0, mS), big-group members to [end - mB, end); mB = *mB_dev
template <int GS>
__global__ void __launch_bounds__(RG_THREADS)
k_split_apply(const u64* __restrict__ key, const u32* __restrict__ val, const u32* __restrict__ slot, u32 m, u32 gcap,
              const u32* __restrict__ psum, const u32* __restrict__ mB_dev, u32 end,
              u64* __restrict__ key_out, u32* __restrict__ val_out, u32* __restrict__ slot_out) {
    __shared__ u32 ws[33];
    if (m == 0) return;                                          // warm-up launch (api.cu: nlz_ctx_create)
    const u32 tile_start = blockIdx.x * RG_TILE;
    const u32 i0 = tile_start + threadIdx.x * RG_ITEMS;          // RG_ITEMS consecutive members per thread
    bool big[RG_ITEMS];
    u32 c = 0;
#pragma unroll
    for (int q = 0; q < RG_ITEMS; ++q) {
        const u32 e = i0 + q;
        big[q] = e < m && bg_is_big<GS>(key, slot, m, e, gcap);
        c += big[q] ? 1u : 0u;
    }
    u32 total;
    u32 before = psum[blockIdx.x] + cta_excl_scan(c, ws, total);   // big members before i0
    const u32 b0 = end - *mB_dev;
#pragma unroll
    for (int q = 0; q < RG_ITEMS; ++q) {
        const u32 e = i0 + q;
        if (e < m) {
            const u32 dst = big[q] ? b0 + before : e - before;
            key_out[dst] = key[e];
            val_out[dst] = val[e];
            slot_out[dst] = slot[e];
            before += big[q] ? 1u : 0u;
        }
    }
}

// ---- one CTA per big group -----------------------------------------------------------------------------------
struct StreamOut {
    u64* key_next; u32* val_next; u32* slot_next;   // next round's lists (whole buffers)
    u32 end;                                        // capacity of the lists: B grows downwards from here
    u32* mS;                                        // next S length (shared with k_tile_sort's reservations)
    u32* maxgS;                                     // largest group appended to S
    u32* mB;                                        // next B length
    u32* fallback;                                  // set when a group cannot be handled here
    // two-size launch: the SMALL pass appends the chunks whose group overflowed its buffer (ovf_out / ovf_cnt_out);
    // the BIG pass (persistent CTAs) works through that list (ovf_in / ovf_cnt_in).  All null: one pass over all chunks.
    u32* ovf_out; u32* ovf_cnt_out;
    const u32* ovf_in; const u32* ovf_cnt_in;
};

template <int GS, u32 CAP>
__device__ __forceinline__ void gs_process_chunk(const u64* __restrict__ key_in, const u32* __restrict__ val_in, const u32* __restrict__ slot_in,
                                                 u32 b0, u32 mB, u32 gcap, u32* __restrict__ SA, const RankDst& RANK, const StreamOut& out, int dbg,
                                                 u32 chunk_idx, unsigned char* gs_smem_base) {
    constexpr u32 GS_OUT_CAP = CAP;
    u64* okv = reinterpret_cast<u64*>(gs_smem_base);
    unsigned short* ohead = reinterpret_cast<unsigned short*>(gs_smem_base + gs_off_head(CAP));
    unsigned short* osize = reinterpret_cast<unsigned short*>(gs_smem_base + gs_off_size(CAP));
    u32* misc = reinterpret_cast<u32*>(gs_smem_base + gs_off_misc(CAP));
    using KH = KeyHalves<GS>;
    using RT = typename KH::RT;
    constexpr int OS = OkvLayout<GS>::SHIFT;
    constexpr u64 OMASK = (1ull << OS) - 1;
    __shared__ unsigned long long s_piv;
    u32& s_head = misc[0];
    u32& s_probe = misc[1];
    u32& s_endpos = misc[2];
    u32& s_nout = misc[3];
    u32& s_nlt = misc[4];
    u32& s_eq = misc[5];       // pass B: pivot-equal members placed so far
    u32& s_baseS = misc[6];
    u32& s_baseEq = misc[7];
    u32& s_baseU = misc[8];
    u32& s_nbig = misc[9];
    u32& s_actS = misc[10];
    u32* ws = misc + 32;       // 33 words of scan scratch
    u32* big_head = misc + 128;
    u32* big_base = misc + 128 + GS_MAX_BIGSUB;

    const u32 tid = threadIdx.x, lane = tid & 31;
    const u32 bend = b0 + mB;
    const u32 c0 = b0 + chunk_idx * gcap;
    if (c0 >= bend) return;
    u32 c1 = c0 + gcap;
    if (c1 > bend) c1 = bend;
    if (tid == 0) { s_head = 0xFFFFFFFFu; s_probe = 0xFFFFFFFFu; s_endpos = 0xFFFFFFFFu; s_nout = 0; s_nlt = 0; s_eq = 0; s_nbig = 0; }
    __syncthreads();
    for (u32 j = c0 + tid; j < c1; j += GS_THREADS)
        if (j == b0 || gs_grp<GS>(key_in[j - 1]) != gs_grp<GS>(key_in[j])) s_head = j;    // at most one head per chunk
    __syncthreads();
    const u32 head = s_head;
    if (head == 0xFFFFFFFFu) return;
    const u32 g = gs_grp<GS>(key_in[head]);        // the group's rank: a slot inside the group's slot range (see below)
    const u32 gbase = slot_in[head];            // first slot of the group
    // group end: first probe position (head + t * gcap) outside the group, then the exact end inside that stride
    for (u32 tb = 0;; tb += GS_THREADS) {
        const u64 p = (u64)head + (u64)(tb + tid + 1) * gcap;
        const bool outside = p >= bend || gs_grp<GS>(key_in[p]) != g;
        if (outside) atomicMin(&s_probe, tb + tid + 1);
        __syncthreads();
        if (s_probe != 0xFFFFFFFFu) break;
        __syncthreads();
    }
    {
        const u32 t = s_probe;                                     // end in (head + (t-1) * gcap, head + t * gcap]
        const u64 lo = (u64)head + (u64)(t - 1) * gcap + 1;
        for (u32 o = tid; o < gcap; o += GS_THREADS) {
            const u64 p = lo + o;
            if (p >= bend || gs_grp<GS>(key_in[p]) != g) atomicMin(&s_endpos, (u32)(p < bend ? p : bend));
        }
        __syncthreads();
    }
    const u32 gend = s_endpos;
    const u32 M = gend - head;                                     // > gcap
    // pivot: the most frequent of 32 evenly spaced samples.  (Three samples are not enough here: the members of a
    // group are in no particular order after a streamed round, and a pivot that is itself an outlier turns the whole
    // group into outliers -- more than fit in shared memory.)
    if (tid < 32) {
        const RT k2 = KH::rank(key_in[head + (u32)(((u64)M * lane) >> 5)]);
        const u32 votes = __popc(__match_any_sync(0xffffffffu, k2));
        const u32 best = __reduce_max_sync(0xffffffffu, votes);
        const u32 who = __ffs(__ballot_sync(0xffffffffu, votes == best)) - 1;
        const RT pv = __shfl_sync(0xffffffffu, k2, who);
        if (lane == 0) s_piv = (unsigned long long)pv;
    }
    __syncthreads();
    const RT piv = (RT)s_piv;
    // ---- pass A: outliers to shared memory, members below the pivot counted
    for (u32 jb = head; jb < gend; jb += GS_THREADS) {
        const u32 j = jb + tid;
        bool outl = false, less = false;
        RT k2 = 0;
        if (j < gend) {
            k2 = KH::rank(key_in[j]);
            outl = k2 != piv;
            less = k2 < piv;
        }
        const u32 om = __ballot_sync(0xffffffffu, outl);
        const u32 lm = __ballot_sync(0xffffffffu, less);
        if (om) {
            u32 base = 0;
            if (lane == 0) { base = atomicAdd(&s_nout, __popc(om)); if (lm) atomicAdd(&s_nlt, __popc(lm)); }
            base = __shfl_sync(0xffffffffu, base, 0);
            if (outl) {
                const u32 q = base + __popc(om & lanemask_lt());
                if (q < GS_OUT_CAP) okv[q] = ((u64)k2 << OS) | (u64)val_in[j];
            }
        }
    }
    __syncthreads();
    const u32 nout = s_nout, nlt = s_nlt, neq = M - nout;
    if (nout > GS_OUT_CAP || (dbg & 8)) {
        // does not fit: nothing has been written yet.  SMALL pass: leave the group to the BIG pass; otherwise: fallback
        if (tid == 0) {
            if (out.ovf_out && !(dbg & 8) && nout <= GS_CAP_BIG) out.ovf_out[atomicAdd(out.ovf_cnt_out, 1u)] = chunk_idx;
            else atomicExch(out.fallback, 1u);
        }
        return;
    }
    // ---- sort the outliers by (key2, suffix)
    if (nout > 1) {
        u32 ns = 32;
        while (ns < nout) ns <<= 1;
        for (u32 o = nout + tid; o < ns; o += GS_THREADS) okv[o] = ~0ull;
        __syncthreads();
        for (u32 k = 2; k <= ns; k <<= 1) {
            for (u32 j = k >> 1; j > 0; j >>= 1) {
                for (u32 t = tid; t < (ns >> 1); t += GS_THREADS) {
                    const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const u32 l = i + j;
                    const bool asc = (i & k) == 0;
                    const u64 ki = okv[i], kl = okv[l];
                    if ((ki > kl) == asc) { okv[i] = kl; okv[l] = ki; }
                }
                __syncthreads();
            }
        }
    }
    // ---- sub-groups of the outliers: heads (max-scan), sizes, still-tied members bound for S
    constexpr u32 PER = GS_OUT_CAP / GS_THREADS;                   // 16 consecutive outliers per thread
    u32 actS_before = 0;                                           // still-tied outliers bound for S before this thread's range
    if (nout) {
        const u32 q0 = tid * PER;
        u32 run = 0;                                               // head index + 1 of the running sub-group
        u32 loc[PER];
#pragma unroll
        for (u32 i = 0; i < PER; ++i) {
            const u32 q = q0 + i;
            if (q < nout) {
                const bool hd = q == 0 || q == nlt || (okv[q] >> OS) != (okv[q - 1] >> OS);
                if (hd) run = q + 1;
            }
            loc[i] = run;
        }
        // CTA-wide exclusive max-scan of `run`
        u32 inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 x = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (u32)o) inc = max(inc, x);
        }
        if (lane == 31) ws[tid >> 5] = inc;
        __syncthreads();
        u32 carry = 0;
        for (u32 i = 0; i < (tid >> 5); ++i) carry = max(carry, ws[i]);
        u32 prevl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) prevl = 0;
        carry = max(carry, prevl);
#pragma unroll
        for (u32 i = 0; i < PER; ++i) {
            const u32 q = q0 + i;
            if (q < nout) ohead[q] = (unsigned short)(max(loc[i], carry) - 1);
        }
        __syncthreads();
        // sizes at the heads (written by the last member of every sub-group)
#pragma unroll
        for (u32 i = 0; i < PER; ++i) {
            const u32 q = q0 + i;
            if (q < nout) {
                const bool last = q + 1 == nout || q + 1 == nlt || (okv[q + 1] >> OS) != (okv[q] >> OS);
                if (last) {
                    const u32 hq = ohead[q], sz = q - hq + 1;
                    osize[hq] = (unsigned short)sz;
                    if (sz > gcap) {
                        const u32 k = atomicAdd(&s_nbig, 1u);
                        if (k < GS_MAX_BIGSUB) big_head[k] = hq;
                    }
                }
            }
        }
        __syncthreads();
        if (s_nbig > GS_MAX_BIGSUB) {
            if (tid == 0) atomicExch(out.fallback, 1u);
            return;
        }
        // exclusive count of the still-tied outliers that go to S (sub-group size 2..gcap)
        u32 mine = 0, gmax = 0;
#pragma unroll
        for (u32 i = 0; i < PER; ++i) {
            const u32 q = q0 + i;
            if (q < nout) {
                const u32 sz = osize[ohead[q]];
                if (sz >= 2 && sz <= gcap) { ++mine; gmax = max(gmax, sz); }
            }
        }
        u32 total;
        actS_before = cta_excl_scan(mine, ws, total);
        gmax = __reduce_max_sync(0xffffffffu, gmax);
        if (lane == 0 && gmax) atomicMax(out.maxgS, gmax);
        if (tid == 0) s_actS = total;
        __syncthreads();
    } else if (tid == 0) {
        s_actS = 0;
    }
    __syncthreads();
    // ---- reservations: S (outlier sub-groups, then the pivot block if it is small), B (pivot block / big sub-groups),
    // records of changed ranks
    const bool eq_act = neq >= 2, eq_big = neq > gcap;
    // Ranks.  A group's rank only has to be a slot inside the group's own slot range [gbase, gbase + M): ranges of
    // different groups are disjoint, so any such representative orders the groups correctly, and a singleton's
    // rank is its slot (the inverse suffix array in the end).  Outlier sub-groups take their head slot.  The pivot
    // block KEEPS the group's rank whenever that slot still lies inside the block's range, and otherwise takes the
    // slot in the middle of its range -- a tandem-array group sheds outliers from the same side in every round,
    // and with the head slot as its name all of its members (and, on several GPUs, one record each for every
    // replica) would be renamed every round; named by a middle slot it is renamed O(log) times.
    const u32 rel = g - gbase;
    const u32 eq_rank = (rel >= nlt && rel < nlt + neq) ? g : gbase + nlt + (neq >> 1);
    const bool eq_changed = eq_rank != g;
    // the one outlier sub-group that keeps the rank: its head slot is the old rank
    u32 qu = 0xFFFFFFFFu, usz = 0;
    {
        const u32 qc = rel < nlt ? rel : (rel >= nlt + neq ? rel - neq : 0xFFFFFFFFu);
        if (qc < nout && ohead[qc] == qc) { qu = qc; usz = osize[qc]; }
    }
    const u32 n_changed = (eq_changed ? neq : 0u) + nout - usz;
    if (tid == 0) {
        const u32 toS = s_actS + ((eq_act && !eq_big) ? neq : 0u);
        s_baseS = toS ? atomicAdd(out.mS, toS) : 0u;
        if (eq_act && !eq_big) atomicMax(out.maxgS, neq);
        s_baseEq = 0;
        if (eq_big) s_baseEq = out.end - (atomicAdd(out.mB, neq) + neq);
        for (u32 k = 0; k < s_nbig; ++k) {
            const u32 sz = osize[big_head[k]];
            big_base[k] = out.end - (atomicAdd(out.mB, sz) + sz);
        }
        s_baseU = (RANK.upd && n_changed) ? atomicAdd(RANK.upd_count, n_changed) : 0u;
    }
    __syncthreads();
    const u32 baseS = s_baseS, baseU = s_baseU;
    const u32 baseEq = eq_big ? s_baseEq : baseS + s_actS;         // list position of the pivot block's first member
    // ---- outliers out
    if (nout) {
        const u32 q0 = tid * PER;
        u32 sidx = actS_before;
#pragma unroll 1
        for (u32 i = 0; i < PER; ++i) {
            const u32 q = q0 + i;
            if (q >= nout) break;
            const u64 kv = okv[q];
            const u32 s = (u32)(kv & OMASK);
            const u32 hq = ohead[q], sz = osize[hq];
            const u32 pos = q < nlt ? q : neq + q;                  // position inside the group
            const u32 hpos = hq < nlt ? hq : neq + hq;
            const u32 slot = gbase + pos, newrank = gbase + hpos;
            SA[slot - RANK.base] = s;
            if (newrank != g) {
                if (RANK.rank) RANK.rank[s] = newrank;
                // record index: rank among the changed outliers (sorted order minus the sub-group that kept the rank)
                if (RANK.upd) RANK.upd[baseU + (q - ((qu != 0xFFFFFFFFu && q >= qu + usz) ? usz : 0u))] = ((u64)newrank << 32) | (u64)s;
            }
            if (sz >= 2) {
                u32 dst;
                if (sz <= gcap) dst = baseS + sidx++;
                else {
                    u32 k = 0;
                    while (big_head[k] != hq) ++k;
                    dst = big_base[k] + (q - hq);
                }
                out.key_next[dst] = (u64)newrank << GS;
                out.val_next[dst] = s;
                out.slot_next[dst] = slot;
            }
        }
    }
    // ---- pass B: the pivot block (any order among its members is a valid order of equal keys)
    {
        const u32 newrank = eq_rank;
        const bool changed = eq_changed;
        const u32 slot0 = gbase + nlt;
        for (u32 jb = head; jb < gend; jb += GS_THREADS) {
            const u32 j = jb + tid;
            bool eq = false;
            if (j < gend) eq = KH::rank(key_in[j]) == piv;
            const u32 em = __ballot_sync(0xffffffffu, eq);
            if (!em) continue;
            u32 base = 0;
            if (lane == 0) base = atomicAdd(&s_eq, __popc(em));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (eq) {
                const u32 e = base + __popc(em & lanemask_lt());
                const u32 s = val_in[j];
                const u32 slot = slot0 + e;
                SA[slot - RANK.base] = s;
                if (changed) {
                    if (RANK.rank) RANK.rank[s] = newrank;
                    // records: [changed outliers | pivot block]
                    if (RANK.upd) RANK.upd[baseU + (nout - usz) + e] = ((u64)newrank << 32) | (u64)s;
                }
                if (eq_act) {
                    const u32 dst = baseEq + e;
                    out.key_next[dst] = (u64)newrank << GS;
                    out.val_next[dst] = s;
                    out.slot_next[dst] = slot;
                }
            }
        }
    }
}

// One CTA per chunk of gcap list positions (a chunk holds at most one group head), or -- BIG pass of the two-size
// launch -- persistent CTAs over the overflow list.
template <int GS, u32 CAP>
__global__ void __launch_bounds__(GS_THREADS, CAP <= GS_CAP_SMALL ? 2 : 1)
k_group_stream(const u64* __restrict__ key_in, const u32* __restrict__ val_in, const u32* __restrict__ slot_in,
               u32 b0, u32 mB, u32 gcap, u32* __restrict__ SA, RankDst RANK, StreamOut out, int dbg) {
    extern __shared__ __align__(16) unsigned char gs_shared[];
    if (out.ovf_in) {
        const u32 n = *out.ovf_cnt_in;
        for (u32 q = blockIdx.x; q < n; q += gridDim.x) {
            gs_process_chunk<GS, CAP>(key_in, val_in, slot_in, b0, mB, gcap, SA, RANK, out, dbg, out.ovf_in[q], gs_shared);
            __syncthreads();
        }
        return;
    }
    gs_process_chunk<GS, CAP>(key_in, val_in, slot_in, b0, mB, gcap, SA, RANK, out, dbg, blockIdx.x, gs_shared);
}

}  // namespace nlz
