// One text across G GPUs (configs[3]/[4] of BASELINE.json).
//
// The reference builds ONE suffix tree serially even in its parallel mode
// (/root/reference/src/cpp/parallel_factorizer.cpp:78-84).  Here the suffix array is partitioned by
// RANK RANGE: a histogram of the leading key bits (the first 12 bases) of every suffix gives G bucket
// ranges of equal size; GPU g keeps, sorts and refines exactly the suffixes of its range.  Tie groups of
// prefix doubling are contiguous in rank order and never leave their range, so every doubling round is
// local -- the only thing a round needs from other GPUs is RANK[s+h], and that array is kept as a full
// replica on every GPU: the kernels that refine ranks store them into all replicas through peer pointers
// over NVLink (fused compute + exchange; no staging buffers, no NCCL call on the data path).
//   * text:  replicated (<= 1 byte per base);        RANK (= ISA): replicated, written by every GPU;
//   * SA, LCP, node tables, sort buffers: partitioned by rank range;
//   * LCP:   Kasai needs text order -> positions are dealt in G equal slices; the owner of rank r sends
//            Phi[SA[r]] = SA[r-1] to the owner of position SA[r], which sends LCP[r] back (peer stores);
//   * stage 3 runs in rank space on the local range.  LCP intervals that cross a range edge have string
//            depth < 12 (they are separated by the bucket prefix); each GPU publishes, for either edge,
//            a "staircase" -- per depth v the minimum forward start / maximum rc start of the ranks that
//            connect to the edge at depth v -- and every GPU appends its neighbours' staircases to its
//            local arrays as <= 64 virtual ranks per side (boundary-exchange pass).  Seen from a local
//            leaf, a virtual rank is indistinguishable from the block of remote suffixes it stands for.
//   * per-position results are pushed to GPU 0, which extracts the chain.
// Synchronisation is a stream-ordered flag barrier in peer memory (k_dist_barrier); it also carries the
// small per-round payloads (active counts, edge staircases), so the round loop needs no host collective.
#pragma once
#include "common.cuh"
#include "lpnf.cuh"
#include "sa.cuh"

namespace nlz {

constexpr int DIST_XCH_WORDS = 512;        // payload words per rank per exchange
constexpr int DIST_VIRT = 64;              // virtual ranks reserved on either side of the local range
constexpr int DIST_MAX_DEPTH = 24;         // crossing nodes are shallower than this many symbols

// control block at the start of every rank's shared segment
struct DistCtl {
    u32 flags[MAX_PEERS];                              // flags[g] = last barrier epoch rank g has reached
    u32 error;                                         // set by a barrier that timed out
    u32 pad[7];
    u32 xch[2][MAX_PEERS][DIST_XCH_WORDS];             // exchange slots, double-buffered by exchange parity
    u32 packed[MAX_PEERS * DIST_XCH_WORDS + 8];        // after a payload barrier: [g * nwords + w] of every rank, then the time-out flag
};

struct DistPeers {
    DistCtl* ctl[MAX_PEERS];
    int n;
    int me;
};

// All ranks launch this kernel at the same point of their (identical) launch sequences.  Posts
// `nwords` words from `src` into slot [parity][me] of every peer, raises this rank's flag at every peer
// to `epoch`, then waits until every peer's flag here has reached `epoch`.  Stream order gives the
// rest: everything this rank launched before the barrier has completed (its peer stores are performed),
// and nothing launched after it starts before all ranks have arrived.
__global__ void __launch_bounds__(256)
k_dist_barrier(DistPeers peers, u32 epoch, int parity, const u32* __restrict__ src, u32 nwords, u32 timeout_s) {
    const int G = peers.n, me = peers.me;
    for (u32 i = threadIdx.x; i < nwords * (u32)G; i += blockDim.x) {
        const int g = (int)(i / nwords);
        const u32 wd = i % nwords;
        peers.ctl[g]->xch[parity][me][wd] = src[wd];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < (u32)G) {
        const int g = (int)threadIdx.x;
        volatile u32* theirs = &peers.ctl[g]->flags[me];
        *theirs = epoch;
        __threadfence_system();
        volatile u32* mine = &peers.ctl[me]->flags[g];
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while ((int)(*mine - epoch) < 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > (unsigned long long)timeout_s * 1000000000ull) { peers.ctl[me]->error = epoch; break; }   // a peer died: do not hang the GPU
            __nanosleep(200);
        }
    }
    __threadfence_system();
    if (nwords) {                                       // pack everybody's words for ONE host readback
        __syncthreads();
        DistCtl* my = peers.ctl[me];
        for (u32 i = threadIdx.x; i < nwords * (u32)G; i += blockDim.x)
            my->packed[i] = *(volatile u32*)&my->xch[parity][i / nwords][i % nwords];
        if (threadIdx.x == 0) my->packed[nwords * (u32)G] = *(volatile u32*)&my->error;
    }
}

// splitters from the scanned prefix histogram: rank g owns buckets [split[g], split[g+1]) = global
// ranks [base[g], base[g+1]).  cum = exclusive scan of the histogram (nb entries), total = n1.
__global__ void k_dist_splitters(const u32* __restrict__ cum, u32 nb, u32 n1, int G, u32* __restrict__ split,
                                 u32* __restrict__ base) {
    const int g = threadIdx.x;
    if (g > G) return;
    if (g == 0) { split[0] = 0; base[0] = 0; return; }
    if (g == G) { split[G] = nb; base[G] = n1; return; }
    const u64 target = (u64)n1 * g / G;
    u32 lo = 0, hi = nb - 1;                       // largest bucket with cum[bucket] <= target
    while (lo < hi) {
        const u32 mid = lo + (hi - lo + 1) / 2;
        if (cum[mid] <= target) lo = mid; else hi = mid - 1;
    }
    split[g] = lo;
    base[g] = cum[lo];
}

// ---- bucketed pair exchange -----------------------------------------------------------------------
// Phi and LCP change owners (rank owner <-> position owner).  Scattered peer stores collapse when most of
// them are remote, so every GPU first buckets its (destination-local index, value) pairs by destination
// GPU into a contiguous staging list (order inside a bucket is irrelevant), the buckets travel as bulk
// copies, and the receiver scatters them locally (k_dist_apply_pairs).
//   SRC 0 (Phi):  item = local rank r;      index = SA[r] - owner * chunk, value = SA[r-1]   (left_sa for r = 0)
//   SRC 1 (LCP):  item = local position t;  index = RANK[pos0 + t] - base[owner], value = PLCP[t]
struct PairSrc {
    const u32* SA; u32 left_sa; u32 chunk;            // SRC 0
    const u32* RANK; const u32* PLCP; u32 pos0;      // SRC 1
    u32 base[MAX_PEERS + 1];
    int G;
    u32 count;                                        // items
};
template <int SRC>
__device__ __forceinline__ u64 pair_of(const PairSrc& ps, u32 t, u32& dest) {
    if (SRC == 0) {
        const u32 s = ps.SA[t];
        const u32 prev = t ? ps.SA[t - 1] : ps.left_sa;
        dest = s / ps.chunk;
        return ((u64)prev << 32) | (u64)(s - dest * ps.chunk);
    }
    const u32 r = ps.RANK[ps.pos0 + t];
    u32 g = 0;
    while ((int)g + 1 < ps.G && r >= ps.base[g + 1]) ++g;
    dest = g;
    return ((u64)ps.PLCP[t] << 32) | (u64)(r - ps.base[g]);
}
// PASS 0: counts[dest] += ...;  PASS 1: staging[cursor[dest]++] = pair (cursor initialised to the bucket starts)
template <int SRC, int PASS>
__global__ void __launch_bounds__(256)
k_dist_bucket_pairs(PairSrc ps, u32* __restrict__ counts_or_cursor, u64* __restrict__ staging) {
    __shared__ u32 cnt[MAX_PEERS], basev[MAX_PEERS];
    if (threadIdx.x < MAX_PEERS) cnt[threadIdx.x] = 0;
    __syncthreads();
    const u32 t = blockIdx.x * 256 + threadIdx.x;
    const bool valid = t < ps.count;
    u32 dest = 0xFFu;
    u64 pr = 0;
    if (valid) pr = pair_of<SRC>(ps, t, dest);
    const u32 same = __match_any_sync(0xffffffffu, dest);
    const u32 leader = __ffs(same) - 1, lane = threadIdx.x & 31;
    u32 off = 0;
    if (valid && lane == leader) off = atomicAdd(&cnt[dest], __popc(same));
    off = __shfl_sync(0xffffffffu, off, leader) + __popc(same & lanemask_lt());
    __syncthreads();
    if (threadIdx.x < MAX_PEERS && cnt[threadIdx.x]) basev[threadIdx.x] = atomicAdd(&counts_or_cursor[threadIdx.x], cnt[threadIdx.x]);
    if (PASS == 0) return;
    __syncthreads();
    if (valid) staging[basev[dest] + off] = pr;
}
__global__ void k_dist_bucket_starts(const u32* __restrict__ counts, u32* __restrict__ cursor, int G) {
    if (threadIdx.x == 0) { u32 run = 0; for (int g = 0; g < G; ++g) { cursor[g] = run; run += counts[g]; } }
}
__global__ void __launch_bounds__(256)
k_dist_apply_pairs(const u64* __restrict__ pairs, u32 cnt, u32* __restrict__ dst) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u64 u = pairs[e];
    dst[(u32)u] = (u32)(u >> 32);
}

// Edge staircases of the local rank range (real ranks [0, m) of the arrays in T, with LCP[0] and LCP[m]
// temporarily 0).  For depth v in 1..K (v = K stands for ">= K"):
//   right edge (seen by the ranges to the right): the ranks that connect to rank m-1 at depth >= v are
//     [q_v, m-1], q_v = largest index <= m-1 with LCP < v;   block(v) = [q_v, q_{v+1}-1], block(K) = [q_K, m-1]
//   left edge (seen by the ranges to the left): the ranks that connect to rank 0 at depth >= v are
//     [0, p_v-1], p_v = smallest index >= 1 with LCP < v;     block(v) = [p_{v+1}, p_v-1], block(K) = [0, p_K-1]
// out: for side s (0 = right edge, 1 = left edge) and v: out[(s*K + v-1)*4 + {0,1,2,3}] =
//      {ranks in the block, min forward start, max rc start, q_v == 0 / p_v == m (the whole range connects)}
template <bool RC>
__global__ void k_dist_edges(Trees T, WalkParams p, u32 m, int K, u32* __restrict__ out) {
    const int t = threadIdx.x;
    if (t >= 2 * K) return;
    const int side = t / K;
    const u32 v = (u32)(t % K) + 1;
    u32 lo, hi, whole;     // block = [lo, hi] (empty when lo > hi)
    if (side == 0) {
        const u32 qv = find_prev_less<false>(T, m - 1, v);
        const u32 qn = ((int)v == K) ? m : find_prev_less<false>(T, m - 1, v + 1);
        lo = qv; hi = qn - 1; whole = qv == 0;
    } else {
        const u32 pv = (m > 1) ? find_next_less<false>(T, 1, v) : m;
        const u32 pn = ((int)v == K) ? 0u : ((m > 1) ? find_next_less<false>(T, 1, v + 1) : m);
        lo = pn; hi = pv - 1; whole = pv == m;
    }
    u32 fmin = NONE_MIN, rmax = 0, cnt = 0;
    if (lo <= hi && hi < m) {
        cnt = hi - lo + 1;
        agg_range<RC, RC, false>(T, p, (i64)lo, (i64)hi, fmin, rmax);
    }
    u32* o = out + (size_t)t * 4;
    o[0] = cnt; o[1] = fmin; o[2] = rmax; o[3] = whole;
}

// this GPU's (suffix, rank) records of the round (*cnt of them) -> the inbox slice of every other GPU
struct UpdDst {
    u64* p[MAX_PEERS];
    int n, me;
};
__global__ void __launch_bounds__(256)
k_dist_push_ranks(const u64* __restrict__ src, const u32* __restrict__ cnt, UpdDst dst) {
    const u32 m = *cnt;
    for (u32 e = blockIdx.x * 256 + threadIdx.x; e < m; e += gridDim.x * 256) {
        const u64 u = src[e];
        for (int g = 0; g < dst.n; ++g)
            if (g != dst.me) dst.p[g][e] = u;
    }
}

// (suffix, rank) records received from another GPU -> local RANK replica
__global__ void __launch_bounds__(256)
k_dist_apply_ranks(const u64* __restrict__ upd, u32 cnt, u32* __restrict__ RANK) {
    const u32 e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cnt) return;
    const u64 u = upd[e];
    RANK[(u32)u] = (u32)(u >> 32);
}

// per-position results of the local ranks, in rank order (bulk-copied to GPU 0 together with SA)
__global__ void __launch_bounds__(256)
k_dist_pack_lr(const u32* __restrict__ SA, u32 m, u32 nfac, const u64* __restrict__ LRloc, u64* __restrict__ lval) {
    const u32 r = blockIdx.x * 256 + threadIdx.x;
    if (r >= m) return;
    const u32 i = SA[r];
    lval[r] = i < nfac ? LRloc[i] : 0ull;
}
// GPU 0: LR[pos[r]] = lval[r]
__global__ void __launch_bounds__(256)
k_dist_apply_lr(const u32* __restrict__ pos, const u64* __restrict__ lval, u32 cnt, u32 nfac, u64* __restrict__ LR) {
    const u32 r = blockIdx.x * 256 + threadIdx.x;
    if (r >= cnt) return;
    const u32 i = pos[r];
    if (i < nfac) LR[i] = lval[r];
}

// Exclusive scan of a large u32 array in three launches: per-tile sums, single-CTA scan of the tile
// sums (k_scan_u32_single_cta), per-tile scan with the tile offset.  Tile = 4096 words, 1024 threads.
constexpr int SCAN_TILE = 4096;
template <bool APPLY>
__global__ void __launch_bounds__(1024)
k_scan_tiles(u32* __restrict__ data, u32 count, u32* __restrict__ tile_sums) {
    __shared__ u32 wsum[32];
    const u32 base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    u32 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (base + j < count) ? data[base + j] : 0u;
    const u32 s = v[0] + v[1] + v[2] + v[3];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (u32)o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    u32 carry = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const u32 x = wsum[i]; if (i < (int)w) carry += x; tot += x; }
    if (!APPLY) {
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
        return;
    }
    u32 run = tile_sums[blockIdx.x] + carry + inc - s;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (base + j < count) data[base + j] = run;
        run += v[j];
    }
}

__global__ void k_copy_words(const u32* __restrict__ a, u32* __restrict__ b, u32 n) {
    for (u32 i = threadIdx.x; i < n; i += blockDim.x) b[i] = a[i];
}

}  // namespace nlz
