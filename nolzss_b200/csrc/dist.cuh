// Pieces shared by the distributed path (dist2.cuh / dist2_host.cuh): the stream-ordered flag barrier in peer memory
// with its payload exchange, the edge staircases of the boundary-exchange pass of stage 3, and a three-launch exclusive
// scan.  (Round 1's distributed design -- RANK replicated on every GPU, per-position results gathered on GPU 0,
// 32-bit only -- lived here; round 2 replaced it, see dist2.cuh.)
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
