// Stage 4: extract the greedy factor chain  i -> i + len(i)  from start_pos, without host round trips.
//
// Replaces the sequential "emit, then next_leaf(lambda, len)" loop of the reference
// (/root/reference/src/cpp/factorizer_core.hpp:66,111-115 and :241,376-379; next_leaf =
// factorizer_helpers.hpp:38-44).  The reference's own parallel mode already relies on the chain
// being a function of the position only (parallel_factorizer.cpp:499-569, convergence of chunked
// chains); here that idea is made work-efficient and exact:
//   1. every chunk of 1024 positions resolves, in shared memory (pointer jumping, 10 rounds), the
//      first chain position OUTSIDE the chunk reached from each of its positions (EXIT);
//   2. the distinct exit values (few per chunk, because chains converge) form the only nodes whose
//      pointers must be doubled globally: ceil(log2(#chunks))+1 tiny rounds mark every chunk's true
//      entry point reachable from start_pos;
//   3. every chunk walks from its entry in shared memory, publishes a bitmask of chain positions
//      and its factor count; after a scan the triples (start, length, ref) are written in order.
#pragma once
#include "common.cuh"
#include "lpnf.cuh"
#include "sa.cuh"

namespace nlz {

constexpr int CH_THREADS = 256;
constexpr int CH_CHUNK = 1024;
constexpr int CH_PER_THREAD = CH_CHUNK / CH_THREADS;
constexpr int CH_ROUNDS = 10;  // 2^10 = CH_CHUNK hops
constexpr int CH_HASH_BITS = 11, CH_HASH = 1 << CH_HASH_BITS;

// One text across G GPUs (dist2.cuh): the positions [0, nfac) are dealt in G slices of `chunk` positions (a multiple
// of CH_CHUNK); GPU g holds LR / FLAGS / MASK of its slice [t0, t1) and the slice of the chain-node arrays (exit
// pointers J, their double buffer, reach flags).  Steps 1 and 3 are local to a slice; the exit-node doubling of step 2
// follows pointers into other slices through peer memory (a few nodes per chunk: the traffic is tiny), one
// cross-GPU barrier per doubling round.  One GPU: G = 1, t0 = 0, t1 = nfac.
struct ChainDom {
    u32 t0, t1, nfac;     // local slice and global end
    u32 chunk;            // positions per GPU
    int G;
    u32* J[MAX_PEERS];    // exit pointers, slice of GPU g (double buffer A)
    u32* J2[MAX_PEERS];   // double buffer B
    u8* REACH[MAX_PEERS];
};
__device__ __forceinline__ u32 cd_owner(const ChainDom& d, u32 x, u32& off) {
    if (d.G == 1) { off = x; return 0; }
    const u32 g = x / d.chunk;
    off = x - g * d.chunk;
    return g;
}

// LR, EXIT: this GPU's slice (index = position - t0)
__global__ void __launch_bounds__(CH_THREADS)
k_chain_exit(const u64* __restrict__ LR, u32 t0, u32 t1, u32 nfac, u32* __restrict__ EXIT, u32* __restrict__ alist,
             u32* __restrict__ acount) {
    __shared__ u32 nx[2][CH_CHUNK];
    __shared__ u32 hs[CH_HASH];            // exit values already listed by this chunk
    __shared__ u32 s_cnt, s_base;
    const u32 base = t0 + blockIdx.x * CH_CHUNK;
    u32 end = base + CH_CHUNK;
    if (end > t1) end = t1;
    if (threadIdx.x == 0) s_cnt = 0;
    for (int i = threadIdx.x; i < CH_HASH; i += CH_THREADS) hs[i] = 0xFFFFFFFFu;
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t) {
        u32 o = t * CH_THREADS + threadIdx.x;
        u32 e = base + o;
        nx[0][o] = (e < t1) ? e + (u32)LR[e - t0] : 0xFFFFFFFFu;
    }
    __syncthreads();
    int cur = 0;
#pragma unroll 1
    for (int round = 0; round < CH_ROUNDS; ++round) {
#pragma unroll
        for (int t = 0; t < CH_PER_THREAD; ++t) {
            u32 o = t * CH_THREADS + threadIdx.x;
            u32 v = nx[cur][o];
            if (v < end) v = nx[cur][v - base];
            nx[cur ^ 1][o] = v;
        }
        __syncthreads();
        cur ^= 1;
    }
    // publish exits; collect the DISTINCT in-range exit values.  The chains of a chunk converge onto a handful of exits,
    // but neighbouring positions sit on different chains, so equal values interleave: a small shared-memory hash set
    // keeps each value once (round 1 de-duplicated adjacent positions only and listed ~50 nodes per chunk; every listed
    // node costs three scattered accesses in each of the ~20 doubling rounds)
    u32 myv[CH_PER_THREAD];
    u32 myslot[CH_PER_THREAD];
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t) {
        u32 o = t * CH_THREADS + threadIdx.x;
        u32 e = base + o;
        myslot[t] = 0xFFFFFFFFu;
        myv[t] = 0;
        if (e < t1) {
            u32 v = nx[cur][o];
            EXIT[e - t0] = v;
            bool fresh = v < nfac && ((o == 0) || (nx[cur][o - 1] != v));
            if (fresh) {
                u32 slot = (v * 2654435761u) >> (32 - CH_HASH_BITS);
                for (int probe = 0; probe < 8; ++probe) {
                    const u32 old = atomicCAS(&hs[slot], 0xFFFFFFFFu, v);
                    if (old == 0xFFFFFFFFu) break;                   // first to list v
                    if (old == v) { fresh = false; break; }          // already listed
                    slot = (slot + 1) & (CH_HASH - 1);               // (table full around here: list v again, harmless)
                }
            }
            if (fresh) { myv[t] = v; myslot[t] = atomicAdd(&s_cnt, 1u); }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = s_cnt ? atomicAdd(acount, s_cnt) : 0u;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t)
        if (myslot[t] != 0xFFFFFFFFu) alist[s_base + myslot[t]] = myv[t];
}

// the owner of start_pos seeds its node list with it (reach flag set); every other GPU starts with an empty list
__global__ void k_chain_init(u32* __restrict__ alist, u32* __restrict__ acount, u8* __restrict__ REACH, u32 start_pos,
                             u32 t0, u32 t1) {
    if (start_pos >= t0 && start_pos < t1) {
        alist[0] = start_pos;
        *acount = 1;
        REACH[start_pos - t0] = 1;
    } else {
        *acount = 0;
    }
}

// one doubling round over the exit-value nodes: propagate reachability, then J <- J o J
// `swap`: the roles of the two pointer buffers in this round (J read, J2 written, or the reverse)
__global__ void __launch_bounds__(256)
k_chain_double(const u32* __restrict__ alist, const u32* __restrict__ acount, ChainDom d, int swap) {
    const u32 cnt = *acount;
    for (u32 idx = blockIdx.x * 256 + threadIdx.x; idx < cnt; idx += gridDim.x * 256) {
        const u32 x = alist[idx];
        u32 xo, jo;
        const u32 xg = cd_owner(d, x, xo);
        const u32* Jx = swap ? d.J2[xg] : d.J[xg];
        u32* Jnx = swap ? d.J[xg] : d.J2[xg];
        const u32 j = Jx[xo];
        if (j < d.nfac) {
            const u32 jg = cd_owner(d, j, jo);
            if (d.REACH[xg][xo]) d.REACH[jg][jo] = 1;
            Jnx[xo] = (swap ? d.J2[jg] : d.J[jg])[jo];
        } else {
            Jnx[xo] = j;
        }
    }
}

// LR, REACH, MASK: this GPU's slice (local indices); n_loc = t1 - t0
__global__ void __launch_bounds__(CH_THREADS)
k_chain_mark(const u64* __restrict__ LR, u32 n_loc, const u8* __restrict__ REACH, u32* __restrict__ MASK,
             u32* __restrict__ CNT) {
    __shared__ u32 nx[CH_CHUNK];
    __shared__ u8 on[CH_CHUNK];
    __shared__ u32 s_entry;
    const u32 base = blockIdx.x * CH_CHUNK;
    u32 valid = n_loc - base;
    if (valid > CH_CHUNK) valid = CH_CHUNK;
    if (threadIdx.x == 0) s_entry = 0xFFFFFFFFu;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t) {
        u32 o = t * CH_THREADS + threadIdx.x;
        on[o] = 0;
        if (o < valid) {
            nx[o] = o + (u32)LR[base + o];
            if (REACH[base + o]) atomicMin(&s_entry, o);
        } else {
            nx[o] = 0xFFFFFFFFu;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 o = s_entry, c = 0;
        while (o < valid) { on[o] = 1; ++c; o = nx[o]; }
        CNT[blockIdx.x] = c;
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t) {
        u32 o = t * CH_THREADS + threadIdx.x;
        u32 bits = __ballot_sync(0xffffffffu, on[o] != 0);
        if ((threadIdx.x & 31) == 0) MASK[(base + o) >> 5] = bits;
    }
}

// BATCH: positions are translated to record-local coordinates; the literal factor of every interior
// sentinel (one per record but the last) is dropped -- record b's factors are preceded by exactly b of
// them, so the output index is idx - b -- and its index is published so that the host can derive the
// per-record factor counts.
// LR, FLAGS, MASK: this GPU's slice (local indices); t0 = first position of the slice; the output index is local too.
template <bool RC, bool BATCH>
__global__ void __launch_bounds__(CH_THREADS)
k_chain_emit(const u64* __restrict__ LR, const u8* __restrict__ FLAGS, u32 t0, const u32* __restrict__ MASK,
             const u32* __restrict__ OFF, u64* __restrict__ out, u64 out_capacity, BatchView bv, u32* __restrict__ sentidx) {
    __shared__ u32 wpre[CH_CHUNK / 32];
    const u32 base = blockIdx.x * CH_CHUNK;
    const u32 lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        u32 wbits = MASK[(base >> 5) + lane];
        u32 c = __popc(wbits), inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        wpre[lane] = inc - c;
    }
    __syncthreads();
    const u64 off = OFF[blockIdx.x];
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t) {
        u32 o = t * CH_THREADS + threadIdx.x;
        u32 wi = o >> 5;
        u32 wbits = MASK[(base >> 5) + wi];
        if ((wbits >> lane) & 1u) {
            const u32 lp = base + o;                      // local index
            u32 pos = t0 + lp;
            u64 idx = off + wpre[wi] + __popc(wbits & lanemask_lt());
            u32 shift = 0;
            if (BATCH) {
                const u32 b = bv.REC[pos];
                shift = bv.fstart[b];
                if (pos == shift + bv.flen[b]) { sentidx[b] = (u32)idx; continue; }
                idx -= b;
            }
            if (idx < out_capacity) {
                u64 lr = LR[lp];
                u32 ref32 = (u32)(lr >> 32);
                const bool is_rc = RC && (FLAGS[lp] & FLAG_RC) != 0;
                if (BATCH) { ref32 -= shift; pos -= shift; }
                u64 ref = (u64)ref32 | (is_rc ? (1ULL << 63) : 0ULL);
                out[3 * idx + 0] = pos;
                out[3 * idx + 1] = (u32)lr;
                out[3 * idx + 2] = ref;
            }
        }
    }
}

}  // namespace nlz
