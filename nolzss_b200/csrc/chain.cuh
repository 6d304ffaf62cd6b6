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

__global__ void __launch_bounds__(CH_THREADS)
k_chain_exit(const u64* __restrict__ LR, u32 nfac, u32* __restrict__ EXIT, u32* __restrict__ alist,
             u32* __restrict__ acount) {
    __shared__ u32 nx[2][CH_CHUNK];
    __shared__ u32 s_cnt, s_base;
    const u32 base = blockIdx.x * CH_CHUNK;
    u32 end = base + CH_CHUNK;
    if (end > nfac) end = nfac;
    if (threadIdx.x == 0) s_cnt = 0;
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t) {
        u32 o = t * CH_THREADS + threadIdx.x;
        u32 e = base + o;
        nx[0][o] = (e < nfac) ? e + (u32)LR[e] : 0xFFFFFFFFu;
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
    // publish exits; collect the distinct in-range exit values (adjacent de-duplication)
    u32 myv[CH_PER_THREAD];
    u32 myslot[CH_PER_THREAD];
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t) {
        u32 o = t * CH_THREADS + threadIdx.x;
        u32 e = base + o;
        myslot[t] = 0xFFFFFFFFu;
        myv[t] = 0;
        if (e < nfac) {
            u32 v = nx[cur][o];
            EXIT[e] = v;
            bool fresh = (o == 0) || (nx[cur][o - 1] != v);
            if (v < nfac && fresh) { myv[t] = v; myslot[t] = atomicAdd(&s_cnt, 1u); }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base = s_cnt ? atomicAdd(acount, s_cnt) : 0u;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CH_PER_THREAD; ++t)
        if (myslot[t] != 0xFFFFFFFFu) alist[s_base + myslot[t]] = myv[t];
}

__global__ void k_chain_init(u32* __restrict__ alist, u32* __restrict__ acount, u8* __restrict__ REACH, u32 start_pos) {
    alist[0] = start_pos;
    *acount = 1;
    REACH[start_pos] = 1;
}

// one doubling round over the exit-value nodes: propagate reachability, then J <- J o J
__global__ void __launch_bounds__(256)
k_chain_double(const u32* __restrict__ alist, const u32* __restrict__ acount, const u32* __restrict__ J,
               u32* __restrict__ Jn, u8* __restrict__ REACH, u32 nfac) {
    const u32 cnt = *acount;
    for (u32 idx = blockIdx.x * 256 + threadIdx.x; idx < cnt; idx += gridDim.x * 256) {
        u32 x = alist[idx];
        u32 j = J[x];
        if (j < nfac) {
            if (REACH[x]) REACH[j] = 1;
            Jn[x] = J[j];
        } else {
            Jn[x] = j;
        }
    }
}

__global__ void __launch_bounds__(CH_THREADS)
k_chain_mark(const u64* __restrict__ LR, u32 nfac, const u8* __restrict__ REACH, u32* __restrict__ MASK,
             u32* __restrict__ CNT) {
    __shared__ u32 nx[CH_CHUNK];
    __shared__ u8 on[CH_CHUNK];
    __shared__ u32 s_entry;
    const u32 base = blockIdx.x * CH_CHUNK;
    u32 valid = nfac - base;
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
template <bool RC, bool BATCH>
__global__ void __launch_bounds__(CH_THREADS)
k_chain_emit(const u64* __restrict__ LR, u32 nfac, const u32* __restrict__ MASK, const u32* __restrict__ OFF,
             u64* __restrict__ out, u64 out_capacity, BatchView bv, u32* __restrict__ sentidx) {
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
            u32 pos = base + o;
            u64 idx = off + wpre[wi] + __popc(wbits & lanemask_lt());
            u32 shift = 0;
            if (BATCH) {
                const u32 b = bv.REC[pos];
                shift = bv.fstart[b];
                if (pos == shift + bv.flen[b]) { sentidx[b] = (u32)idx; continue; }
                idx -= b;
            }
            if (idx < out_capacity) {
                u64 lr = LR[pos];
                u32 ref32 = (u32)(lr >> 32);
                if (BATCH) { ref32 -= shift; pos -= shift; }
                u64 ref = RC ? ((u64)(ref32 & ~LR_RC_FLAG) | ((ref32 & LR_RC_FLAG) ? (1ULL << 63) : 0ULL))
                             : (u64)ref32;
                out[3 * idx + 0] = pos;
                out[3 * idx + 1] = (u32)lr;
                out[3 * idx + 2] = ref;
            }
        }
    }
}

}  // namespace nlz
