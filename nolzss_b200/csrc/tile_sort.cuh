// Doubling round, fused: gather the second key half RANK[s+h] and sort every tie group by it, in
// shared memory, in ONE pass over the active set (instead of 6 LSD radix passes over HBM).
//
// The active list is a concatenation of tie groups (equal high key half g = group head slot), each
// contiguous.  CTA t owns the groups whose head lies in [t*tile, (t+1)*tile); because no group is
// larger than `maxg` (checked on the host from the previous regroup), its elements fit in
// tile + maxg <= 4096 shared-memory slots.  A bitonic network on the 64-bit composite key
// (g, RANK[s+h]) sorts all owned groups at once; foreign slots are padded with the maximum key.
// Rounds whose largest group exceeds the capacity fall back to the global radix sort.
#pragma once
#include "common.cuh"

namespace nlz {

constexpr int TSORT_THREADS = 256;
constexpr int TSORT_SLOTS = 4096;
constexpr size_t TSORT_SMEM = (size_t)TSORT_SLOTS * 12 + 16;
constexpr u32 TSORT_ALLPAIRS_MAX = 256;   // CTA-local group size up to which counting beats the bitonic network

__global__ void __launch_bounds__(TSORT_THREADS)
k_tile_sort(const u64* __restrict__ key_in, const u32* __restrict__ val_in, u32 m,
            const u32* __restrict__ RANK, u64 h, u32 n1, u32 tile, u32 maxg,
            u64* __restrict__ key_out, u32* __restrict__ val_out) {
    extern __shared__ __align__(16) unsigned char tsort_smem[];
    u64* skey = reinterpret_cast<u64*>(tsort_smem);
    u32* sval = reinterpret_cast<u32*>(tsort_smem + (size_t)TSORT_SLOTS * 8);
    u32& s_first = *reinterpret_cast<u32*>(tsort_smem + (size_t)TSORT_SLOTS * 12);
    u32& s_end = *reinterpret_cast<u32*>(tsort_smem + (size_t)TSORT_SLOTS * 12 + 4);
    const u32 a = blockIdx.x * tile;
    u32 b = a + tile;
    if (b > m) b = m;
    u32 load_end = b + maxg;          // a group headed before b ends before b + maxg
    if (load_end > m) load_end = m;
    const u32 nload = load_end - a;   // <= TSORT_SLOTS
    if (threadIdx.x == 0) { s_first = 0xFFFFFFFFu; s_end = load_end; }
    __syncthreads();
    // pass 1: find the first owned head (>= a) and the first foreign head (>= b)
    for (u32 o = threadIdx.x; o < nload; o += TSORT_THREADS) {
        const u32 j = a + o;
        const u32 g = (u32)(key_in[j] >> 32);
        const bool head = (j == 0) || ((u32)(key_in[j - 1] >> 32) != g);
        if (head) {
            if (j < b) atomicMin(&s_first, j);
            else atomicMin(&s_end, j);
        }
    }
    __syncthreads();
    const u32 first = s_first, end = s_end;
    if (first == 0xFFFFFFFFu) return;             // no group starts in this tile
    const u32 cnt = end - first;                  // owned elements [first, end)
    u32 ns = 32;
    while (ns < cnt) ns <<= 1;
    // pass 2: build composite keys for the owned elements, pad the rest
    for (u32 o = threadIdx.x; o < ns; o += TSORT_THREADS) {
        u64 k = ~0ull;
        u32 v = 0;
        if (o < cnt) {
            const u32 j = first + o;
            v = val_in[j];
            const u64 p = (u64)v + h;
            const u32 r = p < n1 ? RANK[p] : 0u;
            k = (key_in[j] & 0xFFFFFFFF00000000ull) | (u64)r;
        }
        skey[o] = k;
        sval[o] = v;
    }
    __syncthreads();
    // largest owned group (heads measure their group by a forward scan)
    u32& s_gmax = *reinterpret_cast<u32*>(tsort_smem + (size_t)TSORT_SLOTS * 12 + 8);
    if (threadIdx.x == 0) s_gmax = 0;
    __syncthreads();
    for (u32 o = threadIdx.x; o < cnt; o += TSORT_THREADS) {
        const u32 g = (u32)(skey[o] >> 32);
        if (o == 0 || (u32)(skey[o - 1] >> 32) != g) {
            u32 e = o + 1;
            while (e < cnt && (u32)(skey[e] >> 32) == g) ++e;
            atomicMax(&s_gmax, e - o);
        }
    }
    __syncthreads();
    if (s_gmax <= TSORT_ALLPAIRS_MAX) {
        // small groups: rank every element inside its group by counting (no barriers, ~s steps)
        for (u32 o = threadIdx.x; o < cnt; o += TSORT_THREADS) {
            const u64 k = skey[o];
            const u32 g = (u32)(k >> 32);
            u32 before = 0, smaller = 0;
            for (u32 q = o; q-- > 0;) {
                const u64 kq = skey[q];
                if ((u32)(kq >> 32) != g) break;
                ++before;
                smaller += (kq <= k) ? 1u : 0u;       // ties keep their current order
            }
            for (u32 q = o + 1; q < cnt; ++q) {
                const u64 kq = skey[q];
                if ((u32)(kq >> 32) != g) break;
                smaller += (kq < k) ? 1u : 0u;
            }
            const u32 dst = first + (o - before) + smaller;
            key_out[dst] = k;
            val_out[dst] = sval[o];
        }
        return;
    }
    // bitonic sort of ns slots by key
    for (u32 k = 2, lk = 1; k <= ns; k <<= 1, ++lk) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            for (u32 t = threadIdx.x; t < (ns >> 1); t += TSORT_THREADS) {
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
    for (u32 o = threadIdx.x; o < cnt; o += TSORT_THREADS) {
        key_out[first + o] = skey[o];
        val_out[first + o] = sval[o];
    }
}

}  // namespace nlz
