// Hand-written stable LSD radix sort of (key, u32 value) pairs for sm_100a.
//
// Suffix sorting here is HBM-bound integer work: every pass streams the pairs once for the digit
// histogram and once for the scatter.  Layout choices:
//   * persistent CTAs: at most 4 CTAs per SM (148 SMs), each owning a contiguous run of tiles, so
//     the per-pass histogram table is [256 digits][<=592 CTAs] and is scanned by one CTA;
//   * warp-striped coalesced loads, MATCH.ANY warp ranking (stable), and a shared-memory staged
//     reorder so that every digit run leaves the CTA as consecutive, coalesced stores;
//   * passes are described by a DigitPlan so callers sort only the bit ranges that can differ.
#pragma once
#include "common.cuh"
#include "prof.cuh"

namespace nlz {

struct DigitPlan {
    int npass = 0;
    int shift[16];
    int bits[16];
};

// Cover key bits [lo, hi) with the fewest passes of <= 8 bits, evenly sized.
static inline void plan_add_range(DigitPlan& p, int lo, int hi) {
    int nb = hi - lo;
    if (nb <= 0) return;
    int np = (nb + 7) / 8, base = nb / np, extra = nb % np, s = lo;
    for (int i = 0; i < np; ++i) {
        int b = base + (i < extra ? 1 : 0);
        p.shift[p.npass] = s;
        p.bits[p.npass] = b;
        p.npass++;
        s += b;
    }
}

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;
constexpr int RS_MAX_CTAS = kNumSM * 4;

template <typename KeyT> struct RsCfg;
template <> struct RsCfg<u32> { static constexpr int ITEMS = 16; };
template <> struct RsCfg<u64> { static constexpr int ITEMS = 8; };

// Exclusive scan of one value per thread across a 256-thread CTA. `ws` holds >= 8 words.
__device__ __forceinline__ u32 block_excl_scan_256(u32 v, u32* ws, u32* total = nullptr) {
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    u32 wp = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < RS_WARPS; ++i) {
        u32 x = ws[i];
        if (i < (int)w) wp += x;
        tot += x;
    }
    if (total) *total = tot;
    __syncthreads();
    return wp + inc - v;
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const KeyT* __restrict__ keys, u32 m, int shift, u32 mask, u32 tiles_per_cta,
          u32* __restrict__ hist) {
    constexpr u32 TS = RS_THREADS * RsCfg<KeyT>::ITEMS;
    __shared__ u32 h[RS_WARPS][RS_BINS];
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const u64 begin = (u64)blockIdx.x * tiles_per_cta * TS;
    u64 end = begin + (u64)tiles_per_cta * TS;
    if (end > m) end = m;
    const u32 w = threadIdx.x >> 5;
    for (u64 i = begin + threadIdx.x; i < end; i += RS_THREADS) {
        u32 d = (u32)(keys[i] >> shift) & mask;
        atomicAdd(&h[w][d], 1u);
    }
    __syncthreads();
    u32 s = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) s += h[ww][threadIdx.x];
    hist[threadIdx.x * gridDim.x + blockIdx.x] = s;
}

// Exclusive scan, in place, of `count` words by ONE CTA of 1024 threads (count <= 256*592).
__global__ void __launch_bounds__(1024) k_scan_u32_single_cta(u32* __restrict__ data, u32 count,
                                                             u32* __restrict__ total_out) {
    __shared__ u32 wsum[32];
    const u32 per = (count + 1023) / 1024;
    u32 b = threadIdx.x * per, e = b + per;
    if (b > count) b = count;
    if (e > count) e = count;
    u32 s = 0;
    for (u32 i = b; i < e; ++i) s += data[i];
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 v = wsum[lane], vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        wsum[lane] = vi - v;
        if (lane == 31 && total_out) *total_out = vi;
    }
    __syncthreads();
    u32 run = wsum[w] + inc - s;
    for (u32 i = b; i < e; ++i) {
        u32 t = data[i];
        data[i] = run;
        run += t;
    }
}

// Histogram table is [256 digits][ctas].  One CTA per digit row: exclusive scan across the CTAs of
// that row (coalesced), row total to row_tot[digit].  The cross-digit prefix is taken by every
// scatter CTA itself from the 256 row totals, so no serial whole-table scan remains.
__global__ void __launch_bounds__(1024) k_rs_scan_rows(u32* __restrict__ hist, u32 ctas, u32* __restrict__ row_tot) {
    __shared__ u32 wsum[32];
    u32* row = hist + (size_t)blockIdx.x * ctas;
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 carry = 0;
    for (u32 base = 0; base < ctas; base += 1024) {
        u32 i = base + threadIdx.x;
        u32 v = i < ctas ? row[i] : 0u;
        u32 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        u32 wp = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            u32 x = wsum[k];
            if (k < (int)w) wp += x;
            tot += x;
        }
        if (i < ctas) row[i] = carry + wp + inc - v;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_tot[blockIdx.x] = carry;
}

template <typename KeyT>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const KeyT* __restrict__ kin, const u32* __restrict__ vin, KeyT* __restrict__ kout,
             u32* __restrict__ vout, u32 m, int shift, u32 mask, u32 tiles_per_cta,
             const u32* __restrict__ hist, const u32* __restrict__ row_tot) {
    constexpr int ITEMS = RsCfg<KeyT>::ITEMS;
    constexpr u32 TS = RS_THREADS * ITEMS;
    __shared__ u32 warp_cnt[RS_WARPS][RS_BINS + 1];  // bin 256 collects out-of-range lanes
    __shared__ u32 digit_base[RS_BINS];
    __shared__ u32 tile_excl[RS_BINS];
    __shared__ u32 tile_tot[RS_BINS];
    __shared__ u32 ws[RS_WARPS];
    __shared__ KeyT skey[TS];
    __shared__ u32 sval[TS];

    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    {
        u32 dig_excl = block_excl_scan_256(row_tot[tid], ws);   // start of digit `tid` in the output
        digit_base[tid] = dig_excl + hist[tid * gridDim.x + blockIdx.x];
    }
    const u32 ntiles = (u32)(((u64)m + TS - 1) / TS);
    const u32 tile0 = blockIdx.x * tiles_per_cta;
    u32 tile1 = tile0 + tiles_per_cta;
    if (tile1 > ntiles) tile1 = ntiles;

    for (u32 tile = tile0; tile < tile1; ++tile) {
        const u64 base = (u64)tile * TS;
        const u32 valid = (u32)((m - base) < (u64)TS ? (m - base) : (u64)TS);
        KeyT key[ITEMS];
        u32 val[ITEMS];
        u32 rnk[ITEMS];
#pragma unroll
        for (int t = 0; t < ITEMS; ++t) {
            u32 idx = w * (32 * ITEMS) + t * 32 + lane;
            if (idx < valid) {
                key[t] = kin[base + idx];
                val[t] = vin[base + idx];
            } else {
                key[t] = 0;
                val[t] = 0;
            }
        }
        for (int i = tid; i < RS_WARPS * (RS_BINS + 1); i += RS_THREADS) (&warp_cnt[0][0])[i] = 0;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < ITEMS; ++t) {
            u32 idx = w * (32 * ITEMS) + t * 32 + lane;
            u32 d = idx < valid ? ((u32)(key[t] >> shift) & mask) : (u32)RS_BINS;
            u32 mm = __match_any_sync(0xffffffffu, d);
            u32 leader = __ffs(mm) - 1;
            u32 pre = __popc(mm & lanemask_lt());
            u32 old = 0;
            if (lane == leader) {
                old = warp_cnt[w][d];
                warp_cnt[w][d] = old + __popc(mm);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rnk[t] = old + pre;
            __syncwarp();
        }
        __syncthreads();
        {
            u32 run = 0;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ++ww) {
                u32 c = warp_cnt[ww][tid];
                warp_cnt[ww][tid] = run;
                run += c;
            }
            tile_tot[tid] = run;
            u32 ex = block_excl_scan_256(run, ws);
            tile_excl[tid] = ex;
        }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < ITEMS; ++t) {
            u32 idx = w * (32 * ITEMS) + t * 32 + lane;
            if (idx < valid) {
                u32 d = (u32)(key[t] >> shift) & mask;
                u32 p = tile_excl[d] + warp_cnt[w][d] + rnk[t];
                skey[p] = key[t];
                sval[p] = val[t];
            }
        }
        __syncthreads();
        for (u32 q = tid; q < valid; q += RS_THREADS) {
            KeyT k = skey[q];
            u32 d = (u32)(k >> shift) & mask;
            u32 dst = digit_base[d] + (q - tile_excl[d]);
            kout[dst] = k;
            vout[dst] = sval[q];
        }
        __syncthreads();
        digit_base[tid] += tile_tot[tid];
    }
}

// Sorts m pairs held in (k[0], v[0]); buffers (k[1], v[1]) are scratch.  Returns through *res the
// index of the buffer pair that holds the sorted output.  `d_hist` needs 256*RS_MAX_CTAS + 256 words.
template <typename KeyT>
int radix_sort_pairs(KeyT* const k[2], u32* const v[2], u32 m, const DigitPlan& plan, u32* d_hist,
                     cudaStream_t st, int* res, Profiler& P) {
    int cur = 0;
    if (m > 1) {
        constexpr u32 TS = RS_THREADS * RsCfg<KeyT>::ITEMS;
        u32 nt = ceil_div_u32(m, TS);
        u32 ctas = nt < (u32)RS_MAX_CTAS ? nt : (u32)RS_MAX_CTAS;
        u32 tpc = ceil_div_u32(nt, ctas);
        ctas = ceil_div_u32(nt, tpc);
        const u64 kb = sizeof(KeyT);
        for (int p = 0; p < plan.npass; ++p) {
            u32 mask = (1u << plan.bits[p]) - 1u;
            KL(P, KC_RS_HIST, (u64)m * kb, st,
               (k_rs_hist<KeyT><<<ctas, RS_THREADS, 0, st>>>(k[cur], m, plan.shift[p], mask, tpc, d_hist)));
            KL(P, KC_RS_SCAN, (u64)RS_BINS * ctas * 8, st,
               (k_rs_scan_rows<<<RS_BINS, 1024, 0, st>>>(d_hist, ctas, d_hist + (size_t)RS_BINS * RS_MAX_CTAS)));
            KL(P, KC_RS_SCATTER, (u64)m * 2 * (kb + 4), st,
               (k_rs_scatter<KeyT><<<ctas, RS_THREADS, 0, st>>>(k[cur], v[cur], k[cur ^ 1], v[cur ^ 1], m,
                                                               plan.shift[p], mask, tpc, d_hist,
                                                               d_hist + (size_t)RS_BINS * RS_MAX_CTAS)));
            cur ^= 1;
        }
        NLZ_CK(cudaGetLastError());
    }
    *res = cur;
    return OK;
}

}  // namespace nlz
