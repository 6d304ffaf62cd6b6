// C ABI + pipeline orchestration of the B200 noLZSS factorizer (see include/nolzss_b200.h).
//
// Pipeline (all on one CUDA stream; host reads back only the byte histogram, the per-round active
// count and the final factor count):
//   S0 prepare   text -> X (raw bytes in HBM, padded); DNA_RC mode builds S = T s0 rc(T) s1 on device
//   S1 SA        keys -> stable LSD radix sort -> prefix doubling with discarding       (sa.cuh)
//   S2 LCP       chunked Kasai on raw bytes, scattered to rank order                   (lcp.cuh)
//   S3 LPnF      32-ary summary trees + per-rank LCP-interval climb -> LR[i]           (lpnf.cuh)
//   S4 chain     chunk exits -> exit-node doubling -> marks -> scan -> triples         (chain.cuh)
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nolzss_b200.h"
#include "chain.cuh"
#include "common.cuh"
#include "lcp.cuh"
#include "lpnf.cuh"
#include "prof.cuh"
#include "radix_sort.cuh"
#include "sa.cuh"
#include "tile_sort.cuh"
#include "big_groups.cuh"

namespace nlz {

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

enum Ev { EV_BEGIN = 0, EV_PREP, EV_KEYS, EV_SORT0, EV_DOUBLING, EV_LCP, EV_LPNF, EV_CHAIN, EV_COUNT };

struct Arena {
    u8* base = nullptr;
    size_t cap = 0, off = 0;
    template <typename T> T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        T* p = reinterpret_cast<T*>(base + off);
        off += bytes;
        return p;
    }
};

struct Workspace {
    u32 n1 = 0;
    u8* X = nullptr;
    u32 *SA = nullptr, *RANK = nullptr, *LCP = nullptr;
    u32* R0 = nullptr;       // RC mode, stage 3: rc-class leaf values (lpnf.cuh: k_leaf_values); F0 overwrites SA in place
    u8* NEED = nullptr;      // by text position: the LCP of this suffix is left to the Kasai kernel (lcp.cuh)
    u64* KEY[2] = {nullptr, nullptr};
    u32* VAL[2] = {nullptr, nullptr};
    u32* SLOT[2] = {nullptr, nullptr};
    u32 *HIST = nullptr, *PMAX = nullptr, *PSUM = nullptr, *CTR = nullptr, *BYTEHIST = nullptr;
    u32* tl[TREE_MAX_LEVELS] = {};
    u32* tf[TREE_MAX_LEVELS] = {};
    u32* tr[TREE_MAX_LEVELS] = {};
    uint4* NODE = nullptr;   // per-node table of stage 3: {parent, min forward start, depth, -}
    // batch mode only: record id per text position, record geometry, factor index of every record's sentinel
    u32 *REC = nullptr, *FSTART = nullptr, *FLEN = nullptr, *INOFF = nullptr, *SENTIDX = nullptr;
    u32* DCNT = nullptr;     // distributed runs: per-CTA counts / offsets of the key compaction
    u32* RING = nullptr;     // pipelined doubling rounds: (list length, largest group) entering every round
};
constexpr int RING_ROUNDS = 48;

}  // namespace nlz

using namespace nlz;

struct nlz_dist;

struct nlz_ctx {
    int device = 0;
    std::mutex mu;
    Arena arena;
    Workspace ws;
    u64* d_out = nullptr;       // device triples for the host-buffer entry points
    size_t d_out_cap = 0;       // in factors
    u32* h_pinned = nullptr;    // 4 KB pinned readback area
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[EV_COUNT];
    cudaEvent_t ring_ev[48];    // pipelined doubling rounds: count readback of every round
    nlz_stats stats;
    Profiler prof;
    Trees trees;                // summary trees of the last stage_lpnf call
    u64 text_suffixes = 0;      // suffixes of the WHOLE text of the current call (a distributed rank sees only its range)
    int debug_flags = 0;        // test hook: 1 force the bitonic tile path, 2 disable the pivot fast path, 4 force counting
};

namespace nlz {

static size_t workspace_bytes_for(u64 n1, u64 nrec) {
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t t = 0;
    t += al(n1 + 192) + al(n1 + 64); // X, NEED
    t += al((n1 + 72) * 4) * 4;      // SA, RANK, LCP, R0 (+ one padded line for whole-line reads)
    t += al((n1 + 72) * 16);         // NODE
    t += al(n1 * 8) * 2;             // KEY
    t += al(n1 * 4) * 4;             // VAL, SLOT
    t += al(((size_t)RS_BINS * RS_MAX_CTAS + RS_BINS) * 4);
    size_t tiles = (n1 + RG_TILE - 1) / RG_TILE + 1;
    t += al(tiles * 4) * 2;
    t += al(64 * 4) + al(256 * 4) + al(2 * RING_ROUNDS * 4);
    u64 c = n1 + 1;
    for (int lev = 1; lev < TREE_MAX_LEVELS; ++lev) {
        c = (c + 31) / 32;
        t += al((c + 72) * 4) * 3;
    }
    if (nrec) t += al((n1 + 72) * 4) + al((nrec + 1) * 4) * 4;   // REC, FSTART, FLEN, INOFF, SENTIDX
    return t + 4096;
}

static int ensure_workspace(nlz_ctx* c, u64 n1, u64 nrec = 0) {
    size_t need = workspace_bytes_for(n1, nrec);
    if (need > c->arena.cap) {
        if (c->arena.base) {
            NLZ_CK(cudaDeviceSynchronize());
            NLZ_CK(cudaFree(c->arena.base));
            c->arena.base = nullptr;
            c->arena.cap = 0;
        }
        size_t want = need + need / 8;
        cudaError_t e = cudaMalloc(&c->arena.base, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = need;
            e = cudaMalloc(&c->arena.base, want);
        }
        if (e != cudaSuccess) {
            set_error("cudaMalloc of %zu workspace bytes failed: %s", want, cudaGetErrorString(e));
            return ERR_CUDA;
        }
        c->arena.cap = want;
    }
    Arena& a = c->arena;
    Workspace& w = c->ws;
    a.off = 0;
    w.n1 = (u32)n1;
    w.X = a.take<u8>(n1 + 192);
    w.NEED = a.take<u8>(n1 + 64);
    w.SA = a.take<u32>(n1 + 72);
    w.RANK = a.take<u32>(n1 + 72);
    w.LCP = a.take<u32>(n1 + 72);
    w.R0 = a.take<u32>(n1 + 72);
    w.NODE = a.take<uint4>(n1 + 72);
    for (int i = 0; i < 2; ++i) w.KEY[i] = a.take<u64>(n1);
    for (int i = 0; i < 2; ++i) w.VAL[i] = a.take<u32>(n1);
    for (int i = 0; i < 2; ++i) w.SLOT[i] = a.take<u32>(n1);
    w.HIST = a.take<u32>((size_t)RS_BINS * RS_MAX_CTAS + RS_BINS);
    size_t tiles = (n1 + RG_TILE - 1) / RG_TILE + 1;
    w.PMAX = a.take<u32>(tiles);
    w.PSUM = a.take<u32>(tiles);
    w.CTR = a.take<u32>(64);
    w.BYTEHIST = a.take<u32>(256);
    w.RING = a.take<u32>(2 * RING_ROUNDS);
    u64 cnt = n1 + 1;
    for (int lev = 1; lev < TREE_MAX_LEVELS; ++lev) {
        cnt = (cnt + 31) / 32;
        w.tl[lev] = a.take<u32>(cnt + 72);
        w.tf[lev] = a.take<u32>(cnt + 72);
        w.tr[lev] = a.take<u32>(cnt + 72);
    }
    if (nrec) {
        w.REC = a.take<u32>(n1 + 72);
        w.FSTART = a.take<u32>(nrec + 1);
        w.FLEN = a.take<u32>(nrec + 1);
        w.INOFF = a.take<u32>(nrec + 1);
        w.SENTIDX = a.take<u32>(nrec + 1);
    }
    c->stats.workspace_bytes = c->arena.cap;
    return OK;
}

// ---------------------------------------------------------------- S0 kernels
__global__ void __launch_bounds__(256)
k_prepare_dna_rc(const u8* __restrict__ T, u32 n, u8* __restrict__ S, u32* __restrict__ first_bad) {
    for (u32 i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        u8 c = T[i];
        u8 u = c, k = 0;
        switch (c) {
            case 'A': case 'a': u = 'A'; k = 'T'; break;
            case 'C': case 'c': u = 'C'; k = 'G'; break;
            case 'G': case 'g': u = 'G'; k = 'C'; break;
            case 'T': case 't': u = 'T'; k = 'A'; break;
            default: atomicMin(first_bad, i); k = c; break;
        }
        S[i] = u;
        S[2 * (u64)n - i] = k;   // rc block starts at n+1; base i maps to n+1+(n-1-i)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S[n] = 1;                 // sentinel 0: factorizer.cpp:110-125 (first byte of 1..255 not in ACGT)
        S[2 * (u64)n + 1] = 2;    // sentinel 1
    }
}

// Distributed runs: every GPU uploads and prepares only its slice [lo, hi) of T (staged in place in its own X) and
// stores both strands into the X of EVERY GPU (coalesced peer stores), instead of G uploads of the whole text.
struct XDst {
    u8* p[MAX_PEERS];
    int n;
};
__global__ void __launch_bounds__(256)
k_prepare_dna_rc_slice(const u8* T, u32 n, u32 lo, u32 hi, XDst X, bool write_sentinels, u32* __restrict__ first_bad) {
    for (u32 i = lo + blockIdx.x * 256 + threadIdx.x; i < hi; i += gridDim.x * 256) {
        u8 c = T[i];
        u8 u = c, k = 0;
        switch (c) {
            case 'A': case 'a': u = 'A'; k = 'T'; break;
            case 'C': case 'c': u = 'C'; k = 'G'; break;
            case 'G': case 'g': u = 'G'; k = 'C'; break;
            case 'T': case 't': u = 'T'; k = 'A'; break;
            default: atomicMin(first_bad, i); k = c; break;
        }
        for (int g = 0; g < X.n; ++g) { X.p[g][i] = u; X.p[g][2 * (u64)n - i] = k; }
    }
    if (write_sentinels && blockIdx.x == 0 && threadIdx.x == 0)
        for (int g = 0; g < X.n; ++g) { X.p[g][n] = 1; X.p[g][2 * (u64)n + 1] = 2; }
}

__global__ void k_zero_pad(u8* __restrict__ X, u64 L) {
    if (threadIdx.x < 128) X[L + threadIdx.x] = 0;
}

__global__ void k_set_u32(u32* p, u32 v) { *p = v; }

// Batch mode (per-sequence FASTA entry points, fasta_processor.cpp:428-561): k independent records in
// ONE indexed text.  Forward half  T1 s T2 s .. Tk s  at [0, N]; with RC the rc half
// rc(Tk) s .. rc(T1) s  at [N+1, 2N+1] -- the record order of prepare_multiple_dna_sequences_w_rc
// (factorizer.cpp:150-169), so the mirror e = 2N - p of the single-text path holds for every record.
// All sentinels carry BATCH_SENT; REC[p] (record id, the leading sort-key field) keeps the records apart.
constexpr u8 BATCH_SENT = 1;
__global__ void __launch_bounds__(256)
k_prepare_batch(const u8* __restrict__ T, const u32* __restrict__ inoff, const u32* __restrict__ fstart,
                const u32* __restrict__ flen, u32 k, u32 N, bool rc, u8* __restrict__ X, u32* __restrict__ REC,
                u32* __restrict__ first_bad) {
    const u32 p = blockIdx.x * 256 + threadIdx.x;
    if (p > N) return;
    u32 lo = 0, hi = k - 1;                      // largest b with fstart[b] <= p
    while (lo < hi) {
        const u32 mid = (lo + hi + 1) >> 1;
        if (fstart[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const u32 b = lo, fs = fstart[b], off = p - fs;
    REC[p] = b;
    if (off == flen[b]) { X[p] = BATCH_SENT; }
    else {
        const u8 c = T[(u64)inoff[b] + off];
        u8 u = c, m = BATCH_SENT;
        switch (c) {
            case 'A': case 'a': u = 'A'; m = 'T'; break;
            case 'C': case 'c': u = 'C'; m = 'G'; break;
            case 'G': case 'g': u = 'G'; m = 'C'; break;
            case 'T': case 't': u = 'T'; m = 'A'; break;
            default: atomicMin(first_bad, p); break;
        }
        X[p] = u;
        if (rc) {
            X[2 * (u64)N - p] = m;
            REC[2 * (u64)N - p] = b;
            if (off == 0) {                       // sentinel that ends rc(T_b): mirror of the sentinel before T_b
                X[2 * (u64)N - fs + 1] = BATCH_SENT;
                REC[2 * (u64)N - fs + 1] = b;
            }
        }
    }
    if (p == 0) REC[rc ? 2 * (u64)N + 2 : (u64)N + 1] = k;   // terminator sorts after every record
}

// ---------------------------------------------------------------- pipeline
struct Problem {
    int mode;
    u64 n_in;       // bytes handed in
    u64 L;          // indexed text length
    u32 n1;         // suffixes
    u32 nfac;       // factorized positions [0, nfac)
    u32 N;          // RC: |S|/2-1
    u64 start_pos;
    bool rc;
    // batch mode (records are independent; see k_prepare_batch)
    u32 nrec = 0;
    const u32 *h_inoff = nullptr, *h_fstart = nullptr, *h_flen = nullptr;   // h_inoff: offsets into the PACKED staging copy
    const u64* h_srcoff = nullptr;      // offset of every record in the caller's buffer (error messages)
    // the caller's buffer may be sparse (records picked from a larger buffer): only the contiguous runs that hold
    // records are copied, back to back, into the staging area
    const u64 *h_run_src = nullptr, *h_run_len = nullptr;
    u32 nruns = 0;
    u64 packed_bytes = 0;
};

static int bits_for(u32 maxval) {
    int nb = 1;
    while (nb < 32 && (maxval >> nb) != 0) ++nb;
    return nb;
}

// Symbols per 64-bit key: enough that a random text keeps its expected fraction of colliding windows below ~1.5 %
// (sigma^W >= 64 n'), rounded up to fill the 8-bit radix passes that many symbols need anyway, at most 29.  Fewer symbols
// = fewer passes of the initial sort; suffixes that tie on the window go through prefix doubling either way (they are
// the repeats), which then starts from h = W.
static int layout_symbols(int sigma, int b, int R, int D, u64 n1) {
    const int wmax = std::min(29, (64 - R - D) / b);
    double space = 1.0;
    int wmin = 0;
    while (wmin < wmax && space < 64.0 * (double)n1) { space *= (double)(sigma > 1 ? sigma : 2); ++wmin; }
    if (wmin < 1) wmin = 1;
    const int passes = (R + wmin * b + D + 7) / 8;
    int W = (passes * 8 - R - D) / b;
    if (W > wmax) W = wmax;
    if (W < wmin) W = wmin;
    return W;
}

static int choose_layout(const u32 hist[256], u64 n1, ClassTable& tab, KeyLayout& lay) {
    int sigma = 0;
    for (int c = 0; c < 256; ++c) {
        if (hist[c] >= 2) tab.cls[c] = (u8)sigma++;
        else tab.cls[c] = (u8)SENT_CLASS;   // unique or absent byte: sentinel class
    }
    int b = 1;
    while ((1 << b) < sigma) ++b;
    // 32-bit keys while the expected number of random collisions stays small, else 64-bit keys
    int w32 = 28 / b;
    if (w32 > 14) w32 = 14;
    bool use32 = false;
    if (w32 >= 1) {
        double space = 1.0;
        for (int i = 0; i < w32; ++i) space *= (double)(sigma > 1 ? sigma : 2);
        use32 = space >= 16.0 * (double)n1;
    }
    if (use32) {
        lay.key_bits = 32; lay.b = b; lay.W = w32; lay.D = 4; lay.R = 0;
    } else {
        lay.key_bits = 64; lay.b = b; lay.D = 5; lay.R = 0;
        lay.W = layout_symbols(sigma, b, 0, 5, n1);
    }
    return OK;
}

// k_group_stream in two sizes (big_groups.cuh): SMALL outlier buffer for every chunk (two CTAs per SM), then persistent
// CTAs with the BIG buffer over the groups that overflowed.  The overflow list lives in the radix-sort histogram table.
template <int GS>
static void launch_group_stream(nlz_ctx* c, cudaStream_t st, const u64* key, const u32* val, const u32* slot, u32 b0, u32 mB,
                                u32 gcap, u32* SA, const RankDst& rdst, StreamOut so, int dbg) {
    Workspace& w = c->ws;
    const u32 chunks = ceil_div_u32(mB, gcap);
    static const bool one_size = getenv("NLZ_GS_ONE_SIZE") != nullptr;
    so.ovf_out = nullptr; so.ovf_cnt_out = nullptr; so.ovf_in = nullptr; so.ovf_cnt_in = nullptr;
    if (chunks <= (u32)RS_BINS * RS_MAX_CTAS && !one_size && !(dbg & 8)) {
        u32* ovf_cnt = w.CTR + 32;
        cudaMemsetAsync(ovf_cnt, 0, 4, st);
        so.ovf_out = w.HIST; so.ovf_cnt_out = ovf_cnt;
        k_group_stream<GS, GS_CAP_SMALL><<<chunks, GS_THREADS, gs_smem(GS_CAP_SMALL), st>>>(key, val, slot, b0, mB, gcap, SA, rdst, so, dbg);
        so.ovf_out = nullptr; so.ovf_cnt_out = nullptr; so.ovf_in = w.HIST; so.ovf_cnt_in = ovf_cnt;
        k_group_stream<GS, GS_CAP_BIG><<<chunks < (u32)kNumSM ? chunks : (u32)kNumSM, GS_THREADS, gs_smem(GS_CAP_BIG), st>>>(key, val, slot, b0, mB, gcap, SA, rdst, so, dbg);
    } else {
        k_group_stream<GS, GS_CAP_BIG><<<chunks, GS_THREADS, gs_smem(GS_CAP_BIG), st>>>(key, val, slot, b0, mB, gcap, SA, rdst, so, dbg);
    }
}

static RankDst local_rank_dst(nlz_ctx* c) {   // one GPU: refined ranks go straight into RANK; no records
    RankDst r;
    r.rank = c->ws.RANK; r.upd = nullptr; r.upd_count = c->ws.CTR + 4; r.base = 0;
    return r;
}

template <typename KeyT>
static int initial_sort_and_regroup(nlz_ctx* c, const Problem& pb, const ClassTable& tab, const KeyLayout& lay,
                                    cudaStream_t st, u32 cnt, int* cur_out, u32* m_out, u32* maxg_out) {
    Workspace& w = c->ws;
    Profiler& P = c->prof;
    const u32 n1 = pb.n1;
    const u64 kb = sizeof(KeyT);
    KeyT* k[2] = {reinterpret_cast<KeyT*>(w.KEY[0]), reinterpret_cast<KeyT*>(w.KEY[1])};
    u32* v[2] = {w.VAL[0], w.VAL[1]};
    KL(P, KC_KEYS, (u64)n1 * (1 + kb + 4), st,
       (k_build_keys<KeyT><<<ceil_div_u32(n1, KB_TP), 256, 0, st>>>(w.X, pb.L, n1, tab, lay, pb.nrec ? w.REC : nullptr,
                                                                   k[0], v[0])));
    NLZ_CK(cudaEventRecord(c->ev[EV_KEYS], st));
    *cur_out = 1; *m_out = 0; *maxg_out = 0;
    if (cnt == 0) { NLZ_CK(cudaEventRecord(c->ev[EV_SORT0], st)); return OK; }
    DigitPlan plan;
    plan_add_range(plan, lay.dshift(), lay.key_bits);   // sentinel-offset field, symbols (, record id): one bit range
    int res = 0;
    NLZ_TRY(radix_sort_pairs<KeyT>(k, v, cnt, plan, w.HIST, st, &res, P));
    NLZ_CK(cudaEventRecord(c->ev[EV_SORT0], st));
    const KeyT dist_mask = (((KeyT)1 << lay.D) - 1) << lay.dshift();
    const u32 tiles = ceil_div_u32(cnt, RG_TILE);
    // compaction target must not alias the sorted buffers: use the other KEY/VAL pair.  The first regroup also seeds
    // the LCP array from adjacent key pairs and marks the suffixes whose LCP needs the text (lcp.cuh).
    LcpSeed seed;
    seed.LCP = w.LCP; seed.NEED = w.NEED; seed.lay = lay; seed.first_pending = false; seed.RANKOUT = nullptr;
    seed.need_in_rankout = false;
    RankDst rdst0 = local_rank_dst(c);
    // RANK beyond the L2 cache: partitioned scatter of the inverse suffix array (see LcpSeed::RANKOUT); the node table
    // (dead until stage 3) lends the three n'-word buffers
    static const bool no_part = getenv("NLZ_NO_PARTITIONED_ISA") != nullptr;
    const bool part_isa = cnt > (48u << 20) && !no_part;
    u32* NR = reinterpret_cast<u32*>(w.NODE);
    if (part_isa) { seed.RANKOUT = NR; rdst0.rank = nullptr; seed.need_in_rankout = n1 < 0x80000000u; }
    NLZ_CK(cudaMemsetAsync(w.NEED, 0, (size_t)n1 + 64, st));
    NLZ_CK(cudaMemsetAsync(w.LCP + n1, 0, 4, st));              // right guard used by the interval walks
    P.begin(st);
    k_regroup_reduce<KeyT, true><<<tiles, RG_THREADS, 0, st>>>(k[res], cnt, dist_mask, w.PMAX, w.PSUM);
    k_regroup_scan_partials<<<1, 1024, 0, st>>>(w.PMAX, w.PSUM, tiles, w.CTR);
    k_regroup_apply<KeyT, true><<<tiles, RG_THREADS, 0, st>>>(k[res], v[res], nullptr, cnt, dist_mask, w.PMAX,
                                                              w.PSUM, w.SA, rdst0, w.KEY[res ^ 1],
                                                              w.VAL[res ^ 1], w.SLOT[0], w.CTR + 3, seed);
    P.end(KC_REGROUP, (u64)cnt * (2 * kb + 4 + 8 + 4), st, 3);
    if (part_isa) {
        u32* pk[2] = {w.SA, NR + (size_t)cnt};                 // one pass: the input pair (SA, new ranks) is only read
        u32* pv[2] = {NR, NR + 2 * (size_t)cnt};
        DigitPlan pp;
        const int nbp = bits_for(n1 - 1);
        plan_add_range(pp, nbp > 8 ? nbp - 8 : 0, nbp);
        int pres = 0;
        NLZ_TRY(radix_sort_pairs<u32>(pk, pv, cnt, pp, w.HIST, st, &pres, P));
        u32 grid = ceil_div_u32(cnt, 256 * 8);
        if (seed.need_in_rankout) {
            KL(P, KC_REGROUP, (u64)cnt * 12, st, (k_scatter_pairs<true><<<grid, 256, 0, st>>>(pk[pres], pv[pres], cnt, w.RANK, w.NEED)));
        } else {
            KL(P, KC_REGROUP, (u64)cnt * 12, st, (k_scatter_pairs<false><<<grid, 256, 0, st>>>(pk[pres], pv[pres], cnt, w.RANK, nullptr)));
        }
    }
    *cur_out = res ^ 1;
    NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 16, cudaMemcpyDeviceToHost, st));
    NLZ_CK(cudaStreamSynchronize(st));
    c->stats.host_syncs += 1;
    *m_out = c->h_pinned[0];
    *maxg_out = c->h_pinned[3];
    return OK;
}

// ---- S1: suffix array of all n1 suffixes on one GPU (the distributed loop is stage_sa_dist in dist2_host.cuh).
// Results: w.SA (rank order), RANK (= ISA).
static int stage_sa(nlz_ctx* c, const Problem& pb, const ClassTable& tab, const KeyLayout& lay, cudaStream_t st) {
    const u32 cnt = pb.n1;
    Workspace& w = c->ws;
    nlz_stats& S = c->stats;
    Profiler& P = c->prof;
    const u32 n1 = pb.n1;
    S.key_bits = lay.key_bits; S.sym_bits = lay.b; S.key_syms = lay.W;
    int cur = 0;
    u32 m = 0, maxg = 0;
    if (lay.key_bits == 32) NLZ_TRY(initial_sort_and_regroup<u32>(c, pb, tab, lay, st, cnt, &cur, &m, &maxg));
    else NLZ_TRY(initial_sort_and_regroup<u64>(c, pb, tab, lay, st, cnt, &cur, &m, &maxg));
    const RankDst rdst = local_rank_dst(c);
    S.lcp_marked = m;
    const int nb = bits_for(n1 - 1);
    DigitPlan plan;
    plan_add_range(plan, 0, nb);
    plan_add_range(plan, 32, 32 + nb);
    u64 h = (u64)lay.W;
    int sc = 0;
    u32 gm = m;
    static const bool no_pipeline = getenv("NLZ_TRACE") != nullptr || getenv("NLZ_NO_PIPELINE") != nullptr;
    // largest tie group the shared-memory tile sort takes (test hook: debug flags >> 8 lower it so that small texts
    // reach the hybrid rounds)
    u32 gcap = (u32)TSORT_SLOTS / 2;
    if (((c->debug_flags >> 8) & 0xFFFF) >= 64 && (u32)((c->debug_flags >> 8) & 0xFFFF) < gcap) gcap = (u32)((c->debug_flags >> 8) & 0xFFFF);
    if (!no_pipeline && m > 0 && maxg <= gcap) {
        // Every remaining round is a fused one (groups only shrink).  The host runs one round AHEAD of the device:
        // round r is launched with grids sized from the counts of round r-1 and reads its true list length from
        // RING[r] on the device, so the per-round count readback no longer leaves the GPU idle.
        u32* RING = w.RING;                              // RING[2r] = m_r, RING[2r+1] = largest group entering round r
        u32* hring = c->h_pinned + 512;
        NLZ_CK(cudaMemsetAsync(RING, 0, 2 * RING_ROUNDS * 4, st));
        hring[0] = m; hring[1] = maxg;
        NLZ_CK(cudaMemcpyAsync(RING, hring, 8, cudaMemcpyHostToDevice, st));
        u32 bm = m, bg = maxg;                           // upper bounds for the round being launched
        for (int r = 0;; ++r) {
            if (r + 1 >= RING_ROUNDS) { set_error("prefix doubling did not converge"); return ERR_RUNTIME; }
            u32 cap = 32;
            while (cap < bg) cap <<= 1;
            const u32 tile = TSORT_SLOTS - cap;
            KL(P, KC_GATHER, (u64)bm * 24, st,
               (k_gather_rank<<<ceil_div_u32(bm, 256), 256, 0, st>>>(w.KEY[cur], w.VAL[cur], bm, RING + 2 * r, w.RANK, h, n1, nullptr)));
            KL(P, KC_TILE_SORT, (u64)bm * (12 + 4 + 8 + 16), st,
               (k_tile_sort<32><<<ceil_div_u32(bm, tile), TSORT_THREADS, TSORT_SMEM, st>>>(
                   w.KEY[cur], w.VAL[cur], w.SLOT[sc], bm, RING + 2 * r, tile, cap, w.SA, rdst, w.KEY[cur ^ 1], w.VAL[cur ^ 1],
                   w.SLOT[sc ^ 1], RING + 2 * (r + 1), RING + 2 * (r + 1) + 1, c->debug_flags)));
            NLZ_CK(cudaMemcpyAsync(hring + 2 * (r + 1), RING + 2 * (r + 1), 8, cudaMemcpyDeviceToHost, st));
            NLZ_CK(cudaEventRecord(c->ring_ev[r + 1], st));
            cur ^= 1; sc ^= 1; h *= 2;
            if (r == 0) { S.doubling_rounds += 1; S.tile_sort_rounds += 1; S.active_sum += m; continue; }
            // the counts entering round r (produced by round r-1, which ran while round r was being launched)
            NLZ_CK(cudaEventSynchronize(c->ring_ev[r]));
            S.host_syncs += 1;
            const u32 mr = hring[2 * r], gr = hring[2 * r + 1];
            if (mr == 0) {                               // round r found nothing to do: r rounds did the work
                P.bytes[KC_GATHER] -= (u64)bm * 24; P.bytes[KC_TILE_SORT] -= (u64)bm * 40;
                break;
            }
            // algorithmic bytes were booked with the bound; correct them to the true list length
            P.bytes[KC_GATHER] -= (u64)(bm - mr) * 24; P.bytes[KC_TILE_SORT] -= (u64)(bm - mr) * 40;
            S.doubling_rounds += 1; S.tile_sort_rounds += 1; S.active_sum += mr;
            bm = mr; bg = gr;
        }
        return OK;
    }
    // Hybrid rounds (big_groups.cuh): once a tie group exceeds the tile-sort capacity the list is split into
    // S (groups <= gcap, left-aligned, k_tile_sort) and B (larger groups, right-aligned, k_group_stream).
    struct { bool on = false, off = false; u32 mS = 0, mB = 0, maxgS = 0; int fallbacks = 0; } hy;
    static const bool no_hybrid = getenv("NLZ_NO_HYBRID") != nullptr;
    hy.off = no_hybrid;
    const u32 END = cnt;                                 // capacity of the active lists (B grows down from here)
    static const bool trace = getenv("NLZ_TRACE") != nullptr;
    // radix sort + regroup of the unified list [0, m) held in the `cur` buffers (keys already gathered)
    auto radix_round = [&](u32 mm, int* rb_out, cudaEvent_t tev1) -> int {
        u64* k[2] = {w.KEY[cur], w.KEY[cur ^ 1]};
        u32* v[2] = {w.VAL[cur], w.VAL[cur ^ 1]};
        int res = 0;
        NLZ_TRY(radix_sort_pairs<u64>(k, v, mm, plan, w.HIST, st, &res, P));
        const int rb = res == 0 ? cur : (cur ^ 1);
        if (tev1) cudaEventRecord(tev1, st);
        const u32 tiles = ceil_div_u32(mm, RG_TILE);
        P.begin(st);
        k_regroup_reduce<u64, false><<<tiles, RG_THREADS, 0, st>>>(w.KEY[rb], mm, 0ull, w.PMAX, w.PSUM);
        k_regroup_scan_partials<<<1, 1024, 0, st>>>(w.PMAX, w.PSUM, tiles, w.CTR);
        k_regroup_apply<u64, false, 32><<<tiles, RG_THREADS, 0, st>>>(w.KEY[rb], w.VAL[rb], w.SLOT[sc], mm, 0ull,
                                                                w.PMAX, w.PSUM, w.SA, rdst, w.KEY[rb ^ 1],
                                                                w.VAL[rb ^ 1], w.SLOT[sc ^ 1], w.CTR + 3);
        P.end(KC_REGROUP, (u64)mm * (16 + 4 + 4 + 8 + 16), st, 3);
        *rb_out = rb;
        return OK;
    };
    for (;;) {
        if (gm == 0) break;
        S.doubling_rounds += 1;
        S.active_sum += m;
        cudaEvent_t tev0 = nullptr, tev1 = nullptr;
        if (trace) { cudaEventCreate(&tev0); cudaEventCreate(&tev1); cudaEventRecord(tev0, st); }
        if (!hy.on && !hy.off && m > 0 && maxg > gcap) {
            // split the list once: small groups to the left, big groups to the right end of the other buffers
            const u32 tiles = ceil_div_u32(m, RG_TILE);
            P.begin(st);
            k_split_count<32><<<tiles, RG_THREADS, 0, st>>>(w.KEY[cur], w.SLOT[sc], m, gcap, w.PSUM);
            k_scan_u32_single_cta<<<1, 1024, 0, st>>>(w.PSUM, tiles, w.CTR + 6);
            k_split_apply<32><<<tiles, RG_THREADS, 0, st>>>(w.KEY[cur], w.VAL[cur], w.SLOT[sc], m, gcap, w.PSUM, w.CTR + 6, END,
                                                        w.KEY[cur ^ 1], w.VAL[cur ^ 1], w.SLOT[sc ^ 1]);
            P.end(KC_REGROUP, (u64)m * 40, st, 3);
            NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 32, cudaMemcpyDeviceToHost, st));
            NLZ_CK(cudaStreamSynchronize(st));
            S.host_syncs += 1;
            hy.on = true;
            hy.mB = c->h_pinned[6]; hy.mS = m - hy.mB; hy.maxgS = gcap;
            cur ^= 1; sc ^= 1;
            if (trace) fprintf(stderr, "[nlz] hybrid rounds from here: S=%u B=%u (groups > %u)\n", hy.mS, hy.mB, gcap);
        }
        int rb = cur;   // physical index of the buffers that hold this round's sorted (key, suffix) pairs
        if (hy.on) {
            const u32 mS = hy.mS, mB = hy.mB, b0 = END - mB;
            NLZ_CK(cudaMemsetAsync(w.CTR, 0, 16, st));                  // [0] next S length, [3] its largest group
            NLZ_CK(cudaMemsetAsync(w.CTR + 6, 0, 8, st));               // [6] next B length, [7] fallback flag
            if (mS) KL(P, KC_GATHER, (u64)mS * 24, st,
                       (k_gather_rank<<<ceil_div_u32(mS, 256), 256, 0, st>>>(w.KEY[cur], w.VAL[cur], mS, nullptr, w.RANK, h, n1, nullptr)));
            if (mB) KL(P, KC_GATHER, (u64)mB * 24, st,
                       (k_gather_rank<<<ceil_div_u32(mB, 256), 256, 0, st>>>(w.KEY[cur] + b0, w.VAL[cur] + b0, mB, nullptr, w.RANK, h, n1, nullptr)));
            if (mS) {
                u32 cap = 32;
                while (cap < hy.maxgS) cap <<= 1;
                const u32 tile = TSORT_SLOTS - cap;
                KL(P, KC_TILE_SORT, (u64)mS * 40, st,
                   (k_tile_sort<32><<<ceil_div_u32(mS, tile), TSORT_THREADS, TSORT_SMEM, st>>>(
                       w.KEY[cur], w.VAL[cur], w.SLOT[sc], mS, nullptr, tile, cap, w.SA, rdst, w.KEY[cur ^ 1], w.VAL[cur ^ 1],
                       w.SLOT[sc ^ 1], w.CTR, w.CTR + 3, c->debug_flags)));
            }
            if (mB) {
                StreamOut so;
                so.key_next = w.KEY[cur ^ 1]; so.val_next = w.VAL[cur ^ 1]; so.slot_next = w.SLOT[sc ^ 1];
                so.end = END; so.mS = w.CTR; so.maxgS = w.CTR + 3; so.mB = w.CTR + 6; so.fallback = w.CTR + 7;
                const int dbg = (c->debug_flags & 8) && S.doubling_rounds >= 3 ? 8 : 0;   // test hook: fail the third round
                KL(P, KC_STREAM, (u64)mB * 44, st,
                   (launch_group_stream<32>(c, st, w.KEY[cur], w.VAL[cur], w.SLOT[sc], b0, mB, gcap, w.SA, rdst, so, dbg)));
            }
            S.tile_sort_rounds += 1;
            if (trace) cudaEventRecord(tev1, st);
            NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 32, cudaMemcpyDeviceToHost, st));
            NLZ_CK(cudaStreamSynchronize(st));
            S.host_syncs += 1;
            if (c->h_pinned[7]) {
                // a group had more outliers than fit in shared memory: unify the lists and redo the round with the
                // radix path (the keys are gathered; ranks and slots written so far are rewritten with equal values)
                if (mB) {
                    NLZ_CK(cudaMemcpyAsync(w.KEY[cur ^ 1], w.KEY[cur] + b0, (size_t)mB * 8, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.KEY[cur] + mS, w.KEY[cur ^ 1], (size_t)mB * 8, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.VAL[cur ^ 1], w.VAL[cur] + b0, (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.VAL[cur] + mS, w.VAL[cur ^ 1], (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.SLOT[sc ^ 1], w.SLOT[sc] + b0, (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.SLOT[sc] + mS, w.SLOT[sc ^ 1], (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                }
                {
                    // The hybrid lists hold their groups in no particular order, but the radix path pairs sorted list
                    // positions with slots BY POSITION, which needs ascending slots: sort the slot list (the groups'
                    // slot intervals are disjoint, so ascending slots line up with the groups in ascending rank order).
                    u32* k[2] = {w.SLOT[sc], w.SLOT[sc ^ 1]};
                    u32* v[2] = {reinterpret_cast<u32*>(w.KEY[cur ^ 1]), reinterpret_cast<u32*>(w.KEY[cur ^ 1]) + END};   // carried along, unused
                    DigitPlan p32;
                    plan_add_range(p32, 0, nb);
                    int res = 0;
                    NLZ_TRY(radix_sort_pairs<u32>(k, v, mS + mB, p32, w.HIST, st, &res, P));
                    if (res) NLZ_CK(cudaMemcpyAsync(w.SLOT[sc], w.SLOT[sc ^ 1], (size_t)(mS + mB) * 4, cudaMemcpyDeviceToDevice, st));
                }
                NLZ_CK(cudaMemsetAsync(w.CTR + 4, 0, 16, st));          // records pushed so far are dropped, B is gone
                hy.on = false; hy.off = ++hy.fallbacks >= 2;            // one more try from a fresh split after this round
                if (trace) fprintf(stderr, "[nlz] round %u: outliers exceed the stream kernel, back to radix rounds\n", S.doubling_rounds);
                NLZ_TRY(radix_round(mS + mB, &rb, nullptr));
                {
                    NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 16, cudaMemcpyDeviceToHost, st));
                    NLZ_CK(cudaStreamSynchronize(st));
                    S.host_syncs += 1;
                    m = c->h_pinned[0]; maxg = c->h_pinned[3]; gm = m;
                }
            } else {
                hy.mS = c->h_pinned[0]; hy.maxgS = c->h_pinned[3]; hy.mB = c->h_pinned[6];
                if (trace) {
                    float tms = 0.f;
                    cudaEventElapsedTime(&tms, tev0, tev1);
                    fprintf(stderr, "[nlz] round %u h=%llu S=%u B=%u sort_ms=%.3f -> S'=%u (maxg %u) B'=%u\n", S.doubling_rounds,
                            (unsigned long long)h, mS, mB, tms, hy.mS, hy.maxgS, hy.mB);
                }
                m = hy.mS + hy.mB; gm = m;
            }
        } else {
            const bool fused = maxg <= gcap;
            if (m > 0)
                KL(P, KC_GATHER, (u64)m * 24, st,
                   (k_gather_rank<<<ceil_div_u32(m, 256), 256, 0, st>>>(w.KEY[cur], w.VAL[cur], m, nullptr, w.RANK, h, n1,
                                                                         fused ? w.CTR : nullptr)));
            else NLZ_CK(cudaMemsetAsync(w.CTR, 0, 32, st));
            if (m > 0 && fused) {
                // every tie group fits in shared memory: segmented sort + regroup in one pass
                u32 cap = 32;
                while (cap < maxg) cap <<= 1;
                const u32 tile = TSORT_SLOTS - cap;
                KL(P, KC_TILE_SORT, (u64)m * (12 + 4 + 8 + 16), st,
                   (k_tile_sort<32><<<ceil_div_u32(m, tile), TSORT_THREADS, TSORT_SMEM, st>>>(
                       w.KEY[cur], w.VAL[cur], w.SLOT[sc], m, nullptr, tile, cap, w.SA, rdst, w.KEY[cur ^ 1], w.VAL[cur ^ 1],
                       w.SLOT[sc ^ 1], w.CTR, w.CTR + 3, c->debug_flags)));
                rb = cur;                       // next round's lists were written to the cur^1 buffers
                S.tile_sort_rounds += 1;
                if (trace) cudaEventRecord(tev1, st);
            } else if (m > 0) {
                NLZ_TRY(radix_round(m, &rb, tev1));
            }
            {
                NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 16, cudaMemcpyDeviceToHost, st));
                NLZ_CK(cudaStreamSynchronize(st));
                S.host_syncs += 1;
                if (trace) {
                    float tms = 0.f;
                    cudaEventElapsedTime(&tms, tev0, tev1);
                    fprintf(stderr, "[nlz] round %u h=%llu m=%u maxg=%u sort_ms=%.3f -> m'=%u maxg'=%u\n", S.doubling_rounds,
                            (unsigned long long)h, m, maxg, tms, c->h_pinned[0], c->h_pinned[3]);
                }
                m = c->h_pinned[0];
                maxg = c->h_pinned[3];
                gm = m;
            }
        }
        if (trace) { cudaEventDestroy(tev0); cudaEventDestroy(tev1); }
        cur = rb ^ 1;
        sc ^= 1;
        h *= 2;
        if (S.doubling_rounds > 40) { set_error("prefix doubling did not converge"); return ERR_RUNTIME; }
    }
    return OK;
}

// ---- S3: per-position factor rule over the rank-ordered arrays (leaf values F0 / R0, LCP) of wp.n1 ranks; in a
// distributed run these hold virtual ranks around the real ones [wp.real_lo, wp.real_hi) and the results are indexed
// by work item (`bylist`, see k_lpnf_rank) instead of by text position.
static int stage_lpnf(nlz_ctx* c, bool rc, cudaStream_t st, const u32* F0, const u32* R0, const u32* LCP, WalkParams wp,
                      const u32* RANK, u64* LR, u8* FLAGS, bool bylist) {
    Workspace& w = c->ws;
    Profiler& P = c->prof;
    const u32 n1 = wp.n1;
    Trees T;
    memset(&T, 0, sizeof(T));
    T.lcp[0] = LCP; T.cntL[0] = n1 + 1;
    T.f[0] = F0; T.r[0] = R0; T.cntS[0] = n1;
    int lev = 0;
    P.begin(st);
    while (T.cntL[lev] > 32 && lev + 1 < TREE_MAX_LEVELS) {
        u32 cl = (T.cntL[lev] + 31) / 32, cs = (T.cntS[lev] + 31) / 32;
        u32 nodes = cl > cs ? cl : cs;
        u32 grid = ceil_div_u32((u64)nodes * 32, 256);
        if (lev == 0) {
            if (rc) k_tree_level1<true><<<grid, 256, 0, st>>>(LCP, T.cntL[0], F0, R0, T.cntS[0], w.tl[1], cl, w.tf[1], w.tr[1], cs);
            else k_tree_level1<false><<<grid, 256, 0, st>>>(LCP, T.cntL[0], F0, R0, T.cntS[0], w.tl[1], cl, w.tf[1], w.tr[1], cs);
        } else {
            if (rc) k_tree_level_up<true><<<grid, 256, 0, st>>>(T.lcp[lev], T.cntL[lev], T.f[lev], T.r[lev], T.cntS[lev], w.tl[lev + 1], cl, w.tf[lev + 1], w.tr[lev + 1], cs);
            else k_tree_level_up<false><<<grid, 256, 0, st>>>(T.lcp[lev], T.cntL[lev], T.f[lev], T.r[lev], T.cntS[lev], w.tl[lev + 1], cl, w.tf[lev + 1], w.tr[lev + 1], cs);
        }
        ++lev;
        T.lcp[lev] = w.tl[lev]; T.f[lev] = w.tf[lev]; T.r[lev] = w.tr[lev];
        T.cntL[lev] = cl; T.cntS[lev] = cs;
    }
    T.nlev = lev + 1;
    P.end(KC_TREE, (u64)n1 * 8 + (u64)n1 / 2, st, (u32)lev);
    c->trees = T;
    if (!LR) return OK;                                  // trees only (edge staircases of a distributed run)
    KL(P, KC_NODES, (u64)n1 * (4 + 8 + 16), st,
       (rc ? k_node_tables<true, true><<<ceil_div_u32((u64)n1 + 1, 256), 256, 0, st>>>(T, wp, w.NODE)
           : k_node_tables<false, false><<<ceil_div_u32((u64)n1 + 1, 256), 256, 0, st>>>(T, wp, w.NODE)));
    unsigned long long* visit_ctr = reinterpret_cast<unsigned long long*>(w.CTR + 16);   // [0] probes, [1] hard
    NLZ_CK(cudaMemsetAsync(visit_ctr, 0, 16, st));
    // flag plane: zero = ordinary forward / literal factor; k_lpnf_rank stores only the hard and the RC marks
    NLZ_CK(cudaMemsetAsync(FLAGS, 0, bylist ? (size_t)(wp.real_hi - wp.real_lo) : (size_t)wp.nfac, st));
    // algorithmic bytes: the leaf value of every rank; per factorized position the two LCP neighbours, the
    // LR store and the flag byte; plus (added after the run, from the probe counter) 16 B per probe
    P.begin(st);
    static const int walk_env = getenv("NLZ_WALK_NODES") ? atoi(getenv("NLZ_WALK_NODES")) : 0;
    const int walk_nodes = walk_env ? walk_env : (c->text_suffixes > WALK_LARGE_TEXT ? WALK_MAX_NODES_LARGE : WALK_MAX_NODES);
    const u32 nreal = wp.real_hi - wp.real_lo;
    u32 nwork;                                           // work items of both kernels when the results are indexed by item
    u32* list = nullptr;
    if (rc) {
        // compact the ranks that hold a forward suffix (about half): no idle lanes in the walk
        list = w.SLOT[1];
        const u32 tiles = ceil_div_u32(n1, FR_TILE);
        k_forward_ranks<1><<<tiles, 256, 0, st>>>(F0, wp, w.PMAX, nullptr);
        k_scan_u32_single_cta<<<1, 1024, 0, st>>>(w.PMAX, tiles, w.CTR + 5);
        k_forward_ranks<2><<<tiles, 256, 0, st>>>(F0, wp, w.PMAX, list);
        nwork = nreal < wp.nfac ? nreal : wp.nfac;
        const u32 grid = ceil_div_u32(nwork ? nwork : 1, 256);
        if (bylist) k_lpnf_rank<true, true><<<grid, 256, 0, st>>>(T, wp, w.NODE, list, w.CTR + 5, walk_nodes, LR, FLAGS, visit_ctr);
        else k_lpnf_rank<true, false><<<grid, 256, 0, st>>>(T, wp, w.NODE, list, w.CTR + 5, walk_nodes, LR, FLAGS, visit_ctr);
    } else {
        nwork = nreal;
        const u32 items = bylist ? nreal : n1;
        const u32 grid = ceil_div_u32(items ? items : 1, 256);
        if (bylist) k_lpnf_rank<false, true><<<grid, 256, 0, st>>>(T, wp, w.NODE, nullptr, nullptr, walk_nodes, LR, FLAGS, visit_ctr);
        else k_lpnf_rank<false, false><<<grid, 256, 0, st>>>(T, wp, w.NODE, nullptr, nullptr, walk_nodes, LR, FLAGS, visit_ctr);
    }
    P.end(KC_WALK, (u64)n1 * 4 + (u64)nreal * 17, st, rc ? 4 : 1);
    {
        const u64 items = bylist ? (u64)nwork : (u64)wp.nfac;
        const u32 grid = ceil_div_u32((u64)ceil_div_u32(items ? items : 1, WALK_Q) * 8, 256);   // one 8-lane tile per run
        P.begin(st);
        if (bylist) {
            if (rc) k_lpnf_hard<true, true><<<grid, 256, 0, st>>>(T, wp, nullptr, list, w.CTR + 5, nwork, LR, FLAGS, visit_ctr);
            else k_lpnf_hard<false, true><<<grid, 256, 0, st>>>(T, wp, nullptr, nullptr, nullptr, nwork, LR, FLAGS, visit_ctr);
        } else {
            if (rc) k_lpnf_hard<true, false><<<grid, 256, 0, st>>>(T, wp, RANK, nullptr, nullptr, 0, LR, FLAGS, visit_ctr);
            else k_lpnf_hard<false, false><<<grid, 256, 0, st>>>(T, wp, RANK, nullptr, nullptr, 0, LR, FLAGS, visit_ctr);
        }
        P.end(KC_WALK_HARD, items, st);
    }
    NLZ_CK(cudaMemcpyAsync(c->h_pinned + 4, visit_ctr, 16, cudaMemcpyDeviceToHost, st));
    return OK;
}

// ---- S4: chain extraction over LR[0, nfac) and emission of the triples.  Scratch: EXIT, J2, alist
// (u32 x nfac), REACH (u8 x nfac), MASK (u32 x (33 x chunks)).
struct ChainScratch { u32 *EXIT, *J2, *alist; u8* REACH; u32* MASK; };
static int stage_chain(nlz_ctx* c, const Problem& pb, cudaStream_t st, const u64* LR, const u8* FLAGS, const ChainScratch& cs,
                       u64* d_out, u64 capacity, bool count_only, u64* out_count) {
    Workspace& w = c->ws;
    nlz_stats& S = c->stats;
    Profiler& P = c->prof;
    const u32 nfac = pb.nfac;
    const u32 nchunks = ceil_div_u32(nfac, CH_CHUNK);
    u32* EXIT = cs.EXIT;
    u32* J2 = cs.J2;
    u32* alist = cs.alist;
    u8* REACH = cs.REACH;
    u32* MASK = cs.MASK;
    u32* CNT = MASK + (size_t)nchunks * 32;
    u32* acount = w.CTR + 1;
    P.begin(st);
    NLZ_CK(cudaMemsetAsync(REACH, 0, nfac, st));
    ChainDom dom;
    memset(&dom, 0, sizeof(dom));
    dom.t0 = 0; dom.t1 = nfac; dom.nfac = nfac; dom.chunk = nfac; dom.G = 1;
    dom.J[0] = EXIT; dom.J2[0] = J2; dom.REACH[0] = REACH;
    k_chain_init<<<1, 1, 0, st>>>(alist, acount, REACH, (u32)pb.start_pos, 0u, nfac);
    k_chain_exit<<<nchunks, CH_THREADS, 0, st>>>(LR, 0u, nfac, nfac, EXIT, alist, acount);
    int rounds = bits_for(nchunks) + 1;
    {
        u32 grid = nchunks < (u32)kNumSM * 2 ? (nchunks ? nchunks : 1) : kNumSM * 2;
        for (int r = 0; r < rounds; ++r) k_chain_double<<<grid, 256, 0, st>>>(alist, acount, dom, r & 1);
    }
    k_chain_mark<<<nchunks, CH_THREADS, 0, st>>>(LR, nfac, REACH, MASK, CNT);
    k_scan_u32_single_cta<<<1, 1024, 0, st>>>(CNT, nchunks, w.CTR + 2);
    P.end(KC_CHAIN, (u64)nfac * (8 + 4 + 1 + 8 + 1), st, (u32)(4 + rounds));
    NLZ_CK(cudaMemcpyAsync(c->h_pinned + 2, w.CTR + 2, 4, cudaMemcpyDeviceToHost, st));
    NLZ_CK(cudaStreamSynchronize(st));
    S.host_syncs += 1;
    const u64 z = c->h_pinned[2];
    *out_count = z;
    S.n_factors = z;
    BatchView bv;
    bv.REC = w.REC; bv.fstart = w.FSTART; bv.flen = w.FLEN; bv.k = pb.nrec; bv.N = pb.rc ? pb.N : 0xFFFFFFFFu;
    if (!count_only || pb.nrec) {   // batch mode always runs the emit pass: it publishes the per-record boundaries
        u64* dst = d_out;
        if (count_only) { dst = nullptr; capacity = 0; }
        else if (!dst) {   // host-buffer entry points: library-owned device output
            if (z > c->d_out_cap) {
                if (c->d_out) NLZ_CK(cudaFree(c->d_out));
                c->d_out = nullptr; c->d_out_cap = 0;
                size_t want = z + z / 4 + 1024;
                NLZ_CK(cudaMalloc(&c->d_out, want * 24));
                c->d_out_cap = want;
            }
            dst = c->d_out;
            capacity = c->d_out_cap;
        }
        if (!count_only && z > capacity) {
            set_error("output capacity %llu factors is too small for %llu factors",
                      (unsigned long long)capacity, (unsigned long long)z);
            return ERR_RUNTIME;
        }
        P.begin(st);
        if (pb.nrec) {
            if (pb.rc) k_chain_emit<true, true><<<nchunks, CH_THREADS, 0, st>>>(LR, FLAGS, 0u, MASK, CNT, dst, capacity, bv, w.SENTIDX);
            else k_chain_emit<false, true><<<nchunks, CH_THREADS, 0, st>>>(LR, FLAGS, 0u, MASK, CNT, dst, capacity, bv, w.SENTIDX);
        } else {
            if (pb.rc) k_chain_emit<true, false><<<nchunks, CH_THREADS, 0, st>>>(LR, FLAGS, 0u, MASK, CNT, dst, capacity, bv, nullptr);
            else k_chain_emit<false, false><<<nchunks, CH_THREADS, 0, st>>>(LR, FLAGS, 0u, MASK, CNT, dst, capacity, bv, nullptr);
        }
        P.end(KC_CHAIN, (u64)nfac / 8 + z * 32, st);
    }
    return OK;
}

// ---- S0: text into X (validated / reverse-complemented on the device), byte histogram, key layout
static int stage_prepare(nlz_ctx* c, const Problem& pb, const void* src, bool src_on_host, cudaStream_t st,
                         u8* staging, ClassTable& tab, KeyLayout& lay) {
    Workspace& w = c->ws;
    nlz_stats& S = c->stats;
    Profiler& P = c->prof;
    const u32 n1 = pb.n1;
    P.begin(st);
    u32 prep_launches = 2;
    if (pb.nrec) {
        u8* tmp = staging;
        if (pb.packed_bytes > 8ull * n1) { set_error("batch staging overflow (%llu bytes)", (unsigned long long)pb.packed_bytes); return ERR_RUNTIME; }
        {
            u64 dst = 0;
            for (u32 r = 0; r < pb.nruns; ++r) {
                NLZ_CK(cudaMemcpyAsync(tmp + dst, static_cast<const u8*>(src) + pb.h_run_src[r], pb.h_run_len[r],
                                       src_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
                dst += pb.h_run_len[r];
            }
        }
        NLZ_CK(cudaMemcpyAsync(w.INOFF, pb.h_inoff, (size_t)pb.nrec * 4, cudaMemcpyHostToDevice, st));
        NLZ_CK(cudaMemcpyAsync(w.FSTART, pb.h_fstart, (size_t)pb.nrec * 4, cudaMemcpyHostToDevice, st));
        NLZ_CK(cudaMemcpyAsync(w.FLEN, pb.h_flen, (size_t)pb.nrec * 4, cudaMemcpyHostToDevice, st));
        k_set_u32<<<1, 1, 0, st>>>(w.CTR + 8, 0xFFFFFFFFu);
        k_prepare_batch<<<ceil_div_u32((u64)pb.N + 1, 256), 256, 0, st>>>(tmp, w.INOFF, w.FSTART, w.FLEN, pb.nrec, pb.N,
                                                                        pb.rc, w.X, w.REC, w.CTR + 8);
        prep_launches += 2;
    } else if (pb.mode == NLZ_MODE_DNA_RC) {
        const u8* dT = static_cast<const u8*>(src);
        if (src_on_host) {
            u8* tmp = staging;
            NLZ_CK(cudaMemcpyAsync(tmp, src, pb.n_in, cudaMemcpyHostToDevice, st));
            dT = tmp;
        }
        k_set_u32<<<1, 1, 0, st>>>(w.CTR + 8, 0xFFFFFFFFu);
        u32 grid = ceil_div_u32(pb.n_in, 256);
        if (grid > (u32)kNumSM * 16) grid = kNumSM * 16;
        k_prepare_dna_rc<<<grid, 256, 0, st>>>(dT, (u32)pb.n_in, w.X, w.CTR + 8);
        prep_launches += 2;
    } else {
        NLZ_CK(cudaMemcpyAsync(w.X, src, pb.n_in, src_on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    }
    k_zero_pad<<<1, 128, 0, st>>>(w.X, pb.L);
    if (pb.mode == NLZ_MODE_DNA_RC && !pb.nrec) {
        // the alphabet is known (A C G T + two unique sentinels): no byte histogram and no host round trip; the
        // validity flag is read back with the first count readback of the suffix sort (check_dna_deferred)
        P.end(KC_PREPARE, pb.n_in + 2 * pb.L, st, prep_launches - 1);
        NLZ_CK(cudaMemcpyAsync(c->h_pinned + 8, w.CTR + 8, 4, cudaMemcpyDeviceToHost, st));
        NLZ_CK(cudaEventRecord(c->ev[EV_PREP], st));
        u32 hist[256];
        for (int ch = 0; ch < 256; ++ch) hist[ch] = 0;
        hist['A'] = hist['C'] = hist['G'] = hist['T'] = 2;
        choose_layout(hist, n1, tab, lay);
        return OK;
    }
    NLZ_CK(cudaMemsetAsync(w.BYTEHIST, 0, 256 * 4, st));
    {
        u32 grid = ceil_div_u32(pb.L / 4 + 1, 256 * 8);
        if (grid > (u32)kNumSM * 8) grid = kNumSM * 8;
        k_byte_hist<<<grid, 256, 0, st>>>(w.X, pb.L, w.BYTEHIST);
    }
    P.end(KC_PREPARE, pb.n_in + 2 * pb.L, st, prep_launches);
    NLZ_CK(cudaMemcpyAsync(c->h_pinned + 16, w.BYTEHIST, 256 * 4, cudaMemcpyDeviceToHost, st));
    if (pb.nrec) NLZ_CK(cudaMemcpyAsync(c->h_pinned + 8, w.CTR + 8, 4, cudaMemcpyDeviceToHost, st));
    NLZ_CK(cudaEventRecord(c->ev[EV_PREP], st));
    NLZ_CK(cudaStreamSynchronize(st));
    S.host_syncs += 1;
    if (pb.nrec && c->h_pinned[8] != 0xFFFFFFFFu) {
        u32 bad = c->h_pinned[8], b = 0;
        while (b + 1 < pb.nrec && pb.h_fstart[b + 1] <= bad) ++b;
        set_error("Invalid nucleotide '%c' found in sequence %u",
                  (char)static_cast<const u8*>(src)[pb.h_srcoff[b] + (bad - pb.h_fstart[b])], b);
        return ERR_RUNTIME;
    }
    if (pb.nrec) {
        // records are told apart by the leading record-id field; every byte but ACGT is a sentinel
        for (int ch = 0; ch < 256; ++ch) tab.cls[ch] = (u8)SENT_CLASS;
        tab.cls['A'] = 0; tab.cls['C'] = 1; tab.cls['G'] = 2; tab.cls['T'] = 3;
        lay.key_bits = 64; lay.b = 2; lay.D = 5; lay.R = bits_for(pb.nrec);
        // windows only collide inside a record: size the window for the longest record
        u32 longest = 1;
        for (u32 q = 0; q < pb.nrec; ++q) longest = std::max(longest, pb.h_flen[q]);
        lay.W = layout_symbols(4, 2, lay.R, 5, (u64)longest * (pb.rc ? 2 : 1) + 2);
    } else {
        choose_layout(c->h_pinned + 16, n1, tab, lay);
    }
    return OK;
}

// DNA_RC mode: the invalid-nucleotide flag of k_prepare_dna_rc, checked after the first host sync of the suffix sort
static int check_dna_deferred(nlz_ctx* c, const Problem& pb, const void* src, bool src_on_host) {
    if (pb.mode != NLZ_MODE_DNA_RC || pb.nrec || c->h_pinned[8] == 0xFFFFFFFFu) return OK;
    const u32 bad = c->h_pinned[8];
    u8 ch = 0;
    if (src_on_host) ch = static_cast<const u8*>(src)[bad];
    else NLZ_CK(cudaMemcpy(&ch, static_cast<const u8*>(src) + bad, 1, cudaMemcpyDeviceToHost));
    // message of prepare_multiple_dna_sequences_w_rc, factorizer.cpp:91-92
    set_error("Invalid nucleotide '%c' found in sequence 0", (char)ch);
    return ERR_RUNTIME;
}

static void account_walk(nlz_ctx* c) {
    unsigned long long visits[2] = {0, 0};
    memcpy(visits, c->h_pinned + 4, 16);
    c->stats.walk_nodes = visits[0];
    c->stats.hard_positions = visits[1];
    c->prof.bytes[KC_WALK] += (u64)visits[0] * 16;
    c->prof.bytes[KC_WALK_HARD] += (u64)visits[1] * (4 + 8 + 8 + 8);
}

static int run_pipeline(nlz_ctx* c, const Problem& pb, const void* src, bool src_on_host, cudaStream_t st,
                        u64* d_out, u64 capacity, bool count_only, bool stop_after_index,
                        bool stop_after_lpnf, u64* out_count) {
    Workspace& w = c->ws;
    Profiler& P = c->prof;
    const u32 n1 = pb.n1;
    NLZ_CK(cudaEventRecord(c->ev[EV_BEGIN], st));
    c->text_suffixes = n1;
    ClassTable tab;
    KeyLayout lay;
    NLZ_TRY(stage_prepare(c, pb, src, src_on_host, st, reinterpret_cast<u8*>(w.KEY[1]), tab, lay));
    NLZ_TRY(stage_sa(c, pb, tab, lay, st));
    NLZ_TRY(check_dna_deferred(c, pb, src, src_on_host));
    NLZ_CK(cudaEventRecord(c->ev[EV_DOUBLING], st));

    // ---- S2: LCP (lcp.cuh): the first regroup has written every LCP value that follows from a key pair; the Kasai pass
    // resolves the marked positions (members of tie groups behind their head)
    BatchView bv;
    bv.REC = w.REC; bv.fstart = w.FSTART; bv.flen = w.FLEN; bv.k = pb.nrec; bv.N = pb.rc ? pb.N : 0xFFFFFFFFu;
    {
        LcpSlice<u32> ld;
        memset(&ld, 0, sizeof(ld));
        ld.NEED = w.NEED; ld.SA = w.SA; ld.RANK = w.RANK; ld.LCP = w.LCP; ld.pos0 = 0; ld.pos1 = n1;
        ld.blocks_per_warp = lcp_blocks_per_warp(n1);
        const u32 gk = ceil_div_u32(ceil_div_u32(n1, (u64)LCP_Q * ld.blocks_per_warp), 256);
        // algorithmic bytes: the marks of every position + 28 bytes per marked position (added from the active count)
        if (pb.nrec) KL(P, KC_LCP, (u64)n1 + (u64)c->stats.lcp_marked * 36, st, (k_lcp_kasai<true, false, u32><<<gk, 256, 0, st>>>(w.X, pb.L, bv, ld)));
        else KL(P, KC_LCP, (u64)n1 + (u64)c->stats.lcp_marked * 28, st, (k_lcp_kasai<false, false, u32><<<gk, 256, 0, st>>>(w.X, pb.L, bv, ld)));
    }
    NLZ_CK(cudaEventRecord(c->ev[EV_LCP], st));
    if (stop_after_index) {
        NLZ_CK(cudaGetLastError());
        return OK;
    }

    // ---- S3: summary trees, node tables, per-rank walk
    WalkParams wp;
    wp.n1 = n1; wp.nfac = pb.nfac; wp.N = pb.N; wp.twoN = 2 * pb.N;
    wp.real_lo = 0; wp.real_hi = n1; wp.rank_add = 0;
    u64* LR = w.KEY[0];
    u8* HARDF = reinterpret_cast<u8*>(w.SLOT[0]);       // flag plane: hard / reverse-complement bits, read again by the emit pass
    // leaf values: general mode reads the suffix array as it is; RC mode splits it into forward starts (in place over
    // SA, which no later stage reads) and rc values
    if (pb.rc)
        KL(P, KC_TREE, (u64)n1 * 12, st,
           (k_leaf_values<true, u32><<<ceil_div_u32(n1, 256), 256, 0, st>>>(w.SA, n1, wp, w.SA, w.R0)));
    NLZ_TRY(stage_lpnf(c, pb.rc, st, w.SA, w.R0, w.LCP, wp, w.RANK, LR, HARDF, false));
    NLZ_CK(cudaEventRecord(c->ev[EV_LPNF], st));
    if (stop_after_lpnf) {
        NLZ_CK(cudaGetLastError());
        return OK;
    }

    // ---- S4: chain
    ChainScratch cs;
    cs.EXIT = w.VAL[0]; cs.J2 = w.VAL[1];
    cs.alist = reinterpret_cast<u32*>(w.NODE);          // the node table is dead after stage 3 (SLOT[0] holds the flag plane)
    cs.REACH = reinterpret_cast<u8*>(w.SLOT[1]);
    cs.MASK = reinterpret_cast<u32*>(w.KEY[1]);
    NLZ_TRY(stage_chain(c, pb, st, LR, HARDF, cs, d_out, capacity, count_only, out_count));
    account_walk(c);
    NLZ_CK(cudaEventRecord(c->ev[EV_CHAIN], st));
    NLZ_CK(cudaGetLastError());
    return OK;
}

// Validates sizes the way the reference entry points do; returns OK with *empty = true when the
// reference would return zero factors without building an index.
static int make_problem(int mode, u64 n, u64 start_pos, Problem& pb, bool* empty) {
    *empty = false;
    pb.mode = mode; pb.n_in = n; pb.start_pos = start_pos; pb.N = 0;
    if (mode == NLZ_MODE_GENERAL) {
        pb.rc = false; pb.L = n; pb.nfac = (u32)n;
        if (n == 0 || start_pos >= n) { *empty = true; return OK; }   // factorizer_core.hpp:66
    } else if (mode == NLZ_MODE_RC_PREPARED) {
        pb.rc = true; pb.L = n;
        if (n < 4) { *empty = true; return OK; }                      // factorizer_core.hpp:180-193
        u64 N = n / 2 - 1;                                            // :195
        if (N == 0) { *empty = true; return OK; }                     // :196-200
        if (start_pos >= N) {                                         // :203-205
            set_error("start_pos must be less than the original sequence length");
            return ERR_INVALID;
        }
        pb.N = (u32)N; pb.nfac = (u32)N;
    } else if (mode == NLZ_MODE_DNA_RC) {
        pb.rc = true;
        if (n == 0) { *empty = true; return OK; }                     // factorizer_core.hpp:143
        pb.L = 2 * n + 2; pb.N = (u32)n; pb.nfac = (u32)n;
        if (start_pos != 0) { set_error("start_pos is not supported in DNA_RC mode"); return ERR_INVALID; }
    } else {
        set_error("unknown mode %d", mode);
        return ERR_INVALID;
    }
    if (pb.L + 1 >= 0xFFFFFFF0ull) {
        set_error("text of %llu symbols exceeds the 32-bit index path of this build",
                  (unsigned long long)pb.L);
        return ERR_RUNTIME;
    }
    pb.n1 = (u32)(pb.L + 1);
    return OK;
}

static void finish_stats(nlz_ctx* c, const Problem& pb) {
    nlz_stats& S = c->stats;
    S.n_text = pb.n_in; S.n_suffixes = pb.n1; S.n_factorized = pb.nfac;
    auto el = [&](int a, int b) { float ms = 0; if (cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess) { cudaGetLastError(); ms = 0; } return ms; };
    S.ms_prepare = el(EV_BEGIN, EV_PREP);
    S.ms_keys = el(EV_PREP, EV_KEYS);
    S.ms_sort0 = el(EV_KEYS, EV_SORT0);
    S.ms_doubling = el(EV_SORT0, EV_DOUBLING);
    S.ms_lcp = el(EV_DOUBLING, EV_LCP);
    S.ms_lpnf = el(EV_LCP, EV_LPNF);
    S.ms_chain = el(EV_LPNF, EV_CHAIN);
    S.ms_total = el(EV_BEGIN, EV_CHAIN);
    c->prof.collect();
    S.kernel_launches = c->prof.total_launches();
}

static void reset_stats(nlz_ctx* c) {
    size_t wsb = c->stats.workspace_bytes;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stats.workspace_bytes = wsb;
    c->prof.reset();
}

static int host_call(nlz_ctx* c, int mode, const u8* text, u64 n, u64 start_pos, u64** out_alloc,
                     u64* out_into, u64 capacity, bool count_only, u64* out_count) {
    if (!c) { set_error("null context"); return ERR_INVALID; }
    if (!out_count) { set_error("null out_count"); return ERR_INVALID; }
    if (n && !text) { set_error("null text"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    *out_count = 0;
    if (out_alloc) *out_alloc = nullptr;
    Problem pb;
    bool empty = false;
    NLZ_TRY(make_problem(mode, n, start_pos, pb, &empty));
    reset_stats(c);
    if (empty) return OK;
    NLZ_TRY(ensure_workspace(c, pb.n1));
    cudaStream_t st = c->own_stream;
    u64 z = 0;
    NLZ_TRY(run_pipeline(c, pb, text, true, st, nullptr, 0, count_only, false, false, &z));
    if (!count_only && z) {
        u64* dst = out_into;
        if (out_alloc) {
            dst = static_cast<u64*>(malloc((size_t)z * 24));
            if (!dst) { set_error("out of host memory for %llu factors", (unsigned long long)z); return ERR_RUNTIME; }
            *out_alloc = dst;
        } else if (z > capacity) {
            *out_count = z;
            NLZ_CK(cudaStreamSynchronize(st));
            set_error("output capacity %llu factors is too small for %llu factors",
                      (unsigned long long)capacity, (unsigned long long)z);
            return ERR_RUNTIME;
        }
        NLZ_CK(cudaMemcpyAsync(dst, c->d_out, (size_t)z * 24, cudaMemcpyDeviceToHost, st));
    }
    NLZ_CK(cudaStreamSynchronize(st));
    finish_stats(c, pb);
    *out_count = z;
    return OK;
}

}  // namespace nlz

template <typename KeyT>
static int debug_sort(nlz_ctx* c, KeyT* keys, uint32_t* vals, uint64_t m, int lo, int hi) {
    if (!c) { set_error("null context"); return ERR_INVALID; }
    if (m == 0) return OK;
    if (m >= 0xFFFFFFF0ull) { set_error("too many pairs"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    NLZ_TRY(ensure_workspace(c, m + 1));
    cudaStream_t st = c->own_stream;
    Workspace& w = c->ws;
    KeyT* k[2] = {reinterpret_cast<KeyT*>(w.KEY[0]), reinterpret_cast<KeyT*>(w.KEY[1])};
    u32* v[2] = {w.VAL[0], w.VAL[1]};
    NLZ_CK(cudaMemcpyAsync(k[0], keys, m * sizeof(KeyT), cudaMemcpyHostToDevice, st));
    NLZ_CK(cudaMemcpyAsync(v[0], vals, m * 4, cudaMemcpyHostToDevice, st));
    DigitPlan plan;
    plan_add_range(plan, lo, hi);
    int res = 0;
    c->prof.reset();
    NLZ_TRY(radix_sort_pairs<KeyT>(k, v, (u32)m, plan, w.HIST, st, &res, c->prof));
    NLZ_CK(cudaMemcpyAsync(keys, k[res], m * sizeof(KeyT), cudaMemcpyDeviceToHost, st));
    NLZ_CK(cudaMemcpyAsync(vals, v[res], m * 4, cudaMemcpyDeviceToHost, st));
    NLZ_CK(cudaStreamSynchronize(st));
    return OK;
}

// =================================================================== one text across G GPUs
#include "dist2_host.cuh"

// =================================================================== C ABI
extern "C" {
static int ctx_init_impl(nlz_ctx* c);
static int ctx_init(nlz_ctx* c) { return ctx_init_impl(c); }

const char* nlz_last_error(void) { return nlz::g_err.c_str(); }
const char* nlz_version(void) { return "1.2.0+b200.r1"; }
void nlz_free(void* p) { free(p); }
int nlz_host_register(void* p, uint64_t bytes) {
    if (!p || !bytes) { set_error("null range"); return ERR_INVALID; }
    // portable: pinned for every CUDA context of the process, whichever device was current at the time of the call
    cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {                       // a read-only mapping (np.load(..., mmap_mode="r")) needs the read-only flag
        cudaGetLastError();
        e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaHostRegister of %llu bytes failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
        return ERR_CUDA;
    }
    return OK;
}
int nlz_host_unregister(void* p) {
    NLZ_CK(cudaHostUnregister(p));
    return OK;
}

int nlz_ctx_create(int device, nlz_ctx** out) {
    if (!out) { set_error("null out"); return ERR_INVALID; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); nolzss_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range [0,%d)", device, count); return ERR_INVALID; }
    NLZ_CK(cudaSetDevice(device));
    nlz_ctx* c = new nlz_ctx();
    c->device = device;
    memset(&c->stats, 0, sizeof(c->stats));
    memset(c->ev, 0, sizeof(c->ev));
    memset(c->ring_ev, 0, sizeof(c->ring_ev));
    c->prof.reset();
    const int rc = ctx_init(c);
    if (rc != OK) { nlz_ctx_destroy(c); return rc; }      // tolerant of partially initialised fields
    *out = c;
    return OK;
}

static int ctx_init_impl(nlz_ctx* c) {
    NLZ_CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    NLZ_CK(cudaMallocHost(&c->h_pinned, 4096));   // words [0, 512): readbacks; [512, 1024): pipelined round counts
    NLZ_CK(cudaFuncSetAttribute(k_tile_sort<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSORT_SMEM));
    NLZ_CK(cudaFuncSetAttribute(k_group_stream<32, GS_CAP_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs_smem(GS_CAP_BIG)));
    NLZ_CK(cudaFuncSetAttribute(k_group_stream<32, GS_CAP_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs_smem(GS_CAP_SMALL)));
    NLZ_CK(cudaFuncSetAttribute(k_tile_sort<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSORT_SMEM));
    NLZ_CK(cudaFuncSetAttribute(k_group_stream<33, GS_CAP_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs_smem(GS_CAP_BIG)));
    NLZ_CK(cudaFuncSetAttribute(k_group_stream<33, GS_CAP_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gs_smem(GS_CAP_SMALL)));
    {
        // First launches of these kernels now, while nothing else runs: the CUDA runtime loads a kernel lazily at
        // its first launch, and that load waits for running kernels.  Inside a distributed doubling round a peer
        // rank of an in-process group (several ranks sharing one device) may already spin in its flag barrier --
        // the load would wait for the barrier and the barrier for this rank (seen as a barrier timeout).
        StreamOut so;
        memset(&so, 0, sizeof(so));
        RankDst rd;
        memset(&rd, 0, sizeof(rd));
        k_group_stream<32, GS_CAP_BIG><<<1, GS_THREADS, gs_smem(GS_CAP_BIG), c->own_stream>>>(nullptr, nullptr, nullptr, 0, 0, 64, nullptr, rd, so, 0);
        k_group_stream<32, GS_CAP_SMALL><<<1, GS_THREADS, gs_smem(GS_CAP_SMALL), c->own_stream>>>(nullptr, nullptr, nullptr, 0, 0, 64, nullptr, rd, so, 0);
        k_split_count<32><<<1, RG_THREADS, 0, c->own_stream>>>(nullptr, nullptr, 0, 64, nullptr);
        k_split_apply<32><<<1, RG_THREADS, 0, c->own_stream>>>(nullptr, nullptr, nullptr, 0, 64, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
        k_group_stream<33, GS_CAP_BIG><<<1, GS_THREADS, gs_smem(GS_CAP_BIG), c->own_stream>>>(nullptr, nullptr, nullptr, 0, 0, 64, nullptr, rd, so, 0);
        k_group_stream<33, GS_CAP_SMALL><<<1, GS_THREADS, gs_smem(GS_CAP_SMALL), c->own_stream>>>(nullptr, nullptr, nullptr, 0, 0, 64, nullptr, rd, so, 0);
        k_split_count<33><<<1, RG_THREADS, 0, c->own_stream>>>(nullptr, nullptr, 0, 64, nullptr);
        k_split_apply<33><<<1, RG_THREADS, 0, c->own_stream>>>(nullptr, nullptr, nullptr, 0, 64, nullptr, nullptr, 0, nullptr, nullptr, nullptr);
        NLZ_CK(cudaStreamSynchronize(c->own_stream));
    }
    for (int i = 0; i < EV_COUNT; ++i) NLZ_CK(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 48; ++i) NLZ_CK(cudaEventCreateWithFlags(&c->ring_ev[i], cudaEventDisableTiming));
    return OK;
}

void nlz_ctx_destroy(nlz_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->arena.base) cudaFree(c->arena.base);
    if (c->d_out) cudaFree(c->d_out);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    for (int i = 0; i < EV_COUNT; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 48; ++i) if (c->ring_ev[i]) cudaEventDestroy(c->ring_ev[i]);
    c->prof.destroy();
    cudaGetLastError();
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int nlz_ctx_device(nlz_ctx* c) { return c ? c->device : -1; }

int nlz_get_stats(nlz_ctx* c, nlz_stats* out) {
    if (!c || !out) { set_error("null argument"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    *out = c->stats;
    return OK;
}

int nlz_set_profiling(nlz_ctx* c, int on) {
    if (!c) { set_error("null context"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    c->prof.timing = on != 0;
    return OK;
}

int nlz_set_debug_flags(nlz_ctx* c, int flags) {
    if (!c) { set_error("null context"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    c->debug_flags = flags;
    return OK;
}

int nlz_kernel_class_count(void) { return KC_COUNT; }

int nlz_get_kernel_stats(nlz_ctx* c, int cls, const char** name, double* ms, uint64_t* bytes, uint32_t* launches) {
    if (!c || cls < 0 || cls >= KC_COUNT) { set_error("bad kernel class"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    if (name) *name = kClassNames[cls];
    if (ms) *ms = c->prof.ms[cls];
    if (bytes) *bytes = c->prof.bytes[cls];
    if (launches) *launches = c->prof.launches[cls];
    return OK;
}

int nlz_factorize_mode(nlz_ctx* c, int mode, const uint8_t* text, uint64_t n, uint64_t start_pos,
                       uint64_t** out_triples, uint64_t* out_count) {
    if (!out_triples) { set_error("null out_triples"); return ERR_INVALID; }
    return host_call(c, mode, text, n, start_pos, out_triples, nullptr, 0, false, out_count);
}
int nlz_factorize_mode_into(nlz_ctx* c, int mode, const uint8_t* text, uint64_t n, uint64_t start_pos,
                            uint64_t* out_triples, uint64_t capacity, uint64_t* out_count) {
    if (!out_triples && capacity) { set_error("null out_triples"); return ERR_INVALID; }
    return host_call(c, mode, text, n, start_pos, nullptr, out_triples, capacity, false, out_count);
}
int nlz_count_mode(nlz_ctx* c, int mode, const uint8_t* text, uint64_t n, uint64_t start_pos, uint64_t* out_count) {
    return host_call(c, mode, text, n, start_pos, nullptr, nullptr, 0, true, out_count);
}

int nlz_factorize_device(nlz_ctx* c, int mode, const void* d_text, uint64_t n, uint64_t start_pos,
                         void* cuda_stream, void* d_out_triples, uint64_t capacity, uint64_t* out_count) {
    if (!c || !out_count) { set_error("null argument"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    *out_count = 0;
    Problem pb;
    bool empty = false;
    NLZ_TRY(make_problem(mode, n, start_pos, pb, &empty));
    reset_stats(c);
    if (empty) return OK;
    NLZ_TRY(ensure_workspace(c, pb.n1));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    u64 z = 0;
    const bool count_only = (d_out_triples == nullptr);
    NLZ_TRY(run_pipeline(c, pb, d_text, false, st, static_cast<u64*>(d_out_triples), capacity, count_only,
                         false, false, &z));
    NLZ_CK(cudaStreamSynchronize(st));
    finish_stats(c, pb);
    *out_count = z;
    return OK;
}

int nlz_factorize_batch(nlz_ctx* c, int with_rc, const uint8_t* concat, const uint64_t* offsets, const uint64_t* lens,
                        uint64_t k, uint64_t** out_triples, uint64_t* per_record_counts, uint64_t* total) {
    if (!c || !total || !per_record_counts || (k && (!offsets || !lens))) { set_error("null argument"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    *total = 0;
    if (out_triples) *out_triples = nullptr;
    std::vector<u32> rec, inoff, fstart, flen;     // non-empty records only (empty ones yield no factors)
    std::vector<u64> srcoff, run_src, run_len;     // the caller's offsets; contiguous runs of the caller's buffer
    u64 fwd = 0, packed = 0;
    for (u64 b = 0; b < k; ++b) {
        per_record_counts[b] = 0;
        if (lens[b] == 0) continue;
        if (lens[b] > 0x7FFFFFF0ull || fwd + lens[b] + 1 > 0x7FFFFFF0ull || offsets[b] + lens[b] < offsets[b]) {
            set_error("batch of %llu records exceeds the 32-bit index path; split it", (unsigned long long)k);
            return ERR_INVALID;
        }
        rec.push_back((u32)b);
        srcoff.push_back(offsets[b]);
        // arbitrary (sparse, unordered, overlapping) offsets are allowed: a record that starts where the previous one
        // ended extends the current run, anything else opens a new run; records are packed back to back on the device
        if (!run_src.empty() && run_src.back() + run_len.back() == offsets[b]) run_len.back() += lens[b];
        else { run_src.push_back(offsets[b]); run_len.push_back(lens[b]); }
        inoff.push_back((u32)packed);
        packed += lens[b];
        fstart.push_back((u32)fwd);
        flen.push_back((u32)lens[b]);
        fwd += lens[b] + 1;
    }
    reset_stats(c);
    if (rec.empty()) return OK;
    if (rec.size() >= (1u << 24)) { set_error("batch of %zu records is too large; split it", rec.size()); return ERR_INVALID; }
    Problem pb;
    pb.mode = with_rc ? NLZ_MODE_RC_PREPARED : NLZ_MODE_GENERAL;
    pb.rc = with_rc != 0;
    pb.n_in = packed;
    pb.N = (u32)(fwd - 1);
    pb.nfac = pb.N;                    // every position but the last forward sentinel (factorizer_core.hpp:195, :241)
    pb.L = with_rc ? 2 * fwd : fwd;
    pb.start_pos = 0;
    if (pb.L + 1 >= 0xFFFFFFF0ull) { set_error("batch exceeds the 32-bit index path; split it"); return ERR_INVALID; }
    pb.n1 = (u32)(pb.L + 1);
    pb.nrec = (u32)rec.size();
    pb.h_inoff = inoff.data(); pb.h_fstart = fstart.data(); pb.h_flen = flen.data();
    pb.h_srcoff = srcoff.data(); pb.h_run_src = run_src.data(); pb.h_run_len = run_len.data();
    pb.nruns = (u32)run_src.size(); pb.packed_bytes = packed;
    NLZ_TRY(ensure_workspace(c, pb.n1, pb.nrec));
    cudaStream_t st = c->own_stream;
    u64 z = 0;
    NLZ_TRY(run_pipeline(c, pb, concat, true, st, nullptr, 0, out_triples == nullptr, false, false, &z));
    std::vector<u32> sentidx(pb.nrec, 0);
    if (pb.nrec > 1) NLZ_CK(cudaMemcpyAsync(sentidx.data(), c->ws.SENTIDX, (size_t)(pb.nrec - 1) * 4, cudaMemcpyDeviceToHost, st));
    const u64 zout = z - (pb.nrec - 1);
    if (out_triples && zout) {
        u64* dst = static_cast<u64*>(malloc((size_t)zout * 24));
        if (!dst) { set_error("out of host memory for %llu factors", (unsigned long long)zout); return ERR_RUNTIME; }
        *out_triples = dst;
        NLZ_CK(cudaMemcpyAsync(dst, c->d_out, (size_t)zout * 24, cudaMemcpyDeviceToHost, st));
    }
    NLZ_CK(cudaStreamSynchronize(st));
    finish_stats(c, pb);
    u64 prev = 0;                      // index (sentinel factors included) of the record's first factor
    for (u32 j = 0; j < pb.nrec; ++j) {
        const u64 end = (j + 1 < pb.nrec) ? sentidx[j] : z;
        per_record_counts[rec[j]] = end - prev;
        prev = end + 1;
    }
    *total = zout;
    return OK;
}

// ---- one text across G GPUs: nlz_dist_* live in dist2_host.cuh ---------------------------------

int nlz_factorize(nlz_ctx* c, const uint8_t* t, uint64_t n, uint64_t sp, uint64_t** o, uint64_t* cnt) {
    return nlz_factorize_mode(c, NLZ_MODE_GENERAL, t, n, sp, o, cnt);
}
int nlz_count_factors(nlz_ctx* c, const uint8_t* t, uint64_t n, uint64_t sp, uint64_t* cnt) {
    return nlz_count_mode(c, NLZ_MODE_GENERAL, t, n, sp, cnt);
}
int nlz_factorize_dna_w_rc(nlz_ctx* c, const uint8_t* t, uint64_t n, uint64_t** o, uint64_t* cnt) {
    return nlz_factorize_mode(c, NLZ_MODE_DNA_RC, t, n, 0, o, cnt);
}
int nlz_count_factors_dna_w_rc(nlz_ctx* c, const uint8_t* t, uint64_t n, uint64_t* cnt) {
    return nlz_count_mode(c, NLZ_MODE_DNA_RC, t, n, 0, cnt);
}
int nlz_factorize_multiple_dna_w_rc(nlz_ctx* c, const uint8_t* t, uint64_t n, uint64_t sp, uint64_t** o, uint64_t* cnt) {
    return nlz_factorize_mode(c, NLZ_MODE_RC_PREPARED, t, n, sp, o, cnt);
}
int nlz_count_factors_multiple_dna_w_rc(nlz_ctx* c, const uint8_t* t, uint64_t n, uint64_t sp, uint64_t* cnt) {
    return nlz_count_mode(c, NLZ_MODE_RC_PREPARED, t, n, sp, cnt);
}

// ---- stage probes ---------------------------------------------------------------------------
int nlz_debug_index(nlz_ctx* c, const uint8_t* text, uint64_t n, uint32_t* sa, uint32_t* isa, uint32_t* lcp) {
    if (!c) { set_error("null context"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    Problem pb;
    bool empty = false;
    NLZ_TRY(make_problem(NLZ_MODE_GENERAL, n, 0, pb, &empty));
    if (empty) { set_error("empty text"); return ERR_INVALID; }
    reset_stats(c);
    NLZ_TRY(ensure_workspace(c, pb.n1));
    cudaStream_t st = c->own_stream;
    u64 z = 0;
    NLZ_TRY(run_pipeline(c, pb, text, true, st, nullptr, 0, true, true, false, &z));
    NLZ_CK(cudaStreamSynchronize(st));
    if (sa) NLZ_CK(cudaMemcpy(sa, c->ws.SA, (size_t)pb.n1 * 4, cudaMemcpyDeviceToHost));
    if (isa) NLZ_CK(cudaMemcpy(isa, c->ws.RANK, (size_t)pb.n1 * 4, cudaMemcpyDeviceToHost));
    if (lcp) NLZ_CK(cudaMemcpy(lcp, c->ws.LCP, ((size_t)pb.n1 + 1) * 4, cudaMemcpyDeviceToHost));
    return OK;
}

int nlz_debug_sort_pairs_u64(nlz_ctx* c, uint64_t* keys, uint32_t* vals, uint64_t m, int lo, int hi) {
    return debug_sort<u64>(c, keys, vals, m, lo, hi);
}
int nlz_debug_sort_pairs_u32(nlz_ctx* c, uint32_t* keys, uint32_t* vals, uint64_t m, int lo, int hi) {
    return debug_sort<u32>(c, keys, vals, m, lo, hi);
}

int nlz_debug_per_position(nlz_ctx* c, int mode, const uint8_t* text, uint64_t n, uint64_t* len_out,
                           uint64_t* ref_out, uint64_t capacity, uint64_t* nfac_out) {
    if (!c || !nfac_out) { set_error("null argument"); return ERR_INVALID; }
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    Problem pb;
    bool empty = false;
    NLZ_TRY(make_problem(mode, n, 0, pb, &empty));
    *nfac_out = 0;
    if (empty) return OK;
    reset_stats(c);
    NLZ_TRY(ensure_workspace(c, pb.n1));
    cudaStream_t st = c->own_stream;
    u64 z = 0;
    NLZ_TRY(run_pipeline(c, pb, text, true, st, nullptr, 0, true, false, true, &z));
    NLZ_CK(cudaStreamSynchronize(st));
    *nfac_out = pb.nfac;
    if (capacity < pb.nfac) { set_error("capacity too small"); return ERR_RUNTIME; }
    std::vector<u64> lr(pb.nfac);
    std::vector<u8> fl(pb.nfac);
    NLZ_CK(cudaMemcpy(lr.data(), c->ws.KEY[0], (size_t)pb.nfac * 8, cudaMemcpyDeviceToHost));
    NLZ_CK(cudaMemcpy(fl.data(), c->ws.SLOT[0], (size_t)pb.nfac, cudaMemcpyDeviceToHost));
    for (u32 i = 0; i < pb.nfac; ++i) {
        u32 ref32 = (u32)(lr[i] >> 32);
        if (len_out) len_out[i] = (u32)lr[i];
        if (ref_out) ref_out[i] = (u64)ref32 | ((pb.rc && (fl[i] & FLAG_RC)) ? NLZ_RC_MASK : 0ull);
    }
    return OK;
}

}  // extern "C"
