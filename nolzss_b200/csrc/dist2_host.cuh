// Host orchestration of the distributed path (kernels and design: dist2.cuh).  Included by api.cu.
#pragma once
#include "dist2.cuh"

namespace nlz {

struct HostBarrier {          // in-process groups only (several ranks of one process, e.g. sharing a GPU in tests)
    std::mutex mu;
    std::condition_variable cv;
    int world = 0, waiting = 0;
    unsigned gen = 0;
    bool broken = false;      // a rank gave up waiting (a peer failed earlier): the group is unusable
    void* shared_ptr[4] = {nullptr, nullptr, nullptr, nullptr};   // rank 0 -> all (e.g. the host output buffer)
    bool arrive() {
        std::unique_lock<std::mutex> lk(mu);
        if (broken) return false;
        const unsigned g = gen;
        if (++waiting == world) { waiting = 0; ++gen; cv.notify_all(); return true; }
        if (!cv.wait_for(lk, std::chrono::seconds(180), [&] { return gen != g || broken; }) || broken) {
            broken = true;
            cv.notify_all();
            return false;
        }
        return true;
    }
};

}  // namespace nlz

struct nlz_dist {
    nlz_ctx* ctx = nullptr;
    int rank = 0, world = 1;
    u64 max_n1 = 0, max_nfac = 0;
    u64 inbox_items = 0;               // capacity of the inbox in 16-byte items
    u64 max_chT = 0;                   // T-positions per GPU at capacity (chain node arrays)
    u8* seg = nullptr;                 // shared segment: DistCtl | X replica | slice histogram | chain nodes | inbox
    size_t seg_bytes = 0, off_x = 0, off_hist = 0, off_cj = 0, off_cj2 = 0, off_creach = 0, off_inbox = 0;
    u8* peer[MAX_PEERS] = {};
    bool ipc_opened[MAX_PEERS] = {};
    bool attached = false;
    u32 epoch = 0;
    int xparity = 0;
    HostBarrier* hb = nullptr;
    bool owns_hb = false;
    u32* h_pin = nullptr;              // pinned: exchange readback
    u64 last_m_loc = 0;                // suffixes of this rank's range in the last call
    u32* SMALL = nullptr;              // device scratch: CTR[64] | PAY[512] | CURSOR[16] | SPLIT[16] | BASE64[2*16] | misc
    Arena arena;                       // per-call private workspace
};

namespace nlz {

struct D2Problem {
    int mode;
    u64 n_in, L, n1, N;
    u32 nfac;
    bool rc;
};

// Host side of a stream sync on the distributed path: busy-poll the stream.  A doubling round of a late round is a few
// hundred microseconds of GPU work between two host decisions on EVERY rank; a blocking / yielding
// cudaStreamSynchronize wakes up tens of microseconds late, each rank at a different moment, and the next barrier waits
// for the latest of them.  (NLZ_DIST_BLOCKING_SYNC=1 restores cudaStreamSynchronize.)
static inline cudaError_t d2_sync(cudaStream_t st) {
    static const bool blocking = getenv("NLZ_DIST_BLOCKING_SYNC") != nullptr;
    if (blocking) return cudaStreamSynchronize(st);
    for (;;) {
        const cudaError_t e = cudaStreamQuery(st);
        if (e != cudaErrorNotReady) return e;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

// per-call runtime of one rank
struct D2 {
    nlz_dist* d;
    nlz_ctx* c;
    cudaStream_t st;
    int G, me;
    u64 chunk, pos0, pos1;             // S-position slices
    u32 ch;                            // positions in my slice
    u32 chunkT, t0, t1;                // T-position slices (results, chain)
    u32 split[MAX_PEERS + 1];
    u64 base[MAX_PEERS + 1];
    u32 m_loc;
    u64 rank_bias;                     // test hook (debug flag 0x1000000): added to every global rank, so that small texts
                                       // carry 33-bit ranks through the keys, the sort kernels and the exchanges
    HandleMap hm;
    u32 *CTR, *PAY, *CURSOR;           // device scratch (PAY[0..8) doubles as the per-destination counts)
    u64* inbox;                        // my inbox (device pointer into the shared segment)
    std::vector<u32> all;              // host copy of the last barrier payloads
};

struct XInfo {
    u32 out_cnt[MAX_PEERS], out_off[MAX_PEERS], in_cnt[MAX_PEERS], in_off[MAX_PEERS], their_off[MAX_PEERS];
    u32 out_total, in_total;
};

// Stream-ordered barrier over all ranks; optionally every rank contributes `nwords` words (device buffer d_src) and
// receives everybody's words in r.all[g * nwords ..] (host sync).
static int d2_barrier(D2& r, const u32* d_src, u32 nwords, bool want_host) {
    nlz_dist* d = r.d;
    cudaStream_t st = r.st;
    DistPeers peers;
    memset(&peers, 0, sizeof(peers));
    for (int g = 0; g < r.G; ++g) peers.ctl[g] = reinterpret_cast<DistCtl*>(d->peer[g]);
    peers.n = r.G; peers.me = r.me;
    d->epoch += 1;
    if (nwords) d->xparity ^= 1;
    static const u32 timeout_s = getenv("NLZ_BARRIER_TIMEOUT_S") ? (u32)atoi(getenv("NLZ_BARRIER_TIMEOUT_S")) : 120u;
    static const bool trace = getenv("NLZ_TRACE_DIST") != nullptr;
    if (trace) fprintf(stderr, "[nlz dist] rank %d barrier %u (%u payload words)\n", r.me, d->epoch, nwords);
    if (d->hb) {
        // In-process groups (ranks that may share ONE device): the ranks first meet on the host, so that no barrier
        // kernel spins on the device while another rank still has work to launch (a kernel launched for the first time
        // is loaded lazily, and that load waits for running kernels; copies queued behind a spinning kernel block the
        // copy queue of the other ranks).  One process per GPU needs no such care.
        NLZ_CK(d2_sync(st));
        if (!d->hb->arrive()) {
            set_error("distributed barrier %u: a rank of the in-process group failed or never arrived (rank %d waited)", d->epoch, r.me);
            return ERR_RUNTIME;
        }
    }
    KL(r.c->prof, KC_BARRIER, (u64)nwords * 4 * r.G, st,
       (k_dist_barrier<<<1, 256, 0, st>>>(peers, d->epoch, d->xparity, d_src, nwords, timeout_s)));
    if (d->hb) NLZ_CK(d2_sync(st));
    if (!want_host) return OK;
    // one readback for everybody's words (the barrier kernel packs them, together with the time-out flag, into a
    // contiguous block of this rank's control area) -- a copy per rank costs more than the barrier itself
    DistCtl* mine = reinterpret_cast<DistCtl*>(d->seg);
    NLZ_CK(cudaMemcpyAsync(d->h_pin, &mine->packed[0], ((size_t)r.G * nwords + 1) * 4, cudaMemcpyDeviceToHost, st));
    NLZ_CK(d2_sync(st));
    r.c->stats.host_syncs += 1;
    if (d->h_pin[(size_t)r.G * nwords] != 0) {
        set_error("distributed barrier %u timed out on rank %d (a peer failed or never arrived)", d->epoch, r.me);
        return ERR_RUNTIME;
    }
    r.all.assign(d->h_pin, d->h_pin + (size_t)r.G * nwords);
    if (trace) {
        fprintf(stderr, "[nlz dist] rank %d passed %u:", r.me, d->epoch);
        for (u32 i = 0; i < (u32)r.G * nwords && i < 40; ++i) fprintf(stderr, " %u", r.all[i]);
        fprintf(stderr, "\n");
    }
    return OK;
}

// ---- bucketed bulk exchange ------------------------------------------------------------------------------------------
// 1. bucket the items by destination into the staging list(s); per-destination counts end up in r.PAY[0..8)
// (`slot`: several exchanges may share one counts barrier; exchange s keeps its counts in r.PAY[8s .. 8s+8))
template <typename F, bool TWO>
static int d2_bucket(D2& r, const F& f, u32 bound, const u32* count_dev, u64* stA, u64* stB, u64 alg_bytes_per_item, int slot = 0) {
    cudaStream_t st = r.st;
    u32* cnts = r.PAY + MAX_PEERS * slot;
    u32* cursor = r.CURSOR + MAX_PEERS * slot;
    NLZ_CK(cudaMemsetAsync(cnts, 0, MAX_PEERS * 4, st));
    if (!bound) return OK;
    const u32 grid = ceil_div_u32(bound, 256);
    Profiler& P = r.c->prof;
    P.begin(st);
    k_d2_bucket<F, 0, TWO><<<grid, 256, 0, st>>>(f, bound, count_dev, cnts, nullptr, nullptr);
    k_d2_bucket_starts<<<1, 32, 0, st>>>(cnts, cursor, r.G);
    k_d2_bucket<F, 1, TWO><<<grid, 256, 0, st>>>(f, bound, count_dev, cursor, stA, stB);
    P.end(KC_XCHG, (u64)bound * alg_bytes_per_item, st, 3);
    return OK;
}
// 2. counts barrier: payload = [counts[8] | nextra extra words from PAY[8..]]; fills the exchange geometry and
//    checks every rank's inbox capacity (the same verdict on every rank)
static int d2_geometry(D2& r, u32 nw, int slot, size_t item_bytes, XInfo& x, bool two_regions);
static int d2_counts(D2& r, u32 nextra, size_t item_bytes, XInfo& x, bool two_regions = false) {
    const u32 nw = MAX_PEERS + nextra;
    NLZ_TRY(d2_barrier(r, r.PAY, nw, true));
    return d2_geometry(r, nw, 0, item_bytes, x, two_regions);
}
// geometry of exchange `slot` from the gathered payloads (nw words per rank)
static int d2_geometry(D2& r, u32 nw, int slot, size_t item_bytes, XInfo& x, bool two_regions) {
    const int G = r.G, me = r.me;
    const size_t so = (size_t)MAX_PEERS * slot;
    memset(&x, 0, sizeof(x));
    for (int dst = 0; dst < G; ++dst) {
        u64 tot = 0;
        for (int g = 0; g < G; ++g) tot += r.all[(size_t)g * nw + so + dst];
        if (tot * item_bytes > r.d->inbox_items * 16ull || tot > 0xFFFFFFF0ull || (two_regions && tot > r.d->inbox_items)) {
            set_error("distributed exchange: rank %d would receive %llu items of %zu bytes, its inbox holds %llu bytes "
                      "(the partition of this text is too unbalanced for %d GPUs)", dst, (unsigned long long)tot, item_bytes,
                      (unsigned long long)(r.d->inbox_items * 16ull), G);
            return ERR_RUNTIME;
        }
    }
    u32 run = 0;
    for (int g = 0; g < G; ++g) {
        x.out_cnt[g] = r.all[(size_t)me * nw + so + g];
        x.out_off[g] = run;
        run += x.out_cnt[g];
    }
    x.out_total = run;
    run = 0;
    for (int g = 0; g < G; ++g) {
        x.in_cnt[g] = r.all[(size_t)g * nw + so + me];
        x.in_off[g] = run;
        run += x.in_cnt[g];
    }
    x.in_total = run;
    for (int g = 0; g < G; ++g) {            // where MY bucket starts in g's inbox: after the buckets of the ranks before me
        u32 o = 0;
        for (int q = 0; q < me; ++q) o += r.all[(size_t)q * nw + so + g];
        x.their_off[g] = o;
    }
    return OK;
}
// 3. one bulk copy per destination into its inbox (at `inbox_byte_off` + the bucket's offset), then a barrier
static int d2_push(D2& r, const XInfo& x, const void* staging, size_t item_bytes, size_t inbox_byte_off) {
    static const bool no_kernel_push = getenv("NLZ_NO_KERNEL_PUSH") != nullptr;
    if (item_bytes == 8 && r.G > 1 && x.out_total && x.out_total <= D2_KERNEL_PUSH_MAX && !no_kernel_push) {
        // one launch instead of G copies: coalesced 8-byte peer stores (the copies of a late doubling round are a few
        // kilobytes each -- their launch and DMA set-up latency, not their size, is what the round pays for)
        PushGeom pg;
        memset(&pg, 0, sizeof(pg));
        for (int g = 0; g < r.G; ++g) {
            pg.dst[g] = reinterpret_cast<u64*>(r.d->peer[g] + r.d->off_inbox + inbox_byte_off) + x.their_off[g];
            pg.off[g] = x.out_off[g];
        }
        for (int g = r.G; g <= MAX_PEERS; ++g) pg.off[g] = x.out_total;
        pg.G = r.G;
        u32 grid = ceil_div_u32(x.out_total, 256 * 4);
        if (grid > (u32)kNumSM * 4) grid = kNumSM * 4;
        k_d2_push<<<grid, 256, 0, r.st>>>(static_cast<const u64*>(staging), x.out_total, pg);
        r.c->prof.bytes[KC_XCHG] += (u64)x.out_total * 8;
        return OK;
    }
    for (int g = 0; g < r.G; ++g) {
        if (!x.out_cnt[g]) continue;
        u8* dst = r.d->peer[g] + r.d->off_inbox + inbox_byte_off + (size_t)x.their_off[g] * item_bytes;
        const u8* src = static_cast<const u8*>(staging) + (size_t)x.out_off[g] * item_bytes;
        NLZ_CK(cudaMemcpyAsync(dst, src, (size_t)x.out_cnt[g] * item_bytes, cudaMemcpyDefault, r.st));
        r.c->prof.bytes[KC_XCHG] += (u64)x.out_cnt[g] * item_bytes;
    }
    return OK;
}
struct VirtBlock { u32 cnt = 0, F = NONE_MIN, R = 0; };

static size_t d2_al(size_t b) { return (b + 255) & ~(size_t)255; }

// ---- S1 on the distributed path: doubling rounds over this GPU's rank range -------------------------------------------
// Lists, slots and group names are LOCAL; refined ranks leave as records (local rank << 32 | handle) in UPD and are
// delivered to the position owners; RANK[s+h] arrives through a request / response exchange.
struct D2Sa {
    u64* UPD;          // records of changed ranks (cnt entries)
    u64* ST;           // exchange staging (cnt entries)
    u64* REQ;          // request staging (cnt entries)
    u64* RANKL;        // my slice of RANK
    u8* NEED;          // marks of my positions: LCP pending (lcp.cuh)
    const u32* OFFIN;  // arrival order: offset of every received suffix in its sender's slice
    ArrivalSegs segs;  // arrival segments per sender
    u32* OFFR;         // handle tables in sorted order (HandleMap)
    u8* SNDR;
};

// One exchange step of a doubling round: (1) the ranks refined in the previous step travel to their position owners and
// (2) the rank requests of THIS round (RANK[pos(val[j]) + h] for the S list [0, mS) and the B list [b0, b0 + mB) of the
// `cur` buffers) travel to the same owners -- both under ONE counts barrier (which also carries every GPU's active
// count: *gm_out = the largest) and ONE data barrier; the owners apply the records, then answer the requests and push
// every answer into its requester's answer region; after a third barrier (answers arrived) they complete the keys:
// key[j] |= RANK[..].  (Three inbox regions: records | requests | answers.)
// With no active suffix anywhere (*gm_out == 0) only the records are delivered.
static int d2_round_exchange(D2& r, const D2Sa& a, u32 nupd_bound, const u32* nupd_dev, int cur, u32 mS, u32 b0, u32 mB, u64 h,
                             u32* gm_out) {
    Workspace& w = r.c->ws;
    cudaStream_t st = r.st;
    Profiler& P = r.c->prof;
    UpdItem ui;
    ui.upd = a.UPD; ui.hm = r.hm; ui.rbase = r.base[r.me] + r.rank_bias;
    NLZ_TRY((d2_bucket<UpdItem, false>(r, ui, nupd_bound, nupd_dev, a.ST, nullptr, 24, 0)));
    ReqItem qi;
    qi.val = w.VAL[cur]; qi.hm = r.hm; qi.h = h; qi.mS = mS; qi.b0 = b0;
    NLZ_TRY((d2_bucket<ReqItem, false>(r, qi, mS + mB, nullptr, a.REQ, nullptr, 20, 1)));
    k_set_u32<<<1, 1, 0, st>>>(r.PAY + 2 * MAX_PEERS, mS + mB);
    const u32 nw = 2 * MAX_PEERS + 1;
    NLZ_TRY(d2_barrier(r, r.PAY, nw, true));
    XInfo xu, xq;
    NLZ_TRY(d2_geometry(r, nw, 0, 8, xu, false));
    NLZ_TRY(d2_geometry(r, nw, 1, 8, xq, false));
    u32 gm = 0;
    for (int g = 0; g < r.G; ++g) gm = std::max(gm, r.all[(size_t)g * nw + 2 * MAX_PEERS]);
    *gm_out = gm;
    // three inbox regions of cap3 items: records | requests | answers
    const u64 cap3 = (r.d->inbox_items * 16 / 24) & ~(u64)31;
    for (int g = 0; g < r.G; ++g) {
        u64 in_u = 0, in_q = 0, out_q = 0;
        for (int q = 0; q < r.G; ++q) {
            in_u += r.all[(size_t)q * nw + g];
            in_q += r.all[(size_t)q * nw + MAX_PEERS + g];
            out_q += r.all[(size_t)g * nw + MAX_PEERS + q];
        }
        if (in_u > cap3 || in_q > cap3 || out_q > cap3) {
            set_error("distributed doubling round: rank %d would exchange %llu records / %llu requests / %llu answers, its inbox regions hold %llu "
                      "(the partition of this text is too unbalanced for %d GPUs)", g, (unsigned long long)in_u, (unsigned long long)in_q,
                      (unsigned long long)out_q, (unsigned long long)cap3, r.G);
            return ERR_RUNTIME;
        }
    }
    const size_t off_b = (size_t)cap3 * 8, off_c = (size_t)cap3 * 16;
    u64* inboxB = reinterpret_cast<u64*>(r.d->seg + r.d->off_inbox + off_b);
    const u64* inboxC = reinterpret_cast<const u64*>(r.d->seg + r.d->off_inbox + off_c);
    NLZ_TRY(d2_push(r, xu, a.ST, 8, 0));
    if (gm) NLZ_TRY(d2_push(r, xq, a.REQ, 8, off_b));
    NLZ_TRY(d2_barrier(r, nullptr, 0, false));                     // records and requests have arrived everywhere
    if (xu.in_total) {
        r.c->stats.rank_records_applied += xu.in_total;
        KL(P, KC_XCHG, (u64)xu.in_total * 16, st,
           (k_d2_apply_ranks<<<ceil_div_u32(xu.in_total, 256), 256, 0, st>>>(r.inbox, xu.in_total, a.RANKL, a.NEED)));
    }
    if (!gm) return OK;
    if (xq.in_total) {
        // answer the requests and push every answer into its requester's answer region, at the place of the request in
        // the requester's staging list (= the start of its bucket for me + the index inside the bucket)
        ServeGeom sg;
        memset(&sg, 0, sizeof(sg));
        sg.G = r.G;
        for (int g = 0; g < r.G; ++g) {
            u64 o = 0;                                             // where g staged its bucket for me: after its buckets for the ranks before me
            for (int q = 0; q < r.me; ++q) o += r.all[(size_t)g * nw + MAX_PEERS + q];
            sg.dst[g] = reinterpret_cast<u64*>(r.d->peer[g] + r.d->off_inbox + off_c) + o;
            sg.in_off[g] = xq.in_off[g];
        }
        for (int g = r.G; g <= MAX_PEERS; ++g) sg.in_off[g] = xq.in_total;
        KL(P, KC_GATHER, (u64)xq.in_total * 24, st,
           (k_d2_serve_push<<<ceil_div_u32(xq.in_total, 256), 256, 0, st>>>(inboxB, xq.in_total, a.RANKL, (u64)r.ch, sg)));
        P.bytes[KC_XCHG] += (u64)xq.in_total * 8;
    }
    NLZ_TRY(d2_barrier(r, nullptr, 0, false));                     // every answer has arrived
    if (xq.out_total)
        KL(P, KC_GATHER, (u64)xq.out_total * 32, st,
           (k_d2_apply_resp<<<ceil_div_u32(xq.out_total, 256), 256, 0, st>>>(a.REQ, inboxC, xq.out_total, w.KEY[cur])));
    // the next exchange writes into the inbox regions only after its own counts barrier, which every rank reaches after
    // its apply kernels have completed (stream order): no extra barrier needed here
    return OK;
}

static int d2_stage_sa(D2& r, const D2Problem& pb, const KeyLayout& lay, const D2Sa& a, u32 cnt) {
    r.hm.off = a.OFFR; r.hm.snd = a.SNDR; r.hm.chunk = r.chunk;
    nlz_ctx* c = r.c;
    Workspace& w = c->ws;
    nlz_stats& S = c->stats;
    Profiler& P = c->prof;
    cudaStream_t st = r.st;
    constexpr int GS = 33;
    S.key_bits = lay.key_bits; S.sym_bits = lay.b; S.key_syms = lay.W;
    RankDst rdst;
    rdst.rank = nullptr; rdst.upd = a.UPD; rdst.upd_count = w.CTR + 4; rdst.base = 0;
    // ---- initial sort of the (key, handle) pairs held in KEY[0] / VAL[0], first regroup
    int cur = 1;
    u32 m = 0, maxg = 0;
    NLZ_CK(cudaMemsetAsync(w.CTR, 0, 64, st));
    if (cnt) {
        u64* k[2] = {w.KEY[0], w.KEY[1]};
        u32* v[2] = {w.VAL[0], w.VAL[1]};
        DigitPlan plan;
        plan_add_range(plan, lay.dshift(), lay.key_bits);
        int res = 0;
        NLZ_TRY(radix_sort_pairs<u64>(k, v, cnt, plan, w.HIST, st, &res, P));
        NLZ_CK(cudaEventRecord(c->ev[EV_SORT0], st));
        // suffix handles = sorted indices from here on: position tables in sorted order
        KL(P, KC_REGROUP, (u64)cnt * 13, st,
           (k_d2_sorted_handles<<<ceil_div_u32(cnt, 256), 256, 0, st>>>(v[res], cnt, a.OFFIN, a.segs, a.OFFR, a.SNDR)));
        const u64 dist_mask = ((1ull << lay.D) - 1) << lay.dshift();
        const u32 tiles = ceil_div_u32(cnt, RG_TILE);
        P.begin(st);
        k_regroup_reduce<u64, true><<<tiles, RG_THREADS, 0, st>>>(k[res], cnt, dist_mask, w.PMAX, w.PSUM);
        k_regroup_scan_partials<<<1, 1024, 0, st>>>(w.PMAX, w.PSUM, tiles, w.CTR);
        LcpSeed seed;                                        // LCP values that follow from adjacent key pairs; the rest is marked
        seed.LCP = w.LCP; seed.NEED = nullptr; seed.lay = lay; seed.first_pending = r.base[r.me] > 0; seed.RANKOUT = nullptr; seed.need_in_rankout = false;
        k_regroup_apply<u64, true, GS><<<tiles, RG_THREADS, 0, st>>>(k[res], nullptr, nullptr, cnt, dist_mask, w.PMAX, w.PSUM,
                                                                     w.SA, rdst, w.KEY[res ^ 1], w.VAL[res ^ 1], w.SLOT[0], w.CTR + 3, seed);
        P.end(KC_REGROUP, (u64)cnt * (2 * 8 + 4 + 8), st, 3);
        cur = res ^ 1;
        k_set_u32<<<1, 1, 0, st>>>(w.CTR + 4, cnt);           // every rank is new: cnt dense records
    } else {
        NLZ_CK(cudaEventRecord(c->ev[EV_SORT0], st));
    }
    NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 32, cudaMemcpyDeviceToHost, st));
    NLZ_CK(d2_sync(st));
    S.host_syncs += 1;
    m = c->h_pinned[0]; maxg = c->h_pinned[3];
    S.lcp_marked = m;
    u32 nupd_bound = cnt;

    const int nbr = r.rank_bias ? 33 : bits_for((u32)std::min<u64>(pb.n1 - 1, 0xFFFFFFFFull)) + (pb.n1 > 0x100000000ull ? 1 : 0);   // bits of a global rank
    const int nbg = bits_for(cnt ? cnt - 1 : 0);                                                                  // bits of a local group name
    DigitPlan plan;
    plan_add_range(plan, 0, nbr > GS ? GS : nbr);
    plan_add_range(plan, GS, GS + nbg);
    u64 h = (u64)lay.W;
    int sc = 0;
    u32 gcap = (u32)TSORT_SLOTS / 2;
    if (((c->debug_flags >> 8) & 0xFFFF) >= 64 && (u32)((c->debug_flags >> 8) & 0xFFFF) < gcap) gcap = (u32)((c->debug_flags >> 8) & 0xFFFF);
    struct { bool on = false, off = false; u32 mS = 0, mB = 0, maxgS = 0; int fallbacks = 0; } hy;
    static const bool no_hybrid = getenv("NLZ_NO_HYBRID") != nullptr;
    hy.off = no_hybrid;
    const u32 END = cnt;
    static const bool trace = getenv("NLZ_TRACE") != nullptr;
    auto radix_round = [&](u32 mm, int* rb_out) -> int {
        u64* k[2] = {w.KEY[cur], w.KEY[cur ^ 1]};
        u32* v[2] = {w.VAL[cur], w.VAL[cur ^ 1]};
        int res = 0;
        NLZ_TRY(radix_sort_pairs<u64>(k, v, mm, plan, w.HIST, st, &res, P));
        const int rb = res == 0 ? cur : (cur ^ 1);
        const u32 tiles = ceil_div_u32(mm, RG_TILE);
        P.begin(st);
        k_regroup_reduce<u64, false><<<tiles, RG_THREADS, 0, st>>>(w.KEY[rb], mm, 0ull, w.PMAX, w.PSUM);
        k_regroup_scan_partials<<<1, 1024, 0, st>>>(w.PMAX, w.PSUM, tiles, w.CTR);
        k_regroup_apply<u64, false, GS><<<tiles, RG_THREADS, 0, st>>>(w.KEY[rb], w.VAL[rb], w.SLOT[sc], mm, 0ull, w.PMAX, w.PSUM,
                                                                      w.SA, rdst, w.KEY[rb ^ 1], w.VAL[rb ^ 1], w.SLOT[sc ^ 1], w.CTR + 3);
        P.end(KC_REGROUP, (u64)mm * (16 + 4 + 4 + 8 + 16), st, 3);
        *rb_out = rb;
        return OK;
    };
    for (;;) {
        // the one-time split into the S / B lists is local: do it before the exchange, which addresses list entries
        if (!hy.on && !hy.off && m > 0 && maxg > gcap) {
            const u32 tiles = ceil_div_u32(m, RG_TILE);
            P.begin(st);
            k_split_count<GS><<<tiles, RG_THREADS, 0, st>>>(w.KEY[cur], w.SLOT[sc], m, gcap, w.PSUM);
            k_scan_u32_single_cta<<<1, 1024, 0, st>>>(w.PSUM, tiles, w.CTR + 6);
            k_split_apply<GS><<<tiles, RG_THREADS, 0, st>>>(w.KEY[cur], w.VAL[cur], w.SLOT[sc], m, gcap, w.PSUM, w.CTR + 6, END,
                                                            w.KEY[cur ^ 1], w.VAL[cur ^ 1], w.SLOT[sc ^ 1]);
            P.end(KC_REGROUP, (u64)m * 40, st, 3);
            NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 32, cudaMemcpyDeviceToHost, st));
            NLZ_CK(d2_sync(st));
            S.host_syncs += 1;
            hy.on = true;
            hy.mB = c->h_pinned[6]; hy.mS = m - hy.mB; hy.maxgS = gcap;
            cur ^= 1; sc ^= 1;
            if (trace) fprintf(stderr, "[nlz] rank %d: hybrid rounds from here: S=%u B=%u (groups > %u)\n", r.me, hy.mS, hy.mB, gcap);
        }
        // deliver the ranks refined in the previous step, fetch RANK[s+h] for this round's lists; learn the largest
        // active count over all GPUs
        u32 gm = 0;
        {
            const u32 mS = hy.on ? hy.mS : m, mB = hy.on ? hy.mB : 0u;
            NLZ_TRY(d2_round_exchange(r, a, nupd_bound, w.CTR + 4, cur, mS, END - mB, mB, h, &gm));
        }
        if (gm == 0) break;
        S.doubling_rounds += 1;
        S.active_sum += hy.on ? hy.mS + hy.mB : m;
        int rb = cur;
        if (hy.on) {
            const u32 mS = hy.mS, mB = hy.mB, b0 = END - mB;
            NLZ_CK(cudaMemsetAsync(w.CTR, 0, 32, st));                  // [0] next S length, [3] its largest group, [4] records, [6] next B length, [7] fallback flag
            if (mS) {
                u32 cap = 32;
                while (cap < hy.maxgS) cap <<= 1;
                const u32 tile = TSORT_SLOTS - cap;
                KL(P, KC_TILE_SORT, (u64)mS * 40, st,
                   (k_tile_sort<GS><<<ceil_div_u32(mS, tile), TSORT_THREADS, TSORT_SMEM, st>>>(
                       w.KEY[cur], w.VAL[cur], w.SLOT[sc], mS, nullptr, tile, cap, w.SA, rdst, w.KEY[cur ^ 1], w.VAL[cur ^ 1],
                       w.SLOT[sc ^ 1], w.CTR, w.CTR + 3, c->debug_flags)));
            }
            if (mB) {
                StreamOut so;
                so.key_next = w.KEY[cur ^ 1]; so.val_next = w.VAL[cur ^ 1]; so.slot_next = w.SLOT[sc ^ 1];
                so.end = END; so.mS = w.CTR; so.maxgS = w.CTR + 3; so.mB = w.CTR + 6; so.fallback = w.CTR + 7;
                const int dbg = (c->debug_flags & 8) && S.doubling_rounds >= 3 ? 8 : 0;
                KL(P, KC_STREAM, (u64)mB * 44, st,
                   (launch_group_stream<GS>(c, st, w.KEY[cur], w.VAL[cur], w.SLOT[sc], b0, mB, gcap, w.SA, rdst, so, dbg)));
            }
            S.tile_sort_rounds += 1;
            NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 32, cudaMemcpyDeviceToHost, st));
            NLZ_CK(d2_sync(st));
            S.host_syncs += 1;
            if (c->h_pinned[7]) {
                // a group had more outliers than fit in shared memory: unify the lists and redo the round with the radix
                // path (the keys are gathered; slots written so far are rewritten with equal values, records are dropped)
                if (mB) {
                    NLZ_CK(cudaMemcpyAsync(w.KEY[cur ^ 1], w.KEY[cur] + b0, (size_t)mB * 8, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.KEY[cur] + mS, w.KEY[cur ^ 1], (size_t)mB * 8, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.VAL[cur ^ 1], w.VAL[cur] + b0, (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.VAL[cur] + mS, w.VAL[cur ^ 1], (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.SLOT[sc ^ 1], w.SLOT[sc] + b0, (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                    NLZ_CK(cudaMemcpyAsync(w.SLOT[sc] + mS, w.SLOT[sc ^ 1], (size_t)mB * 4, cudaMemcpyDeviceToDevice, st));
                }
                {
                    u32* k[2] = {w.SLOT[sc], w.SLOT[sc ^ 1]};
                    u32* v[2] = {reinterpret_cast<u32*>(w.KEY[cur ^ 1]), reinterpret_cast<u32*>(w.KEY[cur ^ 1]) + END};   // carried along, unused
                    DigitPlan p32;
                    plan_add_range(p32, 0, nbg);
                    int res = 0;
                    NLZ_TRY(radix_sort_pairs<u32>(k, v, mS + mB, p32, w.HIST, st, &res, P));
                    if (res) NLZ_CK(cudaMemcpyAsync(w.SLOT[sc], w.SLOT[sc ^ 1], (size_t)(mS + mB) * 4, cudaMemcpyDeviceToDevice, st));
                }
                NLZ_CK(cudaMemsetAsync(w.CTR + 4, 0, 16, st));
                hy.on = false; hy.off = ++hy.fallbacks >= 2;
                if (trace) fprintf(stderr, "[nlz] rank %d round %u: outliers exceed the stream kernel, back to radix rounds\n", r.me, S.doubling_rounds);
                NLZ_TRY(radix_round(mS + mB, &rb));
                NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 32, cudaMemcpyDeviceToHost, st));
                NLZ_CK(d2_sync(st));
                S.host_syncs += 1;
                m = c->h_pinned[0]; maxg = c->h_pinned[3];
            } else {
                hy.mS = c->h_pinned[0]; hy.maxgS = c->h_pinned[3]; hy.mB = c->h_pinned[6];
                m = hy.mS + hy.mB;
            }
            nupd_bound = mS + mB;
        } else {
            const bool fused = maxg <= gcap;
            NLZ_CK(cudaMemsetAsync(w.CTR, 0, 32, st));
            if (m > 0 && fused) {
                u32 cap = 32;
                while (cap < maxg) cap <<= 1;
                const u32 tile = TSORT_SLOTS - cap;
                KL(P, KC_TILE_SORT, (u64)m * 40, st,
                   (k_tile_sort<GS><<<ceil_div_u32(m, tile), TSORT_THREADS, TSORT_SMEM, st>>>(
                       w.KEY[cur], w.VAL[cur], w.SLOT[sc], m, nullptr, tile, cap, w.SA, rdst, w.KEY[cur ^ 1], w.VAL[cur ^ 1],
                       w.SLOT[sc ^ 1], w.CTR, w.CTR + 3, c->debug_flags)));
                rb = cur;
                S.tile_sort_rounds += 1;
            } else if (m > 0) {
                NLZ_TRY(radix_round(m, &rb));
            }
            nupd_bound = m;
            NLZ_CK(cudaMemcpyAsync(c->h_pinned, w.CTR, 32, cudaMemcpyDeviceToHost, st));
            NLZ_CK(d2_sync(st));
            S.host_syncs += 1;
            if (trace) fprintf(stderr, "[nlz] rank %d round %u h=%llu m=%u maxg=%u -> m'=%u maxg'=%u records=%u\n", r.me, S.doubling_rounds,
                               (unsigned long long)h, m, maxg, c->h_pinned[0], c->h_pinned[3], c->h_pinned[4]);
            m = c->h_pinned[0];
            maxg = c->h_pinned[3];
        }
        cur = rb ^ 1;
        sc ^= 1;
        h *= 2;
        if (S.doubling_rounds > 64) { set_error("prefix doubling did not converge"); return ERR_RUNTIME; }
    }
    return OK;
}

// Factorizes one text with all ranks of the group (every rank passes the same text).  Rank 0 receives the factors;
// every rank learns the count.
// `text` may be a host pointer (pageable or pinned) or a device pointer (unified addressing); `out_into` (rank 0) is a
// caller-provided, ideally pinned, buffer of `capacity` factors that replaces the malloc'ed result.
static int run_dist2(nlz_dist* d, const D2Problem& pb, const u8* text, u64** out_alloc, u64* out_into, u64 capacity, u64* out_count) {
    nlz_ctx* c = d->ctx;
    Workspace& w = c->ws;
    Profiler& P = c->prof;
    cudaStream_t st = c->own_stream;
    const int G = d->world, me = d->rank;
    D2 r;
    r.d = d; r.c = c; r.st = st; r.G = G; r.me = me;
    r.rank_bias = (c->debug_flags & 0x1000000) ? 0x100003039ull : 0ull;
    c->text_suffixes = pb.n1;
    w = Workspace();
    w.X = d->seg + d->off_x;
    w.CTR = d->SMALL;
    r.CTR = d->SMALL; r.PAY = d->SMALL + 64; r.CURSOR = d->SMALL + 64 + 512;
    u32* d_split = d->SMALL + 64 + 512 + 16;
    u64* d_base = reinterpret_cast<u64*>(d->SMALL + 64 + 512 + 32);           // 16 u64
    unsigned long long* d_bad = reinterpret_cast<unsigned long long*>(d->SMALL + 64 + 512 + 64);
    u32* d_ovf = d->SMALL + 64 + 512 + 68;
    r.inbox = reinterpret_cast<u64*>(d->seg + d->off_inbox);
    NLZ_CK(cudaEventRecord(c->ev[EV_BEGIN], st));

    // ---- geometry
    const u64 n1 = pb.n1;
    r.chunk = ((n1 + G - 1) / G + KB_TP - 1) / KB_TP * KB_TP;
    r.pos0 = std::min<u64>((u64)me * r.chunk, n1);
    r.pos1 = std::min<u64>((u64)(me + 1) * r.chunk, n1);
    r.ch = (u32)(r.pos1 - r.pos0);
    r.chunkT = (u32)((((u64)pb.nfac + G - 1) / G + CH_CHUNK - 1) / CH_CHUNK * CH_CHUNK);
    if (r.chunkT == 0) r.chunkT = CH_CHUNK;
    r.t0 = (u32)std::min<u64>((u64)me * r.chunkT, pb.nfac);
    r.t1 = (u32)std::min<u64>((u64)(me + 1) * r.chunkT, pb.nfac);
    if (r.chunk > D2_MAX_LOCAL) { set_error("text of %llu suffixes needs more than %d GPUs (at most %u positions per GPU)", (unsigned long long)n1, G, D2_MAX_LOCAL); return ERR_RUNTIME; }

    // ---- S0: text into every replica of X; alphabet; key layout
    ClassTable tab;
    KeyLayout lay;
    P.begin(st);
    u64 bad = ~0ull;
    if (pb.mode == NLZ_MODE_DNA_RC) {
        const u64 n = pb.n_in;
        const u64 per = ((n + G - 1) / G + 255) / 256 * 256;
        const u64 lo = std::min<u64>((u64)me * per, n), hi = std::min<u64>((u64)(me + 1) * per, n);
        XPeers xp;
        memset(&xp, 0, sizeof(xp));
        for (int g = 0; g < G; ++g) xp.p[g] = d->peer[g] + d->off_x;
        xp.n = G;
        k_d2_set_u64<<<1, 1, 0, st>>>(reinterpret_cast<u64*>(d_bad), ~0ull);
        if (hi > lo) {
            NLZ_CK(cudaMemcpyAsync(w.X + lo, text + lo, hi - lo, cudaMemcpyDefault, st));
            u32 grid = ceil_div_u32(hi - lo, 256);
            if (grid > (u32)kNumSM * 16) grid = kNumSM * 16;
            k_d2_prepare_dna_rc_slice<<<grid, 256, 0, st>>>(w.X, n, lo, hi, xp, me == 0, d_bad);
        } else if (me == 0) {
            k_d2_prepare_dna_rc_slice<<<1, 256, 0, st>>>(w.X, n, 0, 0, xp, true, d_bad);
        }
        k_zero_pad<<<1, 128, 0, st>>>(w.X, pb.L);
        P.end(KC_PREPARE, (hi - lo) * (1 + 2 * (u64)G), st, 3);
        // agree on the first invalid nucleotide (every rank must take the same exit)
        NLZ_CK(cudaMemcpyAsync(r.PAY, d_bad, 8, cudaMemcpyDeviceToDevice, st));
        NLZ_TRY(d2_barrier(r, r.PAY, 2, true));
        for (int g = 0; g < G; ++g) {
            const u64 b = (u64)r.all[(size_t)g * 2] | ((u64)r.all[(size_t)g * 2 + 1] << 32);
            bad = std::min(bad, b);
        }
        if (bad != ~0ull) {
            u8 chb = 0;
            NLZ_CK(cudaMemcpy(&chb, text + bad, 1, cudaMemcpyDefault));
            set_error("Invalid nucleotide '%c' found in sequence 0", (char)chb);   // factorizer.cpp:91-92
            return ERR_RUNTIME;
        }
        u32 hist[256];
        for (int ch = 0; ch < 256; ++ch) hist[ch] = 0;
        hist['A'] = hist['C'] = hist['G'] = hist['T'] = 2;
        choose_layout(hist, std::max<u64>(pb.n1, 0x100000000ull), tab, lay);   // forces 64-bit keys
    } else {
        NLZ_CK(cudaMemcpyAsync(w.X, text, pb.n_in, cudaMemcpyDefault, st));
        k_zero_pad<<<1, 128, 0, st>>>(w.X, pb.L);
        u32* BYTEHIST = d->SMALL + 64 + 512 + 128;
        NLZ_CK(cudaMemsetAsync(BYTEHIST, 0, 256 * 4, st));
        u32 grid = ceil_div_u32(pb.L / 4 + 1, 256 * 8);
        if (grid > (u32)kNumSM * 8) grid = kNumSM * 8;
        k_byte_hist<<<grid, 256, 0, st>>>(w.X, pb.L, BYTEHIST);
        P.end(KC_PREPARE, pb.n_in + 2 * pb.L, st, 2);
        NLZ_CK(cudaMemcpyAsync(c->h_pinned + 16, BYTEHIST, 256 * 4, cudaMemcpyDeviceToHost, st));
        NLZ_CK(cudaStreamSynchronize(st));
        choose_layout(c->h_pinned + 16, std::max<u64>(pb.n1, 0x100000000ull), tab, lay);   // forces 64-bit keys
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));                          // (keeps the barrier sequence of both branches aligned)
    }
    NLZ_CK(cudaEventRecord(c->ev[EV_PREP], st));

    // ---- partition: slice histograms, summed through peer memory -> bucket ranges / rank ranges
    int pbits = lay.W * lay.b;
    if (pbits > 24) pbits = 24;
    { int lim = bits_for((u32)std::min<u64>(n1, 0xFFFFFFFFull)) + 2; if (pbits > lim) pbits = lim; }
    pbits -= pbits % lay.b;
    if (pbits < lay.b) pbits = lay.b;
    const int K = pbits / lay.b;                                     // crossing nodes are shallower than K symbols
    const u32 nb = 1u << pbits;
    const u32 nctas = ceil_div_u32(r.ch ? r.ch : 1, KB_TP);
    KeyLayout lay_top = lay;
    lay_top.W = K;
    u32* HISTL = reinterpret_cast<u32*>(d->seg + d->off_hist);        // my slice histogram (peers read it)
    Splits sp;
    memset(&sp, 0, sizeof(sp));
    sp.G = G;
    P.begin(st);
    NLZ_CK(cudaMemsetAsync(HISTL, 0, (size_t)nb * 4, st));
    if (r.ch) k_d2_keys<0><<<nctas, 256, 0, st>>>(w.X, pb.L, r.pos0, r.pos1, tab, lay_top, pbits, sp, HISTL, nctas, nullptr, nullptr);
    P.end(KC_KEYS, (u64)r.ch + (u64)nb * 4, st, 1);
    NLZ_TRY(d2_barrier(r, nullptr, 0, false));
    // private workspace is sized once the rank ranges are known; the global histogram lives in the arena's head
    {
        const size_t need0 = d2_al((size_t)nb * 4) + d2_al(4096 * 8 + 64);
        if (need0 > d->arena.cap) {
            NLZ_CK(cudaStreamSynchronize(st));
            if (d->arena.base) { NLZ_CK(cudaFree(d->arena.base)); d->arena.base = nullptr; d->arena.cap = 0; }
            cudaError_t e = cudaMalloc(&d->arena.base, need0);
            if (e != cudaSuccess) { set_error("cudaMalloc of %zu workspace bytes failed: %s", need0, cudaGetErrorString(e)); return ERR_CUDA; }
            d->arena.cap = need0;
        }
    }
    {
        u32* GH = reinterpret_cast<u32*>(d->arena.base);
        u64* TS = reinterpret_cast<u64*>(d->arena.base + d2_al((size_t)nb * 4));
        HistPeers hp;
        memset(&hp, 0, sizeof(hp));
        for (int g = 0; g < G; ++g) hp.h[g] = reinterpret_cast<const u32*>(d->peer[g] + d->off_hist);
        hp.G = G;
        P.begin(st);
        NLZ_CK(cudaMemsetAsync(d_ovf, 0, 4, st));
        k_d2_hist_sum<<<ceil_div_u32(nb, 256), 256, 0, st>>>(hp, nb, GH, d_ovf);
        k_d2_hist_tiles<<<ceil_div_u32(nb, 4096), 256, 0, st>>>(GH, nb, TS);
        k_d2_splitters<<<1, 288, 0, st>>>(GH, TS, nb, n1, G, d_split, d_base);
        P.end(KC_KEYS, (u64)nb * 4 * (G + 2), st, 3);
        NLZ_CK(cudaMemcpyAsync(d->h_pin, d_split, 64 * 4, cudaMemcpyDeviceToHost, st));   // SPLIT[16] | BASE[16 u64] | bad | - | ovf (word 52)
        NLZ_CK(cudaStreamSynchronize(st));
        c->stats.host_syncs += 1;
        const u64* hb64 = reinterpret_cast<const u64*>(d->h_pin + 16);
        for (int g = 0; g <= G; ++g) { r.split[g] = d->h_pin[g]; r.base[g] = hb64[g]; sp.split[g] = r.split[g]; }
        if (d->h_pin[52]) { set_error("a 12-symbol bucket holds more than 2^32 suffixes: the text is too repetitive for the distributed partition"); return ERR_RUNTIME; }
    }
    for (int g = 0; g < G; ++g) {
        if (r.base[g + 1] - r.base[g] > D2_MAX_LOCAL) {
            set_error("rank range of GPU %d holds %llu suffixes (at most %u per GPU): the partition of this text is too unbalanced for %d GPUs",
                      g, (unsigned long long)(r.base[g + 1] - r.base[g]), D2_MAX_LOCAL, G);
            return ERR_RUNTIME;
        }
    }
    r.m_loc = (u32)(r.base[me + 1] - r.base[me]);
    d->last_m_loc = r.m_loc;
    const u32 m_loc = r.m_loc;
    const u32 ch = r.ch;
    const u32 nT = r.t1 - r.t0;                                      // T-positions of my slice

    // ---- private workspace
    const u64 cap = (u64)std::max(m_loc, ch) + 2 * DIST_VIRT + 72;
    u64 *UPD, *ST, *RANKL, *SA64, *PHI, *STB;
    u8* NEED;
    u32 *OFFIN, *OFFR, *PLCP, *F0buf, *R0buf, *LCPbuf, *DCNT;
    u8* SNDR;
    u64* LRT; u8* FLT; u32 *alist, *MASK;
    u64* LRl; u8* FLl;
    {
        auto tiles_of = [](u64 x) { return (x + RG_TILE - 1) / RG_TILE + 1; };
        const size_t ndc = (size_t)MAX_PEERS * nctas + 8;
        size_t need = 0;
        need += d2_al(cap * 8) * 2 + d2_al(cap * 4) * 4;                 // KEY, VAL, SLOT
        need += d2_al((cap + 72) * 4) + d2_al((cap + 72) * 16);          // SA (handles), NODE
        need += d2_al(((size_t)RS_BINS * RS_MAX_CTAS + RS_BINS) * 4);    // HIST
        need += d2_al(tiles_of(cap) * 4) * 2;                            // PMAX, PSUM
        { u64 cnt = cap + 1; for (int lev = 1; lev < TREE_MAX_LEVELS; ++lev) { cnt = (cnt + 31) / 32; need += d2_al((cnt + 72) * 4) * 3; } }
        need += d2_al(cap * 8) * 3;                                      // UPD (later SA64), ST, STB
        need += d2_al(((size_t)ch + 8) * 8) + d2_al((size_t)ch + 128);   // RANKL, NEED
        need += d2_al(cap * 4) * 2 + d2_al(cap);                         // OFFIN, OFFR, SNDR
        need += d2_al((cap + 72) * 4) * 3;                               // F0, R0, LCP (with virtual ranks)
        need += d2_al(ndc * 4 + SCAN_TILE * 4 + (ndc / SCAN_TILE + 8) * 4);   // DCNT + tile sums
        need += d2_al(((size_t)nT + 8) * 8) + d2_al((size_t)nT + 64);    // LRT, FLT
        need += d2_al(((size_t)nT + 8) * 4);                             // alist
        need += d2_al(((size_t)nT / CH_CHUNK + 2) * 33 * 4);             // MASK + CNT
        need += 8192;
        if (need > d->arena.cap) {
            NLZ_CK(cudaStreamSynchronize(st));
            if (d->arena.base) { NLZ_CK(cudaFree(d->arena.base)); d->arena.base = nullptr; d->arena.cap = 0; }
            const size_t want = need + need / 16;
            cudaError_t e = cudaMalloc(&d->arena.base, want);
            if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&d->arena.base, need); if (e == cudaSuccess) d->arena.cap = need; }
            else d->arena.cap = want;
            if (e != cudaSuccess) { set_error("cudaMalloc of %zu workspace bytes failed: %s", need, cudaGetErrorString(e)); return ERR_CUDA; }
        }
        if (d->hb && !d->hb->arrive()) { set_error("a rank of the in-process group failed before the workspace rendezvous"); return ERR_RUNTIME; }
        Arena& a = d->arena;
        a.off = 0;
        for (int i = 0; i < 2; ++i) w.KEY[i] = a.take<u64>(cap);
        for (int i = 0; i < 2; ++i) w.VAL[i] = a.take<u32>(cap);
        for (int i = 0; i < 2; ++i) w.SLOT[i] = a.take<u32>(cap);
        w.SA = a.take<u32>(cap + 72);
        w.NODE = a.take<uint4>(cap + 72);
        w.HIST = a.take<u32>((size_t)RS_BINS * RS_MAX_CTAS + RS_BINS);
        w.PMAX = a.take<u32>(tiles_of(cap));
        w.PSUM = a.take<u32>(tiles_of(cap));
        { u64 cnt = cap + 1; for (int lev = 1; lev < TREE_MAX_LEVELS; ++lev) { cnt = (cnt + 31) / 32; w.tl[lev] = a.take<u32>(cnt + 72); w.tf[lev] = a.take<u32>(cnt + 72); w.tr[lev] = a.take<u32>(cnt + 72); } }
        UPD = a.take<u64>(cap); ST = a.take<u64>(cap); STB = a.take<u64>(cap);
        RANKL = a.take<u64>((size_t)ch + 8);
        NEED = a.take<u8>((size_t)ch + 128);
        OFFIN = a.take<u32>(cap);
        OFFR = a.take<u32>(cap);
        SNDR = a.take<u8>(cap);
        F0buf = a.take<u32>(cap + 72); R0buf = a.take<u32>(cap + 72); LCPbuf = a.take<u32>(cap + 72);
        DCNT = a.take<u32>(ndc + SCAN_TILE + ndc / SCAN_TILE + 8);
        LRT = a.take<u64>((size_t)nT + 8); FLT = a.take<u8>((size_t)nT + 64);
        alist = a.take<u32>((size_t)nT + 8);
        MASK = a.take<u32>(((size_t)nT / CH_CHUNK + 2) * 33);
        SA64 = UPD;                                  // the record list is dead once the doubling has converged
        PHI = w.KEY[0];                              // the sort buffers are free between the doubling and stage 3
        PLCP = reinterpret_cast<u32*>(w.KEY[1]);
        LRl = w.KEY[0];                              // stage 3 results by work item (PHI is dead by then)
        FLl = reinterpret_cast<u8*>(w.SLOT[0]);
        c->stats.workspace_bytes = d->arena.cap + d->seg_bytes;
    }
    w.n1 = m_loc;

    // ---- (key, position) pairs to the owners of their buckets, in position order
    XInfo xk;
    {
        const size_t ndc = (size_t)MAX_PEERS * nctas;
        u32* tsum = DCNT + ndc + 8;
        u64* KOUT = ST;                                              // staging: keys | offsets
        u32* OOUT = reinterpret_cast<u32*>(STB);
        P.begin(st);
        NLZ_CK(cudaMemsetAsync(DCNT, 0, (ndc + 8) * 4, st));
        if (ch) k_d2_keys<1><<<nctas, 256, 0, st>>>(w.X, pb.L, r.pos0, r.pos1, tab, lay_top, pbits, sp, DCNT, nctas, nullptr, nullptr);   // the destination only needs the prefix
        {
            const u32 cntd = (u32)ndc;
            const u32 nt = ceil_div_u32(cntd, SCAN_TILE);
            if (nt > 1) {
                k_scan_tiles<false><<<nt, 1024, 0, st>>>(DCNT, cntd, tsum);
                k_scan_u32_single_cta<<<1, 1024, 0, st>>>(tsum, nt, nullptr);
                k_scan_tiles<true><<<nt, 1024, 0, st>>>(DCNT, cntd, tsum);
            } else {
                k_scan_u32_single_cta<<<1, 1024, 0, st>>>(DCNT, cntd, nullptr);
            }
        }
        NLZ_CK(cudaMemsetAsync(r.PAY, 0, MAX_PEERS * 4, st));
        k_d2_bucket_totals<<<1, 32, 0, st>>>(DCNT, nctas, ch, G, r.PAY);
        if (ch) k_d2_keys<2><<<nctas, 256, 0, st>>>(w.X, pb.L, r.pos0, r.pos1, tab, lay, pbits, sp, DCNT, nctas, KOUT, OOUT);
        P.end(KC_KEYS, (u64)ch * (2 + 12), st, 5);
        NLZ_TRY(d2_counts(r, 0, 12, xk, true));
        if (xk.in_total != m_loc) { set_error("internal: rank %d received %u pairs for a rank range of %u", me, xk.in_total, m_loc); return ERR_RUNTIME; }
        const size_t off_b = (size_t)d->inbox_items * 8;            // second inbox region (offsets)
        NLZ_TRY(d2_push(r, xk, KOUT, 8, 0));
        NLZ_TRY(d2_push(r, xk, OOUT, 4, off_b));
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));
        if (m_loc)
            KL(P, KC_KEYS, (u64)m_loc * 28, st,
               (k_d2_unpack_keys<<<ceil_div_u32(m_loc, 256), 256, 0, st>>>(r.inbox, reinterpret_cast<const u32*>(d->seg + d->off_inbox + off_b),
                                                                         m_loc, w.KEY[0], w.VAL[0], OFFIN)));
        // (the handle tables r.hm are built from the arrival tables after the sort: d2_stage_sa)
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));                    // everyone has unpacked: the inboxes are free again
    }
    NLZ_CK(cudaEventRecord(c->ev[EV_KEYS], st));

    // ---- S1: local sort + doubling rounds (ranks travel as records, RANK[s+h] by request / response)
    if (ch) NLZ_CK(cudaMemsetAsync(RANKL, 0, ((size_t)ch + 8) * 8, st));
    NLZ_CK(cudaMemsetAsync(NEED, 0, (size_t)ch + 128, st));
    w.LCP = LCPbuf + DIST_VIRT;                                   // seeded by the first regroup
    D2Sa sa;
    sa.UPD = UPD; sa.ST = ST; sa.REQ = STB; sa.RANKL = RANKL; sa.NEED = NEED;
    sa.OFFIN = OFFIN; sa.OFFR = OFFR; sa.SNDR = SNDR;
    memset(&sa.segs, 0, sizeof(sa.segs));
    sa.segs.G = G;
    for (int g = 0; g < G; ++g) sa.segs.seg[g] = xk.in_off[g];
    for (int g = G; g <= MAX_PEERS; ++g) sa.segs.seg[g] = xk.in_total;
    NLZ_TRY(d2_stage_sa(r, pb, lay, sa, m_loc));
    NLZ_CK(cudaEventRecord(c->ev[EV_DOUBLING], st));

    // ---- S2: suffix array as S-positions; Phi to the position owners, Kasai per position slice, PLCP back to the rank owners
    if (m_loc)
        KL(P, KC_LCP, (u64)m_loc * 16, st, (k_d2_sa_positions<<<ceil_div_u32(m_loc, 256), 256, 0, st>>>(w.SA, m_loc, r.hm, SA64)));
    {
        if (m_loc) NLZ_CK(cudaMemcpyAsync(r.PAY, SA64 + (m_loc - 1), 8, cudaMemcpyDeviceToDevice, st));
        else k_d2_set_u64<<<1, 1, 0, st>>>(reinterpret_cast<u64*>(r.PAY), D2_NO_PHI);
        NLZ_TRY(d2_barrier(r, r.PAY, 2, true));
        u64 left_sa = D2_NO_PHI;
        for (int g = me - 1; g >= 0; --g)
            if (r.base[g + 1] > r.base[g]) { left_sa = (u64)r.all[(size_t)g * 2] | ((u64)r.all[(size_t)g * 2 + 1] << 32); break; }
        PhiItem pi;
        pi.SA64 = SA64; pi.LCP = w.LCP; pi.left_sa = left_sa; pi.chunk = r.chunk;
        NLZ_TRY((d2_bucket<PhiItem, false>(r, pi, m_loc, nullptr, ST, nullptr, 24)));
        XInfo x;
        NLZ_TRY(d2_counts(r, 0, 8, x));
        NLZ_TRY(d2_push(r, x, ST, 8, 0));
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));
        if (x.in_total) KL(P, KC_LCP, (u64)x.in_total * 16, st, (k_d2_apply_phi<<<ceil_div_u32(x.in_total, 256), 256, 0, st>>>(r.inbox, x.in_total, PHI)));
        LcpSlice<u64> ld;
        memset(&ld, 0, sizeof(ld));
        ld.NEED = NEED; ld.PHI = PHI; ld.PLCP = PLCP; ld.pos0 = r.pos0; ld.pos1 = r.pos1;
        ld.blocks_per_warp = lcp_blocks_per_warp(ch);
        BatchView bv;
        memset(&bv, 0, sizeof(bv));
        if (ch)
            KL(P, KC_LCP, (u64)ch * 28, st,
               (k_lcp_kasai<false, true, u64><<<ceil_div_u32(ceil_div_u32(ch, (u64)LCP_Q * ld.blocks_per_warp), 256), 256, 0, st>>>(w.X, pb.L, bv, ld)));
        LcpItem li;
        li.RANKL = RANKL; li.PLCP = PLCP; li.NEED = NEED; li.rb.G = G;
        for (int g = 0; g <= G; ++g) li.rb.base[g] = r.base[g] + r.rank_bias;
        for (int g = G + 1; g <= MAX_PEERS; ++g) li.rb.base[g] = r.base[G] + r.rank_bias;
        NLZ_TRY((d2_bucket<LcpItem, false>(r, li, ch, nullptr, ST, nullptr, 20)));
        NLZ_TRY(d2_counts(r, 0, 8, x));
        NLZ_TRY(d2_push(r, x, ST, 8, 0));
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));
        if (x.in_total) KL(P, KC_LCP, (u64)x.in_total * 12, st, (k_d2_apply_lcp<<<ceil_div_u32(x.in_total, 256), 256, 0, st>>>(r.inbox, x.in_total, w.LCP)));
    }
    NLZ_CK(cudaEventRecord(c->ev[EV_LCP], st));

    // ---- S3a: leaf values; boundary exchange (edge staircases of the local range -> virtual ranks)
    WalkParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.n1 = m_loc; wp.nfac = pb.nfac; wp.N = (u32)pb.N; wp.twoN = 0;
    wp.real_lo = 0; wp.real_hi = m_loc; wp.rank_add = 0;
    u32* F0 = F0buf + DIST_VIRT;
    u32* R0 = R0buf + DIST_VIRT;
    u32* d_pay = r.PAY;
    NLZ_CK(cudaMemsetAsync(d_pay, 0, DIST_XCH_WORDS * 4, st));
    if (m_loc) {
        if (pb.rc) KL(P, KC_TREE, (u64)m_loc * 16, st, (k_leaf_values<true, u64><<<ceil_div_u32(m_loc, 256), 256, 0, st>>>(SA64, m_loc, wp, F0, R0)));
        else KL(P, KC_TREE, (u64)m_loc * 12, st, (k_leaf_values<false, u64><<<ceil_div_u32(m_loc, 256), 256, 0, st>>>(SA64, m_loc, wp, F0, R0)));
        NLZ_CK(cudaMemcpyAsync(d_pay, w.LCP, 4, cudaMemcpyDeviceToDevice, st));          // c0 = lcp(previous GPU's last, my first)
        NLZ_CK(cudaMemsetAsync(w.LCP, 0, 4, st));
        NLZ_CK(cudaMemsetAsync(w.LCP + m_loc, 0, 4, st));
        k_set_u32<<<1, 1, 0, st>>>(d_pay + 1, m_loc);
        NLZ_TRY(stage_lpnf(c, pb.rc, st, F0, R0, w.LCP, wp, nullptr, nullptr, nullptr, true));
        if (pb.rc) k_dist_edges<true><<<1, 64, 0, st>>>(c->trees, wp, m_loc, K, d_pay + 4);
        else k_dist_edges<false><<<1, 64, 0, st>>>(c->trees, wp, m_loc, K, d_pay + 4);
    }
    const u32 pay_words = 4 + 8 * (u32)K;
    NLZ_TRY(d2_barrier(r, d_pay, pay_words, true));
    u32 nwork_bound = 0;
    if (m_loc) {
        const std::vector<u32>& all = r.all;
        auto C0 = [&](int g) { return all[(size_t)g * pay_words + 0]; };
        auto M = [&](int g) { return all[(size_t)g * pay_words + 1]; };
        auto E = [&](int g, int side, int v) { return &all[(size_t)g * pay_words + 4 + (size_t)(side * K + v - 1) * 4]; };
        auto minint = [&](int g) { u32 q = 0; for (int v = 1; v <= K; ++v) if (E(g, 0, v)[3]) q = (u32)v; return q; };
        auto prev_ne = [&](int g) { for (--g; g >= 0; --g) if (M(g)) return g; return -1; };
        auto next_ne = [&](int g) { for (++g; g < G; ++g) if (M(g)) return g; return -1; };
        auto merge = [&](std::vector<VirtBlock>& blk, int g, int side, u32 cur) {
            for (int v = 1; v <= K; ++v) {
                const u32* e = E(g, side, v);
                if (!e[0]) continue;
                const u32 eff = (u32)v < cur ? (u32)v : cur;
                VirtBlock& b = blk[eff];
                b.cnt += e[0];
                if (e[1] < b.F) b.F = e[1];
                if (e[2] > b.R) b.R = e[2];
            }
        };
        std::vector<VirtBlock> left(K + 1), right(K + 1);
        const u32 c0 = C0(me);
        {
            u32 cur = c0 < (u32)K ? c0 : (u32)K;
            for (int g = prev_ne(me); g >= 0 && cur > 0; g = prev_ne(g)) {
                merge(left, g, 0, cur);
                cur = std::min(cur, std::min(minint(g), C0(g)));
            }
        }
        {
            int g = next_ne(me);
            u32 cur = g >= 0 ? std::min<u32>(C0(g), (u32)K) : 0u;
            while (g >= 0 && cur > 0) {
                merge(right, g, 1, cur);
                cur = std::min(cur, minint(g));
                g = next_ne(g);
                if (g >= 0) cur = std::min(cur, C0(g));
            }
        }
        // virtual ranks: per block one representative per class (min forward start / max rc value), or a null leaf
        u32* hv = d->h_pin;                  // [F left 64 | R left 64 | LCP left 64 | F right 64 | R right 64 | LCP right 64 + guard]
        u32 *fl = hv, *rl = hv + 64, *ll = hv + 128, *fr = hv + 192, *rr = hv + 256, *lr = hv + 320;
        for (int i = 0; i < 64; ++i) { fl[i] = NONE_MIN; rl[i] = 0; ll[i] = 0; fr[i] = NONE_MIN; rr[i] = 0; lr[i] = 0; }
        lr[64] = 0;
        auto reps_of = [&](const VirtBlock& b, u32 f2[2], u32 r2[2]) {
            int k = 0;
            if (b.F != NONE_MIN) { f2[k] = b.F; r2[k] = 0; ++k; }
            if (b.R != 0) { f2[k] = NONE_MIN; r2[k] = b.R; ++k; }
            if (!k) { f2[k] = NONE_MIN; r2[k] = 0; ++k; }
            return k;
        };
        u32 vl = 0;
        for (int eff = 1; eff <= K; ++eff) if (left[eff].cnt) { u32 f2[2], r2[2]; vl += (u32)reps_of(left[eff], f2, r2); }
        {
            u32 idx = DIST_VIRT - vl, prev_eff = 0;
            for (int eff = 1; eff <= K; ++eff) {
                if (!left[eff].cnt) continue;
                u32 f2[2], r2[2];
                const int k = reps_of(left[eff], f2, r2);
                for (int j = 0; j < k; ++j) { fl[idx] = f2[j]; rl[idx] = r2[j]; ll[idx] = j == 0 ? prev_eff : (u32)eff; ++idx; }
                prev_eff = (u32)eff;
            }
        }
        u32 vr = 0;
        for (int eff = K; eff >= 1; --eff) {
            if (!right[eff].cnt) continue;
            u32 f2[2], r2[2];
            const int k = reps_of(right[eff], f2, r2);
            for (int j = 0; j < k; ++j) { fr[vr] = f2[j]; rr[vr] = r2[j]; lr[vr] = (u32)eff; ++vr; }
        }
        lr[vr] = 0;                                          // right guard
        NLZ_CK(cudaMemcpyAsync(F0buf, fl, 64 * 4, cudaMemcpyHostToDevice, st));
        NLZ_CK(cudaMemcpyAsync(R0buf, rl, 64 * 4, cudaMemcpyHostToDevice, st));
        NLZ_CK(cudaMemcpyAsync(LCPbuf, ll, 64 * 4, cudaMemcpyHostToDevice, st));
        NLZ_CK(cudaMemcpyAsync(F0 + m_loc, fr, 64 * 4, cudaMemcpyHostToDevice, st));
        NLZ_CK(cudaMemcpyAsync(R0 + m_loc, rr, 64 * 4, cudaMemcpyHostToDevice, st));
        NLZ_CK(cudaMemcpyAsync(w.LCP + m_loc, lr, 65 * 4, cudaMemcpyHostToDevice, st));
        u32* c0_src = d->h_pin + 512;
        *c0_src = c0;
        NLZ_CK(cudaMemcpyAsync(w.LCP, c0_src, 4, cudaMemcpyHostToDevice, st));   // restore c0
        NLZ_CK(cudaStreamSynchronize(st));                                        // the pinned staging area is reused by the next barrier

        // ---- S3b: factor rule in rank space over [virtual | real | virtual]; results indexed by work item
        wp.n1 = DIST_VIRT + m_loc + vr;
        wp.real_lo = DIST_VIRT; wp.real_hi = DIST_VIRT + m_loc;
        NLZ_TRY(stage_lpnf(c, pb.rc, st, F0buf, R0buf, LCPbuf, wp, nullptr, LRl, FLl, true));
        nwork_bound = pb.rc ? (m_loc < pb.nfac ? m_loc : pb.nfac) : m_loc;
    }
    // results to the owners of their T-positions
    {
        LrItem it;
        it.LR = LRl; it.FLAGS = FLl; it.list = pb.rc ? w.SLOT[1] : nullptr; it.F0 = F0buf;
        it.real_lo = DIST_VIRT; it.nfac = pb.nfac; it.chunkT = r.chunkT;
        NLZ_TRY((d2_bucket<LrItem, true>(r, it, nwork_bound, pb.rc ? w.CTR + 5 : nullptr, ST, STB, 40)));
        XInfo x;
        NLZ_TRY(d2_counts(r, 0, 16, x, true));
        const size_t off_b = (size_t)d->inbox_items * 8;
        NLZ_TRY(d2_push(r, x, ST, 8, 0));
        NLZ_TRY(d2_push(r, x, STB, 8, off_b));
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));
        if (x.in_total != nT) { set_error("internal: rank %d received %u per-position results for %u positions", me, x.in_total, nT); return ERR_RUNTIME; }
        if (nT)
            KL(P, KC_WALK, (u64)nT * 25, st,
               (k_d2_apply_lr<<<ceil_div_u32(nT, 256), 256, 0, st>>>(r.inbox, reinterpret_cast<const u64*>(d->seg + d->off_inbox + off_b), nT, LRT, FLT)));
    }
    if (m_loc) account_walk(c);
    NLZ_CK(cudaEventRecord(c->ev[EV_LPNF], st));

    // ---- S4: chain over position slices; the exit-node doubling crosses slices through peer memory
    const u32 nchunks = ceil_div_u32(nT ? nT : 1, CH_CHUNK);
    u32* CNT = MASK + (size_t)nchunks * 32;
    u32* acount = w.CTR + 1;
    ChainDom dom;
    memset(&dom, 0, sizeof(dom));
    dom.t0 = r.t0; dom.t1 = r.t1; dom.nfac = pb.nfac; dom.chunk = r.chunkT; dom.G = G;
    for (int g = 0; g < G; ++g) {
        dom.J[g] = reinterpret_cast<u32*>(d->peer[g] + d->off_cj);
        dom.J2[g] = reinterpret_cast<u32*>(d->peer[g] + d->off_cj2);
        dom.REACH[g] = d->peer[g] + d->off_creach;
    }
    u32* EXIT = dom.J[me];
    u8* REACH = dom.REACH[me];
    P.begin(st);
    NLZ_CK(cudaMemsetAsync(REACH, 0, (size_t)nT + 8, st));
    k_chain_init<<<1, 1, 0, st>>>(alist, acount, REACH, 0u, r.t0, r.t1);
    if (nT) k_chain_exit<<<nchunks, CH_THREADS, 0, st>>>(LRT, r.t0, r.t1, pb.nfac, EXIT, alist, acount);
    P.end(KC_CHAIN, (u64)nT * 13, st, 2);
    const int rounds = bits_for(ceil_div_u32(pb.nfac, CH_CHUNK)) + 1;
    NLZ_TRY(d2_barrier(r, nullptr, 0, false));
    for (int q = 0; q < rounds; ++q) {
        const u32 grid = nchunks < (u32)kNumSM * 2 ? nchunks : kNumSM * 2;
        KL(P, KC_CHAIN, 0, st, (k_chain_double<<<grid, 256, 0, st>>>(alist, acount, dom, q & 1)));
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));
    }
    P.begin(st);
    if (nT) k_chain_mark<<<nchunks, CH_THREADS, 0, st>>>(LRT, nT, REACH, MASK, CNT);
    else NLZ_CK(cudaMemsetAsync(CNT, 0, 4, st));
    k_scan_u32_single_cta<<<1, 1024, 0, st>>>(CNT, nT ? nchunks : 0, w.CTR + 2);
    P.end(KC_CHAIN, (u64)nT * 10, st, 2);
    if (!nT) NLZ_CK(cudaMemsetAsync(w.CTR + 2, 0, 4, st));
    NLZ_CK(cudaMemcpyAsync(r.PAY, w.CTR + 2, 4, cudaMemcpyDeviceToDevice, st));
    NLZ_TRY(d2_barrier(r, r.PAY, 1, true));
    u64 z = 0, zoff = 0;
    std::vector<u64> zg(G);
    for (int g = 0; g < G; ++g) { zg[g] = r.all[g]; if (g < me) zoff += zg[g]; z += zg[g]; }
    c->stats.n_factors = z;
    const bool want = out_alloc != nullptr || (me != 0 && d->hb == nullptr && false);
    (void)want;
    // every rank emits the factors of its slice into its own inbox (free by now); rank 0 gathers them
    // (count-only calls skip the emission: rank 0 decides, the others learn it through the barrier payload)
    k_set_u32<<<1, 1, 0, st>>>(r.PAY, (me == 0 && (out_alloc || out_into)) ? 1u : 0u);
    NLZ_TRY(d2_barrier(r, r.PAY, 1, true));
    const bool emit = r.all[0] != 0;
    bool too_small = false;
    if (emit) {
        if (zg[me] * 24 > d->inbox_items * 16ull) { set_error("internal: %llu factors of one slice exceed the inbox", (unsigned long long)zg[me]); return ERR_RUNTIME; }
        BatchView bv;
        memset(&bv, 0, sizeof(bv));
        if (nT && zg[me]) {
            P.begin(st);
            if (pb.rc) k_chain_emit<true, false><<<nchunks, CH_THREADS, 0, st>>>(LRT, FLT, r.t0, MASK, CNT, r.inbox, zg[me], bv, nullptr);
            else k_chain_emit<false, false><<<nchunks, CH_THREADS, 0, st>>>(LRT, FLT, r.t0, MASK, CNT, r.inbox, zg[me], bv, nullptr);
            P.end(KC_CHAIN, (u64)nT / 8 + zg[me] * 32, st);
        }
        NLZ_CK(cudaEventRecord(c->ev[EV_CHAIN], st));
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));
        if (me == 0 && z) {
            u64* dst = out_into;
            if (!dst) {
                dst = static_cast<u64*>(malloc((size_t)z * 24));
                if (!dst) { set_error("out of host memory for %llu factors", (unsigned long long)z); return ERR_RUNTIME; }
                *out_alloc = dst;
            } else if (z > capacity) {
                too_small = true;        // reported after the closing barrier, so that the other ranks are not left waiting
            }
            u64 at = 0;
            for (int g = 0; g < G && !too_small; ++g) {
                if (zg[g]) NLZ_CK(cudaMemcpyAsync(dst + 3 * at, d->peer[g] + d->off_inbox, (size_t)zg[g] * 24, cudaMemcpyDefault, st));
                at += zg[g];
            }
            NLZ_CK(cudaStreamSynchronize(st));
        }
        NLZ_TRY(d2_barrier(r, nullptr, 0, false));          // rank 0 has read every inbox
        if (too_small) {
            set_error("output capacity %llu factors is too small for %llu factors", (unsigned long long)capacity, (unsigned long long)z);
            *out_count = z;
            NLZ_CK(cudaStreamSynchronize(st));
            return ERR_RUNTIME;
        }
    } else {
        NLZ_CK(cudaEventRecord(c->ev[EV_CHAIN], st));
    }
    NLZ_CK(cudaStreamSynchronize(st));
    NLZ_CK(cudaGetLastError());
    *out_count = z;
    return OK;
}

static int d2_make_problem(int mode, u64 n, D2Problem& pb, bool* empty) {
    *empty = false;
    pb.mode = mode; pb.n_in = n; pb.N = 0;
    u64 nfac = 0;
    if (mode == NLZ_MODE_GENERAL) {
        pb.rc = false; pb.L = n; nfac = n;
        if (n == 0) { *empty = true; return OK; }
    } else if (mode == NLZ_MODE_RC_PREPARED) {
        pb.rc = true; pb.L = n;
        if (n < 4) { *empty = true; return OK; }
        const u64 N = n / 2 - 1;
        if (N == 0) { *empty = true; return OK; }
        pb.N = N; nfac = N;
    } else if (mode == NLZ_MODE_DNA_RC) {
        pb.rc = true;
        if (n == 0) { *empty = true; return OK; }
        pb.L = 2 * n + 2; pb.N = n; nfac = n;
    } else {
        set_error("unknown mode %d", mode);
        return ERR_INVALID;
    }
    if (nfac >= 0xFFFFFFF0ull) {
        set_error("text of %llu factorized positions exceeds the 32-bit T-coordinates of this build", (unsigned long long)nfac);
        return ERR_RUNTIME;
    }
    pb.nfac = (u32)nfac;
    pb.n1 = pb.L + 1;
    return OK;
}

}  // namespace nlz

// =================================================================== C ABI of the distributed path
extern "C" {

int nlz_dist_create(nlz_ctx* c, int rank, int world, uint64_t max_text_bytes, int max_mode, nlz_dist** out) {
    if (!c || !out) { set_error("null argument"); return ERR_INVALID; }
    *out = nullptr;
    if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) { set_error("bad rank %d / world %d (at most %d ranks)", rank, world, MAX_PEERS); return ERR_INVALID; }
    const u64 max_n1 = (max_mode == NLZ_MODE_DNA_RC ? 2 * max_text_bytes + 2 : max_text_bytes) + 1;
    if (max_text_bytes >= 0xFFFFFFF0ull) { set_error("text of %llu bytes exceeds the 32-bit T-coordinates of this build", (unsigned long long)max_text_bytes); return ERR_RUNTIME; }
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    nlz_dist* d = new nlz_dist();
    d->ctx = c; d->rank = rank; d->world = world; d->max_n1 = max_n1; d->max_nfac = max_text_bytes;
    const u64 chunk = ((max_n1 + world - 1) / world + KB_TP - 1) / KB_TP * KB_TP;
    // room for a rank range of up to 4x the mean (a text whose suffixes crowd into a few 12-symbol buckets -- long
    // homopolymers, short-period tandem arrays -- cannot be balanced by ANY prefix partition: a tie group is one unit)
    d->inbox_items = 4 * chunk + 65536;
    d->max_chT = ((max_text_bytes + world - 1) / world + CH_CHUNK - 1) / CH_CHUNK * CH_CHUNK + CH_CHUNK;
    d->off_x = d2_al(sizeof(DistCtl));
    d->off_hist = d->off_x + d2_al(max_n1 + 512);
    d->off_cj = d->off_hist + d2_al(((size_t)(1u << 24) + 64) * 4);
    d->off_cj2 = d->off_cj + d2_al((d->max_chT + 64) * 4);
    d->off_creach = d->off_cj2 + d2_al((d->max_chT + 64) * 4);
    d->off_inbox = d->off_creach + d2_al(d->max_chT + 64);
    d->seg_bytes = d->off_inbox + d2_al(d->inbox_items * 16);
    cudaError_t e = cudaMalloc(&d->seg, d->seg_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&d->SMALL, 2048 * 4);
    if (e == cudaSuccess) e = cudaMallocHost(&d->h_pin, ((size_t)MAX_PEERS * DIST_XCH_WORDS + 64) * 4);
    if (e == cudaSuccess) e = cudaMemset(d->seg, 0, d->off_x);
    if (e == cudaSuccess) e = cudaMemset(d->SMALL, 0, 2048 * 4);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error("allocating the distributed workspace (%zu bytes shared) failed: %s", d->seg_bytes, cudaGetErrorString(e));
        cudaGetLastError();
        nlz_dist_destroy(d);
        return ERR_CUDA;
    }
    d->peer[rank] = d->seg;
    *out = d;
    return OK;
}

void nlz_dist_destroy(nlz_dist* d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaDeviceSynchronize();
    for (int g = 0; g < MAX_PEERS; ++g) if (d->ipc_opened[g]) cudaIpcCloseMemHandle(d->peer[g]);
    if (d->seg) cudaFree(d->seg);
    if (d->SMALL) cudaFree(d->SMALL);
    if (d->arena.base) cudaFree(d->arena.base);
    if (d->h_pin) cudaFreeHost(d->h_pin);
    if (d->owns_hb) delete d->hb;
    cudaGetLastError();
    delete d;
}

int nlz_dist_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int nlz_dist_export(nlz_dist* d, uint8_t* handle_out) {
    if (!d || !handle_out) { set_error("null argument"); return ERR_INVALID; }
    NLZ_CK(cudaSetDevice(d->ctx->device));
    cudaIpcMemHandle_t h;
    NLZ_CK(cudaIpcGetMemHandle(&h, d->seg));
    memcpy(handle_out, &h, sizeof(h));
    return OK;
}

int nlz_dist_attach(nlz_dist* d, const uint8_t* all_handles) {
    if (!d || !all_handles) { set_error("null argument"); return ERR_INVALID; }
    NLZ_CK(cudaSetDevice(d->ctx->device));
    for (int g = 0; g < d->world; ++g) {
        if (g == d->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)g * sizeof(h), sizeof(h));
        void* p = nullptr;
        NLZ_CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        d->peer[g] = static_cast<u8*>(p);
        d->ipc_opened[g] = true;
    }
    d->attached = true;
    return OK;
}

int nlz_dist_attach_local(nlz_dist* const* ranks, int world) {
    if (!ranks || world < 1 || world > MAX_PEERS) { set_error("bad local group"); return ERR_INVALID; }
    HostBarrier* hb = new HostBarrier();
    hb->world = world;
    for (int a = 0; a < world; ++a) {
        nlz_dist* d = ranks[a];
        if (!d || d->rank != a || d->world != world) { set_error("local group: rank %d is missing or misnumbered", a); delete hb; return ERR_INVALID; }
        NLZ_CK(cudaSetDevice(d->ctx->device));
        for (int g = 0; g < world; ++g) {
            d->peer[g] = ranks[g]->seg;
            const int dev = ranks[g]->ctx->device;
            if (dev != d->ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    set_error("cannot enable peer access %d -> %d: %s", d->ctx->device, dev, cudaGetErrorString(e));
                    delete hb;
                    return ERR_CUDA;
                }
                cudaGetLastError();
            }
        }
        d->hb = hb;
        d->owns_hb = a == 0;
        d->attached = true;
    }
    return OK;
}

static int dist_factorize_impl(nlz_dist* d, int mode, const uint8_t* text, uint64_t n, uint64_t** out_triples, uint64_t* out_into,
                               uint64_t capacity, uint64_t* out_count) {
    if (!d || !out_count) { set_error("null argument"); return ERR_INVALID; }
    if (n && !text) { set_error("null text"); return ERR_INVALID; }
    if (!d->attached && d->world > 1) { set_error("distributed group is not attached"); return ERR_INVALID; }
    nlz_ctx* c = d->ctx;
    std::lock_guard<std::mutex> lock(c->mu);
    NLZ_CK(cudaSetDevice(c->device));
    *out_count = 0;
    if (out_triples) *out_triples = nullptr;
    D2Problem pb;
    bool empty = false;
    NLZ_TRY(d2_make_problem(mode, n, pb, &empty));
    reset_stats(c);
    if (empty) return OK;
    if (pb.n1 > d->max_n1 || pb.nfac > d->max_nfac) { set_error("text of %llu suffixes exceeds the group's capacity of %llu", (unsigned long long)pb.n1, (unsigned long long)d->max_n1); return ERR_INVALID; }
    u64 z = 0;
    NLZ_TRY(run_dist2(d, pb, text, d->rank == 0 ? out_triples : nullptr, d->rank == 0 ? out_into : nullptr, capacity, &z));
    Problem pstat{};
    pstat.n_in = pb.n_in; pstat.n1 = (u32)std::min<u64>(pb.n1, 0xFFFFFFFFull); pstat.nfac = pb.nfac;
    finish_stats(c, pstat);
    c->stats.n_suffixes = pb.n1;
    c->stats.n_local_suffixes = d->last_m_loc;
    *out_count = z;
    return OK;
}

int nlz_dist_factorize(nlz_dist* d, int mode, const uint8_t* text, uint64_t n, uint64_t** out_triples, uint64_t* out_count) {
    return dist_factorize_impl(d, mode, text, n, out_triples, nullptr, 0, out_count);
}
int nlz_dist_factorize_into(nlz_dist* d, int mode, const void* text, uint64_t n, uint64_t* out_triples, uint64_t capacity,
                            uint64_t* out_count) {
    return dist_factorize_impl(d, mode, static_cast<const uint8_t*>(text), n, nullptr, out_triples, capacity, out_count);
}

}  // extern "C"
