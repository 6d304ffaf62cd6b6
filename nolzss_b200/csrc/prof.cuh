// Per-kernel-class accounting: launches and algorithmic bytes always; CUDA-event timing of every
// launch (on the launching stream) only when profiling is switched on (bench.py's roofline leg).
#pragma once
#include <vector>

#include "common.cuh"

namespace nlz {

enum KClass : int {
    KC_PREPARE = 0,   // text staging, rc(T) construction, byte histogram
    KC_KEYS,          // packed-prefix key construction
    KC_RS_HIST,       // radix sort: per-CTA digit histograms
    KC_RS_SCAN,       // radix sort: histogram scan (one CTA)
    KC_RS_SCATTER,    // radix sort: ranked, smem-staged scatter
    KC_GATHER,        // doubling key: RANK[s+h] gather
    KC_TILE_SORT,     // doubling: fused gather + shared-memory segmented sort
    KC_REGROUP,       // head flags, ranks, SA write-back, compaction
    KC_LCP,           // chunked Kasai LCP
    KC_TREE,          // 32-ary summary trees
    KC_NODES,         // per-node tables: LCP interval named by every rank + its minimum forward start
    KC_RNEAR,         // nearest rc(T) rank on either side of every rank (segmented scans)
    KC_WALK,          // factor rule, rank order (bounded climb + direct RC candidate)
    KC_WALK_HARD,     // factor rule, text order over deep-nesting positions (depth search + carry)
    KC_CHAIN,         // chain extraction (exit, doubling, mark, scan, emit)
    KC_BARRIER,       // distributed runs: flag barrier in peer memory (time = waiting for the slowest GPU)
    KC_STREAM,        // doubling: tie groups beyond the tile, one CTA streaming each (pivot partition + outlier sort)
    KC_XCHG,          // distributed runs: bucketing by destination, bulk copies over NVLink, local apply
    KC_COUNT
};

static const char* const kClassNames[KC_COUNT] = {
    "prepare", "build_keys", "radix_hist", "radix_scan", "radix_scatter", "gather_rank",
    "tile_sort", "regroup", "lcp_kasai", "summary_trees", "node_tables", "rc_neighbours", "lpnf_rank", "lpnf_hard", "chain", "dist_barrier", "group_stream", "dist_exchange"};

struct Profiler {
    bool timing = false;
    u32 launches[KC_COUNT];
    u64 bytes[KC_COUNT];
    double ms[KC_COUNT];
    std::vector<cudaEvent_t> pool;
    size_t used = 0, cur = 0;
    struct Span { int cls; size_t e0, e1; };
    std::vector<Span> spans;

    void reset() {
        for (int i = 0; i < KC_COUNT; ++i) { launches[i] = 0; bytes[i] = 0; ms[i] = 0.0; }
        used = 0;
        spans.clear();
    }
    size_t take() {
        if (used == pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            pool.push_back(e);
        }
        return used++;
    }
    void begin(cudaStream_t st) {
        if (timing) { cur = take(); cudaEventRecord(pool[cur], st); }
    }
    void end(int cls, u64 by, cudaStream_t st, u32 nlaunch = 1) {
        launches[cls] += nlaunch;
        bytes[cls] += by;
        if (timing) {
            size_t e1 = take();
            cudaEventRecord(pool[e1], st);
            spans.push_back({cls, cur, e1});
        }
    }
    void collect() {   // call after the stream is synchronised
        for (const Span& s : spans) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, pool[s.e0], pool[s.e1]) == cudaSuccess) ms[s.cls] += t;
            else cudaGetLastError();
        }
        spans.clear();
        used = 0;
    }
    u32 total_launches() const { u32 t = 0; for (int i = 0; i < KC_COUNT; ++i) t += launches[i]; return t; }
    void destroy() { for (cudaEvent_t e : pool) cudaEventDestroy(e); pool.clear(); }
};

// KL(profiler, class, algorithmic_bytes, stream, launch-statement)
#define KL(P, cls, by, st, ...)   \
    do {                          \
        (P).begin(st);            \
        __VA_ARGS__;              \
        (P).end(cls, by, st);     \
    } while (0)

}  // namespace nlz
