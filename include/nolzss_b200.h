/*
 * nolzss_b200 -- C ABI of the B200-native (sm_100a) non-overlapping LZSS factorizer.
 *
 * This is the drop-in boundary for the hot path of OmerKerner/noLZSS: every entry point below
 * replaces one C++ function that the reference's pybind11 module (src/cpp/bindings.cpp) binds,
 * with plain pointers and sizes (no C++ types, no torch types, no exceptions across the ABI).
 * Citations are relative to /root/reference.
 *
 * Conventions
 *   - status codes: NLZ_OK, or an error; nlz_last_error() returns the message of the last failure
 *     on the calling thread.  NLZ_ERR_RUNTIME corresponds to std::runtime_error (Python
 *     RuntimeError), NLZ_ERR_INVALID to std::invalid_argument (Python ValueError), NLZ_ERR_CUDA to
 *     a CUDA failure (no device / launch error / out of memory) -- there is NO CPU fallback.
 *   - factors are written as the reference's `struct Factor {u64 start, length, ref}`
 *     (src/cpp/factorizer.hpp:147-151), 24 bytes each, host endian; reverse-complement factors
 *     carry NLZ_RC_MASK in `ref` (src/cpp/factorizer.hpp:41).
 *   - buffers returned through `uint64_t** out` are owned by the caller and released with
 *     nlz_free(); inputs are borrowed for the duration of the call.
 *   - a context owns one CUDA device, its workspace in HBM and its statistics; calls on one
 *     context are serialised internally, different contexts may be used concurrently.
 */
#ifndef NOLZSS_B200_H
#define NOLZSS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLZ_OK 0
#define NLZ_ERR_RUNTIME 1
#define NLZ_ERR_INVALID 2
#define NLZ_ERR_CUDA 3

#define NLZ_RC_MASK (1ULL << 63)

/* which core algorithm a generic entry point runs */
#define NLZ_MODE_GENERAL 0     /* detail::nolzss                     factorizer_core.hpp:51-119  */
#define NLZ_MODE_RC_PREPARED 1 /* detail::nolzss_multiple_dna_w_rc   factorizer_core.hpp:177-383 */
#define NLZ_MODE_DNA_RC 2      /* detail::nolzss_dna_w_rc            factorizer_core.hpp:140-151 */

typedef struct nlz_ctx nlz_ctx;

typedef struct nlz_stats {
    uint64_t n_text;        /* bytes handed in */
    uint64_t n_suffixes;    /* indexed suffixes n' (text + terminator; 2n+3 in DNA_RC mode) */
    uint64_t n_factorized;  /* positions whose factor rule was evaluated */
    uint64_t n_factors;     /* z */
    uint64_t active_sum;    /* sum over doubling rounds of suffixes still being sorted */
    uint64_t walk_nodes;    /* suffix-tree path nodes / depth probes evaluated by the per-position rule */
    uint64_t hard_positions;/* positions resolved by the text-order depth search (deep nestings) */
    uint64_t workspace_bytes;
    uint32_t key_bits, sym_bits, key_syms;
    uint32_t doubling_rounds;
    uint32_t tile_sort_rounds; /* doubling rounds sorted entirely in shared memory */
    uint32_t kernel_launches;
    uint32_t host_syncs;
    /* device time per stage (CUDA events on the call's stream), milliseconds */
    float ms_total, ms_prepare, ms_keys, ms_sort0, ms_doubling, ms_lcp, ms_lpnf, ms_chain;
} nlz_stats;

/* ---- context ---------------------------------------------------------------------------- */
int nlz_ctx_create(int device, nlz_ctx** out);
void nlz_ctx_destroy(nlz_ctx* ctx);
const char* nlz_last_error(void);
void nlz_free(void* p);
int nlz_get_stats(nlz_ctx* ctx, nlz_stats* out);
const char* nlz_version(void); /* bindings.cpp:1513-1517 (__version__) */
/* per-kernel-class accounting of the last call: launches and algorithmic bytes always, device time
 * (CUDA events around every launch, on the launching stream) when profiling is switched on */
int nlz_set_profiling(nlz_ctx* ctx, int on);
int nlz_kernel_class_count(void);
int nlz_get_kernel_stats(nlz_ctx* ctx, int cls, const char** name, double* ms, uint64_t* bytes,
                         uint32_t* launches);

/* ---- generic entry points (HOST buffers; H2D/D2H copies happen inside the call) ---------- */
int nlz_factorize_mode(nlz_ctx* ctx, int mode, const uint8_t* text, uint64_t n, uint64_t start_pos,
                       uint64_t** out_triples, uint64_t* out_count);
/* as above, into a caller-provided (ideally pinned) buffer of `capacity` factors; fails with
 * NLZ_ERR_RUNTIME when capacity is too small (*out_count still holds the true count) */
int nlz_factorize_mode_into(nlz_ctx* ctx, int mode, const uint8_t* text, uint64_t n, uint64_t start_pos,
                            uint64_t* out_triples, uint64_t capacity, uint64_t* out_count);
int nlz_count_mode(nlz_ctx* ctx, int mode, const uint8_t* text, uint64_t n, uint64_t start_pos,
                   uint64_t* out_count);

/* ---- device-resident entry point (text and triples already in HBM, on `cuda_stream`) ----- */
int nlz_factorize_device(nlz_ctx* ctx, int mode, const void* d_text, uint64_t n, uint64_t start_pos,
                         void* cuda_stream, void* d_out_triples, uint64_t capacity, uint64_t* out_count);

/* ---- named entry points: one per reference function on the path ------------------------- */
/* noLZSS::factorize(std::string_view, start_pos)                 src/cpp/factorizer.cpp:378-384 */
int nlz_factorize(nlz_ctx* ctx, const uint8_t* text, uint64_t n, uint64_t start_pos,
                  uint64_t** out_triples, uint64_t* out_count);
/* noLZSS::count_factors(std::string_view, start_pos)             src/cpp/factorizer.cpp:337-343 */
int nlz_count_factors(nlz_ctx* ctx, const uint8_t* text, uint64_t n, uint64_t start_pos, uint64_t* out_count);
/* noLZSS::factorize_dna_w_rc(std::string_view)                   src/cpp/factorizer.cpp:519-523 */
int nlz_factorize_dna_w_rc(nlz_ctx* ctx, const uint8_t* text, uint64_t n, uint64_t** out_triples,
                           uint64_t* out_count);
/* noLZSS::count_factors_dna_w_rc(std::string_view)               src/cpp/factorizer.cpp:530-532 */
int nlz_count_factors_dna_w_rc(nlz_ctx* ctx, const uint8_t* text, uint64_t n, uint64_t* out_count);
/* noLZSS::factorize_multiple_dna_w_rc(std::string_view, start)   src/cpp/factorizer.cpp:651-657 */
int nlz_factorize_multiple_dna_w_rc(nlz_ctx* ctx, const uint8_t* prepared, uint64_t n, uint64_t start_pos,
                                    uint64_t** out_triples, uint64_t* out_count);
/* noLZSS::count_factors_multiple_dna_w_rc(std::string_view, st)  src/cpp/factorizer.cpp:676-681 */
int nlz_count_factors_multiple_dna_w_rc(nlz_ctx* ctx, const uint8_t* prepared, uint64_t n, uint64_t start_pos,
                                        uint64_t* out_count);

/* ---- stage probes used by the parity tests (device results copied to host arrays) -------- */
/* suffix array (n+1 entries, terminator included), inverse, and LCP (n+2 entries, LCP[0]=LCP[n+1]=0)
 * of text·$ under this library's symbol order (see csrc/sa.cuh).  Any output may be NULL. */
int nlz_debug_index(nlz_ctx* ctx, const uint8_t* text, uint64_t n, uint32_t* sa, uint32_t* isa, uint32_t* lcp);
/* stable LSD radix sort of (key, value) pairs on bits [lo_bit, hi_bit) */
int nlz_debug_sort_pairs_u64(nlz_ctx* ctx, uint64_t* keys, uint32_t* vals, uint64_t m, int lo_bit, int hi_bit);
int nlz_debug_sort_pairs_u32(nlz_ctx* ctx, uint32_t* keys, uint32_t* vals, uint64_t m, int lo_bit, int hi_bit);
/* per-position rule before chain extraction: len[i], ref[i] (RC flag in bit 63) for i in [0, nfac) */
int nlz_debug_per_position(nlz_ctx* ctx, int mode, const uint8_t* text, uint64_t n, uint64_t* len_out,
                           uint64_t* ref_out, uint64_t capacity, uint64_t* nfac_out);

#ifdef __cplusplus
}
#endif
#endif /* NOLZSS_B200_H */
