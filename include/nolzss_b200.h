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
    /* distributed runs: (suffix, rank) records this GPU received from the others and applied to its replica */
    uint64_t rank_records_applied;
    /* suffixes still tied after the initial key sort (members of tie groups): an upper bound of the positions whose LCP
     * the Kasai kernel computes from the text -- every other LCP value follows from a pair of sort keys */
    uint64_t lcp_marked;
    /* distributed runs: suffixes of this GPU's rank range (the partition's balance: max over ranks / mean) */
    uint64_t n_local_suffixes;
} nlz_stats;

/* ---- context ---------------------------------------------------------------------------- */
int nlz_ctx_create(int device, nlz_ctx** out);
void nlz_ctx_destroy(nlz_ctx* ctx);
int nlz_ctx_device(nlz_ctx* ctx);
const char* nlz_last_error(void);
void nlz_free(void* p);
int nlz_get_stats(nlz_ctx* ctx, nlz_stats* out);
const char* nlz_version(void); /* bindings.cpp:1513-1517 (__version__) */
/* page-locks / releases a caller-owned host range (e.g. the slice of a memory-mapped genome a rank uploads), so that
 * the H2D copy inside an entry point runs at PCIe speed instead of through a pageable staging buffer */
int nlz_host_register(void* p, uint64_t bytes);
int nlz_host_unregister(void* p);
/* per-kernel-class accounting of the last call: launches and algorithmic bytes always, device time
 * (CUDA events around every launch, on the launching stream) when profiling is switched on */
int nlz_set_profiling(nlz_ctx* ctx, int on);
int nlz_kernel_class_count(void);
/* test hook: forces the fallback paths of the shared-memory group sort (1 bitonic network, 2 no pivot fast path,
 * 4 counting only); 8 makes the group-stream kernel of the hybrid doubling rounds give up from the third round
 * on (the round is redone through the radix path); flags >> 8, when in [64, 2048), lowers the largest tie group
 * the tile sort takes, so that small texts run the hybrid rounds; 0x1000000 (distributed path) adds 2^32 + 12345 to every
 * global rank, so that small texts carry 33-bit ranks through the doubling keys, the sort kernels and the exchanges --
 * the widths a 3.1 Gbp text needs; 0 = normal operation */
int nlz_set_debug_flags(nlz_ctx* ctx, int flags);
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

/* ---- batch of independent records in ONE pipeline run (segmented suffix array: the record id is the leading
 * sort-key field) -- the per-record loops of factorize_fasta_dna_{w,no}_rc_per_sequence and their count_/write_
 * variants, src/cpp/fasta_processor.cpp:428-476, :505-561, src/cpp/parallel_fasta_processor.cpp:360-465.
 * Record b is concat[offsets[b], offsets[b] + lens[b]); with_rc != 0 runs nolzss_dna_w_rc on every record, else
 * nolzss.  Factors come back concatenated in record order, in record-local coordinates; per_record_counts
 * (k entries, caller-provided) receives the number of factors of each record.  out_triples == NULL: count only. */
int nlz_factorize_batch(nlz_ctx* ctx, int with_rc, const uint8_t* concat, const uint64_t* offsets, const uint64_t* lens,
                        uint64_t k, uint64_t** out_triples, uint64_t* per_record_counts, uint64_t* total);

/* ---- ONE text across the GPUs of a box (replaces the single serial index build of the reference's parallel mode,
 * src/cpp/parallel_factorizer.cpp:78-84, and its 64-bit index vectors, src/cpp/factorizer_core.hpp:195-232; see
 * csrc/dist2.cuh).  One nlz_dist per rank (GPU).  Every rank calls nlz_dist_factorize with the same text; rank 0
 * receives the factors, every rank the count.  Texts of up to 2^32 - 16 bases are taken (6.2 * 10^9 indexed suffixes for
 * a 3.1 Gbp genome in DNA_RC mode: global ranks and S-positions are 33-bit on this path); a GPU can own at most 2^30
 * suffixes, so such a text needs 8 GPUs.  The ranks exchange data through peer memory: either all ranks live in one
 * process (nlz_dist_attach_local) or one process per GPU exchanges the CUDA IPC handles of the shared segments
 * (nlz_dist_export / nlz_dist_attach; e.g. all-gathered with torch.distributed).  max_text_bytes / max_mode size the
 * shared segment (mode as in nlz_factorize_mode). */
typedef struct nlz_dist nlz_dist;
int nlz_dist_create(nlz_ctx* ctx, int rank, int world, uint64_t max_text_bytes, int max_mode, nlz_dist** out);
void nlz_dist_destroy(nlz_dist* d);
int nlz_dist_ipc_handle_bytes(void);
int nlz_dist_export(nlz_dist* d, uint8_t* handle_out);
int nlz_dist_attach(nlz_dist* d, const uint8_t* all_handles /* world x nlz_dist_ipc_handle_bytes() */);
int nlz_dist_attach_local(nlz_dist* const* ranks, int world);
int nlz_dist_factorize(nlz_dist* d, int mode, const uint8_t* text, uint64_t n, uint64_t** out_triples,
                       uint64_t* out_count);
/* as above; `text` may be a host pointer (pageable or pinned) or a DEVICE pointer (every rank holds the text in its own
 * HBM), and rank 0 receives the factors in a caller-provided (ideally pinned) buffer of `capacity` factors
 * (out_triples == NULL: count only) */
int nlz_dist_factorize_into(nlz_dist* d, int mode, const void* text, uint64_t n, uint64_t* out_triples,
                            uint64_t capacity, uint64_t* out_count);

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

/* ---- host side of the path: text preparation, FASTA, noLZSSv2 factor files ----------------- */
/* prepare_multiple_dna_sequences_w_rc / _no_rc            src/cpp/factorizer.cpp:54-172, :194-294
 * (S = T1 s0 .. Tk s(k-1) rc(Tk) sk .. rc(T1) s(2k-1); no_rc: sentinels only between records).
 * Outputs are malloc'ed (release with nlz_free). */
int nlz_prepare_multiple_dna_sequences_w_rc(const uint8_t* const* seqs, const uint64_t* lens, uint64_t k,
                                            uint8_t** prepared, uint64_t* prepared_len, uint64_t* original_length,
                                            uint64_t** sentinel_positions, uint64_t* n_sentinels);
int nlz_prepare_multiple_dna_sequences_no_rc(const uint8_t* const* seqs, const uint64_t* lens, uint64_t k,
                                             uint8_t** prepared, uint64_t* prepared_len, uint64_t* original_length,
                                             uint64_t** sentinel_positions, uint64_t* n_sentinels);

/* parse_fasta_sequences_and_ids                            src/cpp/fasta_processor.cpp:28-128
 * sanitize_mode: 0 = "remove_ambiguous", 1 = "strict" (fasta_processor.hpp:10-13, bindings.cpp:29-37) */
typedef struct nlz_fasta nlz_fasta;
int nlz_fasta_parse(const char* path, int sanitize_mode, nlz_fasta** out);
uint64_t nlz_fasta_num_sequences(const nlz_fasta* fa);
const char* nlz_fasta_id(const nlz_fasta* fa, uint64_t idx);
const uint8_t* nlz_fasta_sequence(const nlz_fasta* fa, uint64_t idx, uint64_t* len);
void nlz_fasta_free(nlz_fasta* fa);

/* identify_sentinel_factors                                src/cpp/fasta_processor.cpp:131-163 */
int nlz_identify_sentinel_factors(const uint64_t* triples, uint64_t count, const uint64_t* sentinel_positions,
                                  uint64_t n_positions, uint64_t** out_idx, uint64_t* out_n);
/* [factors][meta][FactorFileFooter] with footer_size = 48 + meta_len   src/cpp/factorizer.hpp:64-77 */
int nlz_write_factor_file(const char* out_path, const uint64_t* triples, uint64_t count, const uint8_t* meta,
                          uint64_t meta_len, uint64_t num_sequences, uint64_t num_sentinels, uint64_t total_length);

/* factorize_file / count_factors_file (+ _dna_w_rc, _multiple_dna_w_rc)  src/cpp/factorizer.cpp:359-363, :401-406,
 * :525-535, :659-690: the raw file bytes are the text */
int nlz_factorize_file_mode(nlz_ctx* ctx, int mode, const char* path, uint64_t start_pos, uint64_t** out_triples,
                            uint64_t* out_count);
int nlz_count_file_mode(nlz_ctx* ctx, int mode, const char* path, uint64_t start_pos, uint64_t* out_count);
/* write_factors_binary_file (mode GENERAL), _dna_w_rc (DNA_RC), _multiple_dna_w_rc (RC_PREPARED)
 *                                                          src/cpp/factorizer.cpp:424-459, :597-635, :751-790 */
int nlz_write_factors_binary_file_mode(nlz_ctx* ctx, int mode, const char* in_path, const char* out_path,
                                       uint64_t start_pos, uint64_t* out_count);
/* parallel_factorize_to_file / parallel_factorize_dna_w_rc_to_file  src/cpp/parallel_factorizer.cpp:55-144, :1001-1017
 * (footer: 0 sequences, 0 sentinels, total_length = sum of factor lengths, :754-767) */
int nlz_parallel_factorize_to_file(nlz_ctx* ctx, int mode, const uint8_t* text, uint64_t n, const char* out_path,
                                   uint64_t start_pos, uint64_t* out_count);
/* factorize_dna_w_reference_seq(_file) (dna = 1) / factorize_w_reference(_file) (dna = 0)
 *                                                          src/cpp/factorizer.cpp:825-1021; out_path/out may be NULL */
int nlz_factorize_w_reference(nlz_ctx* ctx, int dna, const uint8_t* ref, uint64_t ref_len, const uint8_t* tgt,
                              uint64_t tgt_len, const char* out_path, uint64_t** out_triples, uint64_t* out_count);
/* factorize_fasta_multiple_dna_{w,no}_rc, factorize_dna_rc_w_ref_fasta_files (ref_fasta != NULL) and the
 * (parallel_)write_factors_binary_file_fasta_multiple_dna_{w,no}_rc /
 * write_factors_dna_w_reference_fasta_files_to_binary writers (out_path != NULL)
 *                                 src/cpp/fasta_processor.cpp:298-423, src/cpp/parallel_fasta_processor.cpp:29-257 */
int nlz_factorize_fasta(nlz_ctx* ctx, const char* ref_fasta, const char* fasta_path, int with_rc, int sanitize_mode,
                        const char* out_path, uint64_t** out_triples, uint64_t* out_count, uint64_t** sentinel_idx,
                        uint64_t* n_sentinel_idx, nlz_fasta** ids_out);
/* factorize_/count_factors_/write_factors_binary_file_fasta_dna_{w,no}_rc_per_sequence (+ parallel_ writers)
 *                                 src/cpp/fasta_processor.cpp:428-561, src/cpp/parallel_fasta_processor.cpp:268-465 */
/* num_threads host threads (0 = min(records, 8)) pull records from a queue, each with its own stream and workspace on
 * ctx's device -- the reference's worker pool over an atomic record index (parallel_fasta_processor.cpp:360-385) */
int nlz_factorize_fasta_per_sequence(nlz_ctx* ctx, const char* fasta_path, int with_rc, int sanitize_mode,
                                     const char* out_dir, int want_factors, int num_threads, uint64_t** out_triples,
                                     uint64_t** per_seq_counts, uint64_t* total_count, nlz_fasta** ids_out);

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
