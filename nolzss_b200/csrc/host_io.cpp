// Host side of the noLZSS path: text preparation, FASTA parsing, noLZSSv2 factor files and the
// file-level entry points that combine them with the GPU factorizer (include/nolzss_b200.h).
//
// Behaviour (sentinel bytes, limits, warnings, error strings, the eight footer variants) follows
// /root/reference/src/cpp/factorizer.cpp, fasta_processor.cpp, parallel_fasta_processor.cpp and
// parallel_factorizer.cpp; each function cites the lines it mirrors.  No factor is ever computed
// on the host: every composite ends in nlz_factorize_mode / nlz_count_mode (CUDA).
#include <cctype>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include "../../include/nolzss_b200.h"

namespace nlz {
void set_error(const char* fmt, ...);   // api.cu
}
extern "C" const char* nlz_last_error(void);
using nlz::set_error;

namespace {

struct Footer {            // factorizer.hpp:64-77 (48 bytes, written raw, little endian hosts)
    char magic[8];
    uint64_t num_factors, num_sequences, num_sentinels, footer_size, total_length;
};
static_assert(sizeof(Footer) == 48, "footer layout");

// factorizer.cpp:110-125 -- index-th byte of 1..255 (wrapping past 255 back to 1) that is not A/C/G/T
uint8_t sentinel_byte(size_t index) {
    uint8_t s = 1;
    size_t count = 0;
    for (;;) {
        if (s != 0 && s != 'A' && s != 'C' && s != 'G' && s != 'T') {
            if (count == index) return s;
            ++count;
        }
        ++s;
        if (s == 0) s = 1;
    }
}

inline bool is_acgt_any_case(uint8_t c) {
    return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't';
}
inline uint8_t upper_ascii(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 'a' + 'A') : c; }
inline uint8_t complement(uint8_t c) {
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; default: return 'A'; }
}

struct Prepared {
    std::string s;
    uint64_t original_length = 0;
    std::vector<uint64_t> sentinels;
};

// factorizer.cpp:54-172 (with_rc) and :194-294 (no_rc)
int prepare(const std::vector<std::string>& seqs, bool with_rc, Prepared& out) {
    out = Prepared();
    if (seqs.empty()) return NLZ_OK;                                   // :55-57
    size_t non_empty = 0, empty = 0;
    for (const auto& q : seqs) (q.empty() ? empty : non_empty)++;
    if (empty > 0)                                                     // :71-73
        std::cerr << "Warning: Skipping " << empty << " empty sequence(s) in prepare_multiple_dna_sequences_"
                  << (with_rc ? "w_rc" : "no_rc") << std::endl;
    if (non_empty == 0) {                                              // :76-78
        set_error("All sequences are empty - cannot prepare for factorization");
        return NLZ_ERR_RUNTIME;
    }
    const size_t limit = with_rc ? 125 : 250;                          // :81-83 / :221-223
    if (non_empty > limit) {
        set_error("Too many sequences: maximum %zu sequences supported (due to sentinel character limitations)", limit);
        return NLZ_ERR_INVALID;
    }
    for (size_t i = 0; i < seqs.size(); ++i)                           // :86-95
        for (unsigned char c : seqs[i])
            if (!is_acgt_any_case(c)) {
                set_error("Invalid nucleotide '%c' found in sequence %zu", (char)c, i);
                return NLZ_ERR_RUNTIME;
            }
    size_t total = 0;
    for (const auto& q : seqs) total += q.size();
    out.s.reserve(with_rc ? 2 * (total + non_empty) : total + non_empty);
    size_t sidx = 0, processed = 0;
    for (const auto& q : seqs) {                                       // :128-145 / :270-289
        if (q.empty()) continue;
        for (unsigned char c : q) out.s.push_back((char)upper_ascii(c));
        ++processed;
        if (with_rc || processed < non_empty) {                        // no_rc: sentinels only BETWEEN records
            out.sentinels.push_back(out.s.size());
            out.s.push_back((char)sentinel_byte(sidx++));
        }
    }
    out.original_length = out.s.size();                                // :147 / :291
    if (with_rc) {
        for (size_t k = seqs.size(); k-- > 0;) {                       // :150-169 reverse record order
            const auto& q = seqs[k];
            if (q.empty()) continue;
            for (size_t j = q.size(); j-- > 0;) out.s.push_back((char)complement(upper_ascii((unsigned char)q[j])));
            out.sentinels.push_back(out.s.size());
            out.s.push_back((char)sentinel_byte(sidx++));
        }
    }
    return NLZ_OK;
}

int read_whole_file(const char* path, const char* what, std::string& data) {
    std::ifstream is(path, std::ios::binary);
    if (!is) { set_error("%s: %s", what, path); return NLZ_ERR_RUNTIME; }
    data.assign((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
    return NLZ_OK;
}

// [factors][meta][footer]; footer_size = 48 + |meta|
// The triples are one contiguous array, so the file is written with writev (factors | metadata + footer) instead of
// a stdio stream: the reference's 1 MiB stream buffer (factorizer.cpp:439) would have to be allocated and zeroed for
// every one of the 10 000 files of a per-sequence run.
int write_factor_file(const char* out_path, const uint64_t* triples, uint64_t count, const std::string& meta,
                      uint64_t num_sequences, uint64_t num_sentinels, uint64_t total_length) {
    const int fd = open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) { set_error("Cannot create output file: %s", out_path); return NLZ_ERR_RUNTIME; }
    Footer ft;
    memcpy(ft.magic, "noLZSSv2", 8);
    ft.num_factors = count;
    ft.num_sequences = num_sequences;
    ft.num_sentinels = num_sentinels;
    ft.footer_size = sizeof(Footer) + meta.size();
    ft.total_length = total_length;
    std::string tail = meta;
    tail.append(reinterpret_cast<const char*>(&ft), sizeof(ft));
    struct iovec iov[2];
    iov[0].iov_base = const_cast<uint64_t*>(triples);
    iov[0].iov_len = (size_t)count * 24;
    iov[1].iov_base = const_cast<char*>(tail.data());
    iov[1].iov_len = tail.size();
    bool ok = true;
    int first = count ? 0 : 1;
    while (first < 2) {
        // at most 1 GiB per call (Linux caps one writev at 2 GiB - 4 KiB)
        struct iovec cur[2];
        int n = 0;
        size_t budget = (size_t)1 << 30;
        for (int k = first; k < 2 && budget; ++k) {
            cur[n] = iov[k];
            if (cur[n].iov_len > budget) cur[n].iov_len = budget;
            budget -= cur[n].iov_len;
            ++n;
        }
        ssize_t w = writev(fd, cur, n);
        if (w < 0) { if (errno == EINTR) continue; ok = false; break; }
        size_t left = (size_t)w;
        while (first < 2 && left >= iov[first].iov_len) { left -= iov[first].iov_len; iov[first].iov_len = 0; ++first; }
        if (first < 2 && left) {
            iov[first].iov_base = static_cast<char*>(iov[first].iov_base) + left;
            iov[first].iov_len -= left;
        }
    }
    if (close(fd) != 0) ok = false;
    if (!ok) { set_error("Error writing output file: %s", out_path); return NLZ_ERR_RUNTIME; }
    return NLZ_OK;
}

// fasta_processor.cpp:131-163
int identify_sentinels(const uint64_t* t, uint64_t count, const std::vector<uint64_t>& pos, bool checks,
                       std::vector<uint64_t>& idx) {
    idx.clear();
    size_t s = 0;
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t start = t[3 * i], len = t[3 * i + 1], ref = t[3 * i + 2];
        while (s < pos.size() && pos[s] < start) ++s;
        if (s < pos.size() && start == pos[s]) {
            if (checks && len != 1) {
                set_error("Sentinel factor has unexpected length: %llu", (unsigned long long)len);
                return NLZ_ERR_RUNTIME;
            }
            if (checks && ref != start) {
                set_error("Sentinel factor reference mismatch: ref=%llu, pos=%llu", (unsigned long long)ref,
                          (unsigned long long)start);
                return NLZ_ERR_RUNTIME;
            }
            idx.push_back(i);
            ++s;
        }
    }
    return NLZ_OK;
}

std::string names_and_indices(const std::vector<std::string>& ids, const std::vector<uint64_t>& sent_idx) {
    std::string meta;                                                  // parallel_fasta_processor.cpp:29-62
    for (const auto& n : ids) { meta.append(n); meta.push_back('\0'); }
    for (uint64_t v : sent_idx) meta.append(reinterpret_cast<const char*>(&v), 8);
    return meta;
}

uint64_t sum_lengths(const uint64_t* t, uint64_t count) {
    uint64_t s = 0;
    for (uint64_t i = 0; i < count; ++i) s += t[3 * i + 1];
    return s;
}

struct Triples {           // RAII for buffers returned by nlz_factorize_mode
    uint64_t* p = nullptr;
    uint64_t n = 0;
    ~Triples() { nlz_free(p); }
};

}  // namespace

// ------------------------------------------------------------------ FASTA handle
struct nlz_fasta {
    std::vector<std::string> ids, seqs;
};

extern "C" {

// fasta_processor.cpp:28-128.  sanitize_mode: 0 = remove_ambiguous, 1 = strict (hpp:10-13)
int nlz_fasta_parse(const char* path, int sanitize_mode, nlz_fasta** out) {
    if (!out || !path) { set_error("null argument"); return NLZ_ERR_INVALID; }
    *out = nullptr;
    if (sanitize_mode != 0 && sanitize_mode != 1) {
        set_error("Invalid sanitize_mode. Expected 'remove_ambiguous' or 'strict'.");   // bindings.cpp:29-37
        return NLZ_ERR_INVALID;
    }
    // The file is mapped and scanned line by line with memchr; a sequence line made only of upper-case ACGT (the
    // common case) is appended with one memcpy -- the reference parses byte by byte into a std::string
    // (fasta_processor.cpp:85-96), which becomes the wall-time floor once the GPU stages take 100 ms.
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { set_error("Cannot open FASTA file: %s", path); return NLZ_ERR_RUNTIME; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) { close(fd); set_error("Cannot open FASTA file: %s", path); return NLZ_ERR_RUNTIME; }
    const size_t fsize = (size_t)sb.st_size;
    const char* map = nullptr;
    if (fsize) {
        void* m = mmap(nullptr, fsize, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); set_error("Cannot open FASTA file: %s", path); return NLZ_ERR_RUNTIME; }
        map = static_cast<const char*>(m);
        madvise(m, fsize, MADV_SEQUENTIAL);
    }
    struct Unmap { const char* m; size_t n; int fd; ~Unmap() { if (m) munmap(const_cast<char*>(m), n); close(fd); } } unmap{map, fsize, fd};
    static const struct Kind { unsigned char k[256]; Kind() {
        for (int c = 0; c < 256; ++c) k[c] = std::isspace(c) ? 2 : 3;          // 0 ACGT, 1 acgt, 2 whitespace, 3 other
        for (const char* u = "ACGT"; *u; ++u) k[(unsigned char)*u] = 0;
        for (const char* l = "acgt"; *l; ++l) k[(unsigned char)*l] = 1;
    } } kind;
    // The file is cut at header lines into one range per host thread; every range is parsed independently (a record
    // never straddles a cut) and the results are concatenated in file order, so ids, sequences, warnings and the
    // first error are exactly those of a serial scan.
    struct Part {
        std::vector<std::string> ids, seqs;
        std::vector<std::string> skipped;        // ids of empty records, in order (warnings)
        size_t removed = 0;
        int err = NLZ_OK;
        std::string msg;
    };
    auto parse_range = [&](const char* p, const char* const end, Part& out) {
        std::string cur_seq, cur_id;
        auto flush = [&]() {
            if (cur_id.empty()) return;
            if (!cur_seq.empty()) {
                const size_t cap = cur_seq.size();
                out.seqs.push_back(std::move(cur_seq));
                out.ids.push_back(cur_id);
                cur_seq = std::string();
                cur_seq.reserve(cap);
            } else {
                out.skipped.push_back(cur_id);
            }
            cur_seq.clear();
        };
        while (p < end) {
            const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
            const char* le = nl ? nl : end;                         // line = [p, le)
            const char* next = nl ? nl + 1 : end;
            while (le > p && std::isspace((unsigned char)le[-1])) --le;
            if (le == p) { p = next; continue; }
            if (*p == '>') {
                flush();
                const char* a = p + 1;
                while (a < le && std::isspace((unsigned char)*a)) ++a;
                const char* b = a;
                while (b < le && !std::isspace((unsigned char)*b)) ++b;
                if (a < le) cur_id.assign(a, b);
                else { out.err = NLZ_ERR_RUNTIME; out.msg = "Empty sequence header in FASTA file"; return; }
            } else {
                const unsigned char* q = reinterpret_cast<const unsigned char*>(p);
                const unsigned char* qe = reinterpret_cast<const unsigned char*>(le);
                while (q < qe) {
                    const unsigned char* r = q;
                    while (r < qe && kind.k[*r] == 0) ++r;           // run of clean bases
                    if (r > q) cur_seq.append(reinterpret_cast<const char*>(q), (size_t)(r - q));
                    if (r == qe) break;
                    const unsigned char c = *r;
                    if (kind.k[c] == 1) cur_seq.push_back((char)upper_ascii(c));
                    else if (kind.k[c] == 3) {
                        if (sanitize_mode == 1) {
                            out.err = NLZ_ERR_RUNTIME;
                            out.msg = std::string("Invalid nucleotide '") + (char)c + "' found in sequence with ID: " + cur_id;
                            return;
                        }
                        ++out.removed;
                    }
                    q = r + 1;
                }
            }
            p = next;
        }
        flush();
    };
    const char* const end = map + fsize;
    size_t nparts = std::thread::hardware_concurrency();
    if (nparts > 32) nparts = 32;
    if (nparts > fsize / (4u << 20)) nparts = fsize / (4u << 20);        // at least 4 MiB per thread
    if (nparts < 1) nparts = 1;
    std::vector<const char*> cut(nparts + 1, end);
    cut[0] = map;
    for (size_t t = 1; t < nparts; ++t) {
        const char* q = map + fsize / nparts * t;
        if (q < cut[t - 1]) q = cut[t - 1];
        const char* hit = end;                                            // next line that starts with '>'
        while (q < end) {
            const char* nl = static_cast<const char*>(memchr(q, '\n', (size_t)(end - q)));
            if (!nl || nl + 1 >= end) break;
            if (nl[1] == '>') { hit = nl + 1; break; }
            q = nl + 1;
        }
        cut[t] = hit;
    }
    std::vector<Part> parts(nparts);
    if (nparts == 1) parse_range(cut[0], cut[1], parts[0]);
    else {
        std::vector<std::thread> pool;
        for (size_t t = 0; t < nparts; ++t)
            pool.emplace_back([&, t] { if (cut[t] < cut[t + 1]) parse_range(cut[t], cut[t + 1], parts[t]); });
        for (auto& th : pool) th.join();
    }
    nlz_fasta* fa = new nlz_fasta();
    size_t empty_count = 0, removed = 0, total = 0;
    for (const Part& pt : parts) total += pt.seqs.size();
    fa->ids.reserve(total);
    fa->seqs.reserve(total);
    for (Part& pt : parts) {
        // warnings and records of this range come before the first error of a later one, as in a serial scan
        for (const std::string& id : pt.skipped) { std::cerr << "Warning: Skipping empty sequence with ID: " << id << std::endl; ++empty_count; }
        if (pt.err != NLZ_OK) { set_error("%s", pt.msg.c_str()); delete fa; return pt.err; }
        removed += pt.removed;
        for (size_t i = 0; i < pt.seqs.size(); ++i) { fa->seqs.push_back(std::move(pt.seqs[i])); fa->ids.push_back(std::move(pt.ids[i])); }
    }
    if (empty_count > 0) std::cerr << "Warning: Skipped " << empty_count << " empty sequence(s) in FASTA file" << std::endl;
    if (sanitize_mode == 0 && removed > 0)
        std::cerr << "Warning: Removed " << removed << " ambiguous nucleotide(s) from FASTA input" << std::endl;
    if (fa->seqs.empty()) { delete fa; set_error("No valid sequences found in FASTA file"); return NLZ_ERR_RUNTIME; }
    *out = fa;
    return NLZ_OK;
}
uint64_t nlz_fasta_num_sequences(const nlz_fasta* fa) { return fa ? fa->seqs.size() : 0; }
const char* nlz_fasta_id(const nlz_fasta* fa, uint64_t i) { return (fa && i < fa->ids.size()) ? fa->ids[i].c_str() : ""; }
const uint8_t* nlz_fasta_sequence(const nlz_fasta* fa, uint64_t i, uint64_t* len) {
    if (!fa || i >= fa->seqs.size()) { if (len) *len = 0; return nullptr; }
    if (len) *len = fa->seqs[i].size();
    return reinterpret_cast<const uint8_t*>(fa->seqs[i].data());
}
void nlz_fasta_free(nlz_fasta* fa) { delete fa; }

// ------------------------------------------------------------------ prepare
static int prepare_abi(const uint8_t* const* seqs, const uint64_t* lens, uint64_t k, bool with_rc, uint8_t** prepared,
                       uint64_t* prepared_len, uint64_t* original_length, uint64_t** sentinels, uint64_t* n_sent) {
    if (!prepared || !prepared_len || !original_length || !sentinels || !n_sent) { set_error("null argument"); return NLZ_ERR_INVALID; }
    std::vector<std::string> v(k);
    for (uint64_t i = 0; i < k; ++i) v[i].assign(reinterpret_cast<const char*>(seqs[i]), (size_t)lens[i]);
    Prepared p;
    int rc = prepare(v, with_rc, p);
    if (rc != NLZ_OK) return rc;
    *prepared = static_cast<uint8_t*>(malloc(p.s.size() + 1));
    *sentinels = static_cast<uint64_t*>(malloc((p.sentinels.size() + 1) * 8));
    if (!*prepared || !*sentinels) { set_error("out of host memory"); return NLZ_ERR_RUNTIME; }
    memcpy(*prepared, p.s.data(), p.s.size());
    if (!p.sentinels.empty()) memcpy(*sentinels, p.sentinels.data(), p.sentinels.size() * 8);
    *prepared_len = p.s.size();
    *original_length = p.original_length;
    *n_sent = p.sentinels.size();
    return NLZ_OK;
}
int nlz_prepare_multiple_dna_sequences_w_rc(const uint8_t* const* seqs, const uint64_t* lens, uint64_t k, uint8_t** prepared,
                                            uint64_t* prepared_len, uint64_t* original_length, uint64_t** sentinels,
                                            uint64_t* n_sent) {
    return prepare_abi(seqs, lens, k, true, prepared, prepared_len, original_length, sentinels, n_sent);
}
int nlz_prepare_multiple_dna_sequences_no_rc(const uint8_t* const* seqs, const uint64_t* lens, uint64_t k, uint8_t** prepared,
                                             uint64_t* prepared_len, uint64_t* original_length, uint64_t** sentinels,
                                             uint64_t* n_sent) {
    return prepare_abi(seqs, lens, k, false, prepared, prepared_len, original_length, sentinels, n_sent);
}

// ------------------------------------------------------------------ helpers exposed to the shim
int nlz_identify_sentinel_factors(const uint64_t* triples, uint64_t count, const uint64_t* sentinel_positions, uint64_t n_pos,
                                  uint64_t** out_idx, uint64_t* out_n) {
    if (!out_idx || !out_n) { set_error("null argument"); return NLZ_ERR_INVALID; }
    std::vector<uint64_t> pos(sentinel_positions, sentinel_positions + n_pos), idx;
    int rc = identify_sentinels(triples, count, pos, true, idx);
    if (rc != NLZ_OK) return rc;
    *out_idx = static_cast<uint64_t*>(malloc((idx.size() + 1) * 8));
    if (!*out_idx) { set_error("out of host memory"); return NLZ_ERR_RUNTIME; }
    if (!idx.empty()) memcpy(*out_idx, idx.data(), idx.size() * 8);
    *out_n = idx.size();
    return NLZ_OK;
}

// generic writer: [factors][meta bytes][footer], footer_size = 48 + meta_len
int nlz_write_factor_file(const char* out_path, const uint64_t* triples, uint64_t count, const uint8_t* meta, uint64_t meta_len,
                          uint64_t num_sequences, uint64_t num_sentinels, uint64_t total_length) {
    std::string m(reinterpret_cast<const char*>(meta), (size_t)meta_len);
    return write_factor_file(out_path, triples, count, m, num_sequences, num_sentinels, total_length);
}

// ------------------------------------------------------------------ file -> factors
// factorize_file / factorize_file_dna_w_rc / factorize_file_multiple_dna_w_rc
// (factorizer.cpp:401-406, :525-529 via :487-492, :659-674): raw file bytes, no newline stripping
int nlz_factorize_file_mode(nlz_ctx* ctx, int mode, const char* path, uint64_t start_pos, uint64_t** out, uint64_t* count) {
    std::string data;
    int rc = read_whole_file(path, mode == NLZ_MODE_RC_PREPARED ? "Cannot open file" : "Cannot open input file", data);
    if (rc != NLZ_OK) return rc;
    return nlz_factorize_mode(ctx, mode, reinterpret_cast<const uint8_t*>(data.data()), data.size(), start_pos, out, count);
}
int nlz_count_file_mode(nlz_ctx* ctx, int mode, const char* path, uint64_t start_pos, uint64_t* count) {
    std::string data;
    int rc = read_whole_file(path, mode == NLZ_MODE_RC_PREPARED ? "Cannot open file" : "Cannot open input file", data);
    if (rc != NLZ_OK) return rc;
    return nlz_count_mode(ctx, mode, reinterpret_cast<const uint8_t*>(data.data()), data.size(), start_pos, count);
}

// write_factors_binary_file (V1, factorizer.cpp:424-459), write_factors_binary_file_dna_w_rc (V2, :597-635),
// write_factors_binary_file_multiple_dna_w_rc (V3, :751-790)
int nlz_write_factors_binary_file_mode(nlz_ctx* ctx, int mode, const char* in_path, const char* out_path, uint64_t start_pos,
                                       uint64_t* count) {
    if (!count) { set_error("null argument"); return NLZ_ERR_INVALID; }
    std::string data;
    int rc = read_whole_file(in_path, "Cannot open input file", data);
    if (rc != NLZ_OK) return rc;
    // the reference opens (truncates) the output before factorizing
    { FILE* f = fopen(out_path, "wb"); if (!f) { set_error("Cannot create output file: %s", out_path); return NLZ_ERR_RUNTIME; } fclose(f); }
    Triples t;
    rc = nlz_factorize_mode(ctx, mode, reinterpret_cast<const uint8_t*>(data.data()), data.size(), start_pos, &t.p, &t.n);
    if (rc != NLZ_OK) return rc;
    *count = t.n;
    if (mode == NLZ_MODE_GENERAL) return write_factor_file(out_path, t.p, t.n, std::string(), 0, 0, data.size());
    if (mode == NLZ_MODE_DNA_RC) return write_factor_file(out_path, t.p, t.n, std::string(1, '\0'), 1, 0, data.size());
    return write_factor_file(out_path, t.p, t.n, std::string(), 0, 0, data.size() - start_pos);
}

// text -> file with the parallel-mode footer (V6: 0/0/48, total = sum of lengths; parallel_factorizer.cpp:754-767)
// mode GENERAL: parallel_factorize (:55-61); mode DNA_RC: parallel_factorize_dna_w_rc (:1001-1017)
int nlz_parallel_factorize_to_file(nlz_ctx* ctx, int mode, const uint8_t* text, uint64_t n, const char* out_path,
                                   uint64_t start_pos, uint64_t* count) {
    if (!count) { set_error("null argument"); return NLZ_ERR_INVALID; }
    *count = 0;
    if (n == 0) return NLZ_OK;                                         // returns 0 without touching the file
    if (mode == NLZ_MODE_GENERAL && start_pos >= n) {
        set_error("start_pos must be less than text length");
        return NLZ_ERR_INVALID;
    }
    Triples t;
    int rc;
    if (mode == NLZ_MODE_DNA_RC && start_pos != 0) {
        std::vector<std::string> one(1, std::string(reinterpret_cast<const char*>(text), (size_t)n));
        Prepared p;
        rc = prepare(one, true, p);
        if (rc != NLZ_OK) return rc;
        rc = nlz_factorize_mode(ctx, NLZ_MODE_RC_PREPARED, reinterpret_cast<const uint8_t*>(p.s.data()), p.s.size(), start_pos, &t.p, &t.n);
    } else {
        rc = nlz_factorize_mode(ctx, mode, text, n, start_pos, &t.p, &t.n);
    }
    if (rc != NLZ_OK) return rc;
    *count = t.n;
    return write_factor_file(out_path, t.p, t.n, std::string(), 0, 0, sum_lengths(t.p, t.n));
}

// reference + target, text variants (V4/V5): factorize_dna_w_reference_seq(_file) factorizer.cpp:825-908,
// factorize_w_reference(_file) :934-1021.  out_path may be NULL (no file); out/count may be NULL when only the file is wanted.
int nlz_factorize_w_reference(nlz_ctx* ctx, int dna, const uint8_t* ref, uint64_t ref_len, const uint8_t* tgt, uint64_t tgt_len,
                              const char* out_path, uint64_t** out, uint64_t* count) {
    if (!count) { set_error("null argument"); return NLZ_ERR_INVALID; }
    FILE* probe = nullptr;
    if (out_path) {
        probe = fopen(out_path, "wb");
        if (!probe) { set_error("Cannot create output file: %s", out_path); return NLZ_ERR_RUNTIME; }
        fclose(probe);
    }
    Triples t;
    int rc;
    const uint64_t start = ref_len + 1;                                // :833 / :945
    if (dna) {
        std::vector<std::string> two = {std::string(reinterpret_cast<const char*>(ref), (size_t)ref_len),
                                        std::string(reinterpret_cast<const char*>(tgt), (size_t)tgt_len)};
        Prepared p;
        rc = prepare(two, true, p);
        if (rc != NLZ_OK) return rc;
        rc = nlz_factorize_mode(ctx, NLZ_MODE_RC_PREPARED, reinterpret_cast<const uint8_t*>(p.s.data()), p.s.size(), start, &t.p, &t.n);
    } else {
        std::string comb(reinterpret_cast<const char*>(ref), (size_t)ref_len);
        comb.push_back('\x01');                                        // :942
        comb.append(reinterpret_cast<const char*>(tgt), (size_t)tgt_len);
        rc = nlz_factorize_mode(ctx, NLZ_MODE_GENERAL, reinterpret_cast<const uint8_t*>(comb.data()), comb.size(), start, &t.p, &t.n);
    }
    if (rc != NLZ_OK) return rc;
    *count = t.n;
    if (out_path) {
        rc = write_factor_file(out_path, t.p, t.n, std::string(), 2, 1, tgt_len);   // :897-906: no names written
        if (rc != NLZ_OK) return rc;
    }
    if (out) { *out = t.p; t.p = nullptr; }
    return NLZ_OK;
}

// ------------------------------------------------------------------ FASTA composites
// concatenated FASTA (fasta_processor.cpp:298-341, parallel_fasta_processor.cpp:64-179): factors, sentinel factor
// indices and, when out_path != NULL, the V7 file [factors][ids\0...][u64 sentinel idx...][footer], total = sum of lengths.
// ref_fasta != NULL selects the reference+target form (fasta_processor.cpp:393-423), RC only.
int nlz_factorize_fasta(nlz_ctx* ctx, const char* ref_fasta, const char* fasta_path, int with_rc, int sanitize_mode,
                        const char* out_path, uint64_t** out, uint64_t* count, uint64_t** sent_idx, uint64_t* n_sent_idx,
                        nlz_fasta** ids_out) {
    if (!count) { set_error("null argument"); return NLZ_ERR_INVALID; }
    nlz_fasta* all = new nlz_fasta();
    uint64_t start = 0;
    const char* paths[2] = {ref_fasta, fasta_path};
    for (int k = 0; k < 2; ++k) {
        if (!paths[k]) continue;
        nlz_fasta* fa = nullptr;
        int rc = nlz_fasta_parse(paths[k], sanitize_mode, &fa);
        if (rc != NLZ_OK) { delete all; return rc; }
        if (k == 0) for (const auto& q : fa->seqs) start += q.size() + 1;    // target_start_index, :236-239
        for (size_t i = 0; i < fa->seqs.size(); ++i) { all->seqs.push_back(std::move(fa->seqs[i])); all->ids.push_back(std::move(fa->ids[i])); }
        delete fa;
    }
    Prepared p;
    int rc = prepare(all->seqs, with_rc != 0, p);
    if (rc != NLZ_OK) { delete all; return rc; }
    Triples t;
    rc = nlz_factorize_mode(ctx, with_rc ? NLZ_MODE_RC_PREPARED : NLZ_MODE_GENERAL,
                            reinterpret_cast<const uint8_t*>(p.s.data()), p.s.size(), start, &t.p, &t.n);
    if (rc != NLZ_OK) { delete all; return rc; }
    std::vector<uint64_t> idx;
    rc = identify_sentinels(t.p, t.n, p.sentinels, out_path == nullptr, idx);   // the file writer does not sanity-check
    if (rc != NLZ_OK) { delete all; return rc; }
    if (out_path) {
        rc = write_factor_file(out_path, t.p, t.n, names_and_indices(all->ids, idx), all->ids.size(), idx.size(),
                               sum_lengths(t.p, t.n));
        if (rc != NLZ_OK) { delete all; return rc; }
    }
    *count = t.n;
    if (sent_idx && n_sent_idx) {
        *sent_idx = static_cast<uint64_t*>(malloc((idx.size() + 1) * 8));
        if (!idx.empty()) memcpy(*sent_idx, idx.data(), idx.size() * 8);
        *n_sent_idx = idx.size();
    }
    if (out) { *out = t.p; t.p = nullptr; }
    if (ids_out) *ids_out = all; else delete all;
    return NLZ_OK;
}

// per-sequence FASTA (fasta_processor.cpp:428-561, parallel_fasta_processor.cpp:268-465).  Results are returned as one
// concatenated triple array plus per-record counts; out_dir != NULL also writes <out_dir>/<sanitized id>.bin (V8).
// The no-RC variants drop the last nucleotide of every record, like the reference (fasta_processor.cpp:469-471).
// Records are independent: instead of the reference's worker pool of per-record index builds
// (parallel_fasta_processor.cpp:360-385) the records are concatenated and run through ONE segmented pipeline per chunk
// (nlz_factorize_batch); `num_threads` host threads (0 = 8) only write the per-record files.
int nlz_factorize_fasta_per_sequence(nlz_ctx* ctx, const char* fasta_path, int with_rc, int sanitize_mode, const char* out_dir,
                                     int want_factors, int num_threads, uint64_t** out, uint64_t** per_seq_counts,
                                     uint64_t* total, nlz_fasta** ids_out) {
    if (!total) { set_error("null argument"); return NLZ_ERR_INVALID; }
    nlz_fasta* fa = nullptr;
    int rc = nlz_fasta_parse(fasta_path, sanitize_mode, &fa);
    if (rc != NLZ_OK) return rc;
    if (out_dir) {
        std::string d(out_dir);
        for (size_t i = 1; i <= d.size(); ++i)
            if (i == d.size() || d[i] == '/') { std::string sub = d.substr(0, i); if (!sub.empty()) mkdir(sub.c_str(), 0777); }
    }
    const size_t k = fa->seqs.size();
    const bool need_triples = want_factors || out_dir;
    std::vector<uint64_t> counts(k, 0);
    std::vector<uint64_t> all;                                     // concatenated record-local triples, record order
    constexpr uint64_t kChunkSuffixes = 1ull << 28, kChunkRecords = 1ull << 20;
    size_t i0 = 0;
    while (i0 < k && rc == NLZ_OK) {
        std::string concat;
        std::vector<uint64_t> offs, lens;
        uint64_t suffixes = 0, bytes = 0;
        size_t i1 = i0;
        while (i1 < k && (i1 == i0 || (suffixes < kChunkSuffixes && i1 - i0 < kChunkRecords))) {
            const std::string& q = fa->seqs[i1];
            const uint64_t n = with_rc ? q.size() : q.size() - 1;  // fasta_processor.cpp:469-471 (no-RC drops the last base)
            offs.push_back(bytes);
            lens.push_back(n);
            bytes += n;
            suffixes += (n + 1) * (with_rc ? 2 : 1);
            ++i1;
        }
        concat.resize(bytes);                                      // one allocation; records copied by a few threads
        {
            const size_t nrec = i1 - i0;
            size_t nt = std::min<size_t>(8, std::max<size_t>(1, bytes >> 22));
            if (nt > nrec) nt = nrec ? nrec : 1;
            auto copy_range = [&](size_t a, size_t b) {
                for (size_t j = a; j < b; ++j) memcpy(&concat[offs[j]], fa->seqs[i0 + j].data(), lens[j]);
            };
            if (nt <= 1) copy_range(0, nrec);
            else {
                std::vector<std::thread> pool;
                for (size_t t = 0; t < nt; ++t) pool.emplace_back(copy_range, nrec * t / nt, nrec * (t + 1) / nt);
                for (auto& th : pool) th.join();
            }
        }
        Triples t;
        rc = nlz_factorize_batch(ctx, with_rc, reinterpret_cast<const uint8_t*>(concat.data()), offs.data(), lens.data(),
                                 i1 - i0, need_triples ? &t.p : nullptr, counts.data() + i0, &t.n);
        if (rc == NLZ_OK && need_triples && t.n) all.insert(all.end(), t.p, t.p + 3 * t.n);
        i0 = i1;
    }
    if (rc != NLZ_OK) { delete fa; return rc; }
    std::vector<uint64_t> first(k + 1, 0);                         // first factor of every record in `all`
    for (size_t i = 0; i < k; ++i) first[i + 1] = first[i] + counts[i];
    if (out_dir) {
        size_t nthreads = num_threads > 0 ? (size_t)num_threads : 8;
        if (nthreads > k) nthreads = k;
        if (nthreads < 1) nthreads = 1;
        std::atomic<size_t> next(0);
        std::atomic<int> failed(NLZ_OK);
        std::vector<std::string> errors(nthreads);
        auto writer = [&](size_t tix) {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= k || failed.load() != NLZ_OK) break;
                std::string safe = fa->ids[i];                         // parallel_fasta_processor.cpp:307-317
                for (char& c : safe)
                    if (c == '/' || c == '\\' || c == ':' || c == '*' || c == '?' || c == '"' || c == '<' || c == '>' || c == '|' || c == ' ') c = '_';
                std::string path = std::string(out_dir) + "/" + safe + ".bin";
                std::string meta = fa->ids[i];
                meta.push_back('\0');
                const uint64_t* tp = all.data() + 3 * first[i];
                int r = write_factor_file(path.c_str(), tp, counts[i], meta, 1, 0, sum_lengths(tp, counts[i]));   // :268-297
                if (r != NLZ_OK) { errors[tix] = nlz_last_error(); failed.store(r); break; }
            }
        };
        if (nthreads == 1) writer(0);
        else {
            std::vector<std::thread> pool;
            for (size_t t = 0; t < nthreads; ++t) pool.emplace_back(writer, t);
            for (auto& th : pool) th.join();
        }
        if (failed.load() != NLZ_OK) {
            for (const auto& e : errors) if (!e.empty()) { set_error("%s", e.c_str()); break; }
            delete fa;
            return failed.load();
        }
    }
    *total = first[k];
    if (per_seq_counts) {
        *per_seq_counts = static_cast<uint64_t*>(malloc((k + 1) * 8));
        memcpy(*per_seq_counts, counts.data(), k * 8);
    }
    if (out) {
        *out = static_cast<uint64_t*>(malloc((all.size() + 1) * 8));
        if (!all.empty()) memcpy(*out, all.data(), all.size() * 8);
    }
    if (ids_out) *ids_out = fa; else delete fa;
    return NLZ_OK;
}

}  // extern "C"
