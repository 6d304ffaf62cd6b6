/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for the noLZSS hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product path (nolzss_b200/) never links or calls it.
 *
 * What it restates (all paths relative to /root/reference):
 *   - the suffix tree of text·$ that the reference builds with SDSL (sdsl-lite v3.0.3, fetched at
 *     build time by CMakeLists.txt:29-43, NOT present on disk) at
 *     src/cpp/factorizer.cpp:320,340,381 and src/cpp/factorizer_core.hpp:208, here as a plain
 *     suffix array (SA-IS) + Kasai LCP array; every CST query on the path has a unique definition
 *     over SA/LCP (see oracle/treewalk_model.py header for the query-by-query map);
 *   - detail::nolzss                       src/cpp/factorizer_core.hpp:51-119   -> nlzo_factorize
 *   - detail::nolzss_multiple_dna_w_rc     src/cpp/factorizer_core.hpp:177-383  -> nlzo_factorize_multiple_dna_w_rc
 *   - lcp()                                src/cpp/factorizer_helpers.hpp:20-24 -> direct_lcp
 *
 * The reference enumerates the path nodes of the current leaf root->leaf (level_anc) and stops at
 * the first node that fails its predicate.  Both predicates are monotone along the path
 * (factorizer_core.hpp:273-277 relies on exactly that), so this file enumerates the same nodes
 * leaf->root as nested LCP intervals and stops at the first node that satisfies the predicate: the
 * same node the reference's loop ends on.  Pinned against the reference's own known-answer tests
 * and against the literal root->leaf model in oracle/treewalk_model.py by tests/test_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* Index width.  The reference is 64-bit end to end (sdsl::int_vector<64>, factorizer_core.hpp:211-213).  The default
 * build uses 32-bit indices (texts below 2^31 suffixes: everything the CPU tests and the bench legs run); -DNLZO_IDX64
 * (libnlz_oracle64.so) widens every index to 64 bits -- same code, checked against the 32-bit build by
 * tests/test_oracle.py -- for texts beyond that (host memory permitting: 24 bytes per suffix). */
#ifdef NLZO_IDX64
typedef int64_t idx_t;
#define NLZO_IDX_LIMIT (1ULL << 62)
#else
typedef int32_t idx_t;
#define NLZO_IDX_LIMIT (1ULL << 31)
#endif
#define RC_MASK (1ULL << 63)

/* ------------------------------------------------------------------ SA-IS (Nong, Zhang, Chan) */
/* s[0..n) over alphabet [0,K), s[n-1] == 0 is the unique smallest symbol. */
static void bucket_bounds(const idx_t *s, idx_t *bkt, idx_t n, idx_t K, int want_end) {
    for (idx_t c = 0; c < K; ++c) bkt[c] = 0;
    for (idx_t i = 0; i < n; ++i) bkt[s[i]]++;
    idx_t sum = 0;
    for (idx_t c = 0; c < K; ++c) {
        sum += bkt[c];
        bkt[c] = want_end ? sum : sum - bkt[c];
    }
}
#define IS_LMS(i) ((i) > 0 && t[(i)] && !t[(i)-1])

static void induce(const idx_t *s, idx_t *SA, const uint8_t *t, idx_t *bkt, idx_t n, idx_t K) {
    bucket_bounds(s, bkt, n, K, 0);
    for (idx_t i = 0; i < n; ++i) {
        idx_t j = SA[i] - 1;
        if (SA[i] > 0 && !t[j]) SA[bkt[s[j]]++] = j;
    }
    bucket_bounds(s, bkt, n, K, 1);
    for (idx_t i = n - 1; i >= 0; --i) {
        idx_t j = SA[i] - 1;
        if (SA[i] > 0 && t[j]) SA[--bkt[s[j]]] = j;
    }
}

static int sais(const idx_t *s, idx_t *SA, idx_t n, idx_t K) {
    if (n == 1) { SA[0] = 0; return 0; }
    uint8_t *t = (uint8_t *)malloc((size_t)n);
    idx_t *bkt = (idx_t *)malloc(sizeof(idx_t) * (size_t)K);
    if (!t || !bkt) { free(t); free(bkt); return -1; }
    t[n - 1] = 1; t[n - 2] = 0;
    for (idx_t i = n - 3; i >= 0; --i)
        t[i] = (s[i] < s[i + 1] || (s[i] == s[i + 1] && t[i + 1])) ? 1 : 0;

    /* stage 1: sort LMS substrings */
    bucket_bounds(s, bkt, n, K, 1);
    for (idx_t i = 0; i < n; ++i) SA[i] = -1;
    for (idx_t i = 1; i < n; ++i) if (IS_LMS(i)) SA[--bkt[s[i]]] = i;
    induce(s, SA, t, bkt, n, K);

    idx_t n1 = 0;
    for (idx_t i = 0; i < n; ++i) if (IS_LMS(SA[i])) SA[n1++] = SA[i];
    for (idx_t i = n1; i < n; ++i) SA[i] = -1;
    idx_t name = 0, prev = -1;
    for (idx_t i = 0; i < n1; ++i) {
        idx_t pos = SA[i];
        int diff = 0;
        for (idx_t d = 0; d < n; ++d) {
            if (prev == -1 || s[pos + d] != s[prev + d] || t[pos + d] != t[prev + d]) { diff = 1; break; }
            if (d > 0 && (IS_LMS(pos + d) || IS_LMS(prev + d))) break;
        }
        if (diff) { name++; prev = pos; }
        SA[n1 + pos / 2] = name - 1;
    }
    for (idx_t i = n - 1, j = n - 1; i >= n1; --i) if (SA[i] >= 0) SA[j--] = SA[i];

    /* stage 2: solve the reduced problem */
    idx_t *SA1 = SA, *s1 = SA + n - n1;
    if (name < n1) {
        if (sais(s1, SA1, n1, name) != 0) { free(t); free(bkt); return -1; }
    } else {
        for (idx_t i = 0; i < n1; ++i) SA1[s1[i]] = i;
    }

    /* stage 3: induce the final order */
    bucket_bounds(s, bkt, n, K, 1);
    for (idx_t i = 1, j = 0; i < n; ++i) if (IS_LMS(i)) s1[j++] = i;
    for (idx_t i = 0; i < n1; ++i) SA1[i] = s1[SA1[i]];
    for (idx_t i = n1; i < n; ++i) SA[i] = -1;
    for (idx_t i = n1 - 1; i >= 0; --i) {
        idx_t j = SA[i];
        SA[i] = -1;
        SA[--bkt[s[j]]] = j;
    }
    induce(s, SA, t, bkt, n, K);
    free(t); free(bkt);
    return 0;
}

#ifndef NLZO_IDX64
int nlzo_suffix_array_i32(const int32_t *s, int32_t n, int32_t K, int32_t *sa) {
    if (n <= 0) return 0;
    return sais(s, sa, n, K);
}
#endif

/* ------------------------------------------------------------------ index over bytes·$ */
typedef struct {
    idx_t n1;          /* number of suffixes = |text| + 1 */
    idx_t *sa, *isa;
    idx_t *lcp;        /* lcp[k] = LCP(sa[k-1], sa[k]); lcp[0] = lcp[n1] = 0 */
} index_t;

static void index_free(index_t *ix) { free(ix->sa); free(ix->isa); free(ix->lcp); memset(ix, 0, sizeof(*ix)); }

static int index_build(const uint8_t *x, uint64_t n, index_t *ix) {
    memset(ix, 0, sizeof(*ix));
    if (n + 1 >= NLZO_IDX_LIMIT) return -2;
    idx_t n1 = (idx_t)n + 1;
    idx_t *s = (idx_t *)malloc(sizeof(idx_t) * (size_t)n1);
    ix->sa = (idx_t *)malloc(sizeof(idx_t) * (size_t)n1);
    ix->isa = (idx_t *)malloc(sizeof(idx_t) * (size_t)n1);
    ix->lcp = (idx_t *)calloc((size_t)n1 + 1, sizeof(idx_t));
    if (!s || !ix->sa || !ix->isa || !ix->lcp) { free(s); index_free(ix); return -1; }
    for (idx_t i = 0; i < n1 - 1; ++i) s[i] = (idx_t)x[i] + 1;   /* byte order, $ = 0 smallest */
    s[n1 - 1] = 0;
    if (sais(s, ix->sa, n1, 257) != 0) { free(s); index_free(ix); return -1; }
    free(s);
    ix->n1 = n1;
    for (idx_t k = 0; k < n1; ++k) ix->isa[ix->sa[k]] = k;
    /* Kasai et al. */
    idx_t l = 0;
    for (idx_t i = 0; i < n1; ++i) {
        idx_t r = ix->isa[i];
        if (r == 0) { l = 0; continue; }
        idx_t j = ix->sa[r - 1];
        while ((uint64_t)(i + l) < n && (uint64_t)(j + l) < n && x[i + l] == x[j + l]) ++l;
        ix->lcp[r] = l;
        if (l > 0) --l;
    }
    ix->lcp[0] = 0; ix->lcp[n1] = 0;
    return 0;
}

#ifndef NLZO_IDX64
int nlzo_sa_lcp_bytes(const uint8_t *text, uint64_t n, int32_t *sa, int32_t *lcp) {
    index_t ix;
    int rc = index_build(text, n, &ix);
    if (rc) return rc;
    memcpy(sa, ix.sa, sizeof(idx_t) * (size_t)ix.n1);
    memcpy(lcp, ix.lcp, sizeof(idx_t) * ((size_t)ix.n1 + 1));
    index_free(&ix);
    return 0;
}
#endif

/* factorizer_helpers.hpp:20-24 : LCA string depth == number of equal leading symbols of the two
 * suffixes of x·$ ($ unique).  `cap` lets callers stop once the answer cannot matter. */
static uint64_t direct_lcp(const uint8_t *x, uint64_t n, uint64_t i, uint64_t j, uint64_t cap) {
    if (i == j) return n + 1 - i;
    uint64_t l = 0;
    while (l < cap && i + l < n && j + l < n && x[i + l] == x[j + l]) ++l;
    return l;
}

typedef struct { uint64_t *v; uint64_t count, cap; } sink_t;
static int sink_push(sink_t *s, uint64_t a, uint64_t b, uint64_t c) {
    if (s->count == s->cap) {
        uint64_t nc = s->cap ? s->cap * 2 : 1024;
        uint64_t *nv = (uint64_t *)realloc(s->v, (size_t)nc * 24);
        if (!nv) return -1;
        s->v = nv; s->cap = nc;
    }
    uint64_t *p = s->v + 3 * s->count++;
    p[0] = a; p[1] = b; p[2] = c;
    return 0;
}

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double g_last_index_s = 0.0, g_last_walk_s = 0.0;
void nlzo_last_timing(double *index_s, double *walk_s) { *index_s = g_last_index_s; *walk_s = g_last_walk_s; }

/* ------------------------------------------------------------------ general mode */
/* factorizer_core.hpp:51-119 */
int nlzo_factorize(const uint8_t *x, uint64_t n, uint64_t start_pos, uint64_t **out, uint64_t *count) {
    *out = NULL; *count = 0;
    if (n == 0) return 0;
    index_t ix;
    double t0 = now_s();
    int rc = index_build(x, n, &ix);
    if (rc) return rc;
    double t1 = now_s();
    sink_t sk = {0, 0, 0};
    const idx_t n1 = ix.n1;
    uint64_t i = start_pos;
    while (i < n) {                                            /* :66 */
        idx_t lo = ix.isa[i], hi = lo;
        uint64_t v_min = i;            /* min leaf of the node below u (the reference's failing v) */
        uint64_t u_min = 0, u_depth = 0;
        int have_u = 0;
        uint64_t m = i;
        for (;;) {
            idx_t d = ix.lcp[lo] > ix.lcp[hi + 1] ? ix.lcp[lo] : ix.lcp[hi + 1];
            if (d <= 0) break;                                 /* parent is the root */
            while (ix.lcp[lo] >= d) { --lo; if ((uint64_t)ix.sa[lo] < m) m = (uint64_t)ix.sa[lo]; }
            while (hi + 1 < n1 && ix.lcp[hi + 1] >= d) { ++hi; if ((uint64_t)ix.sa[hi] < m) m = (uint64_t)ix.sa[hi]; }
            if (m + (uint64_t)d - 1 < i) {                     /* :75 holds at this node => it is u */
                have_u = 1; u_min = m; u_depth = (uint64_t)d;
                break;
            }
            v_min = m;
        }
        uint64_t len, ref;
        if (v_min == i) {                                      /* :82 */
            if (!have_u) { len = 1; ref = i; }                 /* :83-87 */
            else { len = u_depth; ref = u_min; }               /* :89-94 */
        } else {
            uint64_t cap = i - v_min;
            uint64_t L = direct_lcp(x, n, i, v_min, cap);      /* :96-97 */
            if (L > cap) L = cap;
            if (L <= u_depth) { len = u_depth; ref = u_min; }  /* :98-102 */
            else { len = L; ref = v_min; }                     /* :104-107 */
        }
        if (sink_push(&sk, i, len, ref)) { index_free(&ix); free(sk.v); return -1; }
        i += len;                                              /* :113-115 */
    }
    index_free(&ix);
    g_last_index_s = t1 - t0; g_last_walk_s = now_s() - t1;
    *out = sk.v; *count = sk.count;
    return 0;
}

/* ------------------------------------------------------------------ RC mode */
typedef struct { const index_t *ix; const uint8_t *S; uint64_t len_S, N; } rc_ctx_t;

/* One iteration of the reference's main loop (factorizer_core.hpp:241-379): the factor that starts at i. */
static void rc_factor_at(const rc_ctx_t *c, uint64_t i, uint64_t *out_len, uint64_t *out_ref) {
    const index_t *ixp = c->ix;
    const uint8_t *S = c->S;
    const uint64_t len_S = c->len_S, N = c->N;
    const idx_t n1 = ixp->n1;
    const uint64_t INF = UINT64_MAX / 2;
    const uint64_t T_end = N, R_beg = N + 1, R_end = len_S - 1; /* :215-217 */
#define FWD_START(k) ((uint64_t)ixp->sa[k] < T_end ? (uint64_t)ixp->sa[k] : INF)                 /* :221-223 */
#define RC_END(k) (((uint64_t)ixp->sa[k] >= R_beg && (uint64_t)ixp->sa[k] < R_end)                \
                       ? N - ((uint64_t)ixp->sa[k] - R_beg) - 1 : INF)                           /* :224-229 */
    idx_t lo = ixp->isa[i], hi = lo;
    uint64_t jF = FWD_START(lo), eR = RC_END(lo), pR = (uint64_t)ixp->sa[lo];
    int have_fwd = 0, have_rc = 0;
    uint64_t best_fwd_start = 0, best_rc_end = 0, best_rc_posS = 0;
    for (;;) {
        idx_t d = ixp->lcp[lo] > ixp->lcp[hi + 1] ? ixp->lcp[lo] : ixp->lcp[hi + 1];
        if (d <= 0) break;                                 /* :259 root */
        while (ixp->lcp[lo] >= d) {
            --lo;
            uint64_t f = FWD_START(lo), e = RC_END(lo);
            if (f < jF) jF = f;
            if (e < eR) { eR = e; pR = (uint64_t)ixp->sa[lo]; }
        }
        while (hi + 1 < n1 && ixp->lcp[hi + 1] >= d) {
            ++hi;
            uint64_t f = FWD_START(hi), e = RC_END(hi);
            if (f < jF) jF = f;
            if (e < eR) { eR = e; pR = (uint64_t)ixp->sa[hi]; }
        }
        int okF = (jF != INF) && (jF + (uint64_t)d - 1 < i);   /* :266 */
        int okR = (eR != INF) && (eR < i);                     /* :271 */
        if (okF && !have_fwd) { have_fwd = 1; best_fwd_start = jF; }              /* deepest okF: :280-287 */
        if (okR && !have_rc) { have_rc = 1; best_rc_end = eR; best_rc_posS = pR; } /* deepest okR: :290-299 */
        if (have_fwd && have_rc) break;
    }
#undef FWD_START
#undef RC_END
    uint64_t emit_len = 1, emit_ref = i;                   /* :302-303 */
    if (have_fwd || have_rc) {
        uint64_t fwd_true_len = 0, rc_true_len = 0;
        if (have_fwd) {                                    /* :322-326 */
            uint64_t cap = i - best_fwd_start;
            uint64_t L = direct_lcp(S, len_S, i, best_fwd_start, cap);
            fwd_true_len = L < cap ? L : cap;
        }
        if (have_rc) rc_true_len = direct_lcp(S, len_S, i, best_rc_posS, UINT64_MAX); /* :328-330 */
        int use_fwd = 0, use_literal = 0;                  /* :335-352 */
        if (have_fwd && fwd_true_len >= 1) {
            use_fwd = !(have_rc && rc_true_len > fwd_true_len);
        } else {
            if (have_rc && rc_true_len > 1) use_fwd = 0; else use_literal = 1;
        }
        if (use_literal) { emit_len = 1; emit_ref = i; }   /* :354-365 */
        else if (use_fwd) { emit_len = fwd_true_len; emit_ref = best_fwd_start; }
        else { emit_len = rc_true_len; emit_ref = RC_MASK | (best_rc_end - emit_len + 1); }
    }
    *out_len = emit_len; *out_ref = emit_ref;
}

/* factorizer_core.hpp:177-383.  Returns 1 for the start_pos error (std::invalid_argument, :203). */
int nlzo_factorize_multiple_dna_w_rc(const uint8_t *S, uint64_t len_S, uint64_t start_pos,
                                     uint64_t **out, uint64_t *count) {
    *out = NULL; *count = 0;
    if (len_S == 0) return 0;                                  /* :180 */
    if (len_S < 4) return 0;                                   /* :189 */
    const uint64_t N = len_S / 2 - 1;                          /* :195 */
    if (N == 0) return 0;                                      /* :196 */
    if (start_pos >= N) return 1;                              /* :203 */
    index_t ix;
    double t0 = now_s();
    int rc = index_build(S, len_S, &ix);                       /* :208 */
    if (rc) return rc;
    double t1 = now_s();
    rc_ctx_t c = {&ix, S, len_S, N};
    sink_t sk = {0, 0, 0};
    uint64_t i = start_pos;
    while (i < N) {                                            /* :241 */
        uint64_t emit_len, emit_ref;
        rc_factor_at(&c, i, &emit_len, &emit_ref);
        if (sink_push(&sk, i, emit_len, emit_ref)) { index_free(&ix); free(sk.v); return -1; }
        i += emit_len;                                         /* :377-379 */
    }
    index_free(&ix);
    g_last_index_s = t1 - t0; g_last_walk_s = now_s() - t1;
    *out = sk.v; *count = sk.count;
    return 0;
}

/* ------------------------------------------------------------------ RC mode, the reference's CPU "parallel mode"
 * src/cpp/parallel_factorizer.cpp:849-984 (parallel_factorize_dna_w_rc): ONE index built serially (:896-922 --
 * construct_im, fwd_starts/rc_ends, the two RMQs), then the factorized range is cut into num_threads chunks
 * (:86-144; at least MIN_CHARS_PER_THREAD = 100000 characters per thread, parallel_factorizer.hpp:55); every
 * thread follows the factor chain from its chunk start and the chains are merged where they CONVERGE (:499-569):
 * f(i) is a pure function of the text, so once two chains meet at a position they coincide from there on.  Here
 * every thread walks its chunk into its own buffer; the merge keeps thread 0's chain, continues serially from the
 * position where it leaves a chunk until that position is a factor start of the next thread's chain, then switches
 * to that chain.  Output is identical to the serial function (asserted by the reference's
 * tests/test_parallel_fasta.py:294-330 and by tests/test_oracle.py here). */
#include <pthread.h>
typedef struct {
    const rc_ctx_t *c;
    uint64_t begin, end;       /* chain starts at begin, stops at the first position >= end */
    sink_t sk;
    uint64_t exit_pos;
    int rc;
} rc_job_t;

static void *rc_job_run(void *arg) {
    rc_job_t *j = (rc_job_t *)arg;
    uint64_t i = j->begin;
    while (i < j->end) {
        uint64_t l, r;
        rc_factor_at(j->c, i, &l, &r);
        if (sink_push(&j->sk, i, l, r)) { j->rc = -1; return NULL; }
        i += l;
    }
    j->exit_pos = i;
    return NULL;
}

/* index of the factor of job j that starts exactly at pos, or -1 */
static int64_t job_find(const rc_job_t *j, uint64_t pos) {
    int64_t lo = 0, hi = (int64_t)j->sk.count - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        uint64_t s = j->sk.v[3 * mid];
        if (s == pos) return mid;
        if (s < pos) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

int nlzo_parallel_factorize_dna_w_rc(const uint8_t *S, uint64_t len_S, uint64_t start_pos, int num_threads,
                                     uint64_t **out, uint64_t *count, int *threads_used) {
    *out = NULL; *count = 0;
    if (threads_used) *threads_used = 0;
    if (len_S < 4) return 0;
    const uint64_t N = len_S / 2 - 1;
    if (N == 0) return 0;
    if (start_pos >= N) return 1;
    index_t ix;
    double t0 = now_s();
    int rc = index_build(S, len_S, &ix);                       /* serial, as in the reference (:78-84, :896-922) */
    if (rc) return rc;
    double t1 = now_s();
    rc_ctx_t c = {&ix, S, len_S, N};
    uint64_t span = N - start_pos;
    int T = num_threads > 0 ? num_threads : 1;
    if ((uint64_t)T > span / 100000) T = (int)(span / 100000);  /* MIN_CHARS_PER_THREAD */
    if (T < 1) T = 1;
    if (threads_used) *threads_used = T;
    rc_job_t *jobs = (rc_job_t *)calloc((size_t)T, sizeof(rc_job_t));
    pthread_t *th = (pthread_t *)calloc((size_t)T, sizeof(pthread_t));
    if (!jobs || !th) { free(jobs); free(th); index_free(&ix); return -1; }
    for (int t = 0; t < T; ++t) {
        jobs[t].c = &c;
        jobs[t].begin = start_pos + span * (uint64_t)t / (uint64_t)T;
        jobs[t].end = t + 1 == T ? N : start_pos + span * (uint64_t)(t + 1) / (uint64_t)T;
    }
    for (int t = 1; t < T; ++t) pthread_create(&th[t], NULL, rc_job_run, &jobs[t]);
    rc_job_run(&jobs[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
    sink_t sk = {0, 0, 0};
    int fail = 0;
    for (int t = 0; t < T; ++t) fail |= jobs[t].rc;
    /* merge: chain position `pos`; `cur` = the job whose chain we are copying (from factor index `from`) */
    if (!fail) {
        uint64_t pos = start_pos;
        int t = 0;
        while (pos < N && !fail) {
            /* the chunk that holds pos */
            while (t + 1 < T && pos >= jobs[t + 1].begin) ++t;
            int64_t k = job_find(&jobs[t], pos);
            if (k >= 0) {                                      /* converged with job t: take the rest of its chain */
                for (uint64_t q = (uint64_t)k; q < jobs[t].sk.count && !fail; ++q) {
                    const uint64_t *f = jobs[t].sk.v + 3 * q;
                    fail |= sink_push(&sk, f[0], f[1], f[2]);
                }
                pos = jobs[t].exit_pos;
            } else {                                           /* not yet: one more factor of the incoming chain */
                uint64_t l, r;
                rc_factor_at(&c, pos, &l, &r);
                fail |= sink_push(&sk, pos, l, r);
                pos += l;
            }
        }
    }
    for (int t = 0; t < T; ++t) free(jobs[t].sk.v);
    free(jobs); free(th);
    index_free(&ix);
    g_last_index_s = t1 - t0; g_last_walk_s = now_s() - t1;
    if (fail) { free(sk.v); return -1; }
    *out = sk.v; *count = sk.count;
    return 0;
}

void nlzo_free(void *p) { free(p); }

/* LCP array (Kasai) of an int32 string s[0..n) for a GIVEN suffix array: lcp[k] = LCP(sa[k-1], sa[k]),
 * lcp[0] = 0.  Used by the stage-parity tests to check the GPU LCP array under the GPU's own
 * symbol order. */
int nlzo_lcp_from_sa_i32(const int32_t *s, int32_t n, const int32_t *sa, int32_t *lcp) {
    idx_t *isa = (idx_t *)malloc(sizeof(idx_t) * (size_t)n);
    if (!isa) return -1;
    for (idx_t k = 0; k < n; ++k) isa[sa[k]] = k;
    idx_t l = 0;
    for (idx_t i = 0; i < n; ++i) {
        idx_t r = isa[i];
        if (r == 0) { lcp[0] = 0; l = 0; continue; }
        idx_t j = sa[r - 1];
        while (i + l < n && j + l < n && s[i + l] == s[j + l]) ++l;
        lcp[r] = l;
        if (l > 0) --l;
    }
    free(isa);
    return 0;
}
