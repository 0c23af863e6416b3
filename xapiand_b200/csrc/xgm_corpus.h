/* xgm_corpus.h — seeded synthetic Zipfian corpus (header-only C, usable from C, C++ and nvcc).
 *
 * The corpus is the one BASELINE.md §3 / SURVEY.md §8(d) describe: N documents (docid 1..N),
 * V terms "T%06u" (rank 0 = most frequent), document length U[16,112] tokens, each token's rank
 * drawn from Zipf(s=1) over V.  wdf(term, doc) = multiplicity of the term in the doc, so the
 * Xapian doclength (sum of wdf) equals the number of tokens drawn.
 *
 * Every document is generated from its own counter-based RNG stream keyed by (seed, docid), so any
 * shard / docid range can be generated independently and in parallel, and the reference-side writer
 * (oracle/ref_runner.cc, which feeds Xapian::WritableDatabase) and the product-side index builder
 * (xgm_build_synthetic in xgm_host.cc) see exactly the same postings.
 *
 * Zipf sampling uses a Walker alias table built with +,-,*,/ on IEEE doubles only (no libm), so the
 * table — and hence the corpus — is bit-reproducible wherever IEEE-754 double arithmetic is.
 */
#ifndef XGM_CORPUS_H
#define XGM_CORPUS_H

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XGM_CORPUS_MIN_LEN 16u
#define XGM_CORPUS_MAX_LEN 112u

typedef struct xgm_zipf {
    uint32_t V;
    uint32_t* thresh; /* accept own index if (u32 draw) < thresh[i] (thresh==0xffffffff → always) */
    uint32_t* alias;
} xgm_zipf;

static inline uint64_t xgm_splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

/* Build the alias table for p_r ∝ 1/(r+1), r in [0,V). Returns 0 on success. */
static inline int xgm_zipf_init(xgm_zipf* z, uint32_t V) {
    z->V = V;
    z->thresh = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)V);
    z->alias = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)V);
    double* q = (double*)malloc(sizeof(double) * (size_t)V);
    uint32_t* small = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)V);
    uint32_t* large = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)V);
    if (!z->thresh || !z->alias || !q || !small || !large) return -1;
    double H = 0.0;
    for (uint32_t r = V; r-- > 0;) H += 1.0 / (double)(r + 1); /* sum small terms first */
    uint32_t ns = 0, nl = 0;
    for (uint32_t r = 0; r < V; ++r) {
        q[r] = (1.0 / (double)(r + 1)) / H * (double)V;
        z->alias[r] = r;
    }
    /* deterministic order: scan ranks descending so stacks pop in a fixed order */
    for (uint32_t r = V; r-- > 0;) {
        if (q[r] < 1.0) small[ns++] = r; else large[nl++] = r;
    }
    while (ns && nl) {
        uint32_t s = small[--ns], l = large[--nl];
        z->alias[s] = l;
        q[l] = (q[l] + q[s]) - 1.0;
        if (q[l] < 1.0) small[ns++] = l; else large[nl++] = l;
    }
    while (nl) q[large[--nl]] = 1.0;
    while (ns) q[small[--ns]] = 1.0;
    for (uint32_t r = 0; r < V; ++r) {
        double t = q[r] * 4294967296.0;
        z->thresh[r] = (t >= 4294967295.0) ? 0xffffffffu : (uint32_t)t;
    }
    free(q); free(small); free(large);
    return 0;
}

static inline void xgm_zipf_free(xgm_zipf* z) {
    free(z->thresh); free(z->alias);
    z->thresh = z->alias = NULL;
}

static inline uint32_t xgm_zipf_draw(const xgm_zipf* z, uint64_t x) {
    /* high 32 bits pick the column (Lemire range reduction), low 32 bits the accept test */
    uint32_t col = (uint32_t)(((x >> 32) * (uint64_t)z->V) >> 32);
    uint32_t u = (uint32_t)x;
    return (z->thresh[col] == 0xffffffffu || u < z->thresh[col]) ? col : z->alias[col];
}

/* Token ranks of document `docid` (1-based, GLOBAL docid). `out` must hold XGM_CORPUS_MAX_LEN
 * entries. Returns the document length (number of tokens). */
static inline uint32_t xgm_corpus_doc(const xgm_zipf* z, uint64_t seed, uint32_t docid, uint32_t* out) {
    uint64_t s = seed ^ ((uint64_t)docid * 0xd1342543de82ef95ULL);
    (void)xgm_splitmix64(&s);
    uint32_t len = XGM_CORPUS_MIN_LEN +
                   (uint32_t)((xgm_splitmix64(&s) >> 32) % (XGM_CORPUS_MAX_LEN - XGM_CORPUS_MIN_LEN + 1));
    for (uint32_t i = 0; i < len; ++i) out[i] = xgm_zipf_draw(z, xgm_splitmix64(&s));
    return len;
}

/* Collapse token ranks into sorted (rank, wdf) pairs in place. `ranks` is overwritten with the
 * distinct ranks ascending, `wdf` (same capacity) receives the multiplicities. Returns #distinct. */
static inline uint32_t xgm_corpus_collapse(uint32_t* ranks, uint32_t len, uint32_t* wdf) {
    /* insertion sort: len <= 112 */
    for (uint32_t i = 1; i < len; ++i) {
        uint32_t v = ranks[i];
        uint32_t j = i;
        while (j > 0 && ranks[j - 1] > v) { ranks[j] = ranks[j - 1]; --j; }
        ranks[j] = v;
    }
    uint32_t n = 0;
    for (uint32_t i = 0; i < len;) {
        uint32_t j = i + 1;
        while (j < len && ranks[j] == ranks[i]) ++j;
        ranks[n] = ranks[i];
        wdf[n] = j - i;
        ++n;
        i = j;
    }
    return n;
}

/* Per-document synthetic values for BASELINE config C5:
 *   slot 0: 1..3 integer values U[0,1e6), returned sorted ascending in v0[0..n0)
 *   slot 1: one integer sort value U[0,1e6)
 */
static inline uint32_t xgm_corpus_values(uint64_t seed, uint32_t docid, uint64_t v0[3], uint64_t* v1) {
    uint64_t s = (seed + 0x5851f42d4c957f2dULL) ^ ((uint64_t)docid * 0x9e3779b97f4a7c15ULL);
    (void)xgm_splitmix64(&s);
    uint32_t n0 = 1 + (uint32_t)((xgm_splitmix64(&s) >> 32) % 3);
    for (uint32_t i = 0; i < n0; ++i) v0[i] = (xgm_splitmix64(&s) >> 11) % 1000000ULL;
    for (uint32_t i = 1; i < n0; ++i) {
        uint64_t v = v0[i];
        uint32_t j = i;
        while (j > 0 && v0[j - 1] > v) { v0[j] = v0[j - 1]; --j; }
        v0[j] = v;
    }
    *v1 = (xgm_splitmix64(&s) >> 11) % 1000000ULL;
    return n0;
}

/* Term name for a rank: "T%06u" (7 bytes + NUL for rank < 1e6). */
static inline int xgm_corpus_term(uint32_t rank, char out[16]) {
    out[0] = 'T';
    char tmp[12];
    int n = 0;
    uint32_t r = rank;
    do { tmp[n++] = (char)('0' + r % 10); r /= 10; } while (r);
    int pad = n < 6 ? 6 - n : 0;
    int k = 1;
    for (int i = 0; i < pad; ++i) out[k++] = '0';
    while (n) out[k++] = tmp[--n];
    out[k] = 0;
    return k;
}

#ifdef __cplusplus
}
#endif
#endif /* XGM_CORPUS_H */
