/* xgm_kernels.cu — sm_100a kernels of the matcher hot path.
 *
 *   xgm_and_kernel    decode + warp-galloping intersection + fused BM25 (reference:
 *                     GlassPostList::next/skip_to glass_postlist.cc:768-991, MultiAndPostList::
 *                     find_next_match multiandpostlist.cc:179-206, get_weight :149-159,
 *                     BM25Weight::get_sumpart bm25weight.cc:170-181, doclen fetch postlisttree.h:184-195)
 *   xgm_topk_kernel   ProtoMSet top-k + match counting (protomset.h:295-400,484-683; msetcmp.cc:54-98)
 *   xgm_decode_kernel round-trip decode of one term (index self-check)
 *
 * Integer/pointer work: no tensor cores. Packed blocks are staged global→shared with the bulk-copy
 * engine (cp.async.bulk + mbarrier, SASS UBLKCP), deltas are undone with a warp prefix sum, every
 * f64 operation of BM25 is an explicit round-to-nearest intrinsic so nothing is contracted into an
 * FMA and the weights are bit-identical to the reference's x86-64 build.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "xgm_device.h"

#define FULL 0xffffffffu

/* ------------------------------------------------------------------ PTX helpers */

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "XGM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra XGM_DONE;\n"
        "bra XGM_WAIT;\n"
        "XGM_DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

/* ------------------------------------------------------------------ block decode */

#define STAGE_WORDS 132 /* 128 packed words at 32 bits + slack for the funnel-shift's high word; 16B multiple */

__device__ __forceinline__ uint32_t bitmask(uint32_t bits) { return bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u); }

__device__ __forceinline__ uint32_t unpack_sm(const uint32_t* st, uint32_t idx, uint32_t bits, uint32_t mask) {
    uint32_t o = idx * bits;
    uint32_t wi = o >> 5, sh = o & 31;
    return __funnelshift_r(st[wi], st[wi + 1], sh) & mask;
}

/* random access into a packed global block (rare path: only for surviving documents) */
__device__ __forceinline__ uint32_t unpack_gl(const uint4* col, uint32_t off16, uint32_t idx, uint32_t bits) {
    if (bits == 0) return 0;
    const uint32_t* base = reinterpret_cast<const uint32_t*>(col + off16);
    uint32_t o = idx * bits;
    uint32_t wi = o >> 5, sh = o & 31;
    uint32_t lo = __ldg(base + wi);
    uint32_t hi = (sh + bits > 32) ? __ldg(base + wi + 1) : 0u;
    return __funnelshift_r(lo, hi, sh) & bitmask(bits);
}

/* Stage one packed block (16*bits bytes) into the warp's shared buffer with the bulk-copy engine. */
__device__ __forceinline__ void stage_block(const uint4* col, uint32_t off16, uint32_t bits, uint32_t* st,
                                            uint64_t* bar, uint32_t& phase, uint32_t lane) {
    __syncwarp();
    if (bits == 0) return;
    if (lane == 0) {
        mbar_expect_tx(bar, bits * 16u);
        bulk_g2s(st, col + off16, bits * 16u, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
}

/* Decode the 128 docids of a staged block: lane l gets postings 4l..4l+3 (ascending). */
__device__ __forceinline__ void decode_docids(const uint32_t* st, uint32_t bits, uint32_t first, uint32_t lane,
                                              uint32_t d[4]) {
    uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    if (bits) {
        uint32_t mask = bitmask(bits);
        v0 = unpack_sm(st, 4 * lane + 0, bits, mask);
        v1 = unpack_sm(st, 4 * lane + 1, bits, mask);
        v2 = unpack_sm(st, 4 * lane + 2, bits, mask);
        v3 = unpack_sm(st, 4 * lane + 3, bits, mask);
    }
    /* docid_i = docid_{i-1} + v_i + 1, docid_0 = first */
    uint32_t s0 = v0 + (lane ? 1u : 0u);
    uint32_t s1 = s0 + v1 + 1u;
    uint32_t s2 = s1 + v2 + 1u;
    uint32_t s3 = s2 + v3 + 1u;
    uint32_t t = s3;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(FULL, t, o);
        if ((int)lane >= o) t += n;
    }
    uint32_t base = first + (t - s3);
    d[0] = base + s0;
    d[1] = base + s1;
    d[2] = base + s2;
    d[3] = base + s3;
}

/* ------------------------------------------------------------------ BM25 */

/* BM25Weight::get_sumpart, bm25weight.cc:170-181, operation for operation:
 *   normlen = max(len * len_factor, min_normlen)
 *   denom   = k1 * (normlen * b + (1 - b)) + wdf
 *   return termweight * (wdf / denom)                                            */
__device__ __forceinline__ double bm25_sumpart(double termweight, const XgmDevQuery* q, uint32_t wdf, uint32_t len) {
    double normlen = __dmul_rn((double)len, q->len_factor);
    normlen = normlen < q->min_normlen ? q->min_normlen : normlen;
    double wdf_d = (double)wdf;
    double t = __dadd_rn(__dmul_rn(normlen, q->b), q->one_minus_b);
    double denom = __dadd_rn(__dmul_rn(q->k1, t), wdf_d);
    return __dmul_rn(termweight, __ddiv_rn(wdf_d, denom));
}

/* ------------------------------------------------------------------ skip-table search */

/* Largest block index i in [cur, n) with hdr[i].first <= target, given the sentinel at hdr[n].
 * Returns cur when hdr[cur].first > target (caller checks). Warp-galloping: first look at the next 32
 * headers (sequential access pattern of a leapfrog), then a 32-ary search over the rest. */
__device__ __forceinline__ uint32_t warp_seek(const XgmBlockHdr* __restrict__ hdr, uint32_t cur, uint32_t n,
                                              uint32_t target, uint32_t lane) {
    uint32_t i = cur + lane;
    uint32_t f = (i <= n) ? __ldg(&hdr[i].first) : XGM_SENTINEL;
    uint32_t cnt = __popc(__ballot_sync(FULL, f <= target));
    if (cnt < 32) return cnt ? cur + cnt - 1 : cur;
    uint32_t lo = cur + 31, hi = n; /* hdr[lo].first <= target < hdr[hi].first (sentinel) */
    while (hi - lo > 1) {
        uint32_t span = hi - lo;
        uint32_t step = (span + 31) / 32;
        uint32_t p = lo + (lane + 1) * step;
        if (p > hi) p = hi;
        uint32_t fp = __ldg(&hdr[p].first);
        uint32_t c = __popc(__ballot_sync(FULL, fp <= target));
        uint32_t nlo = lo + c * step;
        uint32_t nhi = lo + (c + 1) * step;
        if (nlo > hi) nlo = hi; /* cannot happen: hdr[hi] > target */
        if (nhi > hi) nhi = hi;
        lo = nlo;
        hi = nhi;
    }
    return lo;
}

/* ------------------------------------------------------------------ value-slot predicates (C5) */

__device__ __forceinline__ bool doc_passes_filter(const XgmKernelParams& p, const XgmDevQuery* q, uint32_t did) {
    if (q->filter == 0) return true;
    const XgmDevSlot& s = p.slots[q->filter_slot];
    if (!s.voff) return false;
    uint32_t a = __ldg(&s.voff[did]), b = __ldg(&s.voff[did + 1]);
    if (a == b) return false;
    if (q->filter == 1) { /* stock OP_VALUE_RANGE on the (first) value, valuerangepostlist.cc:132-151 */
        uint64_t v = __ldg(&s.vals[a]);
        return v >= q->range_lo && v <= q->range_hi;
    }
    /* MultipleValueRange::insideRange, src/multivalue/range.cc:351-368: first value >= lo must be <= hi */
    for (uint32_t i = a; i < b; ++i) {
        uint64_t v = __ldg(&s.vals[i]);
        if (v >= q->range_lo) return v <= q->range_hi;
    }
    return false;
}

__device__ __forceinline__ uint64_t doc_sort_key(const XgmKernelParams& p, const XgmDevQuery* q, uint32_t did) {
    if (q->sort_by == 0) return 0;
    const XgmDevSlot& s = p.slots[q->sort_slot];
    if (!s.voff) return 0;
    uint32_t a = __ldg(&s.voff[did]), b = __ldg(&s.voff[did + 1]);
    if (a == b) return 0;
    return __ldg(&s.vals[q->sort_use_max ? b - 1 : a]);
}

/* ------------------------------------------------------------------ sparse AND kernel */

#define AND_WARPS 8

struct __align__(16) WarpScratch {
    uint32_t stage[STAGE_WORDS]; /* packed words of the block being decoded */
    uint32_t dbuf[XGM_BLOCK];    /* decoded docids of the probed block */
    uint64_t bar;
    uint64_t pad;
};

__global__ void __launch_bounds__(AND_WARPS * 32) xgm_and_kernel(XgmKernelParams p) {
    __shared__ WarpScratch scratch[AND_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    if (lane == 0) mbar_init(&ws.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;

    const XgmBlockHdr* __restrict__ hdr = p.hdr;

    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(p.work_counter, 1u);
        item = __shfl_sync(FULL, item, 0);
        if (item >= p.nitems) break;
        const XgmWorkItem wi = p.items[item];
        const XgmDevQuery* q = &p.queries[wi.query];
        const uint32_t nterms = q->nterms;
        /* lane j keeps list j's skip-table cursor */
        uint32_t my_begin = 0, my_nblk = 0, my_cur = 0;
        if (lane < nterms) {
            my_begin = q->terms[lane].blk_begin;
            my_nblk = q->terms[lane].nblocks;
        }
        const uint32_t drv_begin = __shfl_sync(FULL, my_begin, 0);
        const double tw0 = q->terms[0].termweight;

        for (uint32_t db = wi.b0; db < wi.b1; ++db) {
            const XgmBlockHdr dh = hdr[drv_begin + db];
            const uint32_t dcount = XGM_HDR_COUNT(dh.meta);
            uint32_t c[4];
            stage_block(p.docs, dh.doc_off, XGM_HDR_DOC_BITS(dh.meta), ws.stage, &ws.bar, phase, lane);
            decode_docids(ws.stage, XGM_HDR_DOC_BITS(dh.meta), dh.first, lane, c);
            uint32_t alive = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (4 * lane + k < dcount) alive |= 1u << k;
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            uint32_t dl[4] = {0, 0, 0, 0};

            if (nterms == 1) {
                /* single term: every posting matches; decode the wdf block too */
                stage_block(p.tfs, dh.tf_off, XGM_HDR_TF_BITS(dh.meta), ws.stage, &ws.bar, phase, lane);
                const uint32_t tb = XGM_HDR_TF_BITS(dh.meta);
                const uint32_t tmask = bitmask(tb);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (alive >> k & 1u) {
                        uint32_t tf = tb ? unpack_sm(ws.stage, 4 * lane + k, tb, tmask) : 0u;
                        dl[k] = __ldg(&p.doclen[c[k]]);
                        acc[k] = bm25_sumpart(tw0, q, tf, dl[k]);
                    }
                }
            }

            for (uint32_t j = 1; j < nterms; ++j) {
                if (!__any_sync(FULL, alive != 0)) break;
                const uint32_t lbegin = __shfl_sync(FULL, my_begin, j);
                const uint32_t lnblk = __shfl_sync(FULL, my_nblk, j);
                uint32_t cur = __shfl_sync(FULL, my_cur, j);
                const XgmBlockHdr* lh = hdr + lbegin;
                const double twj = q->terms[j].termweight;
                uint32_t unresolved = alive;
                for (;;) {
                    /* smallest unresolved candidate across the warp (candidates ascend with k, lane) */
                    uint32_t m = (unresolved & 1u) ? c[0] : (unresolved & 2u) ? c[1] : (unresolved & 4u) ? c[2]
                                 : (unresolved & 8u) ? c[3] : XGM_SENTINEL;
                    const uint32_t tmin = __reduce_min_sync(FULL, m);
                    if (tmin == XGM_SENTINEL) break;
                    if (cur >= lnblk) { /* list exhausted: nothing else can match */
                        alive &= ~unresolved;
                        unresolved = 0;
                        break;
                    }
                    cur = warp_seek(lh, cur, lnblk, tmin, lane);
                    const XgmBlockHdr bh = lh[cur];
                    const uint32_t next_first = __ldg(&lh[cur + 1].first);
                    if (tmin < bh.first) {
                        /* candidates below this block's first docid fall in a gap: dead */
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if ((unresolved >> k & 1u) && c[k] < bh.first) {
                                unresolved &= ~(1u << k);
                                alive &= ~(1u << k);
                            }
                        continue;
                    }
                    /* decode the block and publish its docids for the per-candidate searches */
                    uint32_t bd[4];
                    stage_block(p.docs, bh.doc_off, XGM_HDR_DOC_BITS(bh.meta), ws.stage, &ws.bar, phase, lane);
                    decode_docids(ws.stage, XGM_HDR_DOC_BITS(bh.meta), bh.first, lane, bd);
                    const uint32_t bcount = XGM_HDR_COUNT(bh.meta);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ws.dbuf[4 * lane + k] = (4 * lane + k < bcount) ? bd[k] : XGM_SENTINEL;
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if ((unresolved >> k & 1u) && c[k] < next_first) {
                            const uint32_t cd = c[k];
                            uint32_t pos = 0;
#pragma unroll
                            for (uint32_t s = 64; s >= 1; s >>= 1)
                                if (ws.dbuf[pos + s - 1] < cd) pos += s;
                            unresolved &= ~(1u << k);
                            if (ws.dbuf[pos] == cd) {
                                if (j == 1) {
                                    /* first confirmation: fetch doclen and the driver's own wdf lazily */
                                    dl[k] = __ldg(&p.doclen[cd]);
                                    uint32_t tf0 = unpack_gl(p.tfs, dh.tf_off, 4 * lane + k, XGM_HDR_TF_BITS(dh.meta));
                                    acc[k] = bm25_sumpart(tw0, q, tf0, dl[k]);
                                }
                                uint32_t tfj = unpack_gl(p.tfs, bh.tf_off, pos, XGM_HDR_TF_BITS(bh.meta));
                                /* MultiAndPostList::get_weight: result += plist[i]->get_weight(), in order */
                                acc[k] = __dadd_rn(acc[k], bm25_sumpart(twj, q, tfj, dl[k]));
                            } else {
                                alive &= ~(1u << k);
                            }
                        }
                    }
                    __syncwarp();
                }
                if (lane == j) my_cur = cur;
            }

            /* value-slot filter (OP_FILTER with a range source), applied to the survivors */
            if (q->filter) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if ((alive >> k & 1u) && !doc_passes_filter(p, q, c[k])) alive &= ~(1u << k);
            }

            /* emit matches: warp-aggregated reservation in the query's match buffer */
            if (__any_sync(FULL, alive != 0)) {
                uint32_t n = __popc(alive);
                uint32_t incl = n;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(FULL, incl, o);
                    if ((int)lane >= o) incl += t;
                }
                uint32_t total = __shfl_sync(FULL, incl, 31);
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&p.match_count[wi.query], total);
                base = __shfl_sync(FULL, base, 0) + (incl - n);
                const size_t qoff = (size_t)wi.query * p.match_cap;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (alive >> k & 1u) {
                        if (base < p.match_cap) {
                            p.match_w[qoff + base] = acc[k];
                            p.match_d[qoff + base] = c[k];
                            if (q->sort_by) p.match_k[qoff + base] = doc_sort_key(p, q, c[k]);
                        }
                        ++base;
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ top-k / ProtoMSet */

/* Strict total order of the reference's comparators (msetcmp.cc:54-98) with ascending docid order:
 * returns true when a ranks before b. */
__device__ __forceinline__ bool ranks_before(uint32_t sort_by, uint32_t reverse, double wa, uint32_t da, uint64_t ka,
                                             double wb, uint32_t db, uint64_t kb) {
    if (sort_by == 1 || sort_by == 2) {
        if (ka > kb) return reverse != 0;
        if (ka < kb) return reverse == 0;
        if (sort_by == 2) return da < db;
    }
    if (wa > wb) return true;
    if (wa < wb) return false;
    if (sort_by == 3) {
        if (ka > kb) return reverse != 0;
        if (ka < kb) return reverse == 0;
    }
    return da < db;
}

#define TOPK_THREADS 128

/* One CTA per query. Rank-sort: every match counts how many matches rank before it; ranks < topk
 * are written to their final position. The same pass counts what ProtoMSet::add would have counted
 * in known_matching_docs while walking the matches in docid order (protomset.h:340-400 together
 * with the `weight < min_weight → continue` of matcher.cc:496-498): a match is counted iff it is
 * among the first max(check_at_least, topk+1) in docid order or fewer than topk earlier matches have
 * a strictly greater weight. */
__global__ void __launch_bounds__(TOPK_THREADS) xgm_topk_kernel(XgmKernelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t qi = blockIdx.x;
    const XgmDevQuery* q = &p.queries[qi];
    if (q->route != 0) return;
    const uint32_t total = p.match_count[qi];
    const uint32_t n = total < p.match_cap ? total : p.match_cap;
    double* sw = reinterpret_cast<double*>(smem_raw);
    uint64_t* sk = reinterpret_cast<uint64_t*>(sw + p.match_cap);
    uint32_t* sd = reinterpret_cast<uint32_t*>(sk + p.match_cap);
    __shared__ uint32_t s_known;
    __shared__ unsigned long long s_maxw;
    const size_t qoff = (size_t)qi * p.match_cap;
    const uint32_t sort_by = q->sort_by, reverse = q->sort_reverse;
    for (uint32_t i = threadIdx.x; i < n; i += TOPK_THREADS) {
        sw[i] = p.match_w[qoff + i];
        sd[i] = p.match_d[qoff + i];
        sk[i] = sort_by ? p.match_k[qoff + i] : 0ull;
    }
    if (threadIdx.x == 0) { s_known = 0; s_maxw = 0ull; }
    __syncthreads();
    const uint32_t topk = q->topk;
    double local_max = 0.0;
    const uint32_t free_count = q->check_at_least > topk + 1 ? q->check_at_least : topk + 1;
    uint32_t known = 0;
    const size_t ooff = (size_t)qi * p.out_stride;
    for (uint32_t i = threadIdx.x; i < n; i += TOPK_THREADS) {
        const double wi = sw[i];
        const uint32_t di = sd[i];
        const uint64_t ki = sk[i];
        uint32_t rank = 0, before = 0, greater_before = 0;
        for (uint32_t j = 0; j < n; ++j) {
            const double wj = sw[j];
            const uint32_t dj = sd[j];
            rank += ranks_before(sort_by, reverse, wj, dj, sk[j], wi, di, ki) ? 1u : 0u;
            const bool earlier = dj < di;
            before += earlier ? 1u : 0u;
            greater_before += (earlier && wj > wi) ? 1u : 0u;
        }
        if (rank < topk) {
            p.out_w[ooff + rank] = wi;
            p.out_d[ooff + rank] = di;
            if (p.out_k) p.out_k[ooff + rank] = ki;
        }
        if (sort_by == 1 || sort_by == 2 || before < free_count || greater_before < topk) ++known;
        local_max = wi > local_max ? wi : local_max;
    }
    atomicAdd(&s_known, known);
    /* weights are >= 0, so the IEEE bit pattern orders like the value */
    atomicMax(&s_maxw, (unsigned long long)__double_as_longlong(local_max));
    __syncthreads();
    if (threadIdx.x == 0) {
        XgmDevResult r;
        r.n = n < topk ? n : topk;
        r.exact = total;
        r.known = s_known;
        r.flags = total > p.match_cap ? 1u : 0u;
        r.max_w = __longlong_as_double((long long)s_maxw);
        r.max_subqs = q->nterms; /* AND: every leaf matches (count_matching_subqs) */
        r.pad = 0;
        p.out_info[qi] = r;
    }
}

/* ------------------------------------------------------------------ decode (round-trip check) */

__global__ void __launch_bounds__(AND_WARPS * 32) xgm_decode_kernel(XgmKernelParams p, uint32_t blk_begin,
                                                                    uint32_t nblocks, uint32_t* out_d,
                                                                    uint32_t* out_w) {
    __shared__ WarpScratch scratch[AND_WARPS];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    if (lane == 0) mbar_init(&ws.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;
    for (uint32_t b = blockIdx.x * AND_WARPS + warp; b < nblocks; b += gridDim.x * AND_WARPS) {
        const XgmBlockHdr h = p.hdr[blk_begin + b];
        const uint32_t count = XGM_HDR_COUNT(h.meta);
        uint32_t d[4];
        stage_block(p.docs, h.doc_off, XGM_HDR_DOC_BITS(h.meta), ws.stage, &ws.bar, phase, lane);
        decode_docids(ws.stage, XGM_HDR_DOC_BITS(h.meta), h.first, lane, d);
        stage_block(p.tfs, h.tf_off, XGM_HDR_TF_BITS(h.meta), ws.stage, &ws.bar, phase, lane);
        const uint32_t tb = XGM_HDR_TF_BITS(h.meta);
        const uint32_t tmask = bitmask(tb);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t idx = 4 * lane + k;
            if (idx < count) {
                out_d[(size_t)b * XGM_BLOCK + idx] = d[k];
                out_w[(size_t)b * XGM_BLOCK + idx] = tb ? unpack_sm(ws.stage, idx, tb, tmask) : 0u;
            }
        }
    }
}

/* ------------------------------------------------------------------ launchers */

cudaError_t xgm_launch_and(const XgmKernelParams& p, int grid, cudaStream_t s) {
    xgm_and_kernel<<<grid, AND_WARPS * 32, 0, s>>>(p);
    return cudaGetLastError();
}

size_t xgm_topk_smem_bytes(uint32_t match_cap) { return (size_t)match_cap * (8 + 8 + 4); }

cudaError_t xgm_launch_topk(const XgmKernelParams& p, uint32_t nq, cudaStream_t s) {
    size_t smem = xgm_topk_smem_bytes(p.match_cap);
    static bool attr_set = false;
    if (!attr_set && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(xgm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    xgm_topk_kernel<<<nq, TOPK_THREADS, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t xgm_launch_decode(const XgmKernelParams& p, uint32_t blk_begin, uint32_t nblocks, uint32_t* out_d,
                              uint32_t* out_w, cudaStream_t s) {
    int grid = (int)((nblocks + AND_WARPS - 1) / AND_WARPS);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    xgm_decode_kernel<<<grid, AND_WARPS * 32, 0, s>>>(p, blk_begin, nblocks, out_d, out_w);
    return cudaGetLastError();
}

int xgm_and_occupancy_blocks_per_sm() {
    int n = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, xgm_and_kernel, AND_WARPS * 32, 0);
    return n;
}
