/* xgm_device.h — kernel parameter block and launcher prototypes (host <-> kernels, internal). */
#ifndef XGM_DEVICE_H
#define XGM_DEVICE_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "xgm_format.h"

#define XGM_MAX_SLOTS 8

struct XgmDevSlot {
    const uint32_t* voff; /* [lastdocid+2] CSR offsets, NULL = slot absent */
    const uint64_t* vals; /* order-preserving numeric keys, ascending per doc */
};

struct XgmDevResult {
    uint32_t n;      /* entries written to out_* (<= topk) */
    uint32_t exact;  /* documents matching the boolean structure */
    uint32_t known;  /* ProtoMSet::known_matching_docs */
    uint32_t flags;  /* bit0: candidates were lost (buffer overflow); bit1: match-count bounds approximate */
    double max_w;    /* dense kernel: best weight seen */
    uint32_t max_subqs;
    uint32_t pad;
};

struct XgmKernelParams {
    /* index */
    const XgmBlockHdr* hdr;
    const uint4* docs;
    const uint4* tfs;
    const uint32_t* doclen;
    const uint32_t* bitmaps;
    const uint32_t* ranks;
    uint32_t lastdocid;
    uint32_t doclen_lb;           /* Database::get_doclength_lower_bound: block-level weight bounds */
    XgmDevSlot slots[XGM_MAX_SLOTS];
    /* batch */
    const XgmDevQuery* queries;
    const XgmWorkItem* items;     /* AND kernel work list */
    const XgmWorkItem* items_or;  /* OR kernel work list */
    const XgmWorkItem* items_bm;  /* bitmap AND kernel work list */
    uint32_t nitems, nitems_or, nitems_bm;
    const uint32_t* nitems_bm_dev; /* non-null: item count of the bitmap AND list lives on the device (range-major expansion) */
    uint32_t nq;
    uint32_t* work_counter;       /* [10],[11] OR items of xgm_or3_kernel (first / second pass); [12],[13] tiles of xgm_or_tile_kernel; [14],[15] items of xgm_or3_kernel<PHASE 1>; [8..9] 64-bit pool reservation; [0] AND items, [1] OR items, [2],[3] same for the second pass, [4] #queries to re-run, [5] overflow-pool entries reserved, [6],[7] bitmap-AND items (first / second pass) */
    uint32_t pass;                /* 0 = first pass, 1 = re-run of overflowed queries with their exact b* */
    XgmQState* qstate;            /* [nq] */
    uint32_t* hist;               /* [nq][XGM_NBINS] */
    /* per-query match buffers */
    uint32_t match_cap;
    uint32_t keep_cap;            /* survivors the top-k kernel can rank in shared memory */
    double* match_w;
    uint32_t* match_d;
    uint64_t* match_k;
    /* overflow pool: second-pass candidates of queries whose first-pass buffer overflowed */
    uint32_t pool_total;
    double* pool_w;
    uint32_t* pool_d;
    uint64_t* pool_k;
    /* results */
    uint32_t out_stride;
    double* out_w;
    uint32_t* out_d;
    uint64_t* out_k;
    XgmDevResult* out_info;
    XgmRaise* raise_log;          /* [nq][XGM_RAISE_LOG] */
    const uint32_t* tileq;        /* queries answered by xgm_or_tile_kernel + xgm_or3_kernel<PHASE 1> */
    uint32_t ntileq;
    /* top-k: queries xgm_topk_small_kernel left to xgm_topk_kernel (work_counter[16] expensive ones from the front
     * of the list, [19] others from its back); null = every query of the batch.  work_counter[17 + pass] hands the
     * entries out. */
    uint32_t* topk_list;
};
#define XGM_CTRL_HDR 128 /* bytes of work counters in front of the per-query state */

cudaError_t xgm_launch_and(const XgmKernelParams& p, int grid, cudaStream_t s);
cudaError_t xgm_launch_and2(const XgmKernelParams& p, int grid, cudaStream_t s);
int xgm_and2_occupancy_blocks_per_sm();
cudaError_t xgm_launch_and_bm(const XgmKernelParams& p, int grid, cudaStream_t s);
int xgm_and_bm_occupancy_blocks_per_sm();
cudaError_t xgm_launch_expand(const XgmWorkItem* seg, uint32_t nseg, uint32_t total, const uint32_t* level_start,
                              uint32_t nlevels, uint32_t bpi, XgmWorkItem* out, cudaStream_t s);
cudaError_t xgm_launch_expand_ranges(const XgmWorkItem* seg, uint32_t nseg, const XgmDevQuery* queries, const XgmBlockHdr* hdr,
                                     uint32_t nranges, uint32_t range_bits, uint32_t bpi, uint32_t* lo, uint32_t* cnt,
                                     uint32_t* off, uint32_t* total, XgmWorkItem* out, cudaStream_t s);
cudaError_t xgm_launch_or(const XgmKernelParams& p, int grid, cudaStream_t s);
int xgm_or_occupancy_blocks_per_sm();
cudaError_t xgm_launch_or3(const XgmKernelParams& p, int grid, int phase, cudaStream_t s);
cudaError_t xgm_launch_or_tile(const XgmKernelParams& p, int grid, cudaStream_t s);
int xgm_or_tile_occupancy_blocks_per_sm();
int xgm_or3_occupancy_blocks_per_sm();
cudaError_t xgm_launch_topk(const XgmKernelParams& p, uint32_t nq, cudaStream_t s);
cudaError_t xgm_launch_decode(const XgmKernelParams& p, uint32_t blk_begin, uint32_t nblocks, uint32_t* out_d,
                              uint32_t* out_w, cudaStream_t s);
size_t xgm_topk_smem_bytes(uint32_t keep_cap);
cudaError_t xgm_launch_merge(const double* gw, const uint32_t* gd, const XgmDevResult* ginfo, size_t part_w, size_t part_d,
                             size_t part_info, uint32_t nparts, uint32_t nq, uint32_t stride, uint32_t k, double* out_w, uint32_t* out_d, uint32_t* out_n,
                             cudaStream_t s);
int xgm_and_occupancy_blocks_per_sm();

#endif
