/* xmem2_b200.h — C ABI of the B200-native XMem++ hot path (libxmem2_b200.so).
 *
 * Every entry point takes plain device pointers + sizes and an explicit cudaStream_t (passed as
 * void*), never allocates, never synchronises, and returns 0 on success or a negative code with
 * the message available from xm_last_error().  No torch types cross this boundary.
 *
 * Reference interfaces replaced (paths relative to the reference repo mbzuai-metaverse/XMem2):
 *   xm_affinity_readout        model/memory_util.py:7-39 (get_similarity) + :41-65 (do_softmax, top-k)
 *                              + inference/memory_manager.py:57-59,61-190 (_readout / match_memory)
 *   xm_query_pack, xm_key_pack model/memory_util.py:22-26 (operand preparation of get_similarity)
 *   xm_conv2d_nhwc             every nn.Conv2d(+BatchNorm2d+ReLU+residual) on the path:
 *                              model/resnet.py:46-114, model/modules.py:22-41,178-211,229-250,
 *                              model/group_modules.py:29-54
 *   xm_* element-wise ops      model/modules.py:63-74,88-99,135-138,166-170,236-247, model/cbam.py:23-77,
 *                              model/group_modules.py:15-26, model/aggregate.py:6-16
 *   xm_resize_argmax           inference/run_on_video.py:165-173 (_post_process), inference/data/mask_mapper.py:56-64
 */
#ifndef XMEM2_B200_H
#define XMEM2_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XM_OK 0
#define XM_ERR_ARG (-1)
#define XM_ERR_CUDA (-2)
#define XM_ERR_WORKSPACE (-3)

#define XM_CK 64        /* key channels  (util/configuration.py:144) */
#define XM_CV 512       /* value channels (util/configuration.py:158) */
#define XM_MAX_GROUPS 8
#define XM_MAX_TOPK 32

const char* xm_last_error(void);
int xm_version(void);
/* diagnostics: {valid, tag, blockIdx.xyz, threadIdx.x, parity} of the last mbarrier-timeout trap */
int xm_debug_last_trap(int* out7);
/* number of kernels this library has launched so far (accounting for bench.py) */
long long xm_launch_count(void);
void xm_add_launch_count(int n);   /* the host mirror reports kernels it replays from a recorded CUDA graph */
/* programmatic dependent launch between this library's kernels (default on) */
void xm_set_pdl(int on);

/* ---------------------------------------------------------------- memory read (K1)
 * xm_affinity_readout launches ONE persistent kernel per object group (one CTA per SM, all co-resident, phases separated by
 * inter-CTA barriers on counters in the workspace): do not run two affinity calls that share a workspace concurrently, and
 * do not share the device with another long-running kernel while it runs. */
/* One memory bank (KeyValueMemoryStore, inference/kv_memory_store.py:4-239) as the kernel sees it. */
typedef struct {
    const void* keys;       /* fp16 [cap][128]: row n = (k[n,:]^2 , k[n,:])  — "packed" keys           */
    const float* shrinkage; /* fp32 [cap]                                                              */
    const void* values;     /* fp16 [n_obj_cap][512][cap]  (reference layout v[g] = [n_g, CV, N_g])    */
    float* usage;           /* fp32 [cap] or NULL; += column sums of the group-0 affinity (memory_util.py:62-63) */
    int64_t cap;            /* allocated columns (multiple of 8)                                       */
    int32_t n_obj_cap;      /* allocated value planes                                                   */
    int32_t size;           /* columns in use                                                           */
} xm_bank_t;

/* An object group reads a SUFFIX of each bank's columns (memory_manager.py:98-128,162-182). */
typedef struct {
    int32_t obj_begin;      /* first value plane (object index, 0-based, background excluded)          */
    int32_t n_obj;          /* planes in this group                                                     */
    int32_t begin[3];       /* first visible column per bank; the range is [begin, bank.size)          */
} xm_group_t;

typedef struct {
    xm_bank_t banks[3];     /* concat order of the reference: long-term, working, permanent (memory_manager.py:82) */
    int32_t n_groups;
    xm_group_t groups[XM_MAX_GROUPS];
    const void* qp;         /* fp16 [hw_pad][128] from xm_query_pack                                    */
    const float* bsq;       /* fp32 [hw_pad]      from xm_query_pack                                    */
    int32_t hw;             /* query positions (h*w)                                                    */
    int32_t hw_pad;         /* hw rounded up to 128                                                     */
    int32_t top_k;          /* <= XM_MAX_TOPK                                                           */
    int32_t n_obj_total;
    void* readout_chw;      /* fp16 [n_obj_total][512][hw]   or NULL (reference layout)                 */
    void* readout_hwc;      /* fp16 [n_obj_total][hw][512]   or NULL (NHWC, feeds the decoder directly) */
    void* workspace;        /* >= xm_affinity_workspace_bytes()                                         */
    int64_t workspace_bytes;
    float* debug_scores;    /* optional fp32 [N_group0][hw_pad] dump of the similarity (tests only)     */
    int32_t plan_is_resident; /* 1: the first 4096 workspace bytes already hold the plan written by xm_affinity_plan
                                 (CUDA-graph replay: bank sizes change without re-recording the launches);
                                 0: build and upload it inside the call                                    */
} xm_affinity_args_t;

int64_t xm_affinity_workspace_bytes(int32_t hw, int32_t n_obj_total);
/* A fresh workspace needs its inter-CTA barrier counters zeroed ONCE before the first xm_affinity_* launch that uses it
 * (the kernel re-arms them on exit, so CUDA-graph replays need nothing). */
int xm_affinity_workspace_init(void* workspace, int64_t workspace_bytes, int32_t hw, int32_t n_obj_total, void* stream);
int xm_affinity_readout(const xm_affinity_args_t* args, void* stream);
/* host only: validate banks/groups and write the 4096-byte column-range plan the kernels read from the head of
 * the workspace (copy it there with any stream-ordered H2D copy).  Launch shapes of xm_affinity_readout depend only
 * on (hw, n_obj_total, group object counts), never on bank sizes.                                          */
int xm_affinity_plan(const xm_affinity_args_t* args, void* host_plan_out, int64_t bytes);
/* diagnostics (synchronous): globaltimer stamps [n_ctas][16] of the phase boundaries of the last launch on this workspace;
 * returns the number of CTAs copied.  Index: 0 start, 1 sweep A done, 2/3 merge A, 4 sweep B done, 5 lists published,
 * 6 merge B done, 7 readout start, 8 readout done, 9 reduce start, 10 end. */
int xm_affinity_debug_timeline(void* workspace, int32_t hw, int32_t n_obj_total, unsigned long long* host_out, int32_t max_ctas);
/* diagnostics, only meaningful when the library was built with -DK1_TRACE: per-tile clock records of CTA 0 ([4][1024] uint64) */
int xm_affinity_debug_counts(void* workspace, int32_t hw, int32_t n_obj_total, int32_t* host_out_per_query);
int xm_affinity_debug_trace(void* workspace, int32_t hw, int32_t n_obj_total, unsigned long long* host_out);

/* T-sharded read: the banks of one long video are split over R ranks by stored frame; every rank
 * has the same query.  The host interleaves the collectives (torch.distributed / NCCL):
 *   stage_a -> all_reduce(MAX) tau_lo[hw_pad] -> stage_b -> all_gather top32[R][hw_pad][32] -> merge (every rank)
 *   -> stage_c -> all_reduce(SUM) readout_f32[n_obj][hw_pad][512] -> cast.   One object group per call (groups[0]);
 * a rank may hold fewer than top_k (even zero) columns.  Same math as xm_affinity_readout (model/memory_util.py:7-65). */
int xm_affinity_tshard_stage_a(const xm_affinity_args_t* args, float* tau_lo_local, void* stream);
int xm_affinity_tshard_stage_b(const xm_affinity_args_t* args, const float* tau_lo_global, float* top32_local, void* stream);
int xm_affinity_tshard_merge(const float* top32_all, int32_t n_ranks, int32_t hw, int32_t hw_pad, int32_t top_k, float* tau, float* inv_den, void* stream);
int xm_affinity_tshard_stage_c(const xm_affinity_args_t* args, const float* tau, const float* inv_den, float* readout_f32, void* stream);
int xm_affinity_tshard_cast(const float* readout_f32, int32_t n_obj, int32_t hw, int32_t hw_pad, void* readout_chw, void* readout_hwc, void* stream);

/* key [hw][64] fp16 (NHWC), selection [hw][64] fp16 -> qp [hw_pad][128] = (-e, 2*k*e), bsq = sum_c e*k^2 */
int xm_query_pack(const void* key_hwc, const void* sel_hwc, int32_t hw, int32_t hw_pad, void* qp, float* bsq, void* stream);
/* key [n][64] fp16 -> rows [n][128] = (k^2, k) written at dst */
int xm_key_pack(const void* key_hwc, int32_t n, void* dst_rows, void* stream);

/* ---------------------------------------------------------------- convolution (K2-K4) */
typedef struct {
    const void* ptr;        /* fp16 NHWC [batch][H][W][C]                                               */
    int32_t channels;       /* multiple of 64                                                           */
    int32_t broadcast;      /* 1: one image shared by every batch entry (MainToGroupDistributor)        */
} xm_conv_src_t;

typedef struct {
    xm_conv_src_t src[3];   /* channel-concatenated inputs (torch.cat(..., dim=channels) without the copy) */
    int32_t n_src;
    int32_t batch, H, W;    /* input spatial size                                                       */
    int32_t ksize;          /* 1 or 3 (7x7 stems go through xm_im2col_stem)                             */
    int32_t stride;         /* 1 or 2 (padding = ksize/2)                                               */
    const void* weight;     /* fp16 [cout_pad][ksize*ksize][cin_total]   (BN folded)                    */
    const float* bias;      /* fp32 [cout_pad]                                                          */
    int32_t cout;           /* real output channels                                                     */
    int32_t cout_pad;       /* multiple of 64                                                           */
    const void* residual;   /* fp16 NHWC [batch|1][Ho][Wo][cout] or NULL, added before the activation   */
    int32_t residual_broadcast;
    int32_t relu;           /* apply ReLU to `out`                                                      */
    void* out;              /* fp16 NHWC [batch][Ho][Wo][out_stride] or NULL                            */
    void* out_relu;         /* optional second copy with ReLU applied (GroupResBlock needs g and relu(g)) */
    int32_t out_stride;     /* channel stride of out/out_relu (>= cout)                                 */
    int32_t out_offset;     /* first channel written                                                    */
    void* workspace;        /* optional split-K scratch: 256 KiB of ZEROED int32 counters followed by fp32 partial
                               tiles; NULL disables split-K.  The kernel leaves the counters zeroed.             */
    int64_t workspace_bytes;
} xm_conv_args_t;

int xm_conv2d_nhwc(const xm_conv_args_t* args, void* stream);

/* ---------------------------------------------------------------- element-wise / small kernels (K5) */
/* 7x7 stride-2 stem patches: image fp32 [3][H][W]; masks fp32 [n][H][W] or NULL (key encoder).
 * out fp16 [n][H/2][W/2][kpad], k = (kh*7+kw)*C + c with C = 3 (image) or 5 (image, mask_b, sum of other masks):
 * model/resnet.py:120 and ValueEncoder.forward model/modules.py:124-135.                                 */
int xm_im2col_stem(const float* image, const float* masks, int32_t n, int32_t H, int32_t W, int32_t kpad, void* out, void* stream);
/* Fused 7x7 / stride-2 stem convolution (+ folded BatchNorm, optional ReLU): the im2col tile is built in shared memory and
 * multiplied on tcgen05 in the same kernel (replaces xm_im2col_stem + a 1x1 xm_conv2d_nhwc on the inference path).
 * Reference: `self.conv1(f)` + `bn1` (+ `relu`) of KeyEncoder (model/modules.py:165-168, model/resnet.py:120) and of ValueEncoder
 * after `torch.cat([image, mask, others], 1)` (model/modules.py:124-137).
 * image fp32 [3][H][W]; masks fp32 [n][H][W] or NULL.  masks == NULL: 3 input channels, n = 1, kpad = 192, K index
 * kh*24 + kw*3 + c; otherwise 5 channels (image, mask[b], sum of the other masks), kpad = 256, K index (kh*7+kw)*5 + c.
 * weight fp16 [64][kpad], bias fp32 [64]; out fp16 NHWC [n][H/2][W/2][64]. */
int xm_stem7x7(const float* image, const float* masks, int32_t n, int32_t H, int32_t W, const void* weight, const float* bias,
               int32_t kpad, int32_t relu, void* out, void* stream);
/* nn.MaxPool2d(3, 2, 1) on NHWC fp16; relu != 0 applies ReLU after the pool (modules.py:137-138).        */
int xm_maxpool3x3s2(const void* in, int32_t B, int32_t H, int32_t W, int32_t C, int32_t relu, void* out, void* stream);
int xm_relu(const void* in, void* out, int64_t n, void* stream);
/* proj fp16 [hw][pstride] = (key 0..63 | d 64 | e 65..128) -> key, selection = sigmoid(e) fp16 [hw][64],
 * shrinkage = d^2+1 fp32 [hw] (modules.py:207-211) and, if qp != NULL, the packed query (see xm_query_pack). */
int xm_keyproj_post(const void* proj, int32_t pstride, int32_t hw, int32_t hw_pad, void* key, void* sel, float* shr,
                    void* qp, float* bsq, void* stream);
/* out = x + CBAM(x) (model/cbam.py:66-77 and the "+ r" of modules.py:38-39); out_relu optional.
 * scratch: (33*B*C + 2*B*H*W) floats.  w1 [C/16][C], b1 [C/16], w2 [C][C/16], b2 [C], w7 [2][7][7].        */
int xm_cbam(const void* x, int32_t B, int32_t H, int32_t W, int32_t C, const float* w1, const float* b1, const float* w2,
            const float* b2, const float* w7, float b7, float* scratch, void* out, void* out_relu, void* stream);
/* out = bilinear_x2(g [B][h][w][C]) + skip [1][2h][2w][C] (modules.py:186-189); out_relu optional.       */
int xm_upsample2x_add(const void* g, const void* skip, int32_t B, int32_t h, int32_t w, int32_t C, void* out, void* out_relu, void* stream);
/* area (mean) down-sampling by f; optional extra single-channel map appended as channel C; zero padded to cpad */
int xm_area_down(const void* in, const void* extra, int32_t B, int32_t H, int32_t W, int32_t C, int32_t f, int32_t cpad, void* out, void* stream);
/* 3x3, pad 1, ONE output channel (decoder.pred, model/modules.py:227,239): weight fp16 [9][C], out fp16 [B][H][W] */
int xm_conv3x3_c1(const void* in, const void* weight_tap_c, float bias, int32_t B, int32_t H, int32_t W, int32_t C, void* out, void* stream);
/* values fp16 [npix][3*hd], h fp32 [npix][hd] -> h' = f*h*(1-u) + u*tanh(v) (modules.py:68-72) as fp32 and fp16 */
int xm_gru(const void* values, const float* h, int64_t npix, int32_t hidden_dim, float* h_out, void* h_out16, void* stream);
/* logits4 fp16 [n][h4][w4] -> bilinear x4, sigmoid, soft aggregation: prob/logits fp32 [n+1][4*h4][4*w4]  */
int xm_upsample4x_aggregate(const void* logits4, int32_t n, int32_t h4, int32_t w4, float* prob, float* logits, void* stream);
/* value fp16 [n_obj][hw][512] (NHWC) -> arena fp16 [n_obj][512][cap] columns [col0, col0+hw)               */
int xm_value_append(const void* value_hwc, int32_t n_obj, int32_t hw, void* arena, int64_t cap, int32_t col0, void* stream);

/* ---------------------------------------------------------------- long-term memory maintenance */
/* torch.topk(use/life, k, sorted=True) of inference/memory_manager.py:355: indices by descending usage, ties by ascending index. */
int xm_usage_topk(const float* use, const float* life, int32_t n, int32_t k, int32_t* out_idx, void* stream);
/* least-used eviction of inference/kv_memory_store.py:160-181: thr = n_remove-th smallest use/life; keep_idx <- ascending columns
 * with use/life > thr, *count <- their number (device int32; the host reads it to update its bookkeeping). */
int xm_usage_evict_list(const float* use, const float* life, int32_t n, int32_t n_remove, int32_t* keep_idx, int32_t* count, void* stream);
/* consolidation (inference/memory_manager.py:349-390): for every prototype q with proto_idx[q] >= col_begin,
 *   aff[q][n] = softmax over n in [col_begin, n) of similarity(candidate n, prototype q)   (model/memory_util.py:7-39,55-60)
 *   shr_out[q] = sum_n s[n] * aff[q][n]                                                    (optional)
 * kp fp16 [n][128] packed candidate keys (xm_key_pack), s fp32 [n], e fp16 [n][64] selection rows or NULL. */
int xm_consolidate_affinity(const void* kp, const float* s, const void* e, int32_t n, const int32_t* proto_idx, int32_t n_proto,
                            int32_t col_begin, float* aff, int64_t aff_stride, float* shr_out, void* stream);
int64_t xm_consolidate_scratch_bytes(int32_t n_obj, int32_t n_valid);
/* out fp16 [n_obj][512][n_valid] = v[o][c][col_begin..n) @ aff[valid_q[j]][col_begin..n)^T; v fp16 planes with column pitch `cap`;
 * valid_q NULL = prototypes 0..n_valid-1. */
int xm_consolidate_values(const void* v, int64_t cap, int32_t n_obj, int32_t col_begin, int32_t n, const float* aff, int64_t aff_stride,
                          const int32_t* valid_q, int32_t n_valid, float* scratch, int64_t scratch_bytes, void* out, void* stream);
/* in-arena compaction of a bank (inference/kv_memory_store.py:125-158,178-181): column i in [first, m) <- column keep_idx[i]
 * (or i + shift when keep_idx is NULL), source >= destination, through the scratch `tmp`. */
int64_t xm_bank_compact_tmp_bytes(int32_t n_obj_cap, int32_t moved);
int xm_bank_compact(void* kp, float* s, void* e, float* use, float* life, void* v, int64_t cap, int32_t n_obj_cap, const int32_t* keep_idx,
                    int32_t shift, int32_t first, int32_t m, void* tmp, int64_t tmp_bytes, void* stream);

/* ---------------------------------------------------------------- annotation-candidate selector */
/* Cycle dissimilarity of ordered frame pairs (inference/frame_selection/frame_selection.py:213-221):
 *   out[p] = mean over [hw, hw] of relu(S_ab - S_ba),  A = pair_a[p], B = pair_b[p], S as in model/memory_util.py:7-39.
 * kp_all/qp_all fp16 [n_frames][hw_pad][128] (xm_key_pack / xm_query_pack rows), bsq_all/ms_all fp32 [n_frames][hw_pad];
 * partial: scratch fp32 [n_pairs][(hw_pad/128)^2].  Deterministic (fixed summation order). */
int xm_pair_dissimilarity(const void* kp_all, const void* qp_all, const float* bsq_all, const float* ms_all, int32_t n_frames,
                          int32_t hw, int32_t hw_pad, const int32_t* pair_a, const int32_t* pair_b, int32_t n_pairs, float* partial,
                          float* out, void* stream);

/* ---------------------------------------------------------------- driver-side post-processing */
/* prob fp32 [channels][in_h][in_w] with element strides (stride_c, stride_h, 1) -> out uint8 [out_h][out_w]:
 * bilinear resize (align_corners = False) + argmax over channels (first maximum) + optional 256-entry label table.
 * Replaces `_post_process` (inference/run_on_video.py:165-173) and MaskMapper.remap_index_mask
 * (inference/data/mask_mapper.py:56-64) in one pass. */
int xm_resize_argmax(const float* prob, int32_t channels, int32_t in_h, int32_t in_w, int64_t stride_c, int64_t stride_h,
                     int32_t out_h, int32_t out_w, const uint8_t* lut, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
