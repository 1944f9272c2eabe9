/*
 * uoc.h -- C ABI of the B200-native UnseenObjectClustering inference hot path.
 *
 * The reference (NVlabs/UnseenObjectClustering @ f5a00c7) is pure Python/PyTorch and has no FFI
 * of its own; its boundary for this path is Python-function level.  The entry
 * points below are what a ctypes binding on the reference side binds to replace those functions;
 * each one cites the reference interface it replaces (paths relative to the reference root).
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every function returns an int status (UOC_OK == 0); on failure uoc_last_error() returns a
 *    thread-local, human readable message.  No exceptions cross the ABI, nothing calls exit().
 *  - all `const float*` / `void*` data arguments are DEVICE pointers unless the name ends in
 *    `_host`.  The library never allocates device memory behind the caller's back except for the
 *    re-packed weights owned by a uoc_backbone handle; callers pass a workspace whose size the
 *    matching *_workspace_bytes() function reports.
 *  - work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed as void*);
 *    functions return after enqueueing unless stated otherwise.
 *  - there is no CPU fallback: without a Blackwell (sm_100) device every compute entry point fails
 *    with UOC_ERR_UNSUPPORTED.
 */
#ifndef UOC_H_
#define UOC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define UOC_API __declspec(dllexport)
#else
#define UOC_API __attribute__((visibility("default")))
#endif

enum {
  UOC_OK = 0,
  UOC_ERR_INVALID = 1,      /* bad argument (shape, alignment, null pointer)            */
  UOC_ERR_CUDA = 2,         /* a CUDA runtime / driver call failed                      */
  UOC_ERR_WORKSPACE = 3,    /* workspace too small                                      */
  UOC_ERR_UNSUPPORTED = 4,  /* no sm_100 device, or a size outside the supported range  */
  UOC_ERR_DEVICE = 5        /* a kernel reported an internal error (pipeline time-out)  */
};

/* flags (bit set) accepted by the compute entry points */
enum {
  UOC_FLAG_LOOP_SIMT = 1,   /* mean-shift iterations on the fp32 SIMT validation kernel instead of tcgen05 */
  UOC_FLAG_CONV_SIMT = 2,   /* backbone convolutions on the fp32 SIMT validation kernel instead of tcgen05 */
  UOC_FLAG_SYNC_CHECK = 4,  /* synchronise the stream and read back the device error word before returning */
  UOC_FLAG_FPS_FP32 = 8,    /* seed selection re-reads the fp32 field in every pass (no bf16 screening pass) */
  UOC_FLAG_EUCLIDEAN = 16   /* metric='euclidean' (cfg.TRAIN.EMBEDDING_METRIC, lib/fcn/config.py:261; the euclidean branches
                               of lib/utils/mean_shift.py:21-24,58-60,101-105,159-160,207-209): distances ||x - z||, weights
                               exp(-kappa ||x - z||^2), update divided by max(sum of weights, 1).  X need not be unit norm.
                               fp32 SIMT kernels (no tensor-core path); SURVEY section 8(f) rank 3. */
};

#define UOC_MAX_SEEDS 128   /* num_seeds upper bound: one tcgen05 M=128 accumulator tile */

typedef void* uoc_stream_t;  /* cudaStream_t */

UOC_API const char* uoc_last_error(void);
UOC_API int uoc_version(void);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches) */
UOC_API unsigned long long uoc_launch_count(void);
/* sm count / compute capability of the current device; UOC_ERR_UNSUPPORTED if it is not sm_100. */
UOC_API int uoc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Parity-test / measurement knobs (not part of the product contract): the same switches the UOC_* environment variables
 * set when the library is loaded -- conv_pair, conv_wres, fps_tc, fps_stream, fps_tmem_tiles, fps_batch_stream, fps_rn_margin,
 * fps_stats, conv_debug, conv_trace, loop_trace, assign_simt (csrc/uoc_common.cuh).  Used by tests/ to run both
 * convolution kernels and every seed-selection variant on the same inputs. */
UOC_API int uoc_set_knob(const char* name, int value);

/* Kernels report pipeline time-outs / bad configurations in a per-device error word instead of hanging; the compute
 * entry points read it only under UOC_FLAG_SYNC_CHECK (that needs a stream synchronisation).  Callers that synchronise
 * anyway (the .cpu() of lib/fcn/test_dataset.py:57, :255) check it there:
 *   uoc_check_device_error        synchronises `stream`, returns UOC_ERR_DEVICE (and clears the word) if it is non-zero
 *   uoc_peek_device_error_async   enqueues a copy of the word into *word_host (pinned host memory; stream-ordered,
 *                                 capturable in a CUDA graph); the caller inspects it after its own synchronisation */
UOC_API int uoc_check_device_error(uoc_stream_t stream);
UOC_API int uoc_peek_device_error_async(uint32_t* word_host, uoc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Clustering: replaces utils.mean_shift.mean_shift_smart_init (lib/utils/mean_shift.py:192-229)
 * as called per batch item by fcn.test_dataset.clustering_features (lib/fcn/test_dataset.py:44-59).
 *
 * X is the embedding field in the reference's own memory layout: for batch item b, point p,
 * channel k the element is X[b*stride_b + k*stride_d + p] (planar NCHW: stride_d = H*W, points
 * contiguous; the reference's `features[j].view(C,-1).t()` view, test_dataset.py:54-55).
 * Rows must be unit-norm (the network emits F.normalize'd features, lib/networks/SEG.py:114).
 * ------------------------------------------------------------------------------------------ */

/* Bytes of device workspace uoc_meanshift_cluster / the stage functions need. */
UOC_API size_t uoc_meanshift_workspace_bytes(int batch, int64_t n, int d, int m);

/*
 * Whole clustering of `batch` independent fields: farthest-point seed selection -> `iters`
 * mean-shift updates -> greedy seed labelling (epsilon) -> nearest-seed pixel labels with the
 * "largest cluster is label 0" swap.
 *   first_seed_host [batch]   the np.random.randint(0, n) draw of mean_shift.py:155, made by the caller
 *   labels_out      [batch,n] int32 (device)    mean_shift_smart_init's cluster_labels
 *   selected_out    [batch,m] int64 (device)    mean_shift_smart_init's selected_indices
 *   seeds_out       [batch,m,d] fp32 (device) or NULL: the converged seeds Z
 *   seed_labels_out [batch,m] int32 (device) or NULL: connected_components labels of the seeds
 *   x_bf16          optional [batch,n,d] bf16 pixel-major copy of X (as written by
 *                   uoc_backbone_forward); NULL -> made internally from X.
 */
UOC_API int uoc_meanshift_cluster(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16,
                                  int batch, int64_t n, int d, int m, float kappa, int iters, float epsilon,
                                  const int64_t* first_seed_host, int32_t* labels_out, int64_t* selected_out,
                                  float* seeds_out, int32_t* seed_labels_out, void* workspace, size_t workspace_bytes,
                                  int flags, uoc_stream_t stream);

/* Same, with the label map also written as float32 (the reference's API type: out_label is a float tensor,
 * lib/fcn/test_dataset.py:48,57) and / or uint8 (the wire type of the multi-GPU label gather; ids < m <= 128) by the same
 * final pass -- either may be NULL. */
UOC_API int uoc_meanshift_cluster_ex(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16,
                                     int batch, int64_t n, int d, int m, float kappa, int iters, float epsilon,
                                     const int64_t* first_seed_host, int32_t* labels_out, float* labels_f32_out,
                                     uint8_t* labels_u8_out, int64_t* selected_out, float* seeds_out,
                                     int32_t* seed_labels_out, void* workspace, size_t workspace_bytes, int flags,
                                     uoc_stream_t stream);

/* Stage entry points (same semantics as the matching reference functions; used by the parity
 * tests and by callers that want one stage only). */

/* select_smart_seeds (lib/utils/mean_shift.py:128-189, cosine): selected_out [batch,m] int64,
 * seeds_out [batch,m,d] fp32.
 * x_bf16 (optional, d = 64/128): the bf16 pixel-major copy of X (round-to-nearest or truncated).  When given, every pass
 * screens the points with it and evaluates the fp32 distance only where the new seed can lower the running minimum;
 * the selected indices are bit-identical either way (fps_tc.cu).  NULL or UOC_FLAG_FPS_FP32: fp32 passes only. */
UOC_API int uoc_select_seeds(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n,
                             int d, int m, const int64_t* first_seed_host, int64_t* selected_out, float* seeds_out,
                             void* workspace, size_t workspace_bytes, int flags, uoc_stream_t stream);

/* select_smart_seeds continued from seeds chosen before (init_seeds / num_init_seeds, lib/utils/mean_shift.py:144-149,
 * :164-169): init_seeds [batch,num_init,d] fp32 (device) are seeds 0 .. num_init-1; their distances enter the running
 * minimum, the remaining m - num_init seeds are sampled as usual.  selected_out is -1 for the given seeds (:140),
 * seeds_out [batch,m,d] holds all m.  fp32 passes (the continuation API is not on the hot path). */
UOC_API int uoc_select_seeds_init(const float* X, int64_t stride_b, int64_t stride_d, int batch, int64_t n, int d, int m,
                                  const float* init_seeds, int num_init, int64_t* selected_out, float* seeds_out,
                                  void* workspace, size_t workspace_bytes, int flags, uoc_stream_t stream);

/* seed_hill_climbing_ball (lib/utils/mean_shift.py:79-109, cosine): Z [batch,m,d] fp32 updated in place. */
UOC_API int uoc_hill_climb(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch,
                           int64_t n, int d, int m, float kappa, int iters, float* Z, void* workspace,
                           size_t workspace_bytes, int flags, uoc_stream_t stream);

/* connected_components (lib/utils/mean_shift.py:41-76, cosine): seed_labels_out [batch,m] int32;
 * num_unique_out [batch] int32 = len(unique(seed labels)) (mean_shift.py:218). */
UOC_API int uoc_label_seeds(const float* Z, int batch, int m, int d, float epsilon, int32_t* seed_labels_out,
                            int32_t* num_unique_out, uoc_stream_t stream);
/* same with flags (UOC_FLAG_EUCLIDEAN: ||z_j - z_i|| <= epsilon, mean_shift.py:58-60) */
UOC_API int uoc_label_seeds_ex(const float* Z, int batch, int m, int d, float epsilon, int flags, int32_t* seed_labels_out,
                               int32_t* num_unique_out, uoc_stream_t stream);

/* nearest-seed assignment + relabel (lib/utils/mean_shift.py:206-227): labels_out [batch,n] int32.
 * x_bf16 (optional): the bf16 pixel-major copy; when given the tensor-core pass with exactness certificate is used. */
UOC_API int uoc_assign_labels(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n,
                              int d, int m, const float* Z, const int32_t* seed_labels, const int32_t* num_unique,
                              int32_t* labels_out, void* workspace, size_t workspace_bytes, uoc_stream_t stream);
/* same with flags (UOC_FLAG_EUCLIDEAN: arg-min of ||x - z_j||, mean_shift.py:207-209; x_bf16 is ignored) */
UOC_API int uoc_assign_labels_ex(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n,
                                 int d, int m, const float* Z, const int32_t* seed_labels, const int32_t* num_unique,
                                 int32_t* labels_out, void* workspace, size_t workspace_bytes, int flags, uoc_stream_t stream);

/* same, with optional float32 / uint8 copies of the label map written by the final pass (see uoc_meanshift_cluster_ex) */
UOC_API int uoc_assign_labels_typed(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch,
                                    int64_t n, int d, int m, const float* Z, const int32_t* seed_labels,
                                    const int32_t* num_unique, int32_t* labels_out, float* labels_f32_out,
                                    uint8_t* labels_u8_out, void* workspace, size_t workspace_bytes, int flags,
                                    uoc_stream_t stream);

/* fp32 planar [batch][d][n] -> bf16 pixel-major [batch][n][d] (the layout the tcgen05 loop streams). */
UOC_API int uoc_pack_bf16(const float* X, int64_t stride_b, int64_t stride_d, int batch, int64_t n, int d,
                          void* x_bf16_out, uoc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Backbone: replaces networks.seg_resnet34_8s_embedding(...).forward for INPUT='RGBD',
 * FUSION_TYPE='add' in eval mode (lib/networks/SEG.py:88-119, :173-176; trunk
 * lib/networks/resnet_dilated.py:287-327 + lib/networks/resnet.py:236-270).
 * ------------------------------------------------------------------------------------------ */

typedef struct uoc_backbone uoc_backbone;

/* One tensor of a reference-format state_dict (keys 'fcn.resnet34_8s.*', 'fcn_depth.resnet34_8s.*'
 * after the 'module.' stripping of SEG.py:145-146); data_host is contiguous fp32 on the host. */
typedef struct {
  const char* name;
  const float* data_host;
  int64_t numel;
} uoc_weight_desc;

/* Folds eval-mode BatchNorm (eps 1e-5) into the convolutions, converts to bf16, re-packs to the
 * K-major [Cout][tap][Cin] layout the implicit-GEMM kernel reads, uploads.  Blocking. */
UOC_API int uoc_backbone_create(uoc_backbone** out, const uoc_weight_desc* tensors, int n_tensors, int num_units);

/* The other input / fusion variants behind the same reference factory (lib/networks/SEG.py:36-37,:69-71,:97-114;
 * SURVEY section 8(f) rank 2), selected by what the reference reads from cfg.INPUT / cfg.TRAIN.FUSION_TYPE /
 * cfg.TRAIN.EMBEDDING_NORMALIZATION:
 *   input_type  UOC_INPUT_RGBD | UOC_INPUT_COLOR (features = fcn(img)) | UOC_INPUT_DEPTH (features = fcn(depth))
 *   fusion_type (RGBD only) UOC_FUSION_ADD (fcn(img) + fcn_depth(depth)) | UOC_FUSION_CAT (channel concatenation,
 *               2 * num_units channels, num_units = 64 only) | UOC_FUSION_EARLY (one trunk on cat(img, depth), 6-channel
 *               stem: seg_resnet34_8s_embedding_early, SEG.py:178-181)
 *   normalize   0 skips the F.normalize of SEG.py:113-114
 * State-dict keys: 'fcn.resnet34_8s.*' always; 'fcn_depth.resnet34_8s.*' for RGBD add / cat.
 * uoc_backbone_create(...) == uoc_backbone_create_ex(..., UOC_INPUT_RGBD, UOC_FUSION_ADD, 1). */
enum { UOC_INPUT_RGBD = 0, UOC_INPUT_COLOR = 1, UOC_INPUT_DEPTH = 2 };
enum { UOC_FUSION_ADD = 0, UOC_FUSION_CAT = 1, UOC_FUSION_EARLY = 2 };
UOC_API int uoc_backbone_create_ex(uoc_backbone** out, const uoc_weight_desc* tensors, int n_tensors, int num_units,
                                   int input_type, int fusion_type, int normalize);
/* channels of the field uoc_backbone_forward writes: num_units, or 2 * num_units for cat fusion */
UOC_API int uoc_backbone_feature_dim(const uoc_backbone* bb);
UOC_API void uoc_backbone_destroy(uoc_backbone* bb);
UOC_API size_t uoc_backbone_workspace_bytes(const uoc_backbone* bb, int N, int H, int W);

/*
 * rgb, xyz: [N,3,H,W] fp32 NCHW (image_color / depth of the reference sample dict,
 * lib/fcn/test_dataset.py:235-239); rgb may be NULL for UOC_INPUT_DEPTH, xyz for UOC_INPUT_COLOR.
 * features_out: [N,C,H,W] fp32 NCHW, C = uoc_backbone_feature_dim(), unit L2 norm over channels (unless
 * normalize == 0).  features_bf16_out (optional, may be NULL): [N,H*W,C] bf16 copy for
 * uoc_meanshift_cluster.  Any H, W >= 16, like the reference (the stride-2 stages round as PyTorch's convolutions do,
 * the head interpolates back to H x W); widths that are multiples of 4 take the TMA-fed stem and 16-byte planar stores.
 */
UOC_API int uoc_backbone_forward(uoc_backbone* bb, const float* rgb, const float* xyz, int N, int H, int W,
                                 float* features_out, void* features_bf16_out, void* workspace,
                                 size_t workspace_bytes, int flags, uoc_stream_t stream);

/* Debug / test hook: copy the trunk output of one branch (0 = rgb, 1 = depth), [N,num_units,H/8,W/8]
 * fp32 NCHW, from the last uoc_backbone_forward's workspace. */
UOC_API int uoc_backbone_read_trunk(uoc_backbone* bb, int branch, int N, int H, int W, const void* workspace,
                                    float* out, uoc_stream_t stream);

/* Generic single convolution on the tcgen05 implicit-GEMM kernel (test hook for the parity tests):
 * x [N,H,W,Cin] bf16 NHWC, w [Cout,KH*KW,Cin] bf16, bias [Cout] fp32, residual (optional) and y
 * [N,Ho,Wo,Cout] bf16 NHWC. Cin % 64 == 0, Cout % 64 == 0. */
UOC_API int uoc_conv2d_bf16(const void* x, const void* w, const float* bias, const void* residual, void* y, int N,
                            int H, int W, int Cin, int Cout, int ksize, int stride, int dilation, int relu, int flags,
                            uoc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Input preparation (the step before the path): replaces read_sample's tensor arithmetic
 * and compute_xyz (tools/test_images.py:96-135; ros/test_images_segmentation.py:38-44,146-159).
 * All pointers are device pointers.  Outputs are bit-identical to the reference's numpy / torch fp32 arithmetic.
 * ------------------------------------------------------------------------------------------ */

/* bgr [N,H,W,3] uint8 (cv2.imread order) -> image_out [N,3,H,W] fp32 = bgr/255 - pixel_means_over_255[c]
 *   (pixel_means_over_255 = float32(cfg.PIXEL_MEANS / 255.0), 3 HOST floats);
 * depth_raw [N,H,W] uint16 -> xyz_out [N,3,H,W] fp32: z = raw / depth_divisor (1000), x = (col - px) z / fx,
 *   y = (row - py) z / fy.   Either input may be NULL (COLOR / DEPTH only). */
UOC_API int uoc_prepare_inputs(const uint8_t* bgr, const uint16_t* depth_raw, int N, int H, int W, float fx, float fy,
                               float px, float py, const float* pixel_means_over_255, float depth_divisor,
                               float* image_out, float* xyz_out, uoc_stream_t stream);

/* compute_xyz (tools/test_images.py:96-102): metric depth [N,H,W] fp32 -> xyz_out [N,3,H,W] fp32. */
UOC_API int uoc_compute_xyz(const float* depth_m, int N, int H, int W, float fx, float fy, float px, float py,
                            float* xyz_out, uoc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Two-stage plumbing (lib/fcn/test_dataset.py:62-198) on the device.  Label ids < 256.
 * ------------------------------------------------------------------------------------------ */

/* bytes of workspace for the functions below (N label maps, K crops) */
UOC_API size_t uoc_refine_workspace_bytes(int N, int K);

/* filter_labels_depth (test_dataset.py:183-198): labels [N,n] int32; depth_z = channel 2 (Z) of the XYZ tensor, item b at
 * depth_z + b * z_stride_b; a non-zero id whose share of pixels with Z > 0 is < threshold becomes 0.  labels_out may alias. */
UOC_API int uoc_filter_labels_depth(const int32_t* labels, const float* depth_z, int64_t z_stride_b, int N, int64_t n,
                                    float threshold, int32_t* labels_out, void* workspace, size_t workspace_bytes,
                                    uoc_stream_t stream);

/* first half of crop_rois (test_dataset.py:68-94 + utils/mask.py:180-195): the non-zero ids of labels [H,W] in ascending
 * order, their tight boxes padded by round(padding_percentage * extent) (half to even) and clamped.
 * count_ids_out [1 + 256] int32: K, then the K ids;  rois_out [256,4] float: x_min, y_min, x_max, y_max (inclusive). */
UOC_API int uoc_crop_boxes(const int32_t* labels, int H, int W, float padding_percentage, int32_t* count_ids_out,
                           float* rois_out, void* workspace, size_t workspace_bytes, uoc_stream_t stream);

/* second half of crop_rois (:96-110): rgb / depth [3,H,W] (depth may be NULL) cropped to the K rois and resized to SxS
 * (bilinear, align_corners); mask_crops = (labels == id) resized with the legacy nearest rule. */
UOC_API int uoc_crop_resize(const float* rgb, const float* depth, const int32_t* labels, int H, int W, const int32_t* ids,
                            const float* rois, int K, int S, float* rgb_crops, float* mask_crops, float* depth_crops,
                            uoc_stream_t stream);

/* match_label_crop (:116-179): labels_crop [K,S,S] int32, mask_crops [K,S,S] (0/1), rois [K,4], depth_crops [K,3,S,S] or
 * NULL.  Clusters overlapping the stage-1 mask by < 50 % of their area become -1 (labels_crop_out); crops are ordered far
 * to near by the mean Z of their kept pixels (roi area without depth), kept clusters renumbered 1, 2, .. in that order and
 * pasted back (legacy nearest) into refined_out [H,W] float, nearer crops overwriting farther ones. */
UOC_API int uoc_match_label_crop(const int32_t* labels_crop, const float* mask_crops, const float* rois,
                                 const float* depth_crops, int K, int S, int H, int W, float* refined_out,
                                 int32_t* labels_crop_out, void* workspace, size_t workspace_bytes, uoc_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation tail: the pixel work of utils.evaluation.multilabel_metrics
 * (lib/utils/evaluation.py:109-257, called per frame by fcn.test_dataset.test_segnet :307-330).  prediction, gt: [H,W]
 * int32 label maps on the device, ids in [0, 254], 0 = background.  All outputs are [256,256] int32 indexed
 * [gt label][predicted label]:
 *   tp_out                 pixels carrying both labels (:186-188)
 *   boundary_prec_tp_out   boundary pixels of the predicted label inside the disk(bound_pix)-dilated boundary of the
 *                          ground-truth label, boundary_rec_tp_out the other way round (seg2bmap :15-70, boundary_overlap :73-106)
 *   boundary_denoms_out    [2] uint64: boundary pixels summed over the predicted / the ground-truth object labels (:206-213)
 * The Hungarian matching and the ratios are host work (evaluation.py of the Python mirror). */
UOC_API size_t uoc_metrics_workspace_bytes(int H, int W);
UOC_API int uoc_multilabel_counts(const int32_t* prediction, const int32_t* gt, int H, int W, int bound_pix, int32_t* tp_out,
                                  int32_t* boundary_prec_tp_out, int32_t* boundary_rec_tp_out,
                                  unsigned long long* boundary_denoms_out, void* workspace, size_t workspace_bytes,
                                  uoc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UOC_H_ */
