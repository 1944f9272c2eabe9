// Internal launch interface of the clustering kernels (see cluster_kernels.cu / meanshift_tc.cu).
#pragma once
#include "uoc_common.cuh"

namespace uoc {

struct ClusterShape {
  int batch;
  int64_t n;         // points per item (H*W)
  int d;             // embedding channels
  int m;             // seeds (<= UOC_MAX_SEEDS)
  int64_t stride_b;  // elements between batch items of X
  int64_t stride_d;  // elements between channels of X (points are contiguous)
};

// Workspace carve-up shared by all stages (device pointers into the caller's workspace).
struct ClusterWorkspace {
  float* r;                       // [batch][n] running min distance of the farthest point sampling
  unsigned long long* keys;       // [batch][m] packed (distance, index) arg-max keys, one slot per pass
  unsigned long long* slots;      // [batch][m][CTAs per item] per-CTA arg-max keys of the second-generation sampler
  size_t slot_bytes;
  unsigned int* barrier;          // grid barrier counter
  long long* first;               // [batch] first seed index (device copy of the caller's host array)
  float* Z;                       // [batch][m][d] current seeds
  float* partials;                // [batch][P][128][d] per-CTA partial sums of the mean-shift update
  int* seed_labels;               // [batch][m]
  int* num_unique;                // [batch]
  int* hist;                      // [batch][m] pixel counts per label
  int* labels_tmp;                // [batch][n] labels before the label-0 swap
  __nv_bfloat16* xb;              // [batch][n][d] bf16 pixel-major copy of X
  float* wsum;                    // [batch][P][128] per-CTA partial sums of the weights (euclidean update only)
  int max_partials;               // P capacity
};

size_t cluster_workspace_bytes(int batch, int64_t n, int d, int m);
int carve_cluster_workspace(void* ws, size_t ws_bytes, int batch, int64_t n, int d, int m, ClusterWorkspace* out);

// K3  farthest point sampling (lib/utils/mean_shift.py:128-189)
// xb != nullptr (d = 64/128): kernel with the bf16 screening pass on tcgen05 (fps_tc.cu), same indices
// metric (all stages): METRIC_COSINE  d(x, z) = 0.5 (1 - x.z);  METRIC_EUCLIDEAN  d(x, z) = ||x - z||  (the 'euclidean'
// branches of lib/utils/mean_shift.py:21-24,58-60,101-105,159-160,207-209; fp32 SIMT kernels only)
enum { METRIC_COSINE = 0, METRIC_EUCLIDEAN = 1 };
int launch_select_seeds(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w,
                        int64_t* selected_out, float* seeds_out, cudaStream_t stream, int metric = METRIC_COSINE);
int launch_select_seeds_init(const float* X, const ClusterShape& s, const ClusterWorkspace& w, const float* init_seeds,
                             int num_init, int64_t* selected_out, float* seeds_out, cudaStream_t stream, int metric = METRIC_COSINE);
int launch_select_seeds_tc(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w,
                           int64_t* selected_out, float* seeds_out, cudaStream_t stream, bool* used);
// K4  mean-shift iterations (lib/utils/mean_shift.py:79-109): fp32 SIMT validation kernel ...
int launch_hill_climb_simt(const float* X, const ClusterShape& s, const ClusterWorkspace& w, float* Z, float kappa,
                           int iters, cudaStream_t stream, int metric = METRIC_COSINE);
// ... and the tcgen05 kernel (meanshift_tc.cu) streaming the bf16 pixel-major copy
int launch_hill_climb_tc(const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w, float* Z,
                         float kappa, int iters, cudaStream_t stream);
// sum the per-CTA partials and L2-normalise rows (F.normalize, mean_shift.py:107)
// wsum != nullptr (euclidean): rows are divided by max(sum of weights, 1) instead (mean_shift.py:101-105)
int launch_reduce_normalize(const float* partials, int batch, int P, int m, int d, int row_stride, float* Z,
                            cudaStream_t stream, const float* wsum = nullptr);
// K5a greedy seed labelling (lib/utils/mean_shift.py:41-76)
int launch_label_seeds(const float* Z, int batch, int m, int d, float epsilon, int* seed_labels, int* num_unique,
                       cudaStream_t stream, int metric = METRIC_COSINE);
// K5b nearest-seed assignment + histogram + label-0 swap (lib/utils/mean_shift.py:206-227)
// xb != nullptr (bf16 pixel-major copy available, d = 64/128): tcgen05 pass with exactness certificate + fp32 fix-up of
// the uncertified points (assign_tc.cu); otherwise the fp32 SIMT kernel.  Identical labels either way.
int launch_assign(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w, const float* Z,
                  const int* seed_labels, const int* num_unique, int* hist, int* labels_tmp, int* labels_out,
                  cudaStream_t stream, int metric = METRIC_COSINE, float* labels_f32_out = nullptr,
                  unsigned char* labels_u8_out = nullptr);
int launch_assign_tc(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w, const float* Z,
                     const int* seed_labels, int* hist, int* labels_tmp, cudaStream_t stream);
// fp32 planar -> bf16 pixel-major
int launch_pack_bf16(const float* X, const ClusterShape& s, __nv_bfloat16* xb, cudaStream_t stream);

}  // namespace uoc
