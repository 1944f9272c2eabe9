// C-ABI entry points of the clustering path (declared in include/uoc.h).
#include "cluster.cuh"

namespace uoc {

static int check_shape(const float* X, int batch, int64_t n, int d, int m, int64_t stride_b, int64_t stride_d) {
  if (!X) return fail(UOC_ERR_INVALID, "X is null");
  if (batch < 1 || n < 1 || d < 1 || m < 1) return fail(UOC_ERR_INVALID, "batch, n, d, m must be positive");
  if (m > UOC_MAX_SEEDS) return fail(UOC_ERR_UNSUPPORTED, "num_seeds > 128 is not supported");
  if (d > 256 || d % 2 != 0) return fail(UOC_ERR_UNSUPPORTED, "d must be even and <= 256");
  if (n >= (int64_t(1) << 31)) return fail(UOC_ERR_UNSUPPORTED, "n must be < 2^31");
  if (stride_d < n) return fail(UOC_ERR_INVALID, "stride_d must be >= n (points contiguous, planar layout)");
  if (batch > 1 && stride_b < int64_t(d) * stride_d && stride_b != 0)
    return fail(UOC_ERR_INVALID, "stride_b overlaps batch items");
  return UOC_OK;
}

static int upload_first(const int64_t* first_host, int batch, int64_t n, const ClusterWorkspace& w, cudaStream_t st) {
  if (!first_host) return fail(UOC_ERR_INVALID, "first_seed_host is null");
  for (int b = 0; b < batch; ++b)
    if (first_host[b] < 0 || first_host[b] >= n) return fail(UOC_ERR_INVALID, "first seed index out of range");
  UOC_CUDA(cudaMemcpyAsync(w.first, first_host, sizeof(int64_t) * batch, cudaMemcpyHostToDevice, st));
  return UOC_OK;
}

static inline int metric_of(int flags) { return (flags & UOC_FLAG_EUCLIDEAN) ? METRIC_EUCLIDEAN : METRIC_COSINE; }


static int hill_climb(const float* X, const void* x_bf16, const ClusterShape& s, const ClusterWorkspace& w, float* Z,
                      float kappa, int iters, int flags, cudaStream_t st, const __nv_bfloat16** xb_used = nullptr) {
  if (xb_used) *xb_used = static_cast<const __nv_bfloat16*>(x_bf16);
  // the euclidean branch (mean_shift.py:21-24,:101-105) runs on the fp32 SIMT kernels only
  if (flags & UOC_FLAG_EUCLIDEAN) return launch_hill_climb_simt(X, s, w, Z, kappa, iters, st, METRIC_EUCLIDEAN);
  if (flags & UOC_FLAG_LOOP_SIMT) return launch_hill_climb_simt(X, s, w, Z, kappa, iters, st);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x_bf16);
  if (!xb) {
    int rc = launch_pack_bf16(X, s, w.xb, st);
    if (rc != UOC_OK) return rc;
    xb = w.xb;
  }
  if (xb_used) *xb_used = xb;
  return launch_hill_climb_tc(xb, s, w, Z, kappa, iters, st);
}

}  // namespace uoc

using namespace uoc;

extern "C" {

size_t uoc_meanshift_workspace_bytes(int batch, int64_t n, int d, int m) {
  if (batch < 1 || n < 1 || d < 1 || m < 1) return 0;
  return cluster_workspace_bytes(batch, n, d, m);
}

int uoc_meanshift_cluster(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n,
                          int d, int m, float kappa, int iters, float epsilon, const int64_t* first_seed_host,
                          int32_t* labels_out, int64_t* selected_out, float* seeds_out, int32_t* seed_labels_out,
                          void* workspace, size_t workspace_bytes, int flags, uoc_stream_t stream) {
  return uoc_meanshift_cluster_ex(X, stride_b, stride_d, x_bf16, batch, n, d, m, kappa, iters, epsilon, first_seed_host, labels_out,
                                  nullptr, nullptr, selected_out, seeds_out, seed_labels_out, workspace, workspace_bytes, flags,
                                  stream);
}

int uoc_meanshift_cluster_ex(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n,
                             int d, int m, float kappa, int iters, float epsilon, const int64_t* first_seed_host,
                             int32_t* labels_out, float* labels_f32_out, uint8_t* labels_u8_out, int64_t* selected_out,
                             float* seeds_out, int32_t* seed_labels_out, void* workspace, size_t workspace_bytes, int flags,
                             uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  rc = check_shape(X, batch, n, d, m, stride_b, stride_d);
  if (rc != UOC_OK) return rc;
  if (!labels_out || !selected_out) return fail(UOC_ERR_INVALID, "labels_out / selected_out is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ClusterWorkspace w;
  rc = carve_cluster_workspace(workspace, workspace_bytes, batch, n, d, m, &w);
  if (rc != UOC_OK) return rc;
  ClusterShape s{batch, n, d, m, stride_b, stride_d};
  rc = upload_first(first_seed_host, batch, n, w, st);
  if (rc != UOC_OK) return rc;
  // the bf16 pixel-major copy serves the screening pass of the seed selection, the tcgen05 loop and the label pass
  const int metric = metric_of(flags);
  const __nv_bfloat16* xb = (metric == METRIC_EUCLIDEAN) ? nullptr : static_cast<const __nv_bfloat16*>(x_bf16);
  if (!xb && !(flags & (UOC_FLAG_LOOP_SIMT | UOC_FLAG_EUCLIDEAN)) && (d == 64 || d == 128)) {
    rc = launch_pack_bf16(X, s, w.xb, st);
    if (rc != UOC_OK) return rc;
    xb = w.xb;
  }
  rc = launch_select_seeds(X, (flags & (UOC_FLAG_FPS_FP32 | UOC_FLAG_LOOP_SIMT)) ? nullptr : xb, s, w, selected_out, w.Z, st,
                           metric);
  if (rc != UOC_OK) return rc;
  rc = hill_climb(X, xb, s, w, w.Z, kappa, iters, flags, st, &xb);
  if (rc != UOC_OK) return rc;
  if (metric == METRIC_EUCLIDEAN) xb = nullptr;
  rc = launch_label_seeds(w.Z, batch, m, d, epsilon, w.seed_labels, w.num_unique, st, metric);
  if (rc != UOC_OK) return rc;
  rc = launch_assign(X, xb, s, w, w.Z, w.seed_labels, w.num_unique, w.hist, w.labels_tmp, labels_out, st, metric,
                     labels_f32_out, labels_u8_out);
  if (rc != UOC_OK) return rc;
  if (seeds_out)
    UOC_CUDA(cudaMemcpyAsync(seeds_out, w.Z, sizeof(float) * size_t(batch) * m * d, cudaMemcpyDeviceToDevice, st));
  if (seed_labels_out)
    UOC_CUDA(cudaMemcpyAsync(seed_labels_out, w.seed_labels, sizeof(int) * size_t(batch) * m, cudaMemcpyDeviceToDevice, st));
  if (flags & UOC_FLAG_SYNC_CHECK) return check_device_error(st);
  return UOC_OK;
}

int uoc_select_seeds(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n, int d,
                     int m, const int64_t* first_seed_host, int64_t* selected_out, float* seeds_out, void* workspace,
                     size_t workspace_bytes, int flags, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  rc = check_shape(X, batch, n, d, m, stride_b, stride_d);
  if (rc != UOC_OK) return rc;
  if (!selected_out || !seeds_out) return fail(UOC_ERR_INVALID, "selected_out / seeds_out is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ClusterWorkspace w;
  rc = carve_cluster_workspace(workspace, workspace_bytes, batch, n, d, m, &w);
  if (rc != UOC_OK) return rc;
  ClusterShape s{batch, n, d, m, stride_b, stride_d};
  rc = upload_first(first_seed_host, batch, n, w, st);
  if (rc != UOC_OK) return rc;
  rc = launch_select_seeds(X, (flags & (UOC_FLAG_FPS_FP32 | UOC_FLAG_EUCLIDEAN)) ? nullptr : static_cast<const __nv_bfloat16*>(x_bf16),
                           s, w, selected_out, seeds_out, st, metric_of(flags));
  if (rc != UOC_OK) return rc;
  if (flags & UOC_FLAG_SYNC_CHECK) return check_device_error(st);
  return UOC_OK;
}

int uoc_select_seeds_init(const float* X, int64_t stride_b, int64_t stride_d, int batch, int64_t n, int d, int m,
                          const float* init_seeds, int num_init, int64_t* selected_out, float* seeds_out, void* workspace,
                          size_t workspace_bytes, int flags, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  rc = check_shape(X, batch, n, d, m, stride_b, stride_d);
  if (rc != UOC_OK) return rc;
  if (!selected_out || !seeds_out) return fail(UOC_ERR_INVALID, "selected_out / seeds_out is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ClusterWorkspace w;
  rc = carve_cluster_workspace(workspace, workspace_bytes, batch, n, d, m, &w);
  if (rc != UOC_OK) return rc;
  ClusterShape s{batch, n, d, m, stride_b, stride_d};
  rc = launch_select_seeds_init(X, s, w, init_seeds, num_init, selected_out, seeds_out, st, metric_of(flags));
  if (rc != UOC_OK) return rc;
  if (flags & UOC_FLAG_SYNC_CHECK) return check_device_error(st);
  return UOC_OK;
}

int uoc_hill_climb(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n, int d,
                   int m, float kappa, int iters, float* Z, void* workspace, size_t workspace_bytes, int flags,
                   uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  rc = check_shape(X, batch, n, d, m, stride_b, stride_d);
  if (rc != UOC_OK) return rc;
  if (!Z) return fail(UOC_ERR_INVALID, "Z is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ClusterWorkspace w;
  rc = carve_cluster_workspace(workspace, workspace_bytes, batch, n, d, m, &w);
  if (rc != UOC_OK) return rc;
  ClusterShape s{batch, n, d, m, stride_b, stride_d};
  rc = hill_climb(X, x_bf16, s, w, Z, kappa, iters, flags, st);
  if (rc != UOC_OK) return rc;
  if (flags & UOC_FLAG_SYNC_CHECK) return check_device_error(st);
  return UOC_OK;
}

int uoc_label_seeds(const float* Z, int batch, int m, int d, float epsilon, int32_t* seed_labels_out,
                    int32_t* num_unique_out, uoc_stream_t stream) {
  return uoc_label_seeds_ex(Z, batch, m, d, epsilon, 0, seed_labels_out, num_unique_out, stream);
}

int uoc_label_seeds_ex(const float* Z, int batch, int m, int d, float epsilon, int flags, int32_t* seed_labels_out,
                       int32_t* num_unique_out, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (!Z || !seed_labels_out || !num_unique_out) return fail(UOC_ERR_INVALID, "null pointer");
  if (batch < 1 || m < 1 || m > UOC_MAX_SEEDS || d < 1 || d > 256) return fail(UOC_ERR_INVALID, "bad batch / m / d");
  return launch_label_seeds(Z, batch, m, d, epsilon, seed_labels_out, num_unique_out, static_cast<cudaStream_t>(stream),
                            metric_of(flags));
}

int uoc_assign_labels(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n, int d,
                      int m, const float* Z, const int32_t* seed_labels, const int32_t* num_unique, int32_t* labels_out,
                      void* workspace, size_t workspace_bytes, uoc_stream_t stream) {
  return uoc_assign_labels_ex(X, stride_b, stride_d, x_bf16, batch, n, d, m, Z, seed_labels, num_unique, labels_out, workspace,
                              workspace_bytes, 0, stream);
}

int uoc_assign_labels_ex(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n, int d,
                         int m, const float* Z, const int32_t* seed_labels, const int32_t* num_unique, int32_t* labels_out,
                         void* workspace, size_t workspace_bytes, int flags, uoc_stream_t stream) {
  return uoc_assign_labels_typed(X, stride_b, stride_d, x_bf16, batch, n, d, m, Z, seed_labels, num_unique, labels_out, nullptr,
                                 nullptr, workspace, workspace_bytes, flags, stream);
}

int uoc_assign_labels_typed(const float* X, int64_t stride_b, int64_t stride_d, const void* x_bf16, int batch, int64_t n, int d,
                            int m, const float* Z, const int32_t* seed_labels, const int32_t* num_unique, int32_t* labels_out,
                            float* labels_f32_out, uint8_t* labels_u8_out, void* workspace, size_t workspace_bytes, int flags,
                            uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  rc = check_shape(X, batch, n, d, m, stride_b, stride_d);
  if (rc != UOC_OK) return rc;
  if (!Z || !seed_labels || !num_unique || !labels_out) return fail(UOC_ERR_INVALID, "null pointer");
  ClusterWorkspace w;
  rc = carve_cluster_workspace(workspace, workspace_bytes, batch, n, d, m, &w);
  if (rc != UOC_OK) return rc;
  ClusterShape s{batch, n, d, m, stride_b, stride_d};
  return launch_assign(X, static_cast<const __nv_bfloat16*>(x_bf16), s, w, Z, seed_labels, num_unique, w.hist, w.labels_tmp,
                       labels_out, static_cast<cudaStream_t>(stream), metric_of(flags), labels_f32_out, labels_u8_out);
}

int uoc_pack_bf16(const float* X, int64_t stride_b, int64_t stride_d, int batch, int64_t n, int d, void* x_bf16_out,
                  uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  rc = check_shape(X, batch, n, d, 1, stride_b, stride_d);
  if (rc != UOC_OK) return rc;
  if (!x_bf16_out) return fail(UOC_ERR_INVALID, "output is null");
  ClusterShape s{batch, n, d, 1, stride_b, stride_d};
  return launch_pack_bf16(X, s, static_cast<__nv_bfloat16*>(x_bf16_out), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
