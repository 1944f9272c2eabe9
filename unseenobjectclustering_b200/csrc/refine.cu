// Two-stage plumbing of lib/fcn/test_dataset.py on the device:
//   uoc_filter_labels_depth   :183-198   drop labels whose pixels mostly have no depth
//   uoc_crop_boxes            :68-94     ids present in the label map, tight boxes (utils/mask.py:180-195), 25 % padding
//   uoc_crop_resize           :96-110    crop + resize to SxS: bilinear align_corners (rgb, depth), legacy nearest (mask)
//   uoc_match_label_crop      :116-179   overlap test, far-to-near ordering, renumbering, nearest paste-back
// The PyTorch formulation of these steps costs ~5 ms of launches and host synchronisations per frame (one small kernel and
// often one .item() per object); here every step is one or two launches for all objects and the only host round trip is
// the number of boxes (the output shapes depend on it).
// Integer work (histograms, boxes, labels) is exact; the resampling arithmetic follows ATen's upsample kernels:
//   bilinear align_corners: scale = (in - 1) / (out - 1), src = scale * dst, i0 = (int)src, l1 = src - i0, l0 = 1 - l1,
//                           v = l0y * (l0x * v00 + l1x * v01) + l1y * (l0x * v10 + l1x * v11)
//   legacy nearest:         src = min((int)floorf(dst * ((float)in / out)), in - 1)
#include <cmath>
#include <cstring>

#include "uoc_common.cuh"

namespace uoc {

namespace {

constexpr int kMaxLabel = 256;      // label ids are seed-label ids (< UOC_MAX_SEEDS) or refined ids (< 256)

// ---------------------------------------------------------------------------------------------- depth filter
__global__ void __launch_bounds__(256) depth_hist_kernel(const int* __restrict__ labels, const float* __restrict__ z,
                                                         long long z_stride_b, long long n, int* __restrict__ hist) {
  __shared__ int s_tot[kMaxLabel], s_good[kMaxLabel];
  const int b = blockIdx.y, tid = threadIdx.x;
  s_tot[tid] = 0; s_good[tid] = 0;
  __syncthreads();
  for (long long p = (long long)blockIdx.x * blockDim.x + tid; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int l = labels[size_t(b) * n + p];
    if (l > 0 && l < kMaxLabel) {
      atomicAdd(&s_tot[l], 1);
      if (z[b * z_stride_b + p] > 0.f) atomicAdd(&s_good[l], 1);
    }
  }
  __syncthreads();
  if (s_tot[tid]) {
    atomicAdd(hist + (size_t(b) * kMaxLabel + tid) * 2, s_tot[tid]);
    atomicAdd(hist + (size_t(b) * kMaxLabel + tid) * 2 + 1, s_good[tid]);
  }
}

__global__ void __launch_bounds__(256) depth_filter_kernel(const int* __restrict__ labels, const int* __restrict__ hist,
                                                           long long n, float threshold, int* __restrict__ out) {
  __shared__ unsigned char s_drop[kMaxLabel];
  const int b = blockIdx.y, tid = threadIdx.x;
  {
    const int tot = hist[(size_t(b) * kMaxLabel + tid) * 2], good = hist[(size_t(b) * kMaxLabel + tid) * 2 + 1];
    // torch.sum(roi_depth > 0).float() / torch.sum(mask) < threshold   (test_dataset.py:194-195), fp32 division
    s_drop[tid] = (tid > 0 && tot > 0 && (float(good) / float(tot)) < threshold) ? 1 : 0;
  }
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + tid;
  if (p < n) {
    const int l = labels[size_t(b) * n + p];
    out[size_t(b) * n + p] = (l > 0 && l < kMaxLabel && s_drop[l]) ? 0 : l;
  }
}

// ---------------------------------------------------------------------------------------------- boxes
__global__ void __launch_bounds__(256) box_kernel(const int* __restrict__ labels, int H, int W, int* __restrict__ box) {
  __shared__ int s_box[kMaxLabel][4];
  const int tid = threadIdx.x;
  s_box[tid][0] = 1 << 30; s_box[tid][1] = 1 << 30; s_box[tid][2] = -1; s_box[tid][3] = -1;
  __syncthreads();
  const long long n = (long long)H * W;
  for (long long p = (long long)blockIdx.x * blockDim.x + tid; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int l = labels[p];
    if (l > 0 && l < kMaxLabel) {
      const int y = int(p / W), x = int(p - (long long)y * W);
      atomicMin(&s_box[l][0], x); atomicMin(&s_box[l][1], y);
      atomicMax(&s_box[l][2], x); atomicMax(&s_box[l][3], y);
    }
  }
  __syncthreads();
  if (s_box[tid][2] >= 0) {
    atomicMin(box + tid * 4 + 0, s_box[tid][0]); atomicMin(box + tid * 4 + 1, s_box[tid][1]);
    atomicMax(box + tid * 4 + 2, s_box[tid][2]); atomicMax(box + tid * 4 + 3, s_box[tid][3]);
  }
}

// one block: ids in ascending order (torch.unique, 0 skipped), padded + clamped boxes (test_dataset.py:82-94)
__global__ void __launch_bounds__(kMaxLabel) roi_kernel(const int* __restrict__ box, int H, int W, float padding,
                                                        int* __restrict__ count_ids /* [1 + kMaxLabel] */,
                                                        float* __restrict__ rois /* [kMaxLabel][4] */) {
  __shared__ int s_pos[kMaxLabel];
  const int tid = threadIdx.x;
  const bool present = tid > 0 && box[tid * 4 + 2] >= 0;
  s_pos[tid] = present ? 1 : 0;
  __syncthreads();
  if (tid == 0) {               // 256-entry exclusive scan: negligible
    int run = 0;
    for (int i = 0; i < kMaxLabel; ++i) { const int v = s_pos[i]; s_pos[i] = run; run += v; }
    count_ids[0] = run;
  }
  __syncthreads();
  if (present) {
    const int k = s_pos[tid];
    int x0 = box[tid * 4 + 0], y0 = box[tid * 4 + 1], x1 = box[tid * 4 + 2], y1 = box[tid * 4 + 3];
    // int(torch.round((x_max - x_min).float() * padding_percentage)): fp32 product, round half to even
    const int xp = int(rintf(float(x1 - x0) * padding)), yp = int(rintf(float(y1 - y0) * padding));
    x0 = max(x0 - xp, 0); x1 = min(x1 + xp, W - 1);
    y0 = max(y0 - yp, 0); y1 = min(y1 + yp, H - 1);
    count_ids[1 + k] = tid;
    rois[k * 4 + 0] = float(x0); rois[k * 4 + 1] = float(y0); rois[k * 4 + 2] = float(x1); rois[k * 4 + 3] = float(y1);
  }
}

// ---------------------------------------------------------------------------------------------- crop + resize
// grid (S*S / 256, K, 1): one thread per output pixel of one crop, all planes
__global__ void __launch_bounds__(256) crop_resize_kernel(const float* __restrict__ rgb, const float* __restrict__ depth,
                                                          const int* __restrict__ labels, int H, int W,
                                                          const int* __restrict__ ids, const float* __restrict__ rois, int S,
                                                          float* __restrict__ rgb_out, float* __restrict__ mask_out,
                                                          float* __restrict__ depth_out) {
  const int k = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= S * S) return;
  const int oy = o / S, ox = o - oy * S;
  const int x0 = int(rois[k * 4 + 0]), y0 = int(rois[k * 4 + 1]), x1 = int(rois[k * 4 + 2]), y1 = int(rois[k * 4 + 3]);
  const int ih = y1 - y0 + 1, iw = x1 - x0 + 1;
  // bilinear, align_corners = True (F.upsample_bilinear, test_dataset.py:104,:109)
  const float sy = (S > 1) ? float(ih - 1) / float(S - 1) : 0.f;
  const float sx = (S > 1) ? float(iw - 1) / float(S - 1) : 0.f;
  const float fy = sy * float(oy), fx = sx * float(ox);
  const int by0 = int(fy), bx0 = int(fx);
  const int by1 = by0 + ((by0 < ih - 1) ? 1 : 0), bx1 = bx0 + ((bx0 < iw - 1) ? 1 : 0);
  const float ly1 = fy - float(by0), lx1 = fx - float(bx0);
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const size_t HW = size_t(H) * W;
  const size_t i00 = size_t(y0 + by0) * W + x0 + bx0, i01 = size_t(y0 + by0) * W + x0 + bx1;
  const size_t i10 = size_t(y0 + by1) * W + x0 + bx0, i11 = size_t(y0 + by1) * W + x0 + bx1;
  const size_t SS = size_t(S) * S;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* src = rgb + c * HW;
    rgb_out[(size_t(k) * 3 + c) * SS + o] = ly0 * (lx0 * src[i00] + lx1 * src[i01]) + ly1 * (lx0 * src[i10] + lx1 * src[i11]);
  }
  if (depth) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* src = depth + c * HW;
      depth_out[(size_t(k) * 3 + c) * SS + o] = ly0 * (lx0 * src[i00] + lx1 * src[i01]) + ly1 * (lx0 * src[i10] + lx1 * src[i11]);
    }
  }
  // mask: legacy nearest (F.upsample_nearest, :106)
  const int ny = min(int(floorf(float(oy) * (float(ih) / float(S)))), ih - 1);
  const int nx = min(int(floorf(float(ox) * (float(iw) / float(S)))), iw - 1);
  mask_out[size_t(k) * SS + o] = (labels[size_t(y0 + ny) * W + x0 + nx] == ids[k]) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------- match_label_crop
// (i) per crop and cluster id: area and overlap with the stage-1 mask crop (test_dataset.py:118-125)
__global__ void __launch_bounds__(256) crop_hist_kernel(const int* __restrict__ lc, const float* __restrict__ mask, int SS,
                                                        int* __restrict__ hist /* [K][kMaxLabel][2] */) {
  __shared__ int s_area[kMaxLabel], s_inter[kMaxLabel];
  const int k = blockIdx.y, tid = threadIdx.x;
  s_area[tid] = 0; s_inter[tid] = 0;
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + tid; p < SS; p += gridDim.x * blockDim.x) {
    const int l = lc[size_t(k) * SS + p];
    if (l >= 0 && l < kMaxLabel) {
      atomicAdd(&s_area[l], 1);
      if (mask[size_t(k) * SS + p] != 0.f) atomicAdd(&s_inter[l], 1);     // the mask crop holds exact 0 / 1
    }
  }
  __syncthreads();
  if (s_area[tid]) {
    atomicAdd(hist + (size_t(k) * kMaxLabel + tid) * 2, s_area[tid]);
    atomicAdd(hist + (size_t(k) * kMaxLabel + tid) * 2 + 1, s_inter[tid]);
  }
}

// (i b) apply the drop (-1) and (ii) the ordering key: mean Z of the kept pixels with Z > 0 (all pixels if nothing is kept)
// one block per crop; fixed-order reduction -> deterministic
__global__ void __launch_bounds__(256) crop_drop_key_kernel(const int* __restrict__ lc, const int* __restrict__ hist,
                                                            const float* __restrict__ depth_crops /* [K][3][SS] or null */,
                                                            const float* __restrict__ rois, int SS, int* __restrict__ lc_out,
                                                            float* __restrict__ keys, int* __restrict__ present /* [K][kMaxLabel] */) {
  __shared__ unsigned char s_drop[kMaxLabel];
  __shared__ float s_sum[2][256];
  __shared__ int s_cnt[2][256];
  const int k = blockIdx.x, tid = threadIdx.x;
  {
    const int area = hist[(size_t(k) * kMaxLabel + tid) * 2], inter = hist[(size_t(k) * kMaxLabel + tid) * 2 + 1];
    // percentage = torch.sum(overlap) / torch.sum(mask) < 0.5
    const bool drop = area > 0 && (float(inter) / float(area)) < 0.5f;
    s_drop[tid] = drop ? 1 : 0;
    present[size_t(k) * kMaxLabel + tid] = (area > 0 && !drop) ? 1 : 0;
  }
  __syncthreads();
  float sum_kept = 0.f, sum_all = 0.f;
  int cnt_kept = 0, cnt_all = 0, any_kept = 0;
  const float* z = depth_crops ? depth_crops + (size_t(k) * 3 + 2) * SS : nullptr;
  for (int p = tid; p < SS; p += 256) {
    int l = lc[size_t(k) * SS + p];
    if (l >= 0 && l < kMaxLabel && s_drop[l]) l = -1;
    lc_out[size_t(k) * SS + p] = l;
    const bool kept = l > -1;
    any_kept |= kept ? 1 : 0;
    if (z) {
      const float zv = z[p];
      if (zv > 0.f) {
        sum_all += zv; cnt_all += 1;
        if (kept) { sum_kept += zv; cnt_kept += 1; }
      }
    }
  }
  s_sum[0][tid] = sum_kept; s_sum[1][tid] = sum_all; s_cnt[0][tid] = cnt_kept; s_cnt[1][tid] = cnt_all;
  __shared__ int s_any;
  if (tid == 0) s_any = 0;
  __syncthreads();
  if (any_kept) atomicOr(&s_any, 1);
  for (int o = 128; o > 0; o >>= 1) {
    __syncthreads();
    if (tid < o) {
      s_sum[0][tid] += s_sum[0][tid + o]; s_sum[1][tid] += s_sum[1][tid + o];
      s_cnt[0][tid] += s_cnt[0][tid + o]; s_cnt[1][tid] += s_cnt[1][tid + o];
    }
  }
  __syncthreads();
  if (tid == 0) {
    float key;
    if (z) {
      const int w = s_any ? 0 : 1;                    // roi_depth = depth[labels > -1] if any, else the whole crop (:131-134)
      key = s_sum[w][0] / float(s_cnt[w][0]);         // mean of the values > 0: NaN when there is none (torch.mean of empty)
    } else {
      key = (rois[k * 4 + 3] - rois[k * 4 + 1] + 1.f) * (rois[k * 4 + 2] - rois[k * 4 + 0] + 1.f);   // roi_size (:137-143)
    }
    keys[k] = key;
  }
}

// (ii b) order = stable descending sort of the keys (NaN first, like torch.argsort(descending=True, stable=True));
// (iii) consecutive new ids for the kept clusters in that order (:150-163).  One block.
__global__ void __launch_bounds__(256) crop_order_kernel(const float* __restrict__ keys, const int* __restrict__ present, int K,
                                                         int* __restrict__ order /* [K] */,
                                                         float* __restrict__ lut /* [K][kMaxLabel + 1], index = label + 1 */) {
  __shared__ int s_order[kMaxLabel];
  __shared__ int s_base[kMaxLabel];
  __shared__ int s_num[kMaxLabel];
  const int tid = threadIdx.x;
  if (tid < K) {
    int c = 0;
    for (int l = 0; l < kMaxLabel; ++l) c += present[size_t(tid) * kMaxLabel + l];
    s_num[tid] = c;
    const float ki = keys[tid];
    const bool ni = ki != ki;
    int rank = 0;
    for (int j = 0; j < K; ++j) {
      const float kj = keys[j];
      const bool nj = kj != kj;
      bool before;                         // does j come before tid?
      if (nj && ni) before = j < tid;
      else if (nj) before = true;
      else if (ni) before = false;
      else before = (kj > ki) || (kj == ki && j < tid);
      rank += before ? 1 : 0;
    }
    s_order[rank] = tid;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int r = 0; r < K; ++r) {
      const int i = s_order[r];
      s_base[i] = run;
      run += s_num[i];
      order[r] = i;
    }
  }
  __syncthreads();
  if (tid < K) {
    // lut[i][l + 1] = base_i + (number of present ids <= l) if present else 0; ids ascending like torch.unique
    const int i = tid;
    int c = s_base[i];
    lut[size_t(i) * (kMaxLabel + 1)] = 0.f;            // label -1
    for (int l = 0; l < kMaxLabel; ++l) {
      if (present[size_t(i) * kMaxLabel + l]) { c += 1; lut[size_t(i) * (kMaxLabel + 1) + l + 1] = float(c); }
      else lut[size_t(i) * (kMaxLabel + 1) + l + 1] = 0.f;
    }
  }
}

// (iii b) paste: crops in far-to-near order, legacy-nearest resize back to the roi, non-zero labels overwrite (:165-177)
__global__ void __launch_bounds__(256) paste_kernel(const int* __restrict__ lc, const float* __restrict__ rois,
                                                    const int* __restrict__ order, const float* __restrict__ lut, int K, int S,
                                                    int H, int W, float* __restrict__ refined) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)H * W) return;
  const int y = int(p / W), x = int(p - (long long)y * W);
  float v = 0.f;
  for (int r = 0; r < K; ++r) {
    const int i = order[r];
    const int x0 = int(rois[i * 4 + 0]), y0 = int(rois[i * 4 + 1]), x1 = int(rois[i * 4 + 2]), y1 = int(rois[i * 4 + 3]);
    if (x < x0 || x > x1 || y < y0 || y > y1) continue;
    const int oh = y1 - y0 + 1, ow = x1 - x0 + 1;
    const int sy = min(int(floorf(float(y - y0) * (float(S) / float(oh)))), S - 1);
    const int sx = min(int(floorf(float(x - x0) * (float(S) / float(ow)))), S - 1);
    const int l = lc[(size_t(i) * S + sy) * S + sx];
    const float nv = lut[size_t(i) * (kMaxLabel + 1) + (l + 1)];
    if (nv != 0.f) v = nv;
  }
  refined[p] = v;
}

}  // namespace

}  // namespace uoc

using namespace uoc;

extern "C" {

size_t uoc_refine_workspace_bytes(int N, int K) {
  const size_t a = size_t(N > 0 ? N : 1) * kMaxLabel * 2 * sizeof(int);                    // depth filter histograms
  const size_t b = size_t(kMaxLabel) * 4 * sizeof(int);                                  // boxes
  const size_t k = size_t(K > 0 ? K : 1);
  const size_t c = k * kMaxLabel * 2 * sizeof(int) + k * sizeof(float) + k * kMaxLabel * sizeof(int) + k * sizeof(int) +
                   k * (kMaxLabel + 1) * sizeof(float) + 1024;                           // match: hist, keys, present, order, lut
  size_t m = a > b ? a : b;
  return (m > c ? m : c) + 256;
}

int uoc_filter_labels_depth(const int32_t* labels, const float* depth_z, int64_t z_stride_b, int N, int64_t n, float threshold,
                            int32_t* labels_out, void* workspace, size_t workspace_bytes, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (!labels || !depth_z || !labels_out || !workspace || N < 1 || n < 1) return fail(UOC_ERR_INVALID, "bad argument");
  if (workspace_bytes < size_t(N) * kMaxLabel * 2 * sizeof(int)) return fail(UOC_ERR_WORKSPACE, "refine workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* hist = static_cast<int*>(workspace);
  UOC_CUDA(cudaMemsetAsync(hist, 0, size_t(N) * kMaxLabel * 2 * sizeof(int), st));
  const int blocks = int((n + 256 * 8 - 1) / (256 * 8));
  depth_hist_kernel<<<dim3(blocks < 1 ? 1 : blocks, N), 256, 0, st>>>(labels, depth_z, z_stride_b, n, hist);
  UOC_CHECK_LAUNCH();
  depth_filter_kernel<<<dim3((unsigned int)((n + 255) / 256), N), 256, 0, st>>>(labels, hist, n, threshold, labels_out);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

int uoc_crop_boxes(const int32_t* labels, int H, int W, float padding_percentage, int32_t* count_ids_out, float* rois_out,
                   void* workspace, size_t workspace_bytes, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (!labels || !count_ids_out || !rois_out || !workspace || H < 1 || W < 1) return fail(UOC_ERR_INVALID, "bad argument");
  if (workspace_bytes < size_t(kMaxLabel) * 4 * sizeof(int)) return fail(UOC_ERR_WORKSPACE, "refine workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* box = static_cast<int*>(workspace);
  // x_min / y_min start at +big (0x3f3f3f3f), x_max / y_max at -1: two byte-pattern memsets of 2-int columns do not fit the
  // interleaved layout, so a tiny init kernel-free trick: memset everything to 0x3f then set the max columns with a 2D memset
  UOC_CUDA(cudaMemsetAsync(box, 0x3f, size_t(kMaxLabel) * 4 * sizeof(int), st));
  UOC_CUDA(cudaMemset2DAsync(box + 2, 4 * sizeof(int), 0xff, 2 * sizeof(int), kMaxLabel, st));
  const long long n = (long long)H * W;
  const int blocks = int((n + 256 * 8 - 1) / (256 * 8));
  box_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(labels, H, W, box);
  UOC_CHECK_LAUNCH();
  roi_kernel<<<1, kMaxLabel, 0, st>>>(box, H, W, padding_percentage, count_ids_out, rois_out);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

int uoc_crop_resize(const float* rgb, const float* depth, const int32_t* labels, int H, int W, const int32_t* ids,
                    const float* rois, int K, int S, float* rgb_crops, float* mask_crops, float* depth_crops,
                    uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (K == 0) return UOC_OK;
  if (!rgb || !labels || !ids || !rois || !rgb_crops || !mask_crops || K < 0 || S < 1) return fail(UOC_ERR_INVALID, "bad argument");
  if (depth && !depth_crops) return fail(UOC_ERR_INVALID, "depth_crops is null");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  crop_resize_kernel<<<dim3((S * S + 255) / 256, K), 256, 0, st>>>(rgb, depth, labels, H, W, ids, rois, S, rgb_crops,
                                                                    mask_crops, depth_crops);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

int uoc_match_label_crop(const int32_t* labels_crop, const float* mask_crops, const float* rois, const float* depth_crops,
                         int K, int S, int H, int W, float* refined_out, int32_t* labels_crop_out, void* workspace,
                         size_t workspace_bytes, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (!refined_out || H < 1 || W < 1) return fail(UOC_ERR_INVALID, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (K == 0) {
    UOC_CUDA(cudaMemsetAsync(refined_out, 0, sizeof(float) * size_t(H) * W, st));
    return UOC_OK;
  }
  if (!labels_crop || !mask_crops || !rois || !labels_crop_out || !workspace || K < 0 || K > kMaxLabel || S < 1)
    return fail(UOC_ERR_INVALID, "bad argument");
  if (workspace_bytes < uoc_refine_workspace_bytes(1, K)) return fail(UOC_ERR_WORKSPACE, "refine workspace too small");
  char* ws = static_cast<char*>(workspace);
  int* hist = reinterpret_cast<int*>(ws);                      ws += size_t(K) * kMaxLabel * 2 * sizeof(int);
  int* present = reinterpret_cast<int*>(ws);                   ws += size_t(K) * kMaxLabel * sizeof(int);
  float* keys = reinterpret_cast<float*>(ws);                  ws += align_up(size_t(K) * sizeof(float), 256);
  int* order = reinterpret_cast<int*>(ws);                     ws += align_up(size_t(K) * sizeof(int), 256);
  float* lut = reinterpret_cast<float*>(ws);
  const int SS = S * S;
  UOC_CUDA(cudaMemsetAsync(hist, 0, size_t(K) * kMaxLabel * 2 * sizeof(int), st));
  crop_hist_kernel<<<dim3((SS + 256 * 8 - 1) / (256 * 8), K), 256, 0, st>>>(labels_crop, mask_crops, SS, hist);
  UOC_CHECK_LAUNCH();
  crop_drop_key_kernel<<<K, 256, 0, st>>>(labels_crop, hist, depth_crops, rois, SS, labels_crop_out, keys, present);
  UOC_CHECK_LAUNCH();
  crop_order_kernel<<<1, 256, 0, st>>>(keys, present, K, order, lut);
  UOC_CHECK_LAUNCH();
  const long long n = (long long)H * W;
  paste_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(labels_crop_out, rois, order, lut, K, S, H, W, refined_out);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

}  // extern "C"
