// Evaluation tail of lib/fcn/test_dataset.py:test_segnet: the pixel work of
// utils.evaluation.multilabel_metrics (lib/utils/evaluation.py:109-257) on the device.
//   * overlap: true-positive counts for every (ground-truth label, predicted label) pair = one joint histogram
//     (the reference builds two boolean masks per pair, :184-202);
//   * boundaries: seg2bmap (:15-70) of every label and boundary_overlap (:73-106) for every pair.  The reference dilates
//     Kp x Kg boundary maps on the CPU with a disk(bound_pix) structuring element; here a pixel carries the (at most four)
//     labels whose one-pixel boundary passes through it, and one pass per pixel looks the labels of the other map up
//     inside the disk -- the dilated maps are never materialised.
// All of it is integer work and exact; the Hungarian matching and the final ratios stay on the host (evaluation.py).
// Label ids must be in [0, 254].
#include <cmath>
#include <cstring>

#include "uoc_common.cuh"

namespace uoc {

namespace {

constexpr int kL = 256;
constexpr unsigned int kNone = 0xFFu;

// joint histogram tp[g][p] (g = ground-truth label, p = predicted label), warp-aggregated
__global__ void __launch_bounds__(256) joint_hist_kernel(const int* __restrict__ pred, const int* __restrict__ gt, long long n,
                                                         int* __restrict__ tp) {
  for (long long q0 = (long long)blockIdx.x * blockDim.x; q0 < n; q0 += (long long)gridDim.x * blockDim.x) {
    const long long q = q0 + threadIdx.x;
    const bool valid = q < n;
    const int key = valid ? ((gt[q] & 0xFF) << 8) | (pred[q] & 0xFF) : -1;
    const unsigned int active = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned int peers = __match_any_sync(active, key);
      if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(tp + key, __popc(peers));
    }
  }
}

// seg2bmap for every label at once: the labels (<= 4) whose boundary map is set at pixel (y, x), packed one per byte
// (0xFF = none).  b = (seg ^ e) | (seg ^ s) | (seg ^ se) with e / s / se the east / south / south-east neighbours (0
// outside); last row: seg ^ e; last column: seg ^ s; bottom-right corner: 0   (evaluation.py:47-57).
__device__ __forceinline__ unsigned int boundary_labels(const int* __restrict__ L, int H, int W, int y, int x) {
  if (y == H - 1 && x == W - 1) return 0xFFFFFFFFu;
  const int a = L[(long long)y * W + x] & 0xFF;
  const int b = (x + 1 < W) ? (L[(long long)y * W + x + 1] & 0xFF) : -1;
  const int c = (y + 1 < H) ? (L[(long long)(y + 1) * W + x] & 0xFF) : -1;
  const int d = (x + 1 < W && y + 1 < H) ? (L[(long long)(y + 1) * W + x + 1] & 0xFF) : -1;
  const bool last_row = (y == H - 1), last_col = (x == W - 1);
  int cand[4] = {a, b, c, d};
  unsigned int out = 0xFFFFFFFFu;
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int l = cand[k];
    if (l < 0) continue;
    bool dup = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) dup = dup || (j < k && cand[j] == l);
    if (dup) continue;
    const bool seg = (a == l), e = (b == l), s = (c == l), se = (d == l);
    bool bit;
    if (last_row) bit = seg != e;
    else if (last_col) bit = seg != s;
    else bit = (seg != e) || (seg != s) || (seg != se);
    if (bit) {
      out = (out & ~(0xFFu << (8 * cnt))) | (static_cast<unsigned int>(l) << (8 * cnt));
      ++cnt;
    }
  }
  return out;
}

__global__ void __launch_bounds__(256) boundary_kernel(const int* __restrict__ pred, const int* __restrict__ gt, int H, int W,
                                                       unsigned int* __restrict__ bpred, unsigned int* __restrict__ bgt,
                                                       unsigned long long* __restrict__ denom /* [2]: pred, gt */) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int np = 0, ng = 0;
  if (q < (long long)H * W) {
    const int y = int(q / W), x = int(q - (long long)y * W);
    const unsigned int bp = boundary_labels(pred, H, W, y, x), bg = boundary_labels(gt, H, W, y, x);
    bpred[q] = bp;
    bgt[q] = bg;
#pragma unroll
    for (int k = 0; k < 4; ++k) {                       // the background label 0 is not an object (evaluation.py:127-133)
      const unsigned int lp = (bp >> (8 * k)) & 0xFFu, lg = (bg >> (8 * k)) & 0xFFu;
      np += (lp != kNone && lp != 0u) ? 1 : 0;
      ng += (lg != kNone && lg != 0u) ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { np += __shfl_xor_sync(0xffffffffu, np, o); ng += __shfl_xor_sync(0xffffffffu, ng, o); }
  if ((threadIdx.x & 31) == 0) {
    if (np) atomicAdd(denom, (unsigned long long)np);
    if (ng) atomicAdd(denom + 1, (unsigned long long)ng);
  }
}

// boundary_overlap for all pairs: fg[g][p] += 1 for every pixel on the boundary of predicted label p that lies inside the
// disk-dilated boundary of ground-truth label g; gtm[g][p] the other way round (evaluation.py:94-106)
__global__ void __launch_bounds__(128) boundary_match_kernel(const unsigned int* __restrict__ bpred,
                                                             const unsigned int* __restrict__ bgt, int H, int W, int R,
                                                             int* __restrict__ fg, int* __restrict__ gtm) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long long)H * W) return;
  const unsigned int mine_p = bpred[q], mine_g = bgt[q];
  const bool has_p = (mine_p != 0xFFFFFFFFu), has_g = (mine_g != 0xFFFFFFFFu);
  if (!has_p && !has_g) return;
  const int y = int(q / W), x = int(q - (long long)y * W);
  unsigned int set_g[8] = {0, 0, 0, 0, 0, 0, 0, 0}, set_p[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // label sets inside the disk
  for (int dy = -R; dy <= R; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= H) continue;
    for (int dx = -R; dx <= R; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= W || dx * dx + dy * dy > R * R) continue;      // skimage.morphology.disk(R)
      const long long qq = (long long)yy * W + xx;
      if (has_p) {
        const unsigned int v = bgt[qq];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned int l = (v >> (8 * k)) & 0xFFu; if (l != kNone) set_g[l >> 5] |= 1u << (l & 31); }
      }
      if (has_g) {
        const unsigned int v = bpred[qq];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const unsigned int l = (v >> (8 * k)) & 0xFFu; if (l != kNone) set_p[l >> 5] |= 1u << (l & 31); }
      }
    }
  }
  set_g[0] &= ~1u; set_p[0] &= ~1u;                                       // background is not an object
  if (has_p) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int p = (mine_p >> (8 * k)) & 0xFFu;
      if (p == kNone || p == 0u) continue;
      for (int w = 0; w < 8; ++w) {
        unsigned int m = set_g[w];
        while (m) { const int g = (w << 5) + __ffs(m) - 1; m &= m - 1; atomicAdd(fg + g * kL + p, 1); }
      }
    }
  }
  if (has_g) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned int g = (mine_g >> (8 * k)) & 0xFFu;
      if (g == kNone || g == 0u) continue;
      for (int w = 0; w < 8; ++w) {
        unsigned int m = set_p[w];
        while (m) { const int p = (w << 5) + __ffs(m) - 1; m &= m - 1; atomicAdd(gtm + g * kL + p, 1); }
      }
    }
  }
}

}  // namespace

}  // namespace uoc

using namespace uoc;

extern "C" {

size_t uoc_metrics_workspace_bytes(int H, int W) {
  if (H < 1 || W < 1) return 0;
  return size_t(H) * W * 2 * sizeof(unsigned int) + 256;
}

int uoc_multilabel_counts(const int32_t* prediction, const int32_t* gt, int H, int W, int bound_pix, int32_t* tp_out,
                          int32_t* boundary_prec_tp_out, int32_t* boundary_rec_tp_out, unsigned long long* boundary_denoms_out,
                          void* workspace, size_t workspace_bytes, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (!prediction || !gt || !tp_out || !boundary_prec_tp_out || !boundary_rec_tp_out || !boundary_denoms_out || !workspace)
    return fail(UOC_ERR_INVALID, "null argument");
  if (H < 1 || W < 1 || bound_pix < 0 || bound_pix > 64) return fail(UOC_ERR_INVALID, "bad H / W / bound_pix");
  if (workspace_bytes < uoc_metrics_workspace_bytes(H, W)) return fail(UOC_ERR_WORKSPACE, "metrics workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)H * W;
  unsigned int* bpred = static_cast<unsigned int*>(workspace);
  unsigned int* bgt = bpred + n;
  UOC_CUDA(cudaMemsetAsync(tp_out, 0, sizeof(int) * kL * kL, st));
  UOC_CUDA(cudaMemsetAsync(boundary_prec_tp_out, 0, sizeof(int) * kL * kL, st));
  UOC_CUDA(cudaMemsetAsync(boundary_rec_tp_out, 0, sizeof(int) * kL * kL, st));
  UOC_CUDA(cudaMemsetAsync(boundary_denoms_out, 0, sizeof(unsigned long long) * 2, st));
  const int blocks = int((n + 255) / 256);
  joint_hist_kernel<<<blocks > 1184 ? 1184 : blocks, 256, 0, st>>>(prediction, gt, n, tp_out);
  UOC_CHECK_LAUNCH();
  boundary_kernel<<<blocks, 256, 0, st>>>(prediction, gt, H, W, bpred, bgt, boundary_denoms_out);
  UOC_CHECK_LAUNCH();
  boundary_match_kernel<<<(unsigned int)((n + 127) / 128), 128, 0, st>>>(bpred, bgt, H, W, bound_pix, boundary_prec_tp_out,
                                                                         boundary_rec_tp_out);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

}  // extern "C"
