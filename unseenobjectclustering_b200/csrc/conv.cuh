// Internal interface of the backbone kernels (conv_tc.cu, backbone_kernels.cu).
#pragma once
#include "uoc_common.cuh"

namespace uoc {

// One convolution problem on NHWC bf16 activations, evaluated as an implicit GEMM
//   M = N*Ho*Wo output pixels (tiles of 8 rows x 16 cols), N = Cout, K = taps*Cin.
// Up to two independent "groups" (the RGB and the depth branch, different weights and buffers, same
// shapes) are evaluated by one launch (gridDim.z).
struct ConvGroup {
  const void* x;          // [N][H][W][Cin] bf16
  const void* w;          // [Cout][taps][Cin] bf16 (BN folded)
  const float* bias;      // [Cout] fp32 (BN folded / fc bias)
  const void* residual;   // [N][Ho][Wo][Cout] bf16 or nullptr
  void* y;                // [N][Ho][Wo][Cout] bf16 (or fp32 when out_fp32)
};

struct ConvProblem {
  int groups;             // 1 or 2
  ConvGroup g[2];
  int N, H, W, Cin, Cout;
  int ksize;              // 1 or 3
  int stride;             // 1 or 2
  int dilation;           // 1, 2, 4
  int relu;               // apply ReLU in the epilogue
  int out_fp32;           // write fp32 instead of bf16
};

static inline int conv_out_dim(int in, int ksize, int stride, int dilation) {
  const int pad = (ksize == 3) ? dilation : 0;   // "full" padding of lib/networks/resnet.py:24-41
  return (in + 2 * pad - dilation * (ksize - 1) - 1) / stride + 1;
}

// tcgen05 implicit-GEMM convolution, first generation: one 128 x BLOCK_N tile per CTA (Cin % 64 == 0, Cout % 64 == 0)
int launch_conv_tc(const ConvProblem& p, cudaStream_t stream);
// second generation (conv_pair.cu): persistent CTA pairs (cta_group::2, M = 256), stream-K, two TMEM accumulators.
// scratch: conv_pair_scratch_bytes() of device memory private to the calling stream (stream-K counters + partials; the
// first kConvCounterBytes must be zero before the first use and are left zero by every launch).
constexpr int kConvMaxPairs = 96;
constexpr size_t kConvCounterBytes = 64 * 1024;
size_t conv_pair_scratch_bytes();
bool conv_pair_supported(const ConvProblem& p);
int launch_conv_pair(const ConvProblem& p, void* scratch, size_t scratch_bytes, cudaStream_t stream);
// third kernel (conv_wres.cu): 3x3 / stride 1 / dilation 1 with Cin = 64 or 128 (layers 1 and 2): weights resident in shared
// memory, one halo fetch per tile, persistent
bool conv_wres_supported(const ConvProblem& p);
int launch_conv_wres(const ConvProblem& p, cudaStream_t stream);
// the default path: the pair kernel (scratch given and UOC_CONV_PAIR != 0), else the first generation
int launch_conv_auto(const ConvProblem& p, void* scratch, size_t scratch_bytes, cudaStream_t stream);
// fp32-accumulate SIMT validation convolution, same interface and data types
int launch_conv_simt(const ConvProblem& p, cudaStream_t stream);

// stem: conv 7x7 s2 p3 (Cin = 3, or 6 for early fusion; fp32 NCHW input) + folded BN + ReLU -> bf16 NHWC [N][H/2][W/2][64]
struct StemGroup {
  const float* x;         // [N][3][H][W] fp32 (channels 0..2)
  const float* x2;        // [N][3][H][W] fp32 (channels 3..5 when cin == 6: torch.cat((img, depth), 1), SEG.py:101-103), else unused
  const float* w;         // [64][7*7*cin] fp32, k = (r*7 + s)*cin + c, BN folded
  const float* bias;      // [64]
  void* y;                // [N][H/2][W/2][64] bf16
};
int launch_stem(const StemGroup* g, int groups, int cin, int N, int H, int W, cudaStream_t stream);
// same on tcgen05 (im2col tile built in shared memory); w_bf16[g]: [64][192 (cin 3) | 320 (cin 6)] bf16 zero-padded, BN folded
int launch_stem_tc(const StemGroup* g, const void* const* w_bf16, int groups, int cin, int N, int H, int W, cudaStream_t stream);
// maxpool 3x3 s2 p1 on bf16 NHWC, C = 64
int launch_maxpool(const void* const* x, void* const* y, int groups, int N, int H, int W, int C, cudaStream_t stream);
// head: trunk outputs [N][h][w][du] fp32 -> bilinear x8 (align_corners) -> (optional) L2 normalise
//       -> fp32 NCHW [N][d][H][W]  and (optional) bf16 pixel-major [N][H*W][d]
//   mode HEAD_ADD: f = a + b (d = du; SEG.py:108)   HEAD_SINGLE: f = a (COLOR / DEPTH / early fusion; b unused)
//   mode HEAD_CAT: f = cat(a, b) over channels (d = 2 du; SEG.py:110)
enum { HEAD_ADD = 0, HEAD_SINGLE = 1, HEAD_CAT = 2 };
int launch_head(const float* a, const float* b, int mode, int normalize, int N, int h, int w, int d, int H, int W,
                float* out_nchw, void* out_bf16, cudaStream_t stream);
// [N][h][w][d] fp32 NHWC -> [N][d][h][w] fp32 NCHW (debug / test hook)
int launch_nhwc_to_nchw(const float* in, int N, int h, int w, int d, float* out, cudaStream_t stream);

}  // namespace uoc
