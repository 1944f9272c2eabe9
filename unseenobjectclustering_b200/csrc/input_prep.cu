// Input preparation: the step immediately before the hot path.
// tools/test_images.py:96-135 (read_sample + compute_xyz; the same arithmetic in ros/test_images_segmentation.py:38-44,146-159):
//   image_color[c][y][x] = float(bgr[y][x][c]) / 255.0f - float32(PIXEL_MEANS[c] / 255.0)          (fp32 ops, channel order kept)
//   z = float(depth_raw[y][x]) / 1000.0f
//   depth[0][y][x] = ((float(x) - px) * z) / fx ;  depth[1] = ((float(y) - py) * z) / fy ;  depth[2] = z
// Every operation is a correctly rounded fp32 op in the reference (numpy / torch CPU), so the kernel uses the _rn
// intrinsics (no FMA contraction): the outputs are BIT-IDENTICAL to the reference's.
// Raw frames are 1.5 MB (uint8 BGR + uint16 depth) instead of 7.4 MB of fp32 over PCIe.
#include "uoc_common.cuh"

namespace uoc {

namespace {

struct PrepParams {
  const uint8_t* bgr;       // [N][H][W][3] or nullptr
  const uint16_t* depth_u16;  // [N][H][W] raw depth, or nullptr
  const float* depth_f32;   // [N][H][W] metric depth (compute_xyz entry point), or nullptr
  float* image_out;         // [N][3][H][W]
  float* xyz_out;           // [N][3][H][W]
  int N, H, W;
  float fx, fy, px, py;
  float mean0, mean1, mean2;   // float32(PIXEL_MEANS / 255.0)
  float depth_divisor;         // 1000.0f
};

__global__ void __launch_bounds__(256) input_prep_kernel(PrepParams p) {
  const long long hw = (long long)p.H * p.W;
  const long long total = hw * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw, pix = i - n * hw;
    const int y = int(pix / p.W), x = int(pix - (long long)y * p.W);
    if (p.bgr) {
      const uint8_t* s = p.bgr + i * 3;
      float* o = p.image_out + n * 3 * hw + pix;
      o[0] = __fsub_rn(__fdiv_rn(float(s[0]), 255.0f), p.mean0);
      o[hw] = __fsub_rn(__fdiv_rn(float(s[1]), 255.0f), p.mean1);
      o[2 * hw] = __fsub_rn(__fdiv_rn(float(s[2]), 255.0f), p.mean2);
    }
    if (p.depth_u16 || p.depth_f32) {
      const float z = p.depth_u16 ? __fdiv_rn(float(p.depth_u16[i]), p.depth_divisor) : p.depth_f32[i];
      float* o = p.xyz_out + n * 3 * hw + pix;
      o[0] = __fdiv_rn(__fmul_rn(__fsub_rn(float(x), p.px), z), p.fx);
      o[hw] = __fdiv_rn(__fmul_rn(__fsub_rn(float(y), p.py), z), p.fy);
      o[2 * hw] = z;
    }
  }
}

int launch_prep(const PrepParams& p, cudaStream_t stream) {
  const long long total = (long long)p.N * p.H * p.W;
  long long blocks = (total + 255) / 256;
  const int cap = sm_count() > 0 ? sm_count() * 16 : 2368;
  if (blocks > cap) blocks = cap;
  input_prep_kernel<<<int(blocks), 256, 0, stream>>>(p);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

}  // namespace
}  // namespace uoc

using namespace uoc;

extern "C" {

int uoc_prepare_inputs(const uint8_t* bgr, const uint16_t* depth_raw, int N, int H, int W, float fx, float fy, float px,
                       float py, const float* pixel_means_over_255, float depth_divisor, float* image_out, float* xyz_out,
                       uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (N < 1 || H < 1 || W < 1) return fail(UOC_ERR_INVALID, "bad frame shape");
  if (!bgr && !depth_raw) return fail(UOC_ERR_INVALID, "neither colour nor depth given");
  if (bgr && (!image_out || !pixel_means_over_255)) return fail(UOC_ERR_INVALID, "image_out / pixel means missing");
  if (depth_raw && !xyz_out) return fail(UOC_ERR_INVALID, "xyz_out missing");
  if (depth_raw && !(fx != 0.f && fy != 0.f && depth_divisor != 0.f)) return fail(UOC_ERR_INVALID, "fx, fy and the depth divisor must be non-zero");
  PrepParams p{};
  p.bgr = bgr; p.depth_u16 = depth_raw; p.depth_f32 = nullptr; p.image_out = image_out; p.xyz_out = xyz_out;
  p.N = N; p.H = H; p.W = W; p.fx = fx; p.fy = fy; p.px = px; p.py = py;
  if (bgr) { p.mean0 = pixel_means_over_255[0]; p.mean1 = pixel_means_over_255[1]; p.mean2 = pixel_means_over_255[2]; }
  p.depth_divisor = depth_divisor;
  return launch_prep(p, static_cast<cudaStream_t>(stream));
}

int uoc_compute_xyz(const float* depth_m, int N, int H, int W, float fx, float fy, float px, float py, float* xyz_out,
                    uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (N < 1 || H < 1 || W < 1 || !depth_m || !xyz_out) return fail(UOC_ERR_INVALID, "bad arguments");
  if (!(fx != 0.f && fy != 0.f)) return fail(UOC_ERR_INVALID, "fx and fy must be non-zero");
  PrepParams p{};
  p.depth_f32 = depth_m; p.xyz_out = xyz_out;
  p.N = N; p.H = H; p.W = W; p.fx = fx; p.fy = fy; p.px = px; p.py = py;
  p.depth_divisor = 1.f;
  return launch_prep(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
