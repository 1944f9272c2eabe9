// SIMT kernels of the backbone: stem (7x7 s2 conv + BN + ReLU), max-pool, the fused head
// (branch add -> bilinear x8 -> L2 normalise -> both output layouts), a layout helper, and the
// fp32-accumulate validation convolution (UOC_FLAG_CONV_SIMT; not the product path).
#include "conv.cuh"

namespace uoc {

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ----------------------------------------------------------------------------------------------
// validation convolution: one thread per (output pixel, output channel)
// ----------------------------------------------------------------------------------------------
struct ConvSimtParams {
  const __nv_bfloat16* x[2];
  const __nv_bfloat16* w[2];
  const float* bias[2];
  const __nv_bfloat16* residual[2];
  void* y[2];
  int N, H, W, Cin, Cout, Ho, Wo, ksize, stride, dil, pad, relu, out_fp32;
};

__global__ void __launch_bounds__(256) conv_simt_kernel(ConvSimtParams p) {
  const int g = blockIdx.z;
  const long long total = (long long)p.N * p.Ho * p.Wo * p.Cout;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = int(idx % p.Cout);
  long long pix = idx / p.Cout;
  const int ox = int(pix % p.Wo);
  const int oy = int((pix / p.Wo) % p.Ho);
  const int n = int(pix / ((long long)p.Wo * p.Ho));
  float acc = 0.f;
  const int taps = p.ksize * p.ksize;
  for (int r = 0; r < p.ksize; ++r) {
    const int iy = oy * p.stride - p.pad + r * p.dil;
    if (iy < 0 || iy >= p.H) continue;
    for (int s = 0; s < p.ksize; ++s) {
      const int ix = ox * p.stride - p.pad + s * p.dil;
      if (ix < 0 || ix >= p.W) continue;
      const uint4* xp = reinterpret_cast<const uint4*>(p.x[g] + ((size_t(n) * p.H + iy) * p.W + ix) * p.Cin);
      const uint4* wp = reinterpret_cast<const uint4*>(p.w[g] + (size_t(co) * taps + (r * p.ksize + s)) * p.Cin);
      for (int c8 = 0; c8 < p.Cin / 8; ++c8) {
        const uint4 xv = __ldg(xp + c8), wv = __ldg(wp + c8);
        acc = fmaf(bf16_lo(xv.x), bf16_lo(wv.x), acc); acc = fmaf(bf16_hi(xv.x), bf16_hi(wv.x), acc);
        acc = fmaf(bf16_lo(xv.y), bf16_lo(wv.y), acc); acc = fmaf(bf16_hi(xv.y), bf16_hi(wv.y), acc);
        acc = fmaf(bf16_lo(xv.z), bf16_lo(wv.z), acc); acc = fmaf(bf16_hi(xv.z), bf16_hi(wv.z), acc);
        acc = fmaf(bf16_lo(xv.w), bf16_lo(wv.w), acc); acc = fmaf(bf16_hi(xv.w), bf16_hi(wv.w), acc);
      }
    }
  }
  acc += p.bias[g][co];
  if (p.residual[g]) acc += __bfloat162float(p.residual[g][pix * p.Cout + co]);
  if (p.relu) acc = fmaxf(acc, 0.f);
  if (p.out_fp32) static_cast<float*>(p.y[g])[pix * p.Cout + co] = acc;
  else static_cast<__nv_bfloat16*>(p.y[g])[pix * p.Cout + co] = __float2bfloat16_rn(acc);
}

int launch_conv_simt(const ConvProblem& c, cudaStream_t stream) {
  if (c.Cin % 8 != 0) return fail(UOC_ERR_UNSUPPORTED, "conv_simt needs Cin % 8 == 0");
  ConvSimtParams p;
  for (int g = 0; g < 2; ++g) {
    const int s = g < c.groups ? g : 0;
    p.x[g] = static_cast<const __nv_bfloat16*>(c.g[s].x);
    p.w[g] = static_cast<const __nv_bfloat16*>(c.g[s].w);
    p.bias[g] = c.g[s].bias;
    p.residual[g] = static_cast<const __nv_bfloat16*>(c.g[s].residual);
    p.y[g] = c.g[s].y;
  }
  p.N = c.N; p.H = c.H; p.W = c.W; p.Cin = c.Cin; p.Cout = c.Cout;
  p.ksize = c.ksize; p.stride = c.stride; p.dil = c.dilation; p.pad = (c.ksize == 3) ? c.dilation : 0;
  p.Ho = conv_out_dim(c.H, c.ksize, c.stride, c.dilation);
  p.Wo = conv_out_dim(c.W, c.ksize, c.stride, c.dilation);
  p.relu = c.relu; p.out_fp32 = c.out_fp32;
  const long long total = (long long)p.N * p.Ho * p.Wo * p.Cout;
  conv_simt_kernel<<<dim3((unsigned int)((total + 255) / 256), 1, c.groups), 256, 0, stream>>>(p);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

// ----------------------------------------------------------------------------------------------
// stem: 7x7 stride-2 pad-3 convolution of the fp32 NCHW input (3 channels), folded BN, ReLU
// (lib/networks/resnet.py:141-145, :237-239).  One block = 16x16 output pixels x 64 channels.
// ----------------------------------------------------------------------------------------------
struct StemParams {
  StemGroup g[2];
  int N, H, W, Ho, Wo;
};

constexpr int kStemPatch = 16 * 2 + 5;  // 37 input rows / cols per 16 output rows / cols

template <int CIN>
__global__ void __launch_bounds__(256) stem_kernel(StemParams p) {
  constexpr int K = 49 * CIN;
  extern __shared__ float sm[];
  float* wsm = sm;                               // [K][64]  (k-major so that 4 channels = one float4)
  float* patch = sm + K * 64;                    // [CIN][37][38]
  const int g = blockIdx.z / p.N, n = blockIdx.z % p.N;
  const int tid = threadIdx.x;
  const float* wg = p.g[g].w;
  for (int e = tid; e < K * 64; e += 256) {
    const int k = e >> 6, co = e & 63;
    wsm[e] = __ldg(wg + co * K + k);
  }
  const int oy0 = blockIdx.y * 16, ox0 = blockIdx.x * 16;
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  for (int e = tid; e < CIN * kStemPatch * kStemPatch; e += 256) {
    const int c = e / (kStemPatch * kStemPatch);
    const int rem = e - c * kStemPatch * kStemPatch;
    const int py = rem / kStemPatch, px = rem - py * kStemPatch;
    const int iy = iy0 + py, ix = ix0 + px;
    const float* xg = (c < 3 ? p.g[g].x : p.g[g].x2) + size_t(n) * 3 * p.H * p.W;
    float v = 0.f;
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = __ldg(xg + (size_t(c < 3 ? c : c - 3) * p.H + iy) * p.W + ix);
    patch[(c * kStemPatch + py) * 38 + px] = v;
  }
  __syncthreads();
  const int ly = tid >> 4, lx = tid & 15;
  float acc[64];
#pragma unroll
  for (int q = 0; q < 64; ++q) acc[q] = 0.f;
  for (int r = 0; r < 7; ++r) {
    for (int s = 0; s < 7; ++s) {
#pragma unroll
      for (int c = 0; c < CIN; ++c) {
        const float xv = patch[(c * kStemPatch + ly * 2 + r) * 38 + lx * 2 + s];
        const float4* wk = reinterpret_cast<const float4*>(wsm + ((r * 7 + s) * CIN + c) * 64);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 wv = wk[q];
          acc[4 * q + 0] = fmaf(xv, wv.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(xv, wv.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, wv.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, wv.w, acc[4 * q + 3]);
        }
      }
    }
  }
  const int oy = oy0 + ly, ox = ox0 + lx;
  if (oy < p.Ho && ox < p.Wo) {
    const float* bias = p.g[g].bias;
    uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.g[g].y) + ((size_t(n) * p.Ho + oy) * p.Wo + ox) * 64);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float f[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) f[h] = fmaxf(acc[8 * e + h] + __ldg(bias + 8 * e + h), 0.f);
      op[e] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
  }
}

template <int CIN>
static int launch_stem_t(const StemParams& p, int groups, cudaStream_t stream) {
  const size_t smem = sizeof(float) * (49 * CIN * 64 + CIN * kStemPatch * 38);
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&stem_kernel<CIN>), int(int(smem)));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  stem_kernel<CIN><<<dim3((p.Wo + 15) / 16, (p.Ho + 15) / 16, groups * p.N), 256, smem, stream>>>(p);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

int launch_stem(const StemGroup* g, int groups, int cin, int N, int H, int W, cudaStream_t stream) {
  if (cin != 3 && cin != 6) return fail(UOC_ERR_UNSUPPORTED, "stem supports 3 or 6 input channels");
  StemParams p;
  for (int i = 0; i < 2; ++i) p.g[i] = g[i < groups ? i : 0];
  p.N = N; p.H = H; p.W = W;
  p.Ho = (H + 6 - 7) / 2 + 1;
  p.Wo = (W + 6 - 7) / 2 + 1;
  return cin == 3 ? launch_stem_t<3>(p, groups, stream) : launch_stem_t<6>(p, groups, stream);
}

// ----------------------------------------------------------------------------------------------
// max-pool 3x3 stride 2 pad 1 (lib/networks/resnet.py:145), bf16 NHWC, 8 channels per thread
// ----------------------------------------------------------------------------------------------
struct PoolParams {
  const __nv_bfloat16* x[2];
  __nv_bfloat16* y[2];
  int N, H, W, C, Ho, Wo;
};

__global__ void __launch_bounds__(256) maxpool_kernel(PoolParams p) {
  const int g = blockIdx.z;
  const int c8n = p.C / 8;
  const long long total = (long long)p.N * p.Ho * p.Wo * c8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = int(idx % c8n);
  long long pix = idx / c8n;
  const int ox = int(pix % p.Wo);
  const int oy = int((pix / p.Wo) % p.Ho);
  const int n = int(pix / ((long long)p.Wo * p.Ho));
  float m[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) m[h] = -INFINITY;
  for (int r = 0; r < 3; ++r) {
    const int iy = oy * 2 - 1 + r;
    if (iy < 0 || iy >= p.H) continue;
    for (int s = 0; s < 3; ++s) {
      const int ix = ox * 2 - 1 + s;
      if (ix < 0 || ix >= p.W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.x[g] + ((size_t(n) * p.H + iy) * p.W + ix) * p.C) + c8);
      m[0] = fmaxf(m[0], bf16_lo(v.x)); m[1] = fmaxf(m[1], bf16_hi(v.x));
      m[2] = fmaxf(m[2], bf16_lo(v.y)); m[3] = fmaxf(m[3], bf16_hi(v.y));
      m[4] = fmaxf(m[4], bf16_lo(v.z)); m[5] = fmaxf(m[5], bf16_hi(v.z));
      m[6] = fmaxf(m[6], bf16_lo(v.w)); m[7] = fmaxf(m[7], bf16_hi(v.w));
    }
  }
  reinterpret_cast<uint4*>(p.y[g] + pix * p.C)[c8] =
      make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]), pack_bf16x2(m[6], m[7]));
}

int launch_maxpool(const void* const* x, void* const* y, int groups, int N, int H, int W, int C, cudaStream_t stream) {
  PoolParams p;
  for (int i = 0; i < 2; ++i) {
    p.x[i] = static_cast<const __nv_bfloat16*>(x[i < groups ? i : 0]);
    p.y[i] = static_cast<__nv_bfloat16*>(y[i < groups ? i : 0]);
  }
  p.N = N; p.H = H; p.W = W; p.C = C;
  p.Ho = (H + 2 - 3) / 2 + 1;
  p.Wo = (W + 2 - 3) / 2 + 1;
  const long long total = (long long)N * p.Ho * p.Wo * (C / 8);
  maxpool_kernel<<<dim3((unsigned int)((total + 255) / 256), 1, groups), 256, 0, stream>>>(p);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

// ----------------------------------------------------------------------------------------------
// K2 head: f = up8(a + b); f /= max(||f||, 1e-12)     (SEG.py:105-108,:114; resnet_dilated.py:325)
// Bilinear with align_corners=True, index arithmetic in fp32 like ATen's upsample_bilinear2d:
//   scale = (in - 1) / (out - 1);  src = scale * dst;  i0 = (int)src;  lambda1 = src - i0.
// Upsampling is linear, so adding the two low-resolution trunk outputs first differs from the
// reference's "upsample each, then add" by rounding only.
// One thread per output pixel, channels in registers.
// ----------------------------------------------------------------------------------------------
// D = channels of the output field; MODE: HEAD_ADD (a + b), HEAD_SINGLE (a), HEAD_CAT (a | b, D/2 channels each)
// One CTA per output ROW.  The two source rows y0, y1 of the (for add fusion already
// summed) trunk output are staged ONCE in shared memory (a thread-per-pixel first generation re-read the four source pixels of
// every output pixel from global memory: 128 LDG.128 per thread, 60 us per frame against 31 us now); every thread then owns PX consecutive output pixels x D/4
// channels: the same interpolation formula from shared memory, the L2 norm over the channels with two warp shuffles,
// and 16-byte stores on BOTH outputs (fp32 planar NCHW: PX pixels of
// one channel; bf16 pixel-major: 8 channels of one pixel).  Lane = quarter * 8 + pixel group, so that the eight lanes of
// a shared-memory phase read at most two different rows (broadcast) and a warp's planar stores are 128-byte runs.
template <int D, int MODE, int PX>
__global__ void __launch_bounds__(256, 2) head_row_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w,
                                                       int H, int W, float sy, float sx, int normalize,
                                                       float* __restrict__ out_nchw, __nv_bfloat16* __restrict__ out_bf16) {
  constexpr int DU = (MODE == HEAD_CAT) ? D / 2 : D;     // channels of one trunk output
  constexpr int CPT = D / 4;                             // channels per thread
  constexpr int PITCH = D + 4;                           // floats per source pixel in shared memory (16-byte aligned, bank shift 4)
  extern __shared__ float vrow[];                        // [2][w][PITCH]: source rows y0, y1
  const int n = blockIdx.y, oy = blockIdx.x, tid = threadIdx.x;
  const float fy = sy * float(oy);
  const int y0 = int(fy);
  const int y1 = y0 + ((y0 < h - 1) ? 1 : 0);
  const float ly1 = fy - float(y0), ly0 = 1.f - ly1;
  const size_t base = size_t(n) * h * w;
  for (int e = tid; e < w * (D / 4); e += 256) {
    const int x = e / (D / 4), k4 = e - x * (D / 4);
    const float* src = (MODE == HEAD_CAT && k4 >= DU / 4) ? b : a;
    const int kk = (MODE == HEAD_CAT && k4 >= DU / 4) ? k4 - DU / 4 : k4;
    const size_t o0 = (base + size_t(y0) * w + x) * DU, o1 = (base + size_t(y1) * w + x) * DU;
    float4 v0 = __ldg(reinterpret_cast<const float4*>(src + o0) + kk), v1 = __ldg(reinterpret_cast<const float4*>(src + o1) + kk);
    if (MODE == HEAD_ADD) {
      const float4 u0 = __ldg(reinterpret_cast<const float4*>(b + o0) + kk), u1 = __ldg(reinterpret_cast<const float4*>(b + o1) + kk);
      v0.x += u0.x; v0.y += u0.y; v0.z += u0.z; v0.w += u0.w;
      v1.x += u1.x; v1.y += u1.y; v1.z += u1.z; v1.w += u1.w;
    }
    *reinterpret_cast<float4*>(vrow + x * PITCH + 4 * k4) = v0;                  // row y0
    *reinterpret_cast<float4*>(vrow + (w + x) * PITCH + 4 * k4) = v1;            // row y1
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const int q = lane >> 3, gi = lane & 7;                // channel quarter, pixel group inside the warp
  const size_t HW = size_t(H) * W;
  const int groups = W / PX;                             // W % PX == 0 (launch_head picks PX = 1 for any other width)
  for (int g0 = warp * 8; g0 < groups; g0 += 8 * 8) {
    const int grp = g0 + gi;
    const bool live = grp < groups;
    float f[PX][CPT];
    float ss[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      const int ox = (live ? grp : 0) * PX + i;
      const float fx = sx * float(ox);
      const int x0 = int(fx);
      const int x1 = x0 + ((x0 < w - 1) ? 1 : 0);
      const float lx1 = fx - float(x0), lx0 = 1.f - lx1;
      const float* r00 = vrow + x0 * PITCH + q * CPT;
      const float* r01 = vrow + x1 * PITCH + q * CPT;
      const float* r10 = vrow + (w + x0) * PITCH + q * CPT;
      const float* r11 = vrow + (w + x1) * PITCH + q * CPT;
      ss[i] = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < CPT / 4; ++c4) {
        const float4 v00 = *reinterpret_cast<const float4*>(r00 + 4 * c4), v01 = *reinterpret_cast<const float4*>(r01 + 4 * c4);
        const float4 v10 = *reinterpret_cast<const float4*>(r10 + 4 * c4), v11 = *reinterpret_cast<const float4*>(r11 + 4 * c4);
        f[i][4 * c4 + 0] = ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x);
        f[i][4 * c4 + 1] = ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y);
        f[i][4 * c4 + 2] = ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z);
        f[i][4 * c4 + 3] = ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w);
      }
    }
    // sum of squares in channel order 0..D-1 (the order of the first generation and of the oracle's reduction is not
    // defined; the tolerance of the embeddings is 1e-3 cosine distance): per-thread partial, then the four quarters
#pragma unroll
    for (int i = 0; i < PX; ++i) {
#pragma unroll
      for (int c = 0; c < CPT; ++c) ss[i] = fmaf(f[i][c], f[i][c], ss[i]);
      ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], 8);
      ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], 16);
    }
    if (!live) continue;
    float inv[PX];
#pragma unroll
    for (int i = 0; i < PX; ++i) inv[i] = normalize ? 1.0f / fmaxf(sqrtf(ss[i]), 1e-12f) : 1.0f;
    const size_t pix0 = size_t(oy) * W + size_t(grp) * PX;
    float* o = out_nchw + size_t(n) * D * HW + size_t(q * CPT) * HW + pix0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (PX == 4) {
        __stcs(reinterpret_cast<float4*>(o + size_t(c) * HW),
               make_float4(f[0][c] * inv[0], f[1][c] * inv[1], f[2 % PX][c] * inv[2 % PX], f[3 % PX][c] * inv[3 % PX]));
      } else if (PX == 2) {
        __stcs(reinterpret_cast<float2*>(o + size_t(c) * HW), make_float2(f[0][c] * inv[0], f[1 % PX][c] * inv[1 % PX]));
      } else {
        __stcs(o + size_t(c) * HW, f[0][c] * inv[0]);       // PX == 1: frames whose width is not a multiple of 4
      }
    }
    if (out_bf16) {
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        uint4* ob = reinterpret_cast<uint4*>(out_bf16 + (size_t(n) * HW + pix0 + i) * D + q * CPT);
#pragma unroll
        for (int e = 0; e < CPT / 8; ++e)
          ob[e] = make_uint4(pack_bf16x2(f[i][8 * e + 0] * inv[i], f[i][8 * e + 1] * inv[i]),
                             pack_bf16x2(f[i][8 * e + 2] * inv[i], f[i][8 * e + 3] * inv[i]),
                             pack_bf16x2(f[i][8 * e + 4] * inv[i], f[i][8 * e + 5] * inv[i]),
                             pack_bf16x2(f[i][8 * e + 6] * inv[i], f[i][8 * e + 7] * inv[i]));
      }
    }
  }
}

template <int D, int MODE, int PX>
static int launch_head_row(const float* a, const float* b, int normalize, int N, int h, int w, int H, int W, float sy, float sx,
                           float* out_nchw, __nv_bfloat16* ob, cudaStream_t stream) {
  const size_t smem = size_t(2) * w * (D + 4) * sizeof(float);
  if (smem > 48 * 1024) {
    const int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&head_row_kernel<D, MODE, PX>), int(smem));
    if (rc != UOC_OK) return rc;
  }
  head_row_kernel<D, MODE, PX><<<dim3(H, N), 256, smem, stream>>>(a, b, h, w, H, W, sy, sx, normalize, out_nchw, ob);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

int launch_head(const float* a, const float* b, int mode, int normalize, int N, int h, int w, int d, int H, int W,
                float* out_nchw, void* out_bf16, cudaStream_t stream) {
  const float sy = (H > 1) ? float(h - 1) / float(H - 1) : 0.f;
  const float sx = (W > 1) ? float(w - 1) / float(W - 1) : 0.f;
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(out_bf16);
  const size_t smem_row = size_t(2) * w * (d + 4) * sizeof(float);
  if (W % 4 == 0 && smem_row <= 200 * 1024) {
    if (mode == HEAD_ADD && d == 64) return launch_head_row<64, HEAD_ADD, 4>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_ADD && d == 128) return launch_head_row<128, HEAD_ADD, 2>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_SINGLE && d == 64) return launch_head_row<64, HEAD_SINGLE, 4>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_SINGLE && d == 128) return launch_head_row<128, HEAD_SINGLE, 2>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_CAT && d == 128) return launch_head_row<128, HEAD_CAT, 2>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    return fail(UOC_ERR_UNSUPPORTED, "head supports 64 or 128 output channels (cat fusion: 2 x 64)");
  }
  if (smem_row <= 200 * 1024) {
    // any other width (the reference takes every frame size, resnet_dilated.py:293,325): one pixel per thread, so that
    // no planar store straddles a row end or needs more than 4-byte alignment (H * W need not be a multiple of 4)
    if (mode == HEAD_ADD && d == 64) return launch_head_row<64, HEAD_ADD, 1>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_ADD && d == 128) return launch_head_row<128, HEAD_ADD, 1>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_SINGLE && d == 64) return launch_head_row<64, HEAD_SINGLE, 1>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_SINGLE && d == 128) return launch_head_row<128, HEAD_SINGLE, 1>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    if (mode == HEAD_CAT && d == 128) return launch_head_row<128, HEAD_CAT, 1>(a, b, normalize, N, h, w, H, W, sy, sx, out_nchw, ob, stream);
    return fail(UOC_ERR_UNSUPPORTED, "head supports 64 or 128 output channels (cat fusion: 2 x 64)");
  }
  return fail(UOC_ERR_UNSUPPORTED, "head: a source row pair must fit in shared memory (W <= ~2900 at 64 channels)");
}

__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, int hw, int d, float* __restrict__ out) {
  const int n = blockIdx.y;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)hw * d) return;
  const int k = int(idx / hw), pp = int(idx % hw);
  out[size_t(n) * hw * d + idx] = in[(size_t(n) * hw + pp) * d + k];
}

int launch_nhwc_to_nchw(const float* in, int N, int h, int w, int d, float* out, cudaStream_t stream) {
  const long long total = (long long)h * w * d;
  nhwc_to_nchw_kernel<<<dim3((unsigned int)((total + 255) / 256), N), 256, 0, stream>>>(in, h * w, d, out);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

}  // namespace uoc

// ----------------------------------------------------------------------------------------------
// stem on the tensor cores: the 7x7 stride-2 convolution as an implicit GEMM with K = 7*7*3 = 147 -> 192.
// A 128-row im2col tile (8 x 16 output pixels) is BUILT in shared memory by the CTA's threads from a staged
// fp32 input patch (bf16, K-major, 128B-swizzled -- the layout TMA would have produced), the weights
// [64][192] bf16 sit in shared memory for the CTA's lifetime, 12 tcgen05.mma (M 128, N 64, K 16) per tile
// accumulate in TMEM, and the same 4 warps run the epilogue (bias, ReLU, bf16 NHWC).  Persistent CTAs,
// two per SM so that one CTA's tile construction overlaps the other's MMA / epilogue.
// ----------------------------------------------------------------------------------------------
namespace uoc {

struct StemTcParams {
  const float* x[2];
  const float* x2[2];            // channels 3..5 (early fusion), else unused
  const __nv_bfloat16* w[2];     // [64][KB*64] bf16, k = (r*7 + s)*CIN + c, zero padded
  const float* bias[2];
  __nv_bfloat16* y[2];
  int N, H, W, Ho, Wo, tiles_x, tiles_y, tiles_per_group, groups;
  unsigned int* err;
  int use_tma;                   // input patches by TMA (fp32 3-D maps over [W][H][3 N], OOB zero fill = the padding), double buffered
  CUtensorMap tmap_x[2];
  CUtensorMap tmap_x2[2];
};

constexpr int kStemPatchH = 8 * 2 + 5;     // 21
constexpr int kStemPatchW = 16 * 2 + 5;    // 37
constexpr int kStemPitch = 44;             // floats per patch row in shared memory: the TMA box starts one pixel early (x0 - 4) so that
                                           // its first byte is 16-byte aligned in global memory, and spans 44 floats = 176 bytes
constexpr int kStemShift = 1;              // column of the patch's own first pixel (x0 - 3) inside a row
constexpr int kStemHalfFloats = ((3 * kStemPatchH * kStemPitch * 4 + 127) / 128) * 128 / 4;   // channels 0..2 | 3..5: 128-byte aligned halves
__host__ __device__ constexpr int stem_patch_index(int c, int row, int col) {
  return (c / 3) * kStemHalfFloats + ((c % 3) * kStemPatchH + row) * kStemPitch + col;
}

template <int CIN>
struct StemTcCfg {
  static constexpr int K = 49 * CIN;                 // 147 | 294
  static constexpr int KB = (K + 63) / 64;           // 3 | 5 K-blocks of 64
  static constexpr int NCH = (K + 7) / 8;            // 19 | 37 16-byte chunks per im2col row that hold data
  static constexpr int kPatchBytes = (CIN / 3) * kStemHalfFloats * 4;     // one patch buffer
  static constexpr int kSmem = 1024 + KB * 16384 + KB * 8192 + 2 * kPatchBytes + 256 /*bias*/ + 64;
};

template <int CIN>
__global__ void __launch_bounds__(128, (CIN == 3 ? 2 : 1)) stem_tc_kernel(const __grid_constant__ StemTcParams p) {
  using Cfg = StemTcCfg<CIN>;
  constexpr int KB = Cfg::KB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_s = smem;                                  // KB x 16 KB : [kb][128 rows][128 B]
  uint8_t* w_s = smem + KB * 16384;                     // KB x  8 KB : [kb][ 64 rows][128 B]
  uint8_t* patch_raw = w_s + KB * 8192;                             // 2 x [CIN][21][40] fp32
  float* s_bias = reinterpret_cast<float*>(patch_raw + 2 * Cfg::kPatchBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_bias + 64);
  uint64_t* p_full = bar + 1;                                       // 2: patch buffers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_full + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int total_tiles = p.groups * p.tiles_per_group;
  int cur_g = -1;

  if (tid == 0) { mbar_init(bar, 1); mbar_init(&p_full[0], 1); mbar_init(&p_full[1], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  for (int e = tid; e < KB * 16384 / 16; e += 128) reinterpret_cast<uint4*>(a_s)[e] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t a_addr = smem_u32(a_s), w_addr = smem_u32(w_s);
  uint32_t phase = 0;
  constexpr uint32_t kPatchTx = CIN * kStemPatchH * kStemPitch * 4;
  // one thread: the input patch of tile t -> patch buffer buf  (a macro, not a lambda: the tensor maps must be addressed
  // in the kernel's parameter space, which a by-reference capture of `p` does not guarantee)
#define UOC_STEM_FETCH(T_, BUF_)                                                                                          \
  do {                                                                                                                    \
    const int fg = (T_) / p.tiles_per_group;                                                                              \
    int fr = (T_) - fg * p.tiles_per_group;                                                                               \
    const int ftx = fr % p.tiles_x; fr /= p.tiles_x;                                                                      \
    const int fty = fr % p.tiles_y;                                                                                       \
    const int fn = fr / p.tiles_y;                                                                                        \
    uint8_t* fdst = patch_raw + (BUF_) * Cfg::kPatchBytes;                                                                \
    mbar_arrive_expect_tx(&p_full[(BUF_)], kPatchTx);                                                                     \
    tma_load_3d(fdst, &p.tmap_x[fg], &p_full[(BUF_)], ftx * 32 - 3 - kStemShift, fty * 16 - 3, fn * 3);                                \
    if (CIN == 6)                                                                                                         \
      tma_load_3d(fdst + kStemHalfFloats * 4, &p.tmap_x2[fg], &p_full[(BUF_)], ftx * 32 - 3 - kStemShift, fty * 16 - 3, fn * 3); \
  } while (0)
  if (p.use_tma && tid == 0 && int(blockIdx.x) < total_tiles) UOC_STEM_FETCH(int(blockIdx.x), 0);
  int it = 0;

  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
    const int g = t / p.tiles_per_group;
    int r = t - g * p.tiles_per_group;
    const int tx = r % p.tiles_x; r /= p.tiles_x;
    const int ty = r % p.tiles_y;
    const int n = r / p.tiles_y;
    const int buf = p.use_tma ? (it & 1) : 0;
    float* patch = reinterpret_cast<float*>(patch_raw + buf * Cfg::kPatchBytes);
    // the next tile's patch travels while this one is processed (its buffer was last read before the barrier that ended
    // the previous iteration)
    if (p.use_tma && tid == 0 && t + int(gridDim.x) < total_tiles) UOC_STEM_FETCH(t + int(gridDim.x), buf ^ 1);
    if (g != cur_g) {                                   // (re)load this branch's weights, K-major 128B-swizzled rows
      const uint4* wg = reinterpret_cast<const uint4*>(p.w[g]);
      for (int e = tid; e < 64 * KB * 8; e += 128) {
        const int row = e / (KB * 8), j = e - row * (KB * 8);
        const int kb = j >> 3, cc = j & 7;
        *reinterpret_cast<uint4*>(w_s + kb * 8192 + row * 128 + ((cc ^ (row & 7)) << 4)) = __ldg(wg + e);
      }
      if (tid < 64) s_bias[tid] = __ldg(p.bias[g] + tid);
      cur_g = g;
    }
    if (p.use_tma) {
      if (!mbar_wait(&p_full[buf], uint32_t(it >> 1) & 1u, p.err)) break;
    } else {
    // stage the fp32 input patch (unaligned input / W % 4 != 0)
    const int iy0 = ty * 16 - 3, ix0 = tx * 32 - 3;
    const float* xa = p.x[g] + size_t(n) * 3 * p.H * p.W;                     // channels 0..2
    const float* xb = (CIN == 6) ? p.x2[g] + size_t(n) * 3 * p.H * p.W : xa;  // channels 3..5 (early fusion)
    for (int e = tid; e < CIN * kStemPatchH * kStemPatchW; e += 128) {
      const int c = e / (kStemPatchH * kStemPatchW);
      const int rem = e - c * kStemPatchH * kStemPatchW;
      const int py = rem / kStemPatchW, px = rem - py * kStemPatchW;
      const int iy = iy0 + py, ix = ix0 + px;
      const float* xg = (CIN == 6 && c >= 3) ? xb + size_t(c - 3) * p.H * p.W : xa + size_t(c) * p.H * p.W;
      float v = 0.f;
      if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = __ldg(xg + size_t(iy) * p.W + ix);
      patch[stem_patch_index(c, py, px + kStemShift)] = v;
    }
    }
    __syncthreads();
    // im2col row of pixel `tid` -> NCH chunks of 8 bf16 (k >= K is zero)
    {
      const int ly = tid >> 4, lx = tid & 15;
      const float* pb = patch + (ly * 2) * kStemPitch + lx * 2 + kStemShift;     // + stem_patch_index(c, kr, ks)
      float vals[8];
#pragma unroll
      for (int j = 0; j < Cfg::NCH; ++j) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = j * 8 + q;                 // compile-time after unrolling: tap / channel offsets fold to constants
          float v = 0.f;
          if (k < Cfg::K) {
            const int tap = k / CIN, c = k - tap * CIN;
            const int kr = tap / 7, ks = tap - kr * 7;
            v = pb[stem_patch_index(c, kr, ks)];
          }
          vals[q] = v;
        }
        const uint4 pk = make_uint4(pack_bf16x2(vals[0], vals[1]), pack_bf16x2(vals[2], vals[3]),
                                    pack_bf16x2(vals[4], vals[5]), pack_bf16x2(vals[6], vals[7]));
        const int kb = j >> 3, cc = j & 7;
        *reinterpret_cast<uint4*>(a_s + kb * 16384 + tid * 128 + ((cc ^ (tid & 7)) << 4)) = pk;
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + kb * 16384 + ks * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(w_addr + kb * 8192 + ks * 32, 16, 1024);
            umma_ss_f16(tmem, ad, bd, idesc, (kb | ks) ? 1u : 0u);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    if (!mbar_wait(bar, phase, p.err)) break;
    phase ^= 1u;
    tc_fence_after();
    {
      const int oy = ty * 8 + (tid >> 4), ox = tx * 16 + (tid & 15);
      const bool inb = (oy < p.Ho) && (ox < p.Wo);
      const uint32_t ta = tmem + (uint32_t(warp * 32) << 16);
      uint4* op = reinterpret_cast<uint4*>(p.y[g] + ((size_t(n) * p.Ho + oy) * p.Wo + ox) * 64);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ta + c * 32, v);
        tmem_wait_ld();
        if (inb) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float f[8];
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c * 32 + 8 * e);          // broadcast
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c * 32 + 8 * e + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int h = 0; h < 8; ++h) f[h] = fmaxf(__uint_as_float(v[8 * e + h]) + bb[h], 0.f);
            op[c * 4 + e] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                       pack_bf16x2(f[6], f[7]));
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();          // TMEM and the A tile are free again
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

#undef UOC_STEM_FETCH

template <int CIN>
static int launch_stem_tc_t(const StemTcParams& p, cudaStream_t stream) {
  using Cfg = StemTcCfg<CIN>;
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&stem_tc_kernel<CIN>), int(Cfg::kSmem));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  const int per_sm = (CIN == 3) ? 2 : 1;
  int grid = per_sm * (sm_count() > 0 ? sm_count() : 148);
  if (grid > p.groups * p.tiles_per_group) grid = p.groups * p.tiles_per_group;
  stem_tc_kernel<CIN><<<grid, 128, Cfg::kSmem, stream>>>(p);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

int launch_stem_tc(const StemGroup* g, const void* const* w_bf16, int groups, int cin, int N, int H, int W, cudaStream_t stream) {
  if (groups < 1 || groups > 2) return fail(UOC_ERR_INVALID, "stem_tc: 1 or 2 branches");
  if (cin != 3 && cin != 6) return fail(UOC_ERR_UNSUPPORTED, "stem supports 3 or 6 input channels");
  StemTcParams p;
  for (int i = 0; i < 2; ++i) {
    const int j = i < groups ? i : 0;
    p.x[i] = g[j].x;
    p.x2[i] = g[j].x2;
    p.w[i] = static_cast<const __nv_bfloat16*>(w_bf16[j]);
    p.bias[i] = g[j].bias;
    p.y[i] = static_cast<__nv_bfloat16*>(g[j].y);
  }
  p.N = N; p.H = H; p.W = W;
  p.Ho = (H + 6 - 7) / 2 + 1;
  p.Wo = (W + 6 - 7) / 2 + 1;
  p.tiles_x = (p.Wo + 15) / 16;
  p.tiles_y = (p.Ho + 7) / 8;
  p.tiles_per_group = N * p.tiles_y * p.tiles_x;
  p.groups = groups;
  p.err = device_error_word();
  if (!p.err) return fail(UOC_ERR_CUDA, "no device error word");
  // input patches by TMA when the planes allow it (16-byte aligned base, rows a multiple of 16 bytes)
  p.use_tma = (W % 4 == 0) ? 1 : 0;
  for (int i = 0; i < groups && p.use_tma; ++i) {
    if (reinterpret_cast<uintptr_t>(g[i].x) % 16 != 0 || (cin == 6 && reinterpret_cast<uintptr_t>(g[i].x2) % 16 != 0)) p.use_tma = 0;
  }
  if (p.use_tma) {
    const uint64_t dims[3] = {uint64_t(W), uint64_t(H), uint64_t(3) * N};
    const uint64_t strides[2] = {uint64_t(W) * 4, uint64_t(H) * W * 4};
    const uint32_t box[3] = {uint32_t(kStemPitch), uint32_t(kStemPatchH), 3};
    for (int i = 0; i < 2; ++i) {
      const int j = i < groups ? i : 0;
      int rc = make_tmap_f32(&p.tmap_x[i], g[j].x, 3, dims, strides, box);
      if (rc != UOC_OK) return rc;
      if (cin == 6) {
        rc = make_tmap_f32(&p.tmap_x2[i], g[j].x2, 3, dims, strides, box);
        if (rc != UOC_OK) return rc;
      }
    }
  }
  return cin == 3 ? launch_stem_tc_t<3>(p, stream) : launch_stem_tc_t<6>(p, stream);
}

}  // namespace uoc
