// K1: convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05), NHWC bf16, fp32 accumulate.
// Replaces the cuDNN convolutions + BatchNorm + ReLU + residual add of lib/networks/resnet.py:57-73,
// :236-270 (BN folded into weights/bias on the host, see backbone.cu).
//
//   M tile = 8 x 16 output pixels (128 rows), N tile = BLOCK_N output channels, K block = 64 input
//   channels of one filter tap.  For tap (r,s) the A tile is the activation box
//   [n][ty*8*stride + r*dil - pad : +8*stride : stride][tx*16*stride + s*dil - pad : ...][cb*64 : +64]
//   fetched by ONE tiled-mode TMA (4-D tensor map over [N][H][W][C], element strides for stride-2
//   convolutions, out-of-bounds -> zero fill implements the padding) straight into the 128B-swizzled
//   K-major layout the UMMA descriptor expects; no im2col buffer exists anywhere.
//   warp 0: TMA producer, warp 1: MMA issuer (one thread), warps 2-5: epilogue
//   (TMEM -> registers -> +bias (+residual) -> ReLU -> bf16/fp32 -> global).
//
//   Two CTAs per SM.  This is the kernel for the SMALL layers (Cout <= 128, and everything short at batch 1); the wide /
//   long convolutions run on the persistent CTA-pair kernel (conv_pair.cu) -- launch_conv_auto picks per layer.
#include <cstdlib>
#include <cstring>

#include "conv.cuh"

namespace uoc {

namespace {

constexpr int kThreads = 192;
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 ch bf16

struct ConvTcParams {
  CUtensorMap tmap_x[2];
  CUtensorMap tmap_w[2];
  const float* bias[2];
  const void* residual[2];
  void* y[2];
  int N, Ho, Wo, Cin, Cout;
  int tiles_x, tiles_y, n_tiles;
  int ksize, stride, dil, pad;
  int relu, out_fp32;
  unsigned int* err;
};

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 128) ? 3 : 4;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
  static constexpr uint32_t kTmemCols = BLOCK_N;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads)
conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using Cfg = ConvCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + Cfg::kStages;
  uint64_t* acc_full = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  // blockIdx.x = m_tile * n_tiles + n_tile
  const int n_tile = int(blockIdx.x) % p.n_tiles;
  int m_tile = int(blockIdx.x) / p.n_tiles;
  const int tx = m_tile % p.tiles_x; m_tile /= p.tiles_x;
  const int ty = m_tile % p.tiles_y;
  const int img = m_tile / p.tiles_y;
  const int n0 = n_tile * BLOCK_N;
  const int cblocks = p.Cin >> 6;
  const int KB = p.ksize * p.ksize * cblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x[g]);
    tma_prefetch_desc(&p.tmap_w[g]);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the
  // previous layer's tail; its activations are complete and visible after the wait.  The next layer may start its own
  // prologue as soon as every CTA of this grid has got here (it waits for this grid's completion before touching memory).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    if (elect_one()) {    // one elected lane: lets the compiler keep descriptors / barriers in uniform registers
      const int x_base = tx * 16 * p.stride - p.pad;
      const int y_base = ty * 8 * p.stride - p.pad;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % Cfg::kStages;
        if (!mbar_wait(&empty[s], ((kb / Cfg::kStages) & 1) ^ 1u, p.err)) break;
        const int tap = kb / cblocks, cb = kb - tap * cblocks;
        const int r = tap / p.ksize, sx = tap - r * p.ksize;
        uint8_t* st = smem + s * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
        tma_load_4d(st, &p.tmap_x[g], &full[s], cb * 64, x_base + sx * p.dil, y_base + r * p.dil, img);
        tma_load_2d(st + kABytes, &p.tmap_w[g], &full[s], kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {    // one elected lane: lets the compiler keep descriptors / barriers in uniform registers
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      bool ok = true;
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % Cfg::kStages;
        if (!mbar_wait(&full[s], (kb / Cfg::kStages) & 1, p.err)) { ok = false; break; }
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = make_smem_desc_sw128(a_addr + ks * 32, 16, 1024);
          const uint64_t bd = make_smem_desc_sw128(b_addr + ks * 32, 16, 1024);
          umma_ss_f16(tmem_base, ad, bd, idesc, (kb | ks) ? 1u : 0u);
        }
        umma_commit(&empty[s]);
      }
      if (ok) umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int oy = ty * 8 + (row >> 4), ox = tx * 16 + (row & 15);
    const bool inb = (oy < p.Ho) && (ox < p.Wo) && (img < p.N);
    const size_t pix = (size_t(img) * p.Ho + oy) * p.Wo + ox;
    const float* bias = p.bias[g] + n0;
    if (mbar_wait(acc_full, 0, p.err)) {
      tc_fence_after();
      const uint32_t ta = tmem_base + (uint32_t(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ta + c * 32, v);
        tmem_wait_ld();
        if (inb) {
          float f[32];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + e * 4));
            f[4 * e + 0] = __uint_as_float(v[4 * e + 0]) + bv.x;
            f[4 * e + 1] = __uint_as_float(v[4 * e + 1]) + bv.y;
            f[4 * e + 2] = __uint_as_float(v[4 * e + 2]) + bv.z;
            f[4 * e + 3] = __uint_as_float(v[4 * e + 3]) + bv.w;
          }
          if (p.residual[g]) {
            const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual[g]) +
                                                             pix * p.Cout + n0 + c * 32);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint4 rv = __ldg(rp + e);
              const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                f[8 * e + 2 * h + 0] += __uint_as_float(w4[h] << 16);
                f[8 * e + 2 * h + 1] += __uint_as_float(w4[h] & 0xFFFF0000u);
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          if (p.out_fp32) {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.y[g]) + pix * p.Cout + n0 + c * 32);
#pragma unroll
            for (int e = 0; e < 8; ++e) op[e] = make_float4(f[4 * e], f[4 * e + 1], f[4 * e + 2], f[4 * e + 3]);
          } else {
            uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.y[g]) + pix * p.Cout + n0 + c * 32);
#pragma unroll
            for (int e = 0; e < 4; ++e)
              op[e] = make_uint4(pack_bf16x2(f[8 * e + 0], f[8 * e + 1]), pack_bf16x2(f[8 * e + 2], f[8 * e + 3]),
                                 pack_bf16x2(f[8 * e + 4], f[8 * e + 5]), pack_bf16x2(f[8 * e + 6], f[8 * e + 7]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int BLOCK_N>
int launch(const ConvTcParams& prm, int m_tiles, int groups, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N>;
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&conv_tc_kernel<BLOCK_N>), int(Cfg::kSmemBytes));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(m_tiles * prm.n_tiles, 1, groups);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous layer's tail
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = 1;
  UOC_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N>, prm));
  count_launch();
  return UOC_OK;
}

}  // namespace

int launch_conv_tc(const ConvProblem& p, cudaStream_t stream) {
  if (p.Cin % 64 != 0 || p.Cout % 64 != 0) return fail(UOC_ERR_UNSUPPORTED, "conv_tc needs Cin % 64 == 0 and Cout % 64 == 0");
  if (p.ksize != 1 && p.ksize != 3) return fail(UOC_ERR_UNSUPPORTED, "conv_tc supports 1x1 and 3x3 filters");
  if (p.stride != 1 && p.stride != 2) return fail(UOC_ERR_UNSUPPORTED, "conv_tc supports stride 1 and 2");
  if (p.groups < 1 || p.groups > 2) return fail(UOC_ERR_INVALID, "groups must be 1 or 2");
  ConvTcParams prm;
  memset(&prm, 0, sizeof(prm));
  const int pad = (p.ksize == 3) ? p.dilation : 0;
  prm.Ho = conv_out_dim(p.H, p.ksize, p.stride, p.dilation);
  prm.Wo = conv_out_dim(p.W, p.ksize, p.stride, p.dilation);
  prm.Cin = p.Cin; prm.Cout = p.Cout; prm.N = p.N;
  prm.tiles_x = (prm.Wo + 15) / 16;
  prm.tiles_y = (prm.Ho + 7) / 8;
  const int block_n = (p.Cout % 128 == 0) ? 128 : 64;
  prm.n_tiles = p.Cout / block_n;
  prm.ksize = p.ksize; prm.stride = p.stride; prm.dil = p.dilation; prm.pad = pad;
  prm.relu = p.relu; prm.out_fp32 = p.out_fp32;
  prm.err = device_error_word();
  if (!prm.err) return fail(UOC_ERR_CUDA, "no device error word");
  const int taps = p.ksize * p.ksize;
  for (int g = 0; g < p.groups; ++g) {
    const uint64_t xd[4] = {uint64_t(p.Cin), uint64_t(p.W), uint64_t(p.H), uint64_t(p.N)};
    const uint64_t xs[3] = {uint64_t(p.Cin) * 2, uint64_t(p.W) * p.Cin * 2, uint64_t(p.H) * p.W * p.Cin * 2};
    const uint32_t xb[4] = {64, uint32_t(16 * p.stride), uint32_t(8 * p.stride), 1};
    const uint32_t xe[4] = {1, uint32_t(p.stride), uint32_t(p.stride), 1};
    int rc = make_tmap_bf16(&prm.tmap_x[g], p.g[g].x, 4, xd, xs, xb, xe);
    if (rc != UOC_OK) return rc;
    const uint64_t wd[2] = {uint64_t(taps) * p.Cin, uint64_t(p.Cout)};
    const uint64_t wsb[1] = {uint64_t(taps) * p.Cin * 2};
    const uint32_t wb[2] = {64, uint32_t(block_n)};
    rc = make_tmap_bf16(&prm.tmap_w[g], p.g[g].w, 2, wd, wsb, wb, nullptr);
    if (rc != UOC_OK) return rc;
    prm.bias[g] = p.g[g].bias;
    prm.residual[g] = p.g[g].residual;
    prm.y[g] = p.g[g].y;
  }
  const int m_tiles = p.N * prm.tiles_y * prm.tiles_x;
  if (block_n == 128) return launch<128>(prm, m_tiles, p.groups, stream);
  return launch<64>(prm, m_tiles, p.groups, stream);
}

int launch_conv_auto(const ConvProblem& p, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  // Two live kernels, chosen per layer by the amount of work (profiles/r02_conv_layers.txt): the persistent CTA-pair
  // stream-K kernel wins where a pair gets a long K range of a wide tile (layers 3 / 4; every layer-3 / 4 convolution when
  // several frames share a launch); the one-tile-per-CTA kernel wins on the small layers (Cout <= 128, and everything
  // short at batch 1), where fixed per-pair costs dominate.  UOC_CONV_PAIR=0 / 1 forces one of them (parity tests).
  // layers 1 / 2 (3x3, stride 1, no dilation, Cin <= 128): the weights-resident halo kernel (UOC_CONV_WRES=0 disables it)
  if (knobs().conv_wres != 0 && conv_wres_supported(p)) return launch_conv_wres(p, stream);
  int use_pair = knobs().conv_pair;
  if (!scratch || !conv_pair_supported(p)) use_pair = 0;
  if (use_pair < 0) {
    const int Ho = conv_out_dim(p.H, p.ksize, p.stride, p.dilation), Wo = conv_out_dim(p.W, p.ksize, p.stride, p.dilation);
    const long long m_pairs = ((long long)p.N * ((Ho + 7) / 8) * ((Wo + 15) / 16) + 1) / 2;
    const long long tiles = m_pairs * (p.Cout / 256) * p.groups;
    const long long units = tiles * p.ksize * p.ksize * (p.Cin / 64);
    int pairs = sm_count() / 2;
    if (pairs < 1) pairs = 1;
    const long long per_pair = units / pairs;
    use_pair = (p.Cout % 256 == 0 && !p.out_fp32 && per_pair >= (p.ksize == 3 ? 32 : 16)) ? 1 : 0;
  }
  if (use_pair) return launch_conv_pair(p, scratch, scratch_bytes, stream);
  return launch_conv_tc(p, stream);
}

}  // namespace uoc
