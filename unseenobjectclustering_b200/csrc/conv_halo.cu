// K1b: 3x3 stride-1 (dilated) convolution with on-chip halo reuse -- the second-generation implicit GEMM.
//
// Measured on B200 (profiles/r01_tma_ingest_rate.txt): one SM ingests at most ~48 B/clk from L2 (TMA, multicast or
// not), while tcgen05 consumes 128x128x64 bf16 MACs in 256 clk.  The first-generation kernel (conv_tc.cu) re-fetches the
// activation tile once per filter tap (9 x 16 KB per 64 channels) plus the weight tile and is therefore ingest-bound at
// ~40 % of the tensor peak.  Here
//   * the CTA's output tile is 16 rows x (8*SUB) columns; ONE TMA box fetches the whole input halo
//     [(16 + 2*dil) rows][pitch columns][64 channels] per 64-channel block, and all nine taps are issued from it:
//     tap (r, s) of sub-tile u is the UMMA operand that starts at row r*dil, column s*dil + 8u of the halo, i.e. a
//     K-major 128B-swizzled matrix whose 8-row groups (one output row = 8 pixels = 1024 B) are `pitch*128` bytes apart
//     (stride-byte-offset), with the descriptor's base-offset field carrying the swizzle phase of the unaligned start;
//   * SUB = 2 sub-tiles (M = 2 x 128) share every weight tile, halving the weight bytes per MAC.
//   XHALO = false is the conservative variant (three boxes per block, one per horizontal tap, each 1024-byte aligned:
//   vertical reuse only) kept for A/B tests.
// Same epilogue as conv_tc.cu (bias, residual, ReLU, bf16 / fp32 NHWC store).
#include <cstdlib>
#include <cstring>

#include "conv.cuh"

namespace uoc {

namespace {

constexpr int kThreads = 192;

struct HaloParams {
  CUtensorMap tmap_x[2];
  CUtensorMap tmap_w[2];
  const float* bias[2];
  const void* residual[2];
  void* y[2];
  int N, Ho, Wo, Cin, Cout;
  int tiles_x, tiles_y, n_tiles;
  int dil, relu, out_fp32;
  int pitch;            // halo columns per row (pixels)
  int rows;             // 16 + 2*dil
  int a_bytes;          // bytes of one activation buffer (all boxes)
  int use_base_offset;  // descriptor base-offset field = swizzle phase of the operand start (A/B knob)
  int a_stride;         // bytes between the two activation buffers (a_bytes rounded up to 1024)
  int b_stages;         // depth of the weight-tile ring (as many as fit next to the two halo buffers)
  int part_rows;        // halo rows per TMA slice
  int a_bufs;           // 2 = double-buffered halo (one big CTA per SM), 1 = single buffer (small CTA, two per SM)
  unsigned int* err;
};

__device__ __forceinline__ uint64_t make_desc_sw128_off(uint32_t smem_addr, uint32_t sbo_bytes, int use_base_offset) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFFu);
  d |= uint64_t(1) << 16;                                  // LBO (unused for swizzled K-major)
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= uint64_t(1) << 46;                                  // version
  if (use_base_offset) d |= uint64_t((smem_addr >> 7) & 0x7u) << 49;   // base offset: swizzle phase of an unaligned start
  d |= uint64_t(2) << 61;                                  // SWIZZLE_128B
  return d;
}

constexpr int kMaxBStages = 12;
constexpr int kSmemBudget = 226 * 1024;      // of the 227 KB a CTA may own

template <int BLOCK_N, int SUB, bool XHALO>
struct HaloCfg {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr uint32_t kTmemCols = (SUB * BLOCK_N <= 128) ? 128 : ((SUB * BLOCK_N <= 256) ? 256 : 512);
};

template <int BLOCK_N, int SUB, bool XHALO>
__global__ void __launch_bounds__(kThreads, 2)
conv_halo_kernel(const __grid_constant__ HaloParams p) {
  using Cfg = HaloCfg<BLOCK_N, SUB, XHALO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NB = p.b_stages;
  uint8_t* abuf = smem;                                  // 2 x a_stride
  uint8_t* bbuf = smem + p.a_bufs * p.a_stride;          // NB x kBBytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(bbuf + NB * Cfg::kBBytes);
  uint64_t* a_full = bars;                // 2
  uint64_t* a_empty = bars + 2;           // 2
  uint64_t* b_full = bars + 4;            // NB
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* acc_full = b_empty + kMaxBStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  const int n_tile = blockIdx.x % p.n_tiles;
  int m_tile = blockIdx.x / p.n_tiles;
  const int tx = m_tile % p.tiles_x; m_tile /= p.tiles_x;
  const int ty = m_tile % p.tiles_y;
  const int img = m_tile / p.tiles_y;
  const int n0 = n_tile * BLOCK_N;
  const int CB = p.Cin >> 6;
  const int x0 = tx * 8 * SUB, y0 = ty * 16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x[g]);
    tma_prefetch_desc(&p.tmap_w[g]);
    for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // The halo of block cb+1 is fetched in slices of `p.part_rows` rows, interleaved with the nine weight tiles of
      // block cb, so that neither stream monopolises the TMA queue.
      const int nparts = XHALO ? (p.rows + p.part_rows - 1) / p.part_rows : 3;
      const int part_bytes = XHALO ? p.pitch * p.part_rows * 128 : p.a_bytes / 3;
      const int AB = p.a_bufs;
      auto issue_part = [&](int cb, int part) {
        const int ab = cb % AB;
        uint8_t* adst = abuf + ab * p.a_stride + part * part_bytes;
        if (XHALO) tma_load_4d(adst, &p.tmap_x[g], &a_full[ab], cb * 64, x0 - p.dil, y0 - p.dil + part * p.part_rows, img);
        else tma_load_4d(adst, &p.tmap_x[g], &a_full[ab], cb * 64, x0 + (part - 1) * p.dil, y0 - p.dil, img);
      };
      // block 0: everything up front
      mbar_arrive_expect_tx(&a_full[0], p.a_bytes);
      for (int part = 0; part < nparts; ++part) issue_part(0, part);
      int bcount = 0;
      bool ok = true;
      for (int cb = 0; cb < CB && ok; ++cb) {
        const bool has_next = cb + 1 < CB;
        const int nab = (cb + 1) % AB;
        int next_part = 0;
        bool next_armed = false;
        for (int tap = 0; tap < 9 && ok; ++tap, ++bcount) {
          const int s = bcount % NB;
          if (!mbar_wait(&b_empty[s], ((bcount / NB) & 1) ^ 1u, p.err)) { ok = false; break; }
          mbar_arrive_expect_tx(&b_full[s], Cfg::kBBytes);
          tma_load_2d(bbuf + s * Cfg::kBBytes, &p.tmap_w[g], &b_full[s], tap * p.Cin + cb * 64, n0);
          if (has_next) {
            if (!next_armed && mbar_try_wait(&a_empty[nab], ((((cb + 1) / AB)) & 1) ^ 1u)) {
              mbar_arrive_expect_tx(&a_full[nab], p.a_bytes);
              next_armed = true;
            }
            if (next_armed) {
              const int upto = (nparts * (tap + 1) + 8) / 9;           // spread the slices over the nine taps
              for (; next_part < upto && next_part < nparts; ++next_part) issue_part(cb + 1, next_part);
            }
          }
        }
        if (ok && has_next) {
          if (!next_armed) {
            if (!mbar_wait(&a_empty[nab], ((((cb + 1) / AB)) & 1) ^ 1u, p.err)) { ok = false; break; }
            mbar_arrive_expect_tx(&a_full[nab], p.a_bytes);
          }
          for (; next_part < nparts; ++next_part) issue_part(cb + 1, next_part);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 0, 0);
      const uint32_t a_addr0 = smem_u32(abuf), b_addr0 = smem_u32(bbuf);
      const uint32_t row_bytes = uint32_t(p.pitch) * 128u;      // one halo row = one output row's stride between 8-pixel groups
      int bcount = 0;
      bool ok = true;
      for (int cb = 0; cb < CB && ok; ++cb) {
        const int ab = cb % p.a_bufs;
        if (!mbar_wait(&a_full[ab], (cb / p.a_bufs) & 1, p.err)) { ok = false; break; }
        const uint32_t a_base = a_addr0 + ab * p.a_stride;
        for (int tap = 0; tap < 9; ++tap, ++bcount) {
          const int s = bcount % NB;
          if (!mbar_wait(&b_full[s], (bcount / NB) & 1, p.err)) { ok = false; break; }
          tc_fence_after();
          const int r = tap / 3, sx = tap - 3 * r;
          const uint32_t b_addr = b_addr0 + s * Cfg::kBBytes;
#pragma unroll
          for (int u = 0; u < SUB; ++u) {
            uint32_t a_tap;
            if (XHALO) a_tap = a_base + uint32_t(r * p.dil) * row_bytes + uint32_t(sx * p.dil + 8 * u) * 128u;
            else a_tap = a_base + uint32_t(sx) * uint32_t(p.a_bytes / 3) + uint32_t(r * p.dil) * row_bytes + uint32_t(8 * u) * 128u;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = make_desc_sw128_off(a_tap + ks * 32, row_bytes, p.use_base_offset);
              const uint64_t bd = make_smem_desc_sw128(b_addr + ks * 32, 16, 1024);
              umma_ss_f16(tmem_base + u * BLOCK_N, ad, bd, idesc, (cb | tap | ks) ? 1u : 0u);
            }
          }
          umma_commit(&b_empty[s]);
        }
        if (ok) umma_commit(&a_empty[ab]);
      }
      if (ok) umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // accumulator row = (output row row>>3, output column row&7 within the sub-tile)
    if (mbar_wait(acc_full, 0, p.err)) {
      tc_fence_after();
      const uint32_t ta = tmem_base + (uint32_t(q * 32) << 16);
      const int oy = y0 + (row >> 3);
#pragma unroll
      for (int u = 0; u < SUB; ++u) {
        const int ox = x0 + 8 * u + (row & 7);
        const bool inb = (oy < p.Ho) && (ox < p.Wo) && (img < p.N);
        const size_t pix = (size_t(img) * p.Ho + oy) * p.Wo + ox;
        const float* bias = p.bias[g] + n0;
#pragma unroll
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(ta + u * BLOCK_N + c * 32, v);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int BLOCK_N, int SUB, bool XHALO>
int launch(HaloParams prm, int ctas, int groups, cudaStream_t stream) {
  using Cfg = HaloCfg<BLOCK_N, SUB, XHALO>;
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&conv_halo_kernel<BLOCK_N, SUB, XHALO>), int(kSmemBudget));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  prm.a_stride = (prm.a_bytes + 1023) / 1024 * 1024;
  int budget = kSmemBudget;
  prm.a_bufs = 2;
  if (const char* e = getenv("UOC_CONV_HALO_SMALL")) {
    if (atoi(e) != 0) { budget = 112 * 1024; prm.a_bufs = 1; }     // two CTAs per SM hide each other's halo-fetch bubble
  }
  int nb = (budget - 1024 - prm.a_bufs * prm.a_stride - 512) / Cfg::kBBytes;
  if (nb > kMaxBStages) nb = kMaxBStages;
  if (const char* e = getenv("UOC_CONV_HALO_BSTAGES")) { int v = atoi(e); if (v >= 2 && v < nb) nb = v; }
  if (nb < 2) return fail(UOC_ERR_UNSUPPORTED, "conv_halo: tile does not fit in shared memory");
  prm.b_stages = nb;
  const int smem_bytes = 1024 + prm.a_bufs * prm.a_stride + nb * Cfg::kBBytes + 512;
  conv_halo_kernel<BLOCK_N, SUB, XHALO><<<dim3(ctas, 1, groups), kThreads, smem_bytes, stream>>>(prm);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

template <bool XHALO>
int dispatch(const HaloParams& prm, int block_n, int sub, int ctas, int groups, cudaStream_t st) {
  if (block_n == 128 && sub == 2) return launch<128, 2, XHALO>(prm, ctas, groups, st);
  if (block_n == 128 && sub == 1) return launch<128, 1, XHALO>(prm, ctas, groups, st);
  if (block_n == 64 && sub == 2) return launch<64, 2, XHALO>(prm, ctas, groups, st);
  if (block_n == 64 && sub == 1) return launch<64, 1, XHALO>(prm, ctas, groups, st);
  return fail(UOC_ERR_UNSUPPORTED, "conv_halo: unsupported tile configuration");
}

}  // namespace

bool conv_halo_supported(const ConvProblem& p) {
  return p.ksize == 3 && p.stride == 1 && (p.dilation == 1 || p.dilation == 2 || p.dilation == 4) && p.Cin % 64 == 0 &&
         p.Cout % 64 == 0 && p.groups >= 1 && p.groups <= 2;
}

int launch_conv_halo(const ConvProblem& p, cudaStream_t stream) {
  if (!conv_halo_supported(p)) return fail(UOC_ERR_UNSUPPORTED, "conv_halo supports 3x3 stride-1 convolutions with dilation 1/2/4");
  int xhalo = 1;
  if (const char* e = getenv("UOC_CONV_XHALO")) xhalo = atoi(e);
  HaloParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.N = p.N; prm.Ho = p.H; prm.Wo = p.W; prm.Cin = p.Cin; prm.Cout = p.Cout;
  prm.dil = p.dilation; prm.relu = p.relu; prm.out_fp32 = p.out_fp32;
  const int block_n = (p.Cout % 128 == 0) ? 128 : 64;
  prm.n_tiles = p.Cout / block_n;
  // sub-tiles: two (16 x 16 outputs) unless that leaves most SMs without a tile
  int sub = 2;
  {
    const long long tiles2 = (long long)p.N * ((p.H + 15) / 16) * ((p.W + 15) / 16) * prm.n_tiles * p.groups;
    if (tiles2 < 64) sub = 1;
  }
  if (const char* e = getenv("UOC_CONV_SUB")) { int v = atoi(e); if (v == 1 || v == 2) sub = v; }
  if (!xhalo) sub = 1;                       // three aligned boxes per block only fit for one sub-tile
  prm.tiles_x = (p.W + 8 * sub - 1) / (8 * sub);
  prm.tiles_y = (p.H + 15) / 16;
  prm.rows = 16 + 2 * p.dilation;
  prm.pitch = xhalo ? (sub == 2 ? 24 : 16) : 8 * sub;
  prm.a_bytes = (xhalo ? 1 : 3) * prm.pitch * prm.rows * 128;
  prm.part_rows = 2;                              // rows = 18 / 20 / 24 are all even
  if (const char* e = getenv("UOC_CONV_HALO_PART_ROWS")) { int v = atoi(e); if (v >= 1 && prm.rows % v == 0) prm.part_rows = v; }
  prm.err = device_error_word();
  if (!prm.err) return fail(UOC_ERR_CUDA, "no device error word");
  prm.use_base_offset = 0;
  if (const char* e = getenv("UOC_CONV_BASE_OFFSET")) prm.use_base_offset = atoi(e);
  for (int g = 0; g < p.groups; ++g) {
    const uint64_t xd[4] = {uint64_t(p.Cin), uint64_t(p.W), uint64_t(p.H), uint64_t(p.N)};
    const uint64_t xs[3] = {uint64_t(p.Cin) * 2, uint64_t(p.W) * p.Cin * 2, uint64_t(p.H) * p.W * p.Cin * 2};
    const uint32_t xb[4] = {64, uint32_t(prm.pitch), uint32_t(xhalo ? prm.part_rows : prm.rows), 1};
    int rc = make_tmap_bf16(&prm.tmap_x[g], p.g[g].x, 4, xd, xs, xb, nullptr);
    if (rc != UOC_OK) return rc;
    const uint64_t wd[2] = {uint64_t(9) * p.Cin, uint64_t(p.Cout)};
    const uint64_t wsb[1] = {uint64_t(9) * p.Cin * 2};
    const uint32_t wb[2] = {64, uint32_t(block_n)};
    rc = make_tmap_bf16(&prm.tmap_w[g], p.g[g].w, 2, wd, wsb, wb, nullptr);
    if (rc != UOC_OK) return rc;
    prm.bias[g] = p.g[g].bias;
    prm.residual[g] = p.g[g].residual;
    prm.y[g] = p.g[g].y;
  }
  const int ctas = p.N * prm.tiles_y * prm.tiles_x * prm.n_tiles;
  return xhalo ? dispatch<true>(prm, block_n, sub, ctas, p.groups, stream) : dispatch<false>(prm, block_n, sub, ctas, p.groups, stream);
}

}  // namespace uoc
