// K1'' : the 3x3 / stride-1 / dilation-1 convolutions with few input channels (ResNet34 layer1: 64 -> 64, layer2: 128 -> 128;
// lib/networks/resnet.py:57-73 inside :236-270) with the WEIGHTS RESIDENT in shared memory and ONE halo fetch per tile.
//
// Why (measured, profiles/r02_conv_layers.md): the general kernels fetch the activation tile once per filter tap
// (9 x 16 KB per 128 pixels and 64 channels) and sit on the ~48 B/clk an SM can ingest through TMA; for these layers that
// is 0.3 PFLOP/s.  Here
//   * a CTA is persistent and owns ONE (branch, 64-channel output half) for the whole launch: its 9 x Cin x 64 weights
//     (72 KB at Cin = 64, 144 KB at Cin = 128) are loaded once and stay in shared memory;
//   * the output tile is 16 rows x 8 columns; ONE TMA box fetches its input halo [18 rows][10 columns][64 channels]
//     (22.5 KB instead of 9 x 16 KB) per 64-channel block, and all nine taps are issued from it: tap (r, s) is the UMMA
//     operand that starts at halo pixel (r, s), a K-major 128B-swizzled matrix whose 8-row groups (one output row = 8
//     pixels = 1024 B) are 10 pixels = 1280 B apart (stride-byte offset).  The hardware derives the swizzle phase from
//     absolute shared-memory address bits, so operands may start at any 128-byte row of a TMA-written box
//     (profiles/r01_conv_halo_vs_tc.md);
//   * four TMEM accumulators (64 columns each): the epilogue of a tile overlaps the main loops of the next ones;
//   * epilogue per warp: tcgen05.ld -> +bias -> +residual (TMA load of the warp's 32 pixels into its staging block) -> ReLU
//     -> bf16 -> 128B-swizzled staging -> TMA store (clips ragged tiles).
// The main loop is then bound by the shared-memory operand reads of the tensor core (M 128 x N 64 x K 16 SS: 48 clk per
// instruction, profiles/r01_mma_issue_rate.txt): 36 instructions = 1 728 clk per tile and 64-channel block.
//   warp 0: TMA producer, warp 1: MMA issuer (one elected thread), warps 2..: epilogue groups of four warps.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv.cuh"

namespace uoc {

namespace {

constexpr int kTileH = 16, kTileW = 8;
constexpr int kHaloH = kTileH + 2, kHaloW = kTileW + 2;
constexpr int kHaloBytes = kHaloH * kHaloW * 128;     // 23 040
constexpr int kHaloStride = 23 * 1024;                // stage pitch (1024-byte aligned)
constexpr int kWTileBytes = 64 * 128;                 // [64 output channels][64 input channels] bf16
constexpr int kAcc = 4;                               // TMEM accumulators of 64 columns

struct WresParams {
  CUtensorMap tmap_x[2];
  CUtensorMap tmap_w[2];
  CUtensorMap tmap_y[2];  // bf16 output [N][Ho][Wo][Cout], box [64 ch][8][4][1] = one epilogue warp's 32 pixels
  CUtensorMap tmap_r[2];  // residual, same shape and box
  const float* bias[2];
  int has_res;
  int tiles_x, tiles_y, tiles;   // per group: N * tiles_y * tiles_x
  int n_halves, groups, ctas_per_combo;
  int relu;
  unsigned int* err;
  long long* trace;       // [16] clock sums of CTA 0 (knob conv_trace): see launch_conv_wres
};

template <int KBLK>
struct WresCfg {
  static constexpr int kStages = (KBLK == 1) ? 4 : 2;
  static constexpr int kEpiGroups = (KBLK == 1) ? 2 : 1;
  static constexpr int kThreads = 64 + 128 * kEpiGroups;
  static constexpr int kWBytes = 9 * KBLK * kWTileBytes;
  static constexpr int kOutBytes = kEpiGroups * 16384;
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = 1024 + kWBytes + kStages * kHaloStride + kOutBytes + 256 /*bias*/ + kBarBytes;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

template <int KBLK>
__global__ void __launch_bounds__(WresCfg<KBLK>::kThreads, 1)
conv_wres_kernel(const __grid_constant__ WresParams p) {
  using Cfg = WresCfg<KBLK>;
  constexpr int S = Cfg::kStages, G = Cfg::kEpiGroups;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wbuf = smem;
  uint8_t* stages = wbuf + Cfg::kWBytes;
  uint8_t* outbuf = stages + S * kHaloStride;
  float* s_bias = reinterpret_cast<float*>(outbuf + Cfg::kOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(outbuf + Cfg::kOutBytes + 256);
  uint64_t* full = bars;
  uint64_t* empty = full + S;
  uint64_t* w_full = empty + S;              // one per (K block, tap): the first tile starts as soon as its first taps are in
  uint64_t* acc_full = w_full + 9 * KBLK;
  uint64_t* acc_empty = acc_full + kAcc;
  uint64_t* res_full = acc_empty + kAcc;     // one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 4 * G);
  static_assert((2 * S + 9 * KBLK + 2 * kAcc + 4 * G) * 8 + 8 <= Cfg::kBarBytes, "barrier region");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncombo = p.groups * p.n_halves;
  const int combo = int(blockIdx.x) % ncombo;
  const int g = combo / p.n_halves;
  const int n0 = (combo - g * p.n_halves) * 64;
  const int r = int(blockIdx.x) / ncombo, R = p.ctas_per_combo;
  const int my_tiles = (r < p.tiles) ? (p.tiles - r + R - 1) / R : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x[g]);
    tma_prefetch_desc(&p.tmap_w[g]);
    tma_prefetch_desc(&p.tmap_y[g]);
    if (p.has_res) tma_prefetch_desc(&p.tmap_r[g]);
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < 9 * KBLK; ++i) mbar_init(&w_full[i], 1);
    for (int a = 0; a < kAcc; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    for (int i = 0; i < 4 * G; ++i) mbar_init(&res_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kAcc * 64);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128) s_bias[threadIdx.x - 64] = __ldg(p.bias[g] + n0 + (threadIdx.x - 64));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // the weights are constants of the model: fetch them before waiting for the previous layer (PDL)
      for (int kb = 0; kb < KBLK; ++kb)
        for (int tap = 0; tap < 9; ++tap) {
          const int idx = kb * 9 + tap;
          mbar_arrive_expect_tx(&w_full[idx], kWTileBytes);
          tma_load_2d(wbuf + idx * kWTileBytes, &p.tmap_w[g], &w_full[idx], (tap * KBLK + kb) * 64, n0);
        }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      long long* trc = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
      const long long t0 = trc ? clock64() : 0;
      long long t_wait = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int t = r + i * R;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int img = t / p.tiles_y;
        bool ok = true;
        for (int kb = 0; kb < KBLK; ++kb) {
          const int j = i * KBLK + kb, s = j % S;
          const long long w0 = trc ? clock64() : 0;
          if (!mbar_wait(&empty[s], ((j / S) & 1) ^ 1u, p.err)) { ok = false; break; }
          if (trc) t_wait += clock64() - w0;
          mbar_arrive_expect_tx(&full[s], kHaloBytes);
          tma_load_4d(stages + s * kHaloStride, &p.tmap_x[g], &full[s], kb * 64, tx * kTileW - 1, ty * kTileH - 1, img);
        }
        if (!ok) break;
      }
      if (trc) { trc[0] = t_wait; trc[1] = clock64() - t0; trc[2] = my_tiles; }
    }
  } else if (warp == 1) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint32_t w_addr = smem_u32(wbuf), st_addr = smem_u32(stages);
      bool ok = true;
      long long* trc = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
      const long long t0 = trc ? clock64() : 0;
      long long t_acc = 0, t_full = 0, t_w = 0;
      for (int i = 0; i < my_tiles && ok; ++i) {
        const int a = i % kAcc;
        const long long w0 = trc ? clock64() : 0;
        if (!mbar_wait(&acc_empty[a], ((i / kAcc) & 1) ^ 1u, p.err)) break;
        if (trc) t_acc += clock64() - w0;
        tc_fence_after();
        const uint32_t d_addr = tmem_base + uint32_t(a * 64);
        for (int kb = 0; kb < KBLK && ok; ++kb) {
          const int j = i * KBLK + kb, s = j % S;
          const long long w1 = trc ? clock64() : 0;
          if (!mbar_wait(&full[s], (j / S) & 1, p.err)) { ok = false; break; }
          if (trc) t_full += clock64() - w1;
          tc_fence_after();
          // descriptors of (tap 0, k step 0); every other operand of the block is a compile-time offset away
          const uint64_t ad0 = make_smem_desc_sw128(st_addr + s * kHaloStride, 16, kHaloW * 128);
          const uint64_t bd0 = make_smem_desc_sw128(w_addr + kb * 9 * kWTileBytes, 16, 1024);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            if (i == 0) {
              const long long w2 = trc ? clock64() : 0;
              if (!mbar_wait(&w_full[kb * 9 + tap], 0, p.err)) { ok = false; break; }
              if (trc) t_w += clock64() - w2;
            }
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              // (the 14-bit start-address field cannot carry: every operand lies inside the CTA's < 256 KB of shared memory)
              const uint64_t ad = ad0 + uint64_t(((dy * kHaloW + dx) * 128 + ks * 32) >> 4);
              const uint64_t bd = bd0 + uint64_t((tap * kWTileBytes + ks * 32) >> 4);
              umma_ss_f16(d_addr, ad, bd, idesc, (kb | tap | ks) ? 1u : 0u);
            }
          }
          if (ok) umma_commit(&empty[s]);
        }
        if (ok) umma_commit(&acc_full[a]);
      }
      if (trc) { trc[3] = t_acc; trc[4] = t_full; trc[5] = t_w; trc[6] = clock64() - t0; }
    }
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the output buffer may still be read by the previous layer
    const int ew = warp - 2, eg = ew >> 2;
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;               // accumulator row = pixel (row >> 3, row & 7) of the tile
    uint8_t* block = outbuf + eg * 16384 + q * 4096;            // this warp's 32 pixels x 64 channels
    uint8_t* myrow = outbuf + eg * 16384 + row * 128;           // 16-byte unit j sits at ((j ^ (row & 7)) << 4)
    uint32_t uses = 0;
    long long* trc = (p.trace && blockIdx.x == 0 && ew == 0 && lane == 0) ? p.trace : nullptr;
    const long long t0 = trc ? clock64() : 0;
    long long t_st = 0, t_af = 0, t_rs = 0;
    for (int i = eg; i < my_tiles; i += G) {
      const int a = i % kAcc;
      int t = r + i * R;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int img = t / p.tiles_y;
      const int x0 = tx * kTileW, y0 = ty * kTileH + 4 * q;
      const long long w0 = trc ? clock64() : 0;
      if (lane == 0) {
        // the TMA engine has read this warp's previous store out of the staging block (bulk groups are per thread)
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (p.has_res) {
          mbar_arrive_expect_tx(&res_full[ew], 4096);
          tma_load_4d(block, &p.tmap_r[g], &res_full[ew], n0, x0, y0, img);
        }
      }
      __syncwarp();
      const long long w1 = trc ? clock64() : 0;
      if (!mbar_wait(&acc_full[a], (i / kAcc) & 1, p.err)) break;
      tc_fence_after();
      const long long w2 = trc ? clock64() : 0;
      if (p.has_res && !mbar_wait(&res_full[ew], uses & 1u, p.err)) break;
      if (trc) { t_st += w1 - w0; t_af += w2 - w1; t_rs += clock64() - w2; }
      ++uses;
      const uint32_t ta = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(a * 64);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        float f[32];
        tmem_ld_32x32b_x32(ta + h * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 bv = *reinterpret_cast<const float4*>(s_bias + h * 32 + 4 * e);      // broadcast
          f[4 * e + 0] = __uint_as_float(v[4 * e + 0]) + bv.x;
          f[4 * e + 1] = __uint_as_float(v[4 * e + 1]) + bv.y;
          f[4 * e + 2] = __uint_as_float(v[4 * e + 2]) + bv.z;
          f[4 * e + 3] = __uint_as_float(v[4 * e + 3]) + bv.w;
        }
        if (p.has_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 rv = *reinterpret_cast<const uint4*>(myrow + (((4 * h + j) ^ (row & 7)) << 4));
            const uint32_t w4[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              f[8 * j + 2 * k + 0] += __uint_as_float(w4[k] << 16);
              f[8 * j + 2 * k + 1] += __uint_as_float(w4[k] & 0xFFFF0000u);
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) f[e] = fmaxf(f[e], 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(myrow + (((4 * h + j) ^ (row & 7)) << 4)) =
              make_uint4(pack_bf16x2(f[8 * j + 0], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                         pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
      }
      // the accumulator is drained: hand it back to the MMA issuer (one arrival per warp of the group)
      tc_fence_before();
      fence_proxy_async();                            // generic-proxy writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc_empty[a]);
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&p.tmap_y[g])), "r"(smem_u32(block)), "r"(n0), "r"(x0), "r"(y0), "r"(img)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (trc) { trc[7] = t_st; trc[8] = t_af; trc[9] = t_rs; trc[10] = clock64() - t0; }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // this warp's output stores have completed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kAcc * 64);
}

template <int KBLK>
int launch_wres(const WresParams& prm, int grid, cudaStream_t stream) {
  using Cfg = WresCfg<KBLK>;
  int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&conv_wres_kernel<KBLK>), Cfg::kSmemBytes);
  if (rc != UOC_OK) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, 1, 1);
  cfg.blockDim = dim3(Cfg::kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // the weight fetch overlaps the previous layer's tail
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = 1;
  UOC_CUDA(cudaLaunchKernelEx(&cfg, conv_wres_kernel<KBLK>, prm));
  count_launch();
  return UOC_OK;
}

}  // namespace

bool conv_wres_supported(const ConvProblem& p) {
  const int sms = sm_count();
  return p.ksize == 3 && p.stride == 1 && p.dilation == 1 && (p.Cin == 64 || p.Cin == 128) && p.Cout % 64 == 0 &&
         !p.out_fp32 && p.groups >= 1 && p.groups <= 2 && sms > 0 && p.groups * (p.Cout / 64) <= sms;
}

int launch_conv_wres(const ConvProblem& p, cudaStream_t stream) {
  if (!conv_wres_supported(p)) return fail(UOC_ERR_UNSUPPORTED, "conv_wres: 3x3, stride 1, dilation 1, Cin 64 or 128, Cout % 64 == 0");
  WresParams prm;
  memset(&prm, 0, sizeof(prm));
  const int Ho = p.H, Wo = p.W;
  prm.tiles_x = (Wo + kTileW - 1) / kTileW;
  prm.tiles_y = (Ho + kTileH - 1) / kTileH;
  prm.tiles = p.N * prm.tiles_y * prm.tiles_x;
  prm.n_halves = p.Cout / 64;
  prm.groups = p.groups;
  prm.relu = p.relu;
  prm.has_res = p.g[0].residual != nullptr;
  prm.err = device_error_word();
  if (!prm.err) return fail(UOC_ERR_CUDA, "no device error word");
  const int ncombo = p.groups * prm.n_halves;
  int R = sm_count() / ncombo;
  if (R > prm.tiles) R = prm.tiles;
  if (R < 1) R = 1;
  prm.ctas_per_combo = R;
  for (int g = 0; g < p.groups; ++g) {
    if ((p.g[g].residual != nullptr) != (prm.has_res != 0)) return fail(UOC_ERR_INVALID, "conv_wres: residual on one branch only");
    const uint64_t xd[4] = {uint64_t(p.Cin), uint64_t(p.W), uint64_t(p.H), uint64_t(p.N)};
    const uint64_t xs[3] = {uint64_t(p.Cin) * 2, uint64_t(p.W) * p.Cin * 2, uint64_t(p.H) * p.W * p.Cin * 2};
    const uint32_t xb[4] = {64, uint32_t(kHaloW), uint32_t(kHaloH), 1};
    int rc = make_tmap_bf16(&prm.tmap_x[g], p.g[g].x, 4, xd, xs, xb, nullptr);
    if (rc != UOC_OK) return rc;
    const uint64_t wd[2] = {uint64_t(9) * p.Cin, uint64_t(p.Cout)};
    const uint64_t wsb[1] = {uint64_t(9) * p.Cin * 2};
    const uint32_t wb[2] = {64, 64};
    rc = make_tmap_bf16(&prm.tmap_w[g], p.g[g].w, 2, wd, wsb, wb, nullptr);
    if (rc != UOC_OK) return rc;
    const uint64_t yd[4] = {uint64_t(p.Cout), uint64_t(Wo), uint64_t(Ho), uint64_t(p.N)};
    const uint64_t ys[3] = {uint64_t(p.Cout) * 2, uint64_t(Wo) * p.Cout * 2, uint64_t(Ho) * Wo * p.Cout * 2};
    const uint32_t yb[4] = {64, uint32_t(kTileW), 4, 1};
    rc = make_tmap_bf16(&prm.tmap_y[g], p.g[g].y, 4, yd, ys, yb, nullptr);
    if (rc != UOC_OK) return rc;
    if (prm.has_res) {
      rc = make_tmap_bf16(&prm.tmap_r[g], p.g[g].residual, 4, yd, ys, yb, nullptr);
      if (rc != UOC_OK) return rc;
    }
    prm.bias[g] = p.g[g].bias;
  }
  const int grid = ncombo * R;
  const bool want_trace = knobs().conv_trace != 0;
  if (want_trace) {
    UOC_CUDA(cudaMalloc(&prm.trace, 16 * sizeof(long long)));
    UOC_CUDA(cudaMemsetAsync(prm.trace, 0, 16 * sizeof(long long), stream));
  }
  const int rc = (p.Cin == 64) ? launch_wres<1>(prm, grid, stream) : launch_wres<2>(prm, grid, stream);
  if (want_trace && rc == UOC_OK) {
    long long h[16];
    UOC_CUDA(cudaStreamSynchronize(stream));
    UOC_CUDA(cudaMemcpy(h, prm.trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(prm.trace);
    fprintf(stderr, "[conv_wres trace] Cin %d Cout %d N %d: %d tiles per group, grid %d (%d CTAs per combo) | CTA 0: %lld tiles; producer waits "
            "for a free stage %lld of %lld clk | MMA issuer: waits for an accumulator %lld, for a halo %lld, for weights %lld of %lld clk | "
            "epilogue warp 0: waits for its previous store %lld, for an accumulator %lld, for the residual %lld of %lld clk\n",
            p.Cin, p.Cout, p.N, prm.tiles, grid, R, h[2], h[0], h[1], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10]);
  }
  return rc;
}

}  // namespace uoc
