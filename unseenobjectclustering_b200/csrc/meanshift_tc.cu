// K4: one mean-shift update  Z <- normalize_rows( exp(kappa Z X^T) X )  on the 5th-gen tensor cores.
// (lib/utils/mean_shift.py:79-109, cosine branch; W = exp(kappa Z X^T) is never materialised.)
//
// Structure (FlashAttention-like, the seeds play the role of the queries and X is both K and V):
//   per CTA: a strided set of 128-point tiles of the bf16 pixel-major copy of X ([n][d], 128 B per
//   point at d = 64 -> one TMA box [128 points x 64 ch] per 64 channels, 128B-swizzled).
//     warp 0      TMA producer   : X tiles -> shared-memory ring (mbarrier full/empty)
//     warp 1      MMA issuer     : GEMM1  S[128 seeds x 128 pts]  = Zs (smem, K-major) . Xtile^T (smem, K-major)
//                                   GEMM2  O[128 seeds x d]      += P (TMEM, bf16)     . Xtile   (smem, MN-major)
//     warps 2..9  weight warps   : S (TMEM fp32) -> ex2(kappa*log2e*(s-1)) -> P (bf16x2, TMEM, aliasing S)
//                                   two groups of 4 warps ping-pong on even / odd tiles (one S buffer each) so that
//                                   one group's TMEM round trips hide behind the other's MUFU work;
//                                   then the epilogue O (TMEM) -> per-CTA partial sums in global memory
//   S is double buffered in TMEM (2 x 128 columns), O occupies d columns; 512 columns are allocated.
//   A second tiny kernel (reduce_normalize_kernel) adds the per-CTA partials and normalises the rows.
// The factor exp(-kappa) common to all weights cancels in the row normalisation.
//
// Algorithmic traffic per update: n*d*2 bytes of bf16 X (fp32-equivalent: n*d*4), see DESIGN.md.
#include <cstring>

#include "cluster.cuh"

namespace uoc {

namespace {

constexpr int kTile = 128;          // points per tile
constexpr int kThreads = 320;       // warp 0 TMA, warp 1 MMA, warps 2-5 weight group 0 (even tiles), warps 6-9 group 1 (odd tiles)
constexpr int kBoxBytes = 128 * 128;  // one [128 x 64ch] bf16 box, 16 KiB

template <int D>
struct MsCfg {
  static constexpr int kKBlocks = D / 64;
  static constexpr int kStageBytes = kKBlocks * kBoxBytes;
  static constexpr int kStages = (D == 64) ? 6 : 5;
  static constexpr int kZBytes = kKBlocks * kBoxBytes;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kZBytes + kStages * kStageBytes + 256 /*barriers*/;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColS0 = 0, kColS1 = 128, kColO = 256;
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
meanshift_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ Z,
                    float* __restrict__ partials, int m, long long n, float c1, int P, unsigned int* err) {
  using Cfg = MsCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* zs = smem;
  uint8_t* stages = smem + Cfg::kZBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + Cfg::kStages;
  uint64_t* s_full = bars + 2 * Cfg::kStages;
  uint64_t* p_ready = s_full + 2;
  uint64_t* o_full = p_ready + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, b = blockIdx.y;
  const long long tiles_total = (n + kTile - 1) / kTile;
  const int T = (cta < tiles_total) ? int((tiles_total - cta + P - 1) / P) : 0;

  int j_pre = 0;                                        // tiles whose TMA loads are issued before the CTA-wide sync
  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmap_x);
      for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
      mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
      mbar_init(&p_ready[0], 128); mbar_init(&p_ready[1], 128);
      mbar_init(o_full, 1);
      fence_mbar_init();
      asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: the previous kernel's writes (Z, bf16 field) are visible
      // the ring is empty: start streaming X right away, the seeds are staged meanwhile
      for (int j = 0; j < T && j < Cfg::kStages; ++j) {
        mbar_arrive_expect_tx(&x_full[j], Cfg::kStageBytes);
        const int row0 = int((cta + (long long)j * P) * kTile);
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
          tma_load_3d(stages + j * Cfg::kStageBytes + kb * kBoxBytes, &tmap_x, &x_full[j], kb * 64, row0, b);
      }
    }
    j_pre = T < Cfg::kStages ? T : Cfg::kStages;
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: Z (and the bf16 field) of the previous kernel are complete
  if (warp >= 2) {
    // stage the seeds: fp32 [m][D] -> bf16, K-major 128B-swizzled rows (row r, 16B chunk c -> c ^ (r & 7))
    const int r = (threadIdx.x - 64) & 127;        // 0..127
    const int half = (threadIdx.x - 64) >> 7;      // each weight group stages half of the chunks
    const float* zr = Z + (size_t(b) * m + (r < m ? r : 0)) * D;
#pragma unroll
    for (int c = half * (D / 16); c < (half + 1) * (D / 16); ++c) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < m) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(zr + c * 8));
        const float4 bb = __ldg(reinterpret_cast<const float4*>(zr + c * 8 + 4));
        v.x = pack_bf16x2(a.x, a.y); v.y = pack_bf16x2(a.z, a.w);
        v.z = pack_bf16x2(bb.x, bb.y); v.w = pack_bf16x2(bb.z, bb.w);
      }
      const int kb = c >> 3, cc = c & 7;
      *reinterpret_cast<uint4*>(zs + kb * kBoxBytes + r * 128 + ((cc ^ (r & 7)) << 4)) = v;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    if (elect_one()) {    // one elected lane: lets the compiler keep descriptors / barriers in uniform registers
      for (int j = j_pre; j < T; ++j) {
        const int s = j % Cfg::kStages;
        const uint32_t ph = (j / Cfg::kStages) & 1;
        if (!mbar_wait(&x_empty[s], ph ^ 1u, err)) break;
        mbar_arrive_expect_tx(&x_full[s], Cfg::kStageBytes);
        const int row0 = int((cta + (long long)j * P) * kTile);
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
          tma_load_3d(stages + s * Cfg::kStageBytes + kb * kBoxBytes, &tmap_x, &x_full[s], kb * 64, row0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {    // one elected lane: lets the compiler keep descriptors / barriers in uniform registers
      constexpr uint32_t idesc1 = make_idesc_bf16(128, kTile, 0, 0);
      constexpr uint32_t idesc2 = make_idesc_bf16(128, D, 0, 1);
      const uint32_t zs_addr = smem_u32(zs);
      const uint32_t st_addr = smem_u32(stages);
      bool ok = true;
      auto gemm2 = [&](int i) {
        const int s = i % Cfg::kStages, buf = i & 1;
        if (!mbar_wait(&p_ready[buf], (i >> 1) & 1, err)) { ok = false; return; }
        tc_fence_after();
        const uint32_t xb = st_addr + s * Cfg::kStageBytes;
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks) {
          const uint64_t bd = make_smem_desc_sw128(xb + ks * 2048, kBoxBytes, 1024);
          umma_ts_f16(tmem_base + Cfg::kColO, tmem_base + (buf ? Cfg::kColS1 : Cfg::kColS0) + ks * 8, bd, idesc2,
                      (i > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&x_empty[s]);
      };
      for (int j = 0; j < T && ok; ++j) {
        const int s = j % Cfg::kStages, buf = j & 1;
        if (!mbar_wait(&x_full[s], (j / Cfg::kStages) & 1, err)) { ok = false; break; }
        tc_fence_after();
        const uint32_t xb = st_addr + s * Cfg::kStageBytes;
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_smem_desc_sw128(zs_addr + kb * kBoxBytes + ks * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(xb + kb * kBoxBytes + ks * 32, 16, 1024);
            umma_ss_f16(tmem_base + (buf ? Cfg::kColS1 : Cfg::kColS0), ad, bd, idesc1, (kb | ks) ? 1u : 0u);
          }
        }
        umma_commit(&s_full[buf]);
        if (j >= 1) gemm2(j - 1);
      }
      if (ok && T > 0) gemm2(T - 1);
      if (ok) umma_commit(o_full);
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int grp = (warp - 2) >> 2;        // weight group: 0 -> even tiles / S buffer 0, 1 -> odd tiles / S buffer 1
    const int row = q * 32 + lane;          // seed index
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    bool ok = true;
    for (int j = grp; j < T; j += 2) {
      const int buf = grp;
      if (!mbar_wait(&s_full[buf], (j >> 1) & 1, err)) { ok = false; break; }
      tc_fence_after();
      const uint32_t sa = lane_addr + (buf ? Cfg::kColS1 : Cfg::kColS0);
#pragma unroll
      for (int c = 0; c < kTile / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(sa + c * 32, v);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * e]), c1, -c1));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c1, -c1));
          pk[e] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x16(sa + c * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_ready[buf]);
    }
    if (ok && mbar_wait(o_full, 0, err)) {
      tc_fence_after();
      float* dst = partials + ((size_t(b) * P + cta) * 128 + row) * D;
#pragma unroll
      for (int c = grp * (D / 64); c < (grp + 1) * (D / 64); ++c) {   // each group drains half of the columns
        uint32_t v[32];
        tmem_ld_32x32b_x32(lane_addr + Cfg::kColO + c * 32, v);
        tmem_wait_ld();
        if (row < m) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            *reinterpret_cast<uint4*>(dst + c * 32 + e * 4) = make_uint4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int D>
int launch_iter(const CUtensorMap& tmap, const ClusterShape& s, const ClusterWorkspace& w, const float* Z, int P,
                float kappa, cudaStream_t stream) {
  using Cfg = MsCfg<D>;
  static bool attr = false;
  if (!attr) {
    UOC_CUDA(cudaFuncSetAttribute(meanshift_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr = true;
  }
  const float c1 = kappa * 1.4426950408889634f;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(P, s.batch);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous kernel's tail
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = 1;
  const float* Zc = Z;
  float* partials = w.partials;
  int m = s.m;
  long long n = s.n;
  unsigned int* err = device_error_word();
  UOC_CUDA(cudaLaunchKernelEx(&cfg, meanshift_tc_kernel<D>, tmap, Zc, partials, m, n, c1, P, err));
  count_launch();
  return UOC_OK;
}

}  // namespace

int launch_hill_climb_tc(const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w, float* Z,
                         float kappa, int iters, cudaStream_t stream) {
  if (s.d != 64 && s.d != 128) return fail(UOC_ERR_UNSUPPORTED, "tcgen05 mean-shift loop supports d = 64 or 128");
  if (s.m > 128) return fail(UOC_ERR_UNSUPPORTED, "num_seeds > 128 is not supported");
  if (reinterpret_cast<uintptr_t>(xb) % 16 != 0) return fail(UOC_ERR_INVALID, "bf16 field must be 16-byte aligned");
  if (!device_error_word()) return fail(UOC_ERR_CUDA, "no device error word");
  CUtensorMap tmap;
  const uint64_t dims[3] = {uint64_t(s.d), uint64_t(s.n), uint64_t(s.batch)};
  const uint64_t strides[2] = {uint64_t(s.d) * 2, uint64_t(s.n) * s.d * 2};
  const uint32_t box[3] = {64, uint32_t(kTile), 1};
  int rc = make_tmap_bf16(&tmap, xb, 3, dims, strides, box, nullptr);
  if (rc != UOC_OK) return rc;
  const long long tiles = (s.n + kTile - 1) / kTile;
  int P = w.max_partials;
  if (P > tiles) P = int(tiles);
  for (int it = 0; it < iters; ++it) {
    rc = (s.d == 64) ? launch_iter<64>(tmap, s, w, Z, P, kappa, stream) : launch_iter<128>(tmap, s, w, Z, P, kappa, stream);
    if (rc != UOC_OK) return rc;
    rc = launch_reduce_normalize(w.partials, s.batch, P, s.m, s.d, 128, Z, stream);
    if (rc != UOC_OK) return rc;
  }
  return UOC_OK;
}

}  // namespace uoc
