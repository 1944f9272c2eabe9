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
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdio>

#include "cluster.cuh"

namespace uoc {

namespace {

constexpr int kTile = 128;          // points per tile
constexpr int kThreads = 320;       // warp 0 TMA, warp 1 MMA, warps 2-5 weight group 0 (even tiles), warps 6-9 group 1 (odd tiles)
constexpr int kThreadsP = 352;      // persistent kernel: + warp 10, the GEMM2 issuer
constexpr int kIssue2Warp = 10;
constexpr int kIssueBar = 320;      // named barrier 1: the two issuing warps + the eight weight warps
constexpr int kBoxBytes = 128 * 128;  // one [128 x 64ch] bf16 box, 16 KiB
constexpr int kDefaultPoly = 0;       // weights per 32-column chunk evaluated on the FMA pipe (UOC_LOOP_POLY overrides: 0 / 8 / 12)

template <int D>
struct MsCfg {
  static constexpr int kKBlocks = D / 64;
  static constexpr int kStageBytes = kKBlocks * kBoxBytes;
  static constexpr int kStages = (D == 64) ? 12 : 5;   // X tiles in flight: several fields per launch stream from HBM (latency x bandwidth)
  static constexpr int kZBytes = kKBlocks * kBoxBytes;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kZBytes + kStages * kStageBytes + 512 /*barriers*/;
  static_assert((2 * kStages + 3 + 3 + 1 + 3) * 8 + 8 <= 512, "barrier region");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t kColS3 = 128, kColO3 = 384;   // three S/P buffers of 128 columns (tile j -> j % 3), then O
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
  uint64_t* p_ready = s_full + 3;
  uint64_t* o_full = p_ready + 3;
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
      for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); }
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
      // three S/P buffers (tile j -> buffer j % 3): GEMM1 runs two tiles ahead of the weights, GEMM2(j-2) is issued
      // after GEMM1(j), so a weight group finds its next S tile ready when it has published P
      auto gemm2 = [&](int i) {
        const int s = i % Cfg::kStages, buf = i % 3;
        if (!mbar_wait(&p_ready[buf], (i / 3) & 1, err)) { ok = false; return; }
        tc_fence_after();
        const uint32_t xb = st_addr + s * Cfg::kStageBytes;
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks) {
          const uint64_t bd = make_smem_desc_sw128(xb + ks * 2048, kBoxBytes, 1024);
          umma_ts_f16(tmem_base + Cfg::kColO3, tmem_base + buf * Cfg::kColS3 + ks * 8, bd, idesc2,
                      (i > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&x_empty[s]);
      };
      for (int j = 0; j < T && ok; ++j) {
        const int s = j % Cfg::kStages, buf = j % 3;
        if (!mbar_wait(&x_full[s], (j / Cfg::kStages) & 1, err)) { ok = false; break; }
        tc_fence_after();
        const uint32_t xb = st_addr + s * Cfg::kStageBytes;
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_smem_desc_sw128(zs_addr + kb * kBoxBytes + ks * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(xb + kb * kBoxBytes + ks * 32, 16, 1024);
            umma_ss_f16(tmem_base + buf * Cfg::kColS3, ad, bd, idesc1, (kb | ks) ? 1u : 0u);
          }
        }
        umma_commit(&s_full[buf]);
        if (j >= 2) gemm2(j - 2);
      }
      if (ok && T > 1) gemm2(T - 2);
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
      const int buf = j % 3;
      if (!mbar_wait(&s_full[buf], (j / 3) & 1, err)) { ok = false; break; }
      tc_fence_after();
      const uint32_t sa = lane_addr + buf * Cfg::kColS3;
      // software pipeline over the four 32-column chunks: the TMEM load of chunk c+1 is in flight during the MUFU work of chunk c
      uint32_t va[32], vb[32];
      // exp(kappa (s - 1)) * (1 + 2^-9): the factor turns the truncating bf16 pack below into round-to-nearest (within
      // one ulp) and cancels in the row normalisation
      const float c0 = 0.0028150156f - c1;
      auto weights = [&](const uint32_t (&cur)[32], int c) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(cur[2 * e]), c1, c0));
          const float p1 = ex2_approx(fmaf(__uint_as_float(cur[2 * e + 1]), c1, c0));
          pk[e] = pack_bf16x2_trunc(p0, p1);
        }
        tmem_st_32x32b_x16(sa + c * 16, pk);
      };
      static_assert(kTile / 32 == 4, "four chunks per tile");
      tmem_ld_32x32b_x32(sa, va);
      tmem_wait_ld();
      tmem_ld_32x32b_x32(sa + 32, vb);
      weights(va, 0);
      tmem_wait_ld();
      tmem_ld_32x32b_x32(sa + 64, va);
      weights(vb, 1);
      tmem_wait_ld();
      tmem_ld_32x32b_x32(sa + 96, vb);
      weights(va, 2);
      tmem_wait_ld();
      weights(vb, 3);
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
        tmem_ld_32x32b_x32(lane_addr + Cfg::kColO3 + c * 32, v);
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
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&meanshift_tc_kernel<D>), int(Cfg::kSmemBytes));
    if (rc_attr != UOC_OK) return rc_attr;
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


// ----------------------------------------------------------------------------------------------
// Persistent variant: ALL updates of the hill climbing in one cooperative launch.
//   Per update the CTAs exchange through two monotonic counters in global memory instead of kernel boundaries
//   (one polling thread per CTA; per-CTA flag words polled by every thread turned a few L2 lines into a hot spot):
//     parts_done[b] += 1  once a CTA has published its partial sums of update u      (reducers wait for (u+1) * P)
//     rows_done[b]  += 1  once seed row r of Z_{u+1} has been reduced + normalised (by CTA r % P) and stored
//                         (everybody waits for (u+1) * m before staging Z_{u+1})
//   Z is updated in place: row r of Z_{u+2} cannot be produced before every CTA has staged Z_{u+1} (its partials of
//   update u+1 are an input), and the partials of update u+1 are written only after every reducer has finished with
//   those of update u (all rows of Z_{u+1} are an input of the staging).  The TMA producer is free running over
//   iters x T tiles (X does not change), so the next update's first tiles arrive during the exchange.
//   Warps 2..9 double as the 8 reduce warps; warp 1 joins them on named barrier 1 once per update ("Z staged, O drained").
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool poll_flag(const unsigned int* p, unsigned int target, unsigned int* err) {
  for (unsigned int it = 0; it < (1u << 24); ++it)
    if (ld_acquire_u32(p) >= target) return true;
  atomicOr(err, ERR_GRID_BARRIER_TIMEOUT);
  return false;
}

// POLY: of the 32 weights a thread computes per 32-column chunk, POLY are evaluated with ex2_poly on the FMA pipe and
// the rest with MUFU.EX2, interleaved (the two pipes run concurrently; UOC_LOOP_POLY selects the split).
template <int D, int POLY>
__global__ void __launch_bounds__(kThreadsP, 1)
meanshift_tc_persistent_kernel(const __grid_constant__ CUtensorMap tmap_x, float* Z, float* partials,
                               unsigned int* done, unsigned int* rowflag, int m, long long n, float c1, int P,
                               int iters, unsigned int* err, long long* trace) {
  using Cfg = MsCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_part[8][D];
  __shared__ float s_sq[8];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* zs = smem;
  uint8_t* stages = smem + Cfg::kZBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + Cfg::kStages;
  uint64_t* s_full = bars + 2 * Cfg::kStages;
  uint64_t* p_ready = s_full + 3;
  uint64_t* o_full = p_ready + 3;
  uint64_t* s_free = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, b = blockIdx.y;
  const long long tiles_total = (n + kTile - 1) / kTile;
  const int T = (cta < tiles_total) ? int((tiles_total - cta + P - 1) / P) : 0;
  const long long TT = (long long)T * iters;           // tiles this CTA streams over the whole hill climbing

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&s_free[i], 1); }
    mbar_init(o_full, 1);
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

  if (warp == 0) {
    if (elect_one()) {
      for (long long jj = 0; jj < TT; ++jj) {
        const int s = int(jj % Cfg::kStages);
        const uint32_t ph = uint32_t(jj / Cfg::kStages) & 1u;
        if (jj >= Cfg::kStages && !mbar_wait(&x_empty[s], ph ^ 1u, err)) break;
        mbar_arrive_expect_tx(&x_full[s], Cfg::kStageBytes);
        const int j = int(jj % T);
        const int row0 = int((cta + (long long)j * P) * kTile);
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
          tma_load_3d(stages + s * Cfg::kStageBytes + kb * kBoxBytes, &tmap_x, &x_full[s], kb * 64, row0, b);
      }
    }
  } else if (warp == 1) {
    // ---- GEMM1 issuer: S(tile) = Zs . X^T as soon as the X tile has landed and the S/P buffer is free again.
    // GEMM1 and GEMM2 have their OWN issuing threads (warps 1 and 10): with a single in-order issuer every wait for a
    // P tile also held back the next S tile, and tcgen05.mma issue blocks while the queue is full (measured: the
    // single issuer spent 10.8k of 23k clk per update inside issue and 9.3k waiting for P).
    constexpr uint32_t idesc1 = make_idesc_bf16(128, kTile, 0, 0);
    const uint32_t zs_addr = smem_u32(zs);
    const uint32_t st_addr = smem_u32(stages);
    bool ok = true;
    long long* tr2 = (trace && b == 0 && cta == 0) ? trace + 16 * iters : nullptr;   // debug: the issuing thread's waits
    for (int u = 0; u < iters; ++u) {
      named_bar_sync(1, kIssueBar);        // seeds of this update staged, O of the previous one drained
      tc_fence_after();
      if (elect_one()) {
        const long long base = (long long)u * T;
        long long w_x = 0, w_f = 0;
        const long long t_begin = tr2 ? clock64() : 0;
        for (int j = 0; j < T && ok; ++j) {
          const long long jj = base + j;
          const int s = int(jj % Cfg::kStages), buf = int(jj % 3);
          long long tw = tr2 ? clock64() : 0;
          if (!mbar_wait(&x_full[s], uint32_t(jj / Cfg::kStages) & 1u, err)) { ok = false; break; }
          if (tr2) { const long long t = clock64(); w_x += t - tw; tw = t; }
          // the S/P buffer of tile jj was last used by tile jj - 3: its GEMM2 must have consumed P
          if (jj >= 3 && !mbar_wait(&s_free[buf], (uint32_t(jj / 3) + 1u) & 1u, err)) { ok = false; break; }
          if (tr2) w_f += clock64() - tw;
          tc_fence_after();
          const uint32_t xb = st_addr + s * Cfg::kStageBytes;
#pragma unroll
          for (int kb = 0; kb < Cfg::kKBlocks; ++kb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = make_smem_desc_sw128(zs_addr + kb * kBoxBytes + ks * 32, 16, 1024);
              const uint64_t bd = make_smem_desc_sw128(xb + kb * kBoxBytes + ks * 32, 16, 1024);
              umma_ss_f16(tmem_base + buf * Cfg::kColS3, ad, bd, idesc1, (kb | ks) ? 1u : 0u);
            }
          }
          umma_commit(&s_full[buf]);
        }
        if (tr2) { tr2[u * 4 + 0] = w_x; tr2[u * 4 + 1] = w_f; tr2[u * 4 + 2] = clock64() - t_begin; tr2[u * 4 + 3] = T; }
      }
      __syncwarp();
    }
  } else if (warp == kIssue2Warp) {
    // ---- GEMM2 issuer: O += P(tile) . X(tile) as soon as a weight group has published P
    constexpr uint32_t idesc2 = make_idesc_bf16(128, D, 0, 1);
    const uint32_t st_addr = smem_u32(stages);
    bool ok = true;
    for (int u = 0; u < iters; ++u) {
      named_bar_sync(1, kIssueBar);
      tc_fence_after();
      if (elect_one()) {
        const long long base = (long long)u * T;
        for (int j = 0; j < T && ok; ++j) {
          const long long jj = base + j;
          const int s = int(jj % Cfg::kStages), buf = int(jj % 3);
          if (!mbar_wait(&p_ready[buf], uint32_t(jj / 3) & 1u, err)) { ok = false; break; }
          tc_fence_after();
          const uint32_t xb = st_addr + s * Cfg::kStageBytes;
#pragma unroll
          for (int ks = 0; ks < kTile / 16; ++ks) {
            const uint64_t bd = make_smem_desc_sw128(xb + ks * 2048, kBoxBytes, 1024);
            umma_ts_f16(tmem_base + Cfg::kColO3, tmem_base + buf * Cfg::kColS3 + ks * 8, bd, idesc2,
                        (j > 0 || ks > 0) ? 1u : 0u);
          }
          umma_commit(&x_empty[s]);      // GEMM1 of this tile completed before its P existed: the X stage is free
          umma_commit(&s_free[buf]);     // ... and so is the S/P buffer
        }
        if (ok) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int grp = (warp - 2) >> 2;        // weight group
    const int row = q * 32 + lane;          // seed index
    const int w8 = warp - 2;                // reduce warp 0..7
    const int t256 = threadIdx.x - 64;      // 0..255
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    unsigned int* parts_done = done + size_t(b) * 64;       // counters on separate 128-byte lines per batch item
    unsigned int* rows_done = parts_done + 32;
    bool ok = true;
    // debug trace (UOC_LOOP_TRACE): CTA 0 (a reducer) and CTA P-1 stamp 6 phases per update
    long long* tr = (trace && t256 == 0 && b == 0 && (cta == 0 || cta == P - 1)) ? trace + (cta == 0 ? 0 : iters * 8) : nullptr;
    auto stamp = [&](int u, int k) {
      if (tr) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); tr[u * 8 + k] = t; }
    };
    for (int u = 0; u < iters; ++u) {
      stamp(u, 0);
      // ---- stage the seeds Z_u: fp32 [m][D] -> bf16, K-major 128B-swizzled rows
      {
        const int r = t256 & 127;
        const int half = t256 >> 7;
        if (u > 0) {      // all m rows of Z_u are stored: one polling thread per CTA, the named barrier releases the rest
          if (t256 == 0) ok = poll_flag(rows_done, (unsigned int)(u * m), err) && ok;
          named_bar_sync(2, 256);
        }
        const float* zr = Z + (size_t(b) * m + (r < m ? r : 0)) * D;
#pragma unroll
        for (int c = half * (D / 16); c < (half + 1) * (D / 16); ++c) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (r < m) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(zr + c * 8));
            const float4 bb = __ldcg(reinterpret_cast<const float4*>(zr + c * 8 + 4));
            v.x = pack_bf16x2(a.x, a.y); v.y = pack_bf16x2(a.z, a.w);
            v.z = pack_bf16x2(bb.x, bb.y); v.w = pack_bf16x2(bb.z, bb.w);
          }
          const int kb = c >> 3, cc = c & 7;
          *reinterpret_cast<uint4*>(zs + kb * kBoxBytes + r * 128 + ((cc ^ (r & 7)) << 4)) = v;
        }
        fence_proxy_async();
      }
      tc_fence_before();
      named_bar_sync(1, kIssueBar);
      tc_fence_after();
      stamp(u, 1);
      // ---- weights: this group's tiles are those with global parity == grp
      const long long base = (long long)u * T;
      const float c0 = 0.0028150156f - c1;   // * (1 + 2^-9): makes the truncating bf16 pack round to nearest, cancels in the normalisation
      for (int j = int((grp - base) & 1); j < T; j += 2) {
        const long long jj = base + j;
        const int buf = int(jj % 3);
        const long long tw0 = tr ? clock64() : 0;
        if (!mbar_wait(&s_full[buf], uint32_t(jj / 3) & 1u, err)) { ok = false; break; }
        tc_fence_after();
        const long long tw1 = tr ? clock64() : 0;
        const uint32_t sa = lane_addr + buf * Cfg::kColS3;
        uint32_t va[32], vb[32];
        auto weights = [&](const uint32_t (&cur)[32], int c) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float x0 = fmaf(__uint_as_float(cur[2 * e]), c1, c0);
            const float x1 = fmaf(__uint_as_float(cur[2 * e + 1]), c1, c0);
            // spread the polynomial evaluations evenly over the chunk so that both pipes stay busy
            const bool poly0 = ((2 * e) * POLY) / 32 != ((2 * e + 1) * POLY) / 32;
            const bool poly1 = ((2 * e + 1) * POLY) / 32 != ((2 * e + 2) * POLY) / 32;
            const float p0 = poly0 ? ex2_poly(x0) : ex2_approx(x0);
            const float p1 = poly1 ? ex2_poly(x1) : ex2_approx(x1);
            pk[e] = pack_bf16x2_trunc(p0, p1);
          }
          tmem_st_32x32b_x16(sa + c * 16, pk);
        };
        tmem_ld_32x32b_x32(sa, va);
        tmem_wait_ld();
        tmem_ld_32x32b_x32(sa + 32, vb);
        weights(va, 0);
        tmem_wait_ld();
        tmem_ld_32x32b_x32(sa + 64, va);
        weights(vb, 1);
        tmem_wait_ld();
        tmem_ld_32x32b_x32(sa + 96, vb);
        weights(va, 2);
        tmem_wait_ld();
        weights(vb, 3);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&p_ready[buf]);
        if (tr) { tr[u * 8 + 6] += tw1 - tw0; tr[u * 8 + 7] += clock64() - tw1; }
      }
      // ---- epilogue: O -> this CTA's partial sums
      stamp(u, 2);
      if (ok && T > 0 && mbar_wait(o_full, uint32_t(u) & 1u, err)) {
        tc_fence_after();
        stamp(u, 3);
        float* dst = partials + ((size_t(b) * P + cta) * 128 + row) * D;
#pragma unroll
        for (int c = grp * (D / 64); c < (grp + 1) * (D / 64); ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(lane_addr + Cfg::kColO3 + c * 32, v);
          tmem_wait_ld();
          if (row < m) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              __stcg(reinterpret_cast<uint4*>(dst + c * 32 + e * 4), make_uint4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]));
          }
        }
      }
      named_bar_sync(2, 256);
      if (t256 == 0) { __threadfence(); atomicAdd(parts_done, 1u); }
      stamp(u, 4);
      // ---- reduce + normalise the rows this CTA owns
      if (cta < m) {
        if (t256 == 0) ok = poll_flag(parts_done, (unsigned int)((u + 1) * P), err) && ok;
        named_bar_sync(2, 256);
      }
      for (int r = cta; r < m; r += P) {
        float acc[D / 32];
#pragma unroll
        for (int qq = 0; qq < D / 32; ++qq) acc[qq] = 0.f;
        // batches of kRedBatch independent loads per thread: the whole row costs ~2 L2 round trips instead of one per
        // 4 partials (fixed summation order -> deterministic)
        constexpr int kRedBatch = (D == 64) ? 10 : 5;
        for (int part0 = w8; part0 < P; part0 += 8 * kRedBatch) {
          float v[kRedBatch][D / 32];
#pragma unroll
          for (int i = 0; i < kRedBatch; ++i) {
            const int part = part0 + 8 * i;
            const float* src = partials + ((size_t(b) * P + (part < P ? part : w8)) * 128 + r) * D;
#pragma unroll
            for (int qq = 0; qq < D / 32; ++qq) v[i][qq] = (part < P) ? __ldcg(src + lane + 32 * qq) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < kRedBatch; ++i)
#pragma unroll
            for (int qq = 0; qq < D / 32; ++qq) acc[qq] += v[i][qq];
        }
#pragma unroll
        for (int qq = 0; qq < D / 32; ++qq) s_part[w8][lane + 32 * qq] = acc[qq];
        named_bar_sync(2, 256);
        float tot = 0.f;
        if (t256 < D) {
#pragma unroll
          for (int w = 0; w < 8; ++w) tot += s_part[w][t256];
        }
        float sq = tot * tot;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0) s_sq[w8] = sq;
        named_bar_sync(2, 256);
        float all = 0.f;
#pragma unroll
        for (int w = 0; w < D / 32; ++w) all += s_sq[w];
        const float denom = fmaxf(sqrtf(all), 1e-12f);  // F.normalize eps (lib/utils/mean_shift.py:107)
        if (t256 < D) __stcg(Z + (size_t(b) * m + r) * D + t256, tot / denom);
        named_bar_sync(2, 256);
        if (t256 == 0) { __threadfence(); atomicAdd(rows_done, 1u); }
      }
      stamp(u, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int D, int POLY>
int launch_persistent(const CUtensorMap& tmap, const ClusterShape& s, const ClusterWorkspace& w, float* Z, int P,
                      float kappa, int iters, cudaStream_t stream) {
  using Cfg = MsCfg<D>;
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&meanshift_tc_persistent_kernel<D, POLY>), int(Cfg::kSmemBytes));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  unsigned int* done = reinterpret_cast<unsigned int*>(w.slots);   // two counters per batch item, 128 bytes apart
  unsigned int* rowflag = nullptr;
  UOC_CUDA(cudaMemsetAsync(done, 0, 256 * size_t(s.batch), stream));
  float c1 = kappa * 1.4426950408889634f;
  CUtensorMap tm = tmap;
  float* partials = w.partials;
  int m = s.m;
  long long n = s.n;
  unsigned int* err = device_error_word();
  long long* trace = nullptr;
  const bool want_trace = knobs().loop_trace != 0;                 // debug: per-phase timeline on stderr (synchronises)
  if (want_trace) {
    UOC_CUDA(cudaMalloc(&trace, sizeof(long long) * 20 * iters));
    UOC_CUDA(cudaMemsetAsync(trace, 0, sizeof(long long) * 20 * iters, stream));
  }
  void* args[] = {&tm, &Z, &partials, &done, &rowflag, &m, &n, &c1, &P, &iters, &err, &trace};
  UOC_CUDA(cudaLaunchCooperativeKernel(meanshift_tc_persistent_kernel<D, POLY>, dim3(P, s.batch), dim3(kThreadsP), args,
                                       Cfg::kSmemBytes, stream));
  count_launch();
  if (want_trace) {
    std::vector<long long> h(size_t(20) * iters);
    UOC_CUDA(cudaStreamSynchronize(stream));
    UOC_CUDA(cudaMemcpy(h.data(), trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    const long long t0 = h[0];
    for (int c = 0; c < 2; ++c)
      for (int u = 0; u < iters; ++u) {
        const long long* q = h.data() + (size_t(c) * iters + u) * 8;
        fprintf(stderr, "[loop trace] cta %s update %d: start %lld staged %lld weights_done %lld o_full %lld published %lld reduced %lld (ns); group-0 warp: s_full wait %lld clk, weights %lld clk\n",
                c == 0 ? "0" : "P-1", u, q[0] - t0, q[1] - t0, q[2] - t0, q[3] - t0, q[4] - t0, q[5] - t0, q[6], q[7]);
      }
    for (int u = 0; u < iters; ++u) {
      const long long* q = h.data() + size_t(16) * iters + size_t(u) * 4;
      fprintf(stderr, "[loop trace] cta 0 update %d, GEMM1 issuing thread: %lld tiles, waited %lld clk for X tiles (TMA), %lld clk for free S/P buffers, %lld clk in total\n",
              u, q[3], q[0], q[1], q[2]);
    }
  }
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
  // persistent single-launch variant: needs every CTA resident (cooperative launch) and cannot be captured into a CUDA
  // graph; an ongoing stream capture, or more fields than SMs, select the launch-per-update form
  {
    bool persistent = true;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (persistent && cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) persistent = false;
    const int sms = sm_count();
    int Pp = P;
    if (sms > 0 && (long long)Pp * s.batch > sms) Pp = sms / s.batch;
    if (Pp < 1 || iters < 1 || w.slot_bytes < 256 * size_t(s.batch)) persistent = false;
    if (persistent) {
      // (a points x seeds formulation with two seed groups out of phase hides the exchange but is shared-memory bound:
      // profiles/r02_loop_two_groups.md)
      // (POLY > 0 -- part of the exponentials on the FMA pipe -- was measured at -3 %: not instantiated)
      if (s.d == 64) return launch_persistent<64, 0>(tmap, s, w, Z, Pp, kappa, iters, stream);
      return launch_persistent<128, 0>(tmap, s, w, Z, Pp, kappa, iters, stream);
    }
  }
  for (int it = 0; it < iters; ++it) {
    rc = (s.d == 64) ? launch_iter<64>(tmap, s, w, Z, P, kappa, stream) : launch_iter<128>(tmap, s, w, Z, P, kappa, stream);
    if (rc != UOC_OK) return rc;
    rc = launch_reduce_normalize(w.partials, s.batch, P, s.m, s.d, 128, Z, stream);
    if (rc != UOC_OK) return rc;
  }
  return UOC_OK;
}

}  // namespace uoc
