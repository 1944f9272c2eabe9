// K4 (second generation, d = 64): ALL mean-shift updates  Z <- normalize_rows( exp(kappa Z X^T) X )  in one cooperative launch,
// with the SEEDS SPLIT INTO TWO GROUPS that run out of phase.  (lib/utils/mean_shift.py:79-109, cosine branch.)
//
// Why (profiles/r02_ncu_meanshift_tc_persistent_kernel.txt): the first generation (meanshift_tc.cu) keeps the seeds in the
// TMEM lanes (S = Z X^T, 128 lanes for 100 seeds) and every update ends with a ~7 us grid-wide exchange (publish partial
// sums -> reduce + normalise -> re-stage Z) during which the MUFU pipe -- the bound of the weights phase -- idles: 19.4 us
// per update of which 11.5 us are weights.  A seed's next position depends on ITS OWN row only, so the seeds can be cut
// into groups whose updates are independent: while group A exchanges, group B computes.  That needs the seeds in the
// COLUMNS of S (a warp instruction always covers 32 lanes, so lanes cannot be time-shared between groups):
//     GEMM1   S_g[128 points x N_g]   = Xtile (smem, K-major)        . Z_g^T (smem, K-major)           N_g = 48 / 64
//     weights P_g = ex2(kappa log2e (S_g - 1)) -> bf16 -> SHARED MEMORY [128 points][64 seeds] (128B-swizzled rows)
//     GEMM2   O_g[64 seeds x 64 ch]  += P_g^T (smem, MN-major A)     . Xtile (smem, MN-major B)        M = 64
// Only the real seed columns are exponentiated (104 instead of 128 per point for m = 100).
//   warp 0      TMA producer: the CTA's point tiles as a cyclic stream (any T consecutive positions cover every tile once,
//               so a group may begin its sweep at any position)
//   warp 1      GEMM1 issuer: at every stream position, one S tile per ACTIVE group (seeds staged, tiles left)
//   warp 2      GEMM2 issuer
//   warps 3-10  weight warps, two sets of four taking alternate ring entries
//   warps 11-14 exchange warps: drain O_g, publish, reduce + normalise the rows this CTA owns, re-stage Z_g
// The exchange protocol (two monotonic counters per group and field) and its safety argument are those of the first
// generation, per group.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cluster.cuh"

namespace uoc {

namespace {

constexpr int kTile = 128;
constexpr int kD = 64;
constexpr int kBox = kTile * 128;     // one [128 points x 64 ch] bf16 tile, 16 KiB
constexpr int kXStages = 5;
constexpr int kRing = 4;              // S (TMEM, 64 columns) / P (smem, 16 KiB) ring entries
constexpr int kThreads2 = 15 * 32;
constexpr uint32_t kColO = kRing * 64;    // O_A at column 256, O_B at 320

struct Ms2Params {
  CUtensorMap tmap_x;
  float* Z;
  float* partials;
  unsigned int* done;     // per field: parts_done[g] at +16 g, rows_done[g] at +32 + 16 g (uint32 index), 256 bytes per field
  int m, mA;              // seeds, seeds of group A (group B: m - mA)
  long long n;
  float c1;
  int P, iters;
  unsigned int* err;
};

__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool poll_counter(const unsigned int* p, unsigned int target, unsigned int* err) {
  for (unsigned int it = 0; it < (1u << 24); ++it)
    if (ld_acquire(p) >= target) return true;
  atomicOr(err, ERR_GRID_BARRIER_TIMEOUT);
  return false;
}
// wait for a barrier phase, or give up when *stop (>= 0 once the GEMM1 issuer has issued everything) says that the awaited
// item will never come
__device__ __forceinline__ int wait_or_stop(uint64_t* bar, uint32_t parity, const volatile int* stop, int item, unsigned int* err) {
  for (uint32_t it = 0; it < (1u << 26); ++it) {
    if (mbar_try_wait(bar, parity)) return 1;
    const int s = *stop;
    if (s >= 0 && item >= s) return 0;
    if (it > 64) __nanosleep(20);
    if ((it & 0xFFFu) == 0xFFFu && *reinterpret_cast<volatile unsigned int*>(err) != 0u) return -1;
  }
  atomicOr(err, ERR_MBAR_TIMEOUT);
  return -1;
}

__global__ void __launch_bounds__(kThreads2, 1) meanshift_tc2_kernel(const __grid_constant__ Ms2Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_part[4][kD];
  __shared__ float s_sq[2];
  __shared__ int s_total_seq, s_total_pos;
  __shared__ uint32_t s_meta[kRing];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* zs = smem;                                  // 2 x [64 seed rows x 128 B]
  uint8_t* xst = zs + 2 * 8192;                        // kXStages x 16 KiB
  uint8_t* pbuf = xst + kXStages * kBox;               // kRing x 16 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(pbuf + kRing * kBox);
  uint64_t* x_full = bars;
  uint64_t* x_empty = x_full + kXStages;
  uint64_t* s_full = x_empty + kXStages;
  uint64_t* p_ready = s_full + kRing;
  uint64_t* s_free = p_ready + kRing;
  uint64_t* o_full = s_free + kRing;                   // 2
  uint64_t* z_ready = o_full + 2;                      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_ready + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, b = blockIdx.y;
  const int P = p.P, iters = p.iters;
  const long long tiles_total = (p.n + kTile - 1) / kTile;
  const int T = int((tiles_total - cta + P - 1) / P);            // >= 1: the host launches P <= tiles
  const int mg[2] = {p.mA, p.m - p.mA};
  const int ngroups = mg[1] > 0 ? 2 : 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&p.tmap_x);
    for (int s = 0; s < kXStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int k = 0; k < kRing; ++k) { mbar_init(&s_full[k], 1); mbar_init(&p_ready[k], 128); mbar_init(&s_free[k], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&o_full[g], 1); mbar_init(&z_ready[g], 128); }
    s_total_seq = -1;
    s_total_pos = -1;
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---- producer: stream position q carries tile q % T
    if (elect_one()) {
      int issued = 0;
      for (;; ++issued) {
        const int s = issued % kXStages;
        if (issued >= kXStages) {
          const int w = wait_or_stop(&x_empty[s], ((issued / kXStages) & 1) ^ 1u, &s_total_pos, issued, p.err);
          if (w <= 0) break;
        } else {
          const int tp = s_total_pos;
          if (tp >= 0 && issued >= tp) break;
        }
        mbar_arrive_expect_tx(&x_full[s], kBox);
        const int row0 = int((cta + (long long)(issued % T) * P) * kTile);
        tma_load_3d(xst + s * kBox, &p.tmap_x, &x_full[s], 0, row0, b);
      }
      // tiles fetched beyond the last consumed position must have landed before the CTA may exit
      int consumed = s_total_pos;
      if (consumed >= 0)
        for (int q = consumed; q < issued; ++q)
          if (!mbar_wait(&x_full[q % kXStages], (q / kXStages) & 1, p.err)) break;
    }
  } else if (warp == 1) {
    // ---- GEMM1 issuer
    if (elect_one()) {
      const uint32_t zs_addr = smem_u32(zs), x_addr = smem_u32(xst);
      uint32_t idesc1[2];
      for (int g = 0; g < 2; ++g) idesc1[g] = make_idesc_bf16(128, ((mg[g] + 15) / 16) * 16, 0, 0);
      int rem[2] = {T, T}, upd[2] = {0, ngroups == 2 ? 0 : iters};
      const int hold = ngroups == 2 ? T / 2 : 0;       // group B starts half a sweep late: the groups stay out of phase
      int pos = 0, seq = 0;
      bool ok = true;
      unsigned int spins = 0;
      while (ok && (upd[0] < iters || upd[1] < iters)) {
        bool act[2];
        for (int g = 0; g < 2; ++g) {
          act[g] = upd[g] < iters && (g == 0 || pos >= hold || upd[0] >= iters);
          if (act[g] && rem[g] == T) act[g] = mbar_try_wait(&z_ready[g], uint32_t(upd[g]) & 1u);   // seeds of this update staged
        }
        if (!act[0] && !act[1]) {
          if (((++spins) & 0xFFFFu) == 0 && *reinterpret_cast<volatile unsigned int*>(p.err) != 0u) break;
          if (spins > (1u << 28)) { atomicOr(p.err, ERR_MBAR_TIMEOUT); break; }
          continue;
        }
        spins = 0;
        const int s = pos % kXStages;
        if (!mbar_wait(&x_full[s], (pos / kXStages) & 1, p.err)) break;
        tc_fence_after();
        for (int g = 0; g < 2 && ok; ++g) {
          if (!act[g]) continue;
          const int k = seq % kRing;
          if (!mbar_wait(&s_free[k], ((seq / kRing) & 1) ^ 1u, p.err)) { ok = false; break; }
          tc_fence_after();
          const bool last = rem[g] == 1;
          const bool release = !(g == 0 && act[1]);
          s_meta[k] = uint32_t(g) | (uint32_t(s) << 1) | (last ? 0x100u : 0u) | (release ? 0x200u : 0u) | (rem[g] == T ? 0x400u : 0u);
          __threadfence_block();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_smem_desc_sw128(x_addr + s * kBox + ks * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(zs_addr + g * 8192 + ks * 32, 16, 1024);
            umma_ss_f16(tmem_base + uint32_t(k * 64), ad, bd, idesc1[g], ks ? 1u : 0u);
          }
          umma_commit(&s_full[k]);
          ++seq;
          if (--rem[g] == 0) { rem[g] = T; ++upd[g]; }
        }
        ++pos;
      }
      s_total_seq = ok ? seq : 0;
      s_total_pos = ok ? pos : 0;
      __threadfence_block();
    }
  } else if (warp == 2) {
    // ---- GEMM2 issuer: ring entries in order
    if (elect_one()) {
      constexpr uint32_t idesc2 = make_idesc_bf16(64, kD, 1, 1);
      const uint32_t x_addr = smem_u32(xst), p_addr = smem_u32(pbuf);
      for (int seq = 0;; ++seq) {
        const int k = seq % kRing;
        const int w = wait_or_stop(&p_ready[k], (seq / kRing) & 1, &s_total_seq, seq, p.err);
        if (w <= 0) break;
        tc_fence_after();
        const uint32_t meta = *reinterpret_cast<volatile uint32_t*>(&s_meta[k]);
        const int g = meta & 1, s = (meta >> 1) & 0x7F;
        const uint32_t d_addr = tmem_base + kColO + uint32_t(g * 64);
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks) {
          const uint64_t ad = make_smem_desc_sw128(p_addr + k * kBox + ks * 2048, kBox, 1024);
          const uint64_t bd = make_smem_desc_sw128(x_addr + s * kBox + ks * 2048, kBox, 1024);
          umma_ss_f16(d_addr, ad, bd, idesc2, ((meta & 0x400u) && ks == 0) ? 0u : 1u);
        }
        umma_commit(&s_free[k]);
        if (meta & 0x200u) umma_commit(&x_empty[s]);
        if (meta & 0x100u) umma_commit(&o_full[g]);
      }
    }
  } else if (warp < 11) {
    // ---- weight warps: set 0 takes the even ring entries, set 1 the odd ones
    const int set = (warp - 3) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // point of the tile
    const float c1 = p.c1;
    const float c0 = 0.0028150156f - c1;           // * (1 + 2^-9): the truncating bf16 pack rounds to nearest; cancels in the normalisation
    for (int seq = set;; seq += 2) {
      const int k = seq % kRing;
      const int w = wait_or_stop(&s_full[k], (seq / kRing) & 1, &s_total_seq, seq, p.err);
      if (w <= 0) break;
      tc_fence_after();
      const uint32_t meta = *reinterpret_cast<volatile uint32_t*>(&s_meta[k]);
      const int g = meta & 1;
      const int chunks = (mg[g] + 7) >> 3;         // 8-seed (16-byte) units that hold real seeds
      const uint32_t sa = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(k * 64);
      uint8_t* myrow = pbuf + k * kBox + row * 128;
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(sa, va);
      if (chunks > 4) tmem_ld_32x32b_x32(sa + 32, vb);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (j < chunks) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t s0 = (j < 4) ? va[(j & 3) * 8 + 2 * e] : vb[(j & 3) * 8 + 2 * e];
            const uint32_t s1 = (j < 4) ? va[(j & 3) * 8 + 2 * e + 1] : vb[(j & 3) * 8 + 2 * e + 1];
            pk[e] = pack_bf16x2_trunc(ex2_approx(fmaf(__uint_as_float(s0), c1, c0)), ex2_approx(fmaf(__uint_as_float(s1), c1, c0)));
          }
          o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        *reinterpret_cast<uint4*>(myrow + ((j ^ (row & 7)) << 4)) = o;
      }
      fence_proxy_async();                         // generic-proxy writes of P -> visible to the tensor core's operand reads
      tc_fence_before();
      mbar_arrive(&p_ready[k]);
    }
  } else {
    // ---- exchange warps
    const int q = warp & 3;
    const int t128 = threadIdx.x - 11 * 32;        // 0..127
    const int w4 = t128 >> 5;
    unsigned int* done = p.done + size_t(b) * 64;
    bool ok = true;
    auto stage_z = [&](int g) {
      // fp32 rows of group g -> bf16, K-major 128B-swizzled (row r, 16-byte unit c at ((c ^ (r & 7)) << 4))
      const int r = t128 & 63, half = t128 >> 6;
      const int base = g ? p.mA : 0;
      const float* zr = p.Z + (size_t(b) * p.m + base + (r < mg[g] ? r : 0)) * kD;
#pragma unroll
      for (int c = half * 4; c < half * 4 + 4; ++c) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < mg[g]) {
          const float4 a = __ldcg(reinterpret_cast<const float4*>(zr + c * 8));
          const float4 bb = __ldcg(reinterpret_cast<const float4*>(zr + c * 8 + 4));
          v.x = pack_bf16x2(a.x, a.y); v.y = pack_bf16x2(a.z, a.w);
          v.z = pack_bf16x2(bb.x, bb.y); v.w = pack_bf16x2(bb.z, bb.w);
        }
        *reinterpret_cast<uint4*>(zs + g * 8192 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
      }
      fence_proxy_async();
      mbar_arrive(&z_ready[g]);
    };
    for (int g = 0; g < ngroups; ++g) stage_z(g);
    for (int step = 0; step < ngroups * iters && ok; ++step) {
      const int g = ngroups == 2 ? (step & 1) : 0, u = ngroups == 2 ? (step >> 1) : step;
      const int base = g ? p.mA : 0;
      unsigned int* parts_done = done + 16 * g;
      unsigned int* rows_done = done + 32 + 16 * g;
      // ---- drain O_g (M = 64 accumulator: row i lives in lane (i % 16) + 32 (i / 16)) -> this CTA's partial sums
      if (!mbar_wait(&o_full[g], uint32_t(u) & 1u, p.err)) { ok = false; break; }
      tc_fence_after();
      {
        const int rl = 16 * q + lane;
        float* dst = p.partials + ((size_t(b) * P + cta) * 128 + base + rl) * kD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + kColO + uint32_t(g * 64 + c * 32), v);
          tmem_wait_ld();
          if (lane < 16 && rl < mg[g]) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              __stcg(reinterpret_cast<uint4*>(dst + c * 32 + e * 4), make_uint4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]));
          }
        }
      }
      tc_fence_before();
      named_bar(2, 128);
      if (t128 == 0) { __threadfence(); atomicAdd(parts_done, 1u); }
      // ---- reduce + normalise the rows of this group that this CTA owns
      bool polled = false;
      for (int r = cta; r < p.m; r += P) {
        if (r < base || r >= base + mg[g]) continue;
        if (!polled) {
          if (t128 == 0) ok = poll_counter(parts_done, (unsigned int)((u + 1) * P), p.err) && ok;
          named_bar(2, 128);
          polled = true;
        }
        float acc0 = 0.f, acc1 = 0.f;
        constexpr int kRedBatch = 10;              // independent loads in flight per thread (fixed summation order)
        for (int part0 = w4; part0 < P; part0 += 4 * kRedBatch) {
          float v0[kRedBatch], v1[kRedBatch];
#pragma unroll
          for (int i = 0; i < kRedBatch; ++i) {
            const int part = part0 + 4 * i;
            const float* src = p.partials + ((size_t(b) * P + (part < P ? part : w4)) * 128 + r) * kD;
            v0[i] = (part < P) ? __ldcg(src + lane) : 0.f;
            v1[i] = (part < P) ? __ldcg(src + lane + 32) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < kRedBatch; ++i) { acc0 += v0[i]; acc1 += v1[i]; }
        }
        s_part[w4][lane] = acc0;
        s_part[w4][lane + 32] = acc1;
        named_bar(2, 128);
        float tot = 0.f;
        if (t128 < kD) {
#pragma unroll
          for (int w = 0; w < 4; ++w) tot += s_part[w][t128];
        }
        float sq = tot * tot;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if (lane == 0 && w4 < 2) s_sq[w4] = sq;
        named_bar(2, 128);
        const float denom = fmaxf(sqrtf(s_sq[0] + s_sq[1]), 1e-12f);     // F.normalize eps (lib/utils/mean_shift.py:107)
        if (t128 < kD) __stcg(p.Z + (size_t(b) * p.m + r) * kD + t128, tot / denom);
        named_bar(2, 128);
        if (t128 == 0) { __threadfence(); atomicAdd(rows_done, 1u); }
      }
      // ---- re-stage Z_g for the next update
      if (u + 1 < iters) {
        if (t128 == 0) ok = poll_counter(rows_done, (unsigned int)((u + 1) * mg[g]), p.err) && ok;
        named_bar(2, 128);
        stage_z(g);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

constexpr int kSmem2 = 1024 + 2 * 8192 + kXStages * kBox + kRing * kBox + 256;

}  // namespace

bool hill_climb_tc2_supported(const ClusterShape& s, int iters) { return s.d == 64 && s.m >= 1 && s.m <= 128 && iters >= 1; }

// P CTAs per field (P <= tiles, P * batch <= SM count): cooperative launch, one launch = all updates
int launch_hill_climb_tc2(const CUtensorMap& tmap, const ClusterShape& s, const ClusterWorkspace& w, float* Z, int P, float kappa,
                          int iters, cudaStream_t stream) {
  int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&meanshift_tc2_kernel), kSmem2);
  if (rc != UOC_OK) return rc;
  Ms2Params prm;
  memset(&prm, 0, sizeof(prm));
  prm.tmap_x = tmap;
  prm.Z = Z;
  prm.partials = w.partials;
  prm.done = reinterpret_cast<unsigned int*>(w.slots);
  prm.m = s.m;
  // two groups of at most 64 seeds, group A a multiple of 8 (m = 100: 56 + 44 -> 56 + 48 exponentials per point)
  int mA = (((s.m + 1) / 2) + 7) / 8 * 8;
  if (mA > 64) mA = 64;
  if (mA > s.m) mA = s.m;
  if (s.m - mA > 64) return fail(UOC_ERR_UNSUPPORTED, "num_seeds > 128 is not supported");
  prm.mA = mA;
  prm.n = s.n;
  prm.c1 = kappa * 1.4426950408889634f;
  prm.P = P;
  prm.iters = iters;
  prm.err = device_error_word();
  if (!prm.err) return fail(UOC_ERR_CUDA, "no device error word");
  UOC_CUDA(cudaMemsetAsync(prm.done, 0, 256 * size_t(s.batch), stream));
  void* args[] = {&prm};
  UOC_CUDA(cudaLaunchCooperativeKernel(meanshift_tc2_kernel, dim3(P, s.batch), dim3(kThreads2), args, kSmem2, stream));
  count_launch();
  return UOC_OK;
}

}  // namespace uoc
