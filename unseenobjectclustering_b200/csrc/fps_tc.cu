// K3 (third generation): farthest point sampling, screening pass on the 5th-gen tensor cores.
// (lib/utils/mean_shift.py:128-189: seed 0 is given; seed i+1 = argmax_p min_{j<=i} 0.5 (1 - x_p . x_{s_j}).)
//
// The second-generation kernel (cluster_kernels.cu, fps2_kernel) re-reads the whole fp32 field in every one of the
// m-1 passes: 78.6 MB per pass at 640x480x64, served by L2 at the ~48 B/clk/SM ingest ceiling -> ~5.6 us per pass.
// But the running minimum r[p] is only ever LOWERED by a new seed, so a pass needs the exact fp32 distance only for
// the points the new seed can lower -- essentially its Voronoi cell.  Every point is first screened with the bf16
// pixel-major copy of the field (the one the mean-shift loop streams):
//     d~ = 0.5 (1 - bf16(x_p) . s)                       fp32 accumulate; s split in two bf16 terms (error 2^-17 |s|)
//     |d~ - d| <= 2^-8 |x_p| |s| + slop                  bf16 rounding OR truncation of x_p (|x - bf16 x| <= 2^-7 |x|,
//                                                        Cauchy-Schwarz, halved by the 0.5 of the distance)
//     d~ - margin >= r[p]  =>  d >= r[p]  =>  r[p] is unchanged: no fp32 read
// and only the remaining points evaluate the canonical fp32 chain (one fmaf chain over the channels from +0, exactly
// as fps_kernel / fps2_kernel / the oracle), so r[], the arg-max keys and the selected indices are bit-identical.
// Pass 0 evaluates every point exactly and records |x_p| (the bound does not assume unit-norm rows).
// On a bench frame 5% of the points take the fp32 path on average (profiles/r01_fps_tc_trace.txt).
//
// Here the screen of a whole CTA is a handful of tcgen05.mma:  D_t[128 points x 16] = A_t[128 points x d] . B[d x 16],
// B = {bf16(s), bf16(s - bf16(s)), 0, ..} (the seed split in two bf16 terms; N = 16 is the smallest legal N).
//   * The CTA's slice of the bf16 copy is RESIDENT for all m passes and never re-read from L2: as many 128-point tiles
//     as fit next to the accumulators live in TENSOR MEMORY as the A operand (written once with tcgen05.st; .kind::f16
//     reads A from TMEM), the other tiles in shared memory (K-major, 128-byte swizzle).  640x480x64: 17 tiles per CTA
//     = 7 in TMEM (224 columns) + 10 in shared memory (160 KB); accumulators 17 x 16 columns.
//   * MMAs that accumulate into the same tile are ~64 clk apart (measured: sharing one accumulator between 8 tiles
//     through 8 column-shifted copies of B made the screen SLOWER), so every tile has its own 16 columns and the
//     issue order is k-step major: consecutive MMAs touch different accumulators.
//   * One elected thread issues the 4 (d=64) MMAs per tile and ONE commit per pass; every thread then reads the two
//     accumulator columns of its own point with a tcgen05.ld .x2 (TMEM lane == point of the tile).
//   * 128-point tiles are dealt round-robin to the CTAs of an item (the points a seed can lower form a compact blob).
// Per pass: 2 CTA-wide barriers, the all-to-all key exchange of fps2_kernel, seed fetch + B operand by warp 0.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cluster.cuh"

namespace uoc {

namespace {

constexpr int kThreads4 = 512;
constexpr int kWarps4 = kThreads4 / 32;
constexpr int kMaxSlots = 5;               // tiles per warp group (4 groups): <= 20 tiles (2560 points) per CTA
constexpr float kRelMargin4 = 0.0040f;     // 2^-8 (bf16 truncation; halved by the 0.5 of the distance) + slack for split / sum order
constexpr float kAbsMargin4 = 2e-6f;

struct Fps4Params {
  const float* X;
  const uint4* xb;          // bf16 pixel-major copy [batch][n][d], as 16-byte units
  long long sb, sd, n;
  int d, m, batch;
  const long long* first;
  long long* selected_out;
  float* seeds_out;
  unsigned int* err;
  unsigned long long* slots;
  int nb;                   // CTAs per batch item
  int T;                    // tiles of the busiest CTA (accumulator columns: 16 * T)
  int TA;                   // tiles kept in tensor memory (the rest in shared memory)
  float rel_margin;         // screening margin per unit |x||s| (kRelMargin4, or half of it for round-to-nearest copies)
  int prefetch;             // exact chain: prefetch the later channel batches into L1 (UOC_FPS_PREFETCH, A/B)
  unsigned int* stats;      // debug (UOC_FPS_TC_STATS=1): per pass {exact rounds, lanes in them, max rounds of one warp, rounds with <= 4 lanes}
  long long* trace;         // debug (UOC_FPS_TC_TRACE=<cta>): per pass {start, screened, local arg-max, exchanged+seed, fp32 rounds}
  int trace_cta;
};

__device__ __forceinline__ unsigned int orderable4(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_key4(float r, unsigned int idx) {
  return (static_cast<unsigned long long>(orderable4(r)) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ void tmem_ld_32x32b_x2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}

template <int D>
__global__ void __launch_bounds__(kThreads4, 1) fps4_kernel(Fps4Params p) {
  constexpr int KB = D / 64;                 // 64-channel blocks: one 128-byte swizzled row per point and block
  constexpr int NL = D / 8;                  // 16-byte units per point
  constexpr int ACOLS = D / 2;               // tensor-memory columns of one A tile (two bf16 per column)
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_seed[D];
  __shared__ float s_ns;
  __shared__ unsigned long long s_red[kWarps4];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ int s_fail;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bsm = smem;                       // B operand: KB blocks of [16 rows][128 B]
  uint8_t* atiles = smem + KB * 2048;        // shared-memory A tiles: [tile][KB][128 rows][128 B]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, gi = warp >> 2;    // TMEM lane quadrant this warp may access; tile group
  const int nb = p.nb;
  const int b = blockIdx.x / nb, rank = blockIdx.x % nb;
  const int total_tiles = int((p.n + 127) >> 7);
  const int T = rank < total_tiles ? (total_tiles - rank + nb - 1) / nb : 0;     // tiles of this CTA
  const int TA = p.TA < T ? p.TA : T;
  const float* Xb = p.X + b * p.sb;
  const uint4* xbb = p.xb + size_t(b) * p.n * NL;
  auto tile_base = [&](int t) { return ((long long)t * nb + rank) * 128; };       // first point of local tile t
  const uint32_t dcol0 = 0, acol0 = 16u * uint32_t(p.T);

  if (warp == 0) {
    tmem_alloc(&s_tmem, 512);
    tmem_relinquish();
    if (lane == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); s_fail = 0; }
  }
  for (int e = tid; e < KB * 2048 / 16; e += kThreads4) reinterpret_cast<uint4*>(bsm)[e] = make_uint4(0u, 0u, 0u, 0u);
  // shared-memory A tiles: chunk c (8 channels) of row r of block kb goes to ((c ^ (r & 7)) << 4) of its 128-byte row
  for (int e = tid; e < (T - TA) * 128 * NL; e += kThreads4) {
    const int ts = e / (128 * NL), row = (e / NL) % 128, chunk = e % NL;
    const long long pnt = tile_base(ts + TA) + row;
    const uint4 v = (pnt < p.n) ? __ldg(xbb + pnt * NL + chunk) : make_uint4(0u, 0u, 0u, 0u);
    const int kb = chunk >> 3, c = chunk & 7;
    *reinterpret_cast<uint4*>(atiles + (size_t(ts * KB + kb) * 128 + row) * 128 + ((c ^ (row & 7)) << 4)) = v;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
  // tensor-memory A tiles: TMEM lane == point of the tile, column k/2 holds channels k, k+1 (the raw words of the row)
  for (int t = gi; t < TA; t += 4) {
    const long long pnt = tile_base(t) + q * 32 + lane;
#pragma unroll
    for (int c16 = 0; c16 < ACOLS / 16; ++c16) {
      uint32_t w[16];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint4 v = (pnt < p.n) ? __ldg(xbb + pnt * NL + c16 * 4 + u) : make_uint4(0u, 0u, 0u, 0u);
        w[4 * u] = v.x; w[4 * u + 1] = v.y; w[4 * u + 2] = v.z; w[4 * u + 3] = v.w;
      }
      tmem_st_32x32b_x16(lane_addr + acol0 + uint32_t(t) * ACOLS + c16 * 16, w);
    }
  }
  tmem_wait_st();

  unsigned long long mybest = 1ull;
  float r[kMaxSlots], nx[kMaxSlots];
#pragma unroll
  for (int s = 0; s < kMaxSlots; ++s) { r[s] = 0.f; nx[s] = 0.f; }

  // warp 0: fetch seed `idx` (fp32), its norm, and the B operand {bf16(s), bf16(s - bf16(s))} of the screen
  auto stage_seed = [&](long long idx, int i) {
    float sq = 0.f;
#pragma unroll
    for (int h = 0; h < D / 64; ++h) {
      const int c0 = h * 64 + 2 * lane;
      const float v0 = __ldg(Xb + c0 * p.sd + idx), v1 = __ldg(Xb + (c0 + 1) * p.sd + idx);
      s_seed[c0] = v0; s_seed[c0 + 1] = v1;
      sq = fmaf(v0, v0, fmaf(v1, v1, sq));
      const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
      const int chunk = lane >> 2, off = (lane & 3) * 4;          // channels 2*lane, 2*lane+1 of block h
      // row r (= column of B), chunk c of block h at h * 2048 + r * 128 + ((c ^ (r & 7)) << 4)
      *reinterpret_cast<uint32_t*>(bsm + h * 2048 + 0 * 128 + ((chunk ^ 0) << 4) + off) = pack_bf16x2(h0, h1);
      *reinterpret_cast<uint32_t*>(bsm + h * 2048 + 1 * 128 + ((chunk ^ 1) << 4) + off) = pack_bf16x2(v0 - h0, v1 - h1);
      if (rank == 0) {
        float* so = p.seeds_out + (size_t(b) * p.m + i) * D;
        so[c0] = v0; so[c0 + 1] = v1;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) {
      s_ns = sqrtf(sq) * 1.000001f;
      if (rank == 0) p.selected_out[size_t(b) * p.m + i] = idx;
    }
    fence_proxy_async();
  };
  if (warp == 0) stage_seed(p.first[b], 0);
  tc_fence_before();
  __syncthreads();

  constexpr uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
  const uint32_t bsm_addr = smem_u32(bsm), at_addr = smem_u32(atiles);
  const bool trc = p.trace && int(blockIdx.x) == p.trace_cta && tid == 0;

  for (int i = 0; i + 1 < p.m; ++i) {
    if (trc) p.trace[i * 6 + 0] = clock64();
    // ---- screen: all tiles of the CTA against seed i
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {           // k-step major: consecutive MMAs accumulate into different tiles
            const uint64_t bd = make_smem_desc_sw128(bsm_addr + kb * 2048 + ks * 32, 16, 1024);
            const uint32_t accumulate = (kb | ks) ? 1u : 0u;
            for (int t = 0; t < TA; ++t)
              umma_ts_f16(tmem_base + dcol0 + 16u * t, tmem_base + acol0 + uint32_t(t) * ACOLS + kb * 32 + ks * 8, bd, idesc, accumulate);
            for (int t = TA; t < T; ++t) {
              const uint64_t ad = make_smem_desc_sw128(at_addr + uint32_t((t - TA) * KB + kb) * 16384u + ks * 32, 16, 1024);
              umma_ss_f16(tmem_base + dcol0 + 16u * t, ad, bd, idesc, accumulate);
            }
          }
        }
        umma_commit(&s_bar);
      }
      __syncwarp();
    }
    const bool ok = mbar_wait(&s_bar, uint32_t(i) & 1u, p.err);
    tc_fence_after();
    uint32_t c0[kMaxSlots], c1[kMaxSlots];
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
      const int t = gi + 4 * s;
      c0[s] = c1[s] = 0u;
      if (t < T) tmem_ld_32x32b_x2(lane_addr + dcol0 + 16u * uint32_t(t), c0[s], c1[s]);
    }
    tmem_wait_ld();
    if (trc) p.trace[i * 6 + 1] = clock64();
    const float ns = s_ns;
    int n_exact = 0, n_lanes = 0, n_small = 0;
    bool changed = (i == 0);                  // r[] of this thread changed: its cached arg-max key is stale
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
      const int t = gi + 4 * s;
      if (t < T) {                            // warp-uniform
        const long long gp = tile_base(t) + q * 32 + lane;
        const bool valid = gp < p.n;
        bool need = valid;
        if (i > 0 && ok) {
          const float dapprox = 0.5f * (1.0f - (__uint_as_float(c0[s]) + __uint_as_float(c1[s])));
          need = valid && !((dapprox - fmaf(p.rel_margin * nx[s], ns, kAbsMargin4)) >= r[s]);
        }
        if (__any_sync(0xffffffffu, need)) {
          ++n_exact;
          if (p.stats) { const int c = __popc(__ballot_sync(0xffffffffu, need)); n_lanes += c; n_small += (c <= 4) ? 1 : 0; }
          if (need) {
            // canonical fp32 chain (bit-identical to fps_kernel / fps2_kernel / the oracle)
            const float* xp = Xb + gp;
            float acc = 0.f, sq = 0.f;
            constexpr int XB = 32;                    // loads in flight per lane (64 is slower: measured)
            if (p.prefetch) {
              // the later load batches depend on the registers of the first one: pull their lines into L1 now, so that
              // a round costs ONE L2 round trip instead of D / 32
#pragma unroll
              for (int k = XB; k < D; ++k) asm volatile("prefetch.global.L1 [%0];" ::"l"(xp + k * p.sd));
            }
#pragma unroll 1
            for (int k0 = 0; k0 < D; k0 += XB) {
              float x[XB];
#pragma unroll
              for (int k = 0; k < XB; ++k) x[k] = __ldg(xp + (k0 + k) * p.sd);
#pragma unroll
              for (int k = 0; k < XB; ++k) {
                acc = fmaf(x[k], s_seed[k0 + k], acc);
                if (i == 0) sq = fmaf(x[k], x[k], sq);
              }
            }
            const float dist = 0.5f * (1.0f - acc);
            if (i == 0) {
              r[s] = dist;
              nx[s] = sqrtf(sq) * 1.000001f;
            } else {
              changed = changed || dist < r[s];
              r[s] = dist < r[s] ? dist : r[s];
            }
          }
        }
      }
    }
    if (changed) {                            // rare after the first passes: r[] only changes on the fp32 path
      mybest = 1ull;                          // non-zero sentinel: an empty slice still signals arrival
#pragma unroll
      for (int s = 0; s < kMaxSlots; ++s) {
        const int t = gi + 4 * s;
        const long long gp = tile_base(t) + q * 32 + lane;
        if (t < T && gp < p.n) {
          const unsigned long long key = pack_key4(r[s], static_cast<unsigned int>(gp));
          mybest = key > mybest ? key : mybest;
        }
      }
    }
    unsigned long long best = mybest;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_red[warp] = best;
    if (trc) { p.trace[i * 6 + 2] = clock64(); p.trace[i * 6 + 4] = n_exact; }
    if (p.stats && lane == 0 && n_exact) {
      atomicAdd(p.stats + i * 4 + 0, (unsigned int)n_exact); atomicAdd(p.stats + i * 4 + 1, (unsigned int)n_lanes);
      atomicMax(p.stats + i * 4 + 2, (unsigned int)n_exact); atomicAdd(p.stats + i * 4 + 3, (unsigned int)n_small);
    }
    tc_fence_before();
    __syncthreads();                          // everybody is done with s_seed, s_ns and the accumulators of this pass
    if (trc) p.trace[i * 6 + 5] = clock64();
    if (warp == 0) {
      unsigned long long v = (lane < kWarps4) ? s_red[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
      }
      // all-to-all: this CTA's key goes into column `rank` of EVERY CTA's private row; then poll the own row
      unsigned long long* mat = p.slots + (size_t(b) * p.m + (i + 1)) * nb * nb;
#pragma unroll
      for (int c5 = 0; c5 < 5; ++c5) {
        const int c = lane + 32 * c5;
        if (c < nb) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mat + size_t(c) * nb + rank), "l"(v) : "memory");
      }
      const unsigned long long* row = mat + size_t(rank) * nb;
      unsigned long long gmax = 0ull;
      bool done = false;
      for (unsigned int it = 0; it < (1u << 22) && !done; ++it) {
        unsigned long long kv[5];
#pragma unroll
        for (int c5 = 0; c5 < 5; ++c5) {
          const int c = lane + 32 * c5;
          kv[c5] = 1ull;
          if (c < nb) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(kv[c5]) : "l"(row + c));
        }
        bool all = true;
        gmax = 0ull;
#pragma unroll
        for (int c5 = 0; c5 < 5; ++c5) {
          all = all && (kv[c5] != 0ull);
          gmax = kv[c5] > gmax ? kv[c5] : gmax;
        }
        done = __all_sync(0xffffffffu, all);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, gmax, o);
        gmax = other > gmax ? other : gmax;
      }
      if (done && ok) {
        stage_seed(static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(gmax & 0xFFFFFFFFull)), i + 1);
      } else if (lane == 0) {
        s_fail = 1;
        atomicOr(p.err, ERR_GRID_BARRIER_TIMEOUT);
      }
    }
    tc_fence_before();
    __syncthreads();
    if (trc) p.trace[i * 6 + 3] = clock64();
    if (s_fail) break;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}


// ----------------------------------------------------------------------------------------------
// Streaming variant for fields that do NOT fit on chip (BASELINE config 5: 960x720x128, bf16 copy 177 MB): the same
// screen-then-exact scheme, but every pass STREAMS the bf16 pixel-major copy (n*d*2 bytes instead of the n*d*4 of the fp32
// passes of fps2_kernel) and screens on the CUDA cores: each lane reads its own point's row with 16-byte loads (a warp's 32
// rows are contiguous; a shared-memory staged variant with cp.async measured the same: 6.0 vs 6.2 ms at config 5) and
// accumulates bf16(x_p) . s with the fp32 seed.  Config 5: 9.1 ms (fp32 passes) -> 6.1 ms; the rest is the exact chain of
// the new seed's own cluster (bf16 cannot resolve within-cluster distances) reading fp32 rows from HBM.  |d~ - d| <= 2^-8 |x_p||s| + rounding (the seed is exact here, so no
// two-term split is needed); everything the screen cannot prove unchanged runs the canonical fp32 chain -> bit-identical
// indices.  32-point tiles are dealt round-robin to all warps of the item, running minima live in registers.
// ----------------------------------------------------------------------------------------------
constexpr int kMaxSlots5 = 18;             // 32-point tiles per warp: n <= 148 * 16 * 32 * 18 = 1.36M points (4 x 640x480 on 37 CTAs each)

// canonical fp32 chain over the channels of point xp (planar field, channel stride sd): out of line so that the twelve
// unrolled slots of fps5_kernel do not each carry their own 32-register load batch
template <int D>
__device__ __noinline__ float exact_chain5(const float* __restrict__ xp, long long sd, const float* s_seed, bool want_sq,
                                           float* sq_out) {
  float acc = 0.f, sq = 0.f;
  constexpr int XB = 32;
#pragma unroll 1
  for (int k0 = 0; k0 < D; k0 += XB) {
    float x[XB];
#pragma unroll
    for (int k = 0; k < XB; ++k) x[k] = __ldg(xp + (k0 + k) * sd);
#pragma unroll
    for (int k = 0; k < XB; ++k) {
      acc = fmaf(x[k], s_seed[k0 + k], acc);
      if (want_sq) sq = fmaf(x[k], x[k], sq);
    }
  }
  *sq_out = sq;
  return acc;
}

template <int D>
__global__ void __launch_bounds__(kThreads4, 1) fps5_kernel(Fps4Params p) {
  constexpr int NL = D / 8;                  // 16-byte units per point
  __shared__ float s_seed[D];
  __shared__ float s_ns;
  __shared__ unsigned long long s_red[kWarps4];
  __shared__ int s_fail;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = p.nb;
  const int b = blockIdx.x / nb, rank = blockIdx.x % nb;
  const int NW = nb * kWarps4;               // warps of the item
  const int gw = rank * kWarps4 + warp;      // this warp among them
  const long long tiles = (p.n + 31) >> 5;
  const float* Xb = p.X + b * p.sb;
  const uint4* xbb = p.xb + size_t(b) * p.n * NL;

  if (tid == 0) s_fail = 0;
  unsigned long long mybest = 1ull;
  float r[kMaxSlots5], nx[kMaxSlots5];
#pragma unroll
  for (int s = 0; s < kMaxSlots5; ++s) { r[s] = 0.f; nx[s] = 0.f; }

  auto stage_seed = [&](long long idx, int i) {       // warp 0: fetch seed `idx` (fp32) and its norm
    float sq = 0.f;
#pragma unroll
    for (int h = 0; h < D / 64; ++h) {
      const int c0 = h * 64 + 2 * lane;
      const float v0 = __ldg(Xb + c0 * p.sd + idx), v1 = __ldg(Xb + (c0 + 1) * p.sd + idx);
      s_seed[c0] = v0; s_seed[c0 + 1] = v1;
      sq = fmaf(v0, v0, fmaf(v1, v1, sq));
      if (rank == 0) {
        float* so = p.seeds_out + (size_t(b) * p.m + i) * D;
        so[c0] = v0; so[c0 + 1] = v1;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) {
      s_ns = sqrtf(sq) * 1.000001f;
      if (rank == 0) p.selected_out[size_t(b) * p.m + i] = idx;
    }
  };
  if (warp == 0) stage_seed(p.first[b], 0);
  __syncthreads();

  for (int i = 0; i + 1 < p.m; ++i) {
    const float ns = s_ns;
    bool changed = (i == 0);
#pragma unroll
    for (int s = 0; s < kMaxSlots5; ++s) {
      const long long t = (long long)s * NW + gw;
      if (t < tiles) {                          // warp-uniform
        const long long base = t << 5;
        const long long gp = base + lane;
        const bool valid = gp < p.n;
        bool need = valid;
        if (i > 0) {
          // ---- every lane streams its own point's bf16 row: 8 x 16-byte loads in flight per lane; the warp's 32 rows are
          // 32 * D * 2 contiguous bytes and the consecutive loads of a lane hit the L1 lines its first load brought in
          float acc0 = 0.f, acc1 = 0.f;
          if (valid) {
            const uint4* row = xbb + gp * NL;
#pragma unroll
            for (int u0 = 0; u0 < NL; u0 += 8) {
              uint4 v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = __ldg(row + u0 + u);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float* sk = s_seed + (u0 + u) * 8;
                acc0 = fmaf(__uint_as_float(v[u].x << 16), sk[0], acc0); acc1 = fmaf(__uint_as_float(v[u].x & 0xFFFF0000u), sk[1], acc1);
                acc0 = fmaf(__uint_as_float(v[u].y << 16), sk[2], acc0); acc1 = fmaf(__uint_as_float(v[u].y & 0xFFFF0000u), sk[3], acc1);
                acc0 = fmaf(__uint_as_float(v[u].z << 16), sk[4], acc0); acc1 = fmaf(__uint_as_float(v[u].z & 0xFFFF0000u), sk[5], acc1);
                acc0 = fmaf(__uint_as_float(v[u].w << 16), sk[6], acc0); acc1 = fmaf(__uint_as_float(v[u].w & 0xFFFF0000u), sk[7], acc1);
              }
            }
          }
          const float dapprox = 0.5f * (1.0f - (acc0 + acc1));
          need = valid && !((dapprox - fmaf(p.rel_margin * nx[s], ns, kAbsMargin4)) >= r[s]);
        }
        if (__any_sync(0xffffffffu, need)) {
          if (need) {
            // canonical fp32 chain (bit-identical to fps_kernel / fps2_kernel / fps4_kernel / the oracle)
            float sq;
            const float acc = exact_chain5<D>(Xb + gp, p.sd, s_seed, i == 0, &sq);
            const float dist = 0.5f * (1.0f - acc);
            if (i == 0) {
              r[s] = dist;
              nx[s] = sqrtf(sq) * 1.000001f;
            } else {
              changed = changed || dist < r[s];
              r[s] = dist < r[s] ? dist : r[s];
            }
          }
        }
      }
    }
    if (changed) {
      mybest = 1ull;
#pragma unroll
      for (int s = 0; s < kMaxSlots5; ++s) {
        const long long t = (long long)s * NW + gw;
        const long long gp = (t << 5) + lane;
        if (t < tiles && gp < p.n) {
          const unsigned long long key = pack_key4(r[s], static_cast<unsigned int>(gp));
          mybest = key > mybest ? key : mybest;
        }
      }
    }
    unsigned long long best = mybest;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();                          // everybody is done with s_seed / s_ns of this pass
    if (warp == 0) {
      unsigned long long v = (lane < kWarps4) ? s_red[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
      }
      // all-to-all: this CTA's key goes into column `rank` of EVERY CTA's private row; then poll the own row
      unsigned long long* mat = p.slots + (size_t(b) * p.m + (i + 1)) * nb * nb;
#pragma unroll
      for (int c5 = 0; c5 < 5; ++c5) {
        const int c = lane + 32 * c5;
        if (c < nb) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mat + size_t(c) * nb + rank), "l"(v) : "memory");
      }
      const unsigned long long* row = mat + size_t(rank) * nb;
      unsigned long long gmax = 0ull;
      bool done = false;
      for (unsigned int it = 0; it < (1u << 22) && !done; ++it) {
        unsigned long long kv[5];
#pragma unroll
        for (int c5 = 0; c5 < 5; ++c5) {
          const int c = lane + 32 * c5;
          kv[c5] = 1ull;
          if (c < nb) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(kv[c5]) : "l"(row + c));
        }
        bool all = true;
        gmax = 0ull;
#pragma unroll
        for (int c5 = 0; c5 < 5; ++c5) {
          all = all && (kv[c5] != 0ull);
          gmax = kv[c5] > gmax ? kv[c5] : gmax;
        }
        done = __all_sync(0xffffffffu, all);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, gmax, o);
        gmax = other > gmax ? other : gmax;
      }
      if (done) {
        stage_seed(static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(gmax & 0xFFFFFFFFull)), i + 1);
      } else if (lane == 0) {
        s_fail = 1;
        atomicOr(p.err, ERR_GRID_BARRIER_TIMEOUT);
      }
    }
    __syncthreads();
    if (s_fail) break;
  }
}

}  // namespace

int launch_select_seeds_tc(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w,
                           int64_t* selected_out, float* seeds_out, cudaStream_t stream_, bool* used) {
  cudaStream_t stream = stream_;
  *used = false;
  if (!xb || (s.d != 64 && s.d != 128)) return UOC_OK;
  if (reinterpret_cast<uintptr_t>(xb) % 16 != 0) return UOC_OK;
  if (const char* e = getenv("UOC_FPS_TC")) { if (atoi(e) == 0) return UOC_OK; }
  const int sms = sm_count();
  if (sms <= 0 || s.batch > sms) return UOC_OK;
  // a batch of fields that do not fit on chip together: the fields take turns in the resident-slice kernel
  // (launch_select_seeds); UOC_FPS_BATCH_STREAM=1 streams them side by side instead (A/B knob: slower)
  if (s.batch > 1) {
    const char* e = getenv("UOC_FPS_BATCH_STREAM");
    const bool side_by_side = e ? atoi(e) != 0 : false;   // measured (profiles/r02_fps_batch.txt): turns 0.72 ms / field, side by side 1.25 - 1.40
    const long long tiles_per_cta = ((s.n + 127) / 128 + (sms / s.batch) - 1) / (sms / s.batch);
    if (tiles_per_cta > 4 * kMaxSlots && !side_by_side) return UOC_OK;
  }
  const int nb = sms / s.batch;
  if (nb > 160) return UOC_OK;                      // the poll loop reads at most 5 keys per lane
  float rel_margin = kRelMargin4;
  if (const char* e = getenv("UOC_FPS_RN_MARGIN")) { if (atoi(e) != 0) rel_margin = 0.00205f; }   // A/B: copy known to be round-to-nearest
  const long long total_tiles = (s.n + 127) / 128;
  const long long T = (total_tiles + nb - 1) / nb;  // tiles of the busiest CTA
  bool streaming = false;                           // the field does not fit on chip: stream the bf16 copy in every pass
  if (const char* e = getenv("UOC_FPS_STREAM")) streaming = atoi(e) != 0;   // test knob: force the streaming variant
  if (T > 4 * kMaxSlots) streaming = true;
  {
    const int acols0 = s.d / 2, kb0 = s.d / 64;
    int TA0 = int((512 - 16 * T) / acols0);
    if (TA0 > T) TA0 = int(T);
    if (TA0 < 0) TA0 = 0;
    if (1024 + size_t(kb0) * 2048 + size_t(T - TA0) * kb0 * 16384 > 225 * 1024) streaming = true;
  }
  if (streaming) {
    const long long tiles32 = (s.n + 31) / 32;
    if ((tiles32 + (long long)nb * kWarps4 - 1) / ((long long)nb * kWarps4) > kMaxSlots5) return UOC_OK;   // larger still: fp32 passes
    const size_t slot_need5 = size_t(s.batch) * s.m * nb * nb * 8;
    if (slot_need5 > w.slot_bytes) return UOC_OK;
    unsigned int* err5 = device_error_word();
    if (!err5) return fail(UOC_ERR_CUDA, "no device error word");
    void* kern5 = s.d == 64 ? reinterpret_cast<void*>(&fps5_kernel<64>) : reinterpret_cast<void*>(&fps5_kernel<128>);
    const size_t smem5 = 0;
    Fps4Params p5;
    p5.X = X; p5.xb = reinterpret_cast<const uint4*>(xb);
    p5.sb = s.stride_b; p5.sd = s.stride_d; p5.n = s.n; p5.d = s.d; p5.m = s.m; p5.batch = s.batch;
    p5.first = w.first;
    p5.selected_out = reinterpret_cast<long long*>(selected_out);
    p5.seeds_out = seeds_out;
    p5.err = err5;
    p5.slots = w.slots;
    p5.nb = nb; p5.T = 0; p5.TA = 0;
    p5.rel_margin = rel_margin;
    p5.trace = nullptr; p5.trace_cta = 0; p5.stats = nullptr; p5.prefetch = 0;
    UOC_CUDA(cudaMemsetAsync(w.slots, 0, slot_need5, stream_));
    void* args5[] = {&p5};
    UOC_CUDA(cudaLaunchCooperativeKernel(kern5, dim3(nb * s.batch), dim3(kThreads4), args5, smem5, stream_));
    count_launch();
    *used = true;
    return UOC_OK;
  }
  const int acols = s.d / 2, kb = s.d / 64;
  // accumulators: 16 columns per tile; tensor memory takes as many A tiles as fit next to them, shared memory the rest
  int TA = int((512 - 16 * T) / acols);
  if (TA > T) TA = int(T);
  if (TA < 0) TA = 0;
  if (const char* e = getenv("UOC_FPS_TC_TMEM_TILES")) { const int v = atoi(e); if (v >= 0 && v < TA) TA = v; }   // test / A-B knob
  const size_t smem = 1024 + size_t(kb) * 2048 + size_t(T - TA) * kb * 16384;
  if (smem > 225 * 1024) return UOC_OK;
  const size_t slot_need = size_t(s.batch) * s.m * nb * nb * 8;
  if (slot_need > w.slot_bytes) return UOC_OK;
  unsigned int* err = device_error_word();
  if (!err) return fail(UOC_ERR_CUDA, "no device error word");

  void* kern = s.d == 64 ? reinterpret_cast<void*>(&fps4_kernel<64>) : reinterpret_cast<void*>(&fps4_kernel<128>);
  UOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));

  Fps4Params p;
  p.X = X; p.xb = reinterpret_cast<const uint4*>(xb);
  p.sb = s.stride_b; p.sd = s.stride_d; p.n = s.n; p.d = s.d; p.m = s.m; p.batch = s.batch;
  p.first = w.first;
  p.selected_out = reinterpret_cast<long long*>(selected_out);
  p.seeds_out = seeds_out;
  p.err = err;
  p.slots = w.slots;
  p.nb = nb; p.T = int(T); p.TA = TA;
  p.rel_margin = rel_margin;
  p.trace = nullptr;
  p.trace_cta = 0;
  p.prefetch = 0;
  if (const char* e = getenv("UOC_FPS_PREFETCH")) p.prefetch = atoi(e);
  p.stats = nullptr;
  if (getenv("UOC_FPS_TC_STATS")) {
    UOC_CUDA(cudaMalloc(&p.stats, sizeof(unsigned int) * 4 * s.m));
    UOC_CUDA(cudaMemsetAsync(p.stats, 0, sizeof(unsigned int) * 4 * s.m, stream));
  }
  if (const char* e = getenv("UOC_FPS_TC_TRACE")) {
    p.trace_cta = atoi(e);
    UOC_CUDA(cudaMalloc(&p.trace, sizeof(long long) * 6 * s.m));
    UOC_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(long long) * 6 * s.m, stream));
  }
  UOC_CUDA(cudaMemsetAsync(w.slots, 0, slot_need, stream));
  void* args[] = {&p};
  UOC_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nb * s.batch), dim3(kThreads4), args, smem, stream));
  count_launch();
  if (p.stats) {
    std::vector<unsigned int> hs(size_t(4) * s.m);
    UOC_CUDA(cudaStreamSynchronize(stream));
    UOC_CUDA(cudaMemcpy(hs.data(), p.stats, sizeof(unsigned int) * hs.size(), cudaMemcpyDeviceToHost));
    cudaFree(p.stats);
    unsigned long long tr = 0, tl = 0, tsm = 0, tmax = 0; int none = 0;
    for (int i = 1; i + 1 < s.m; ++i) {
      tr += hs[i * 4]; tl += hs[i * 4 + 1]; tmax += hs[i * 4 + 2]; tsm += hs[i * 4 + 3]; none += hs[i * 4] == 0;
      if (i < 6 || i % 10 == 0)
        fprintf(stderr, "[fps tc stats] pass %d: %u exact rounds in the grid, %u lanes, busiest warp %u rounds, %u rounds with <= 4 lanes\n",
                i, hs[i * 4], hs[i * 4 + 1], hs[i * 4 + 2], hs[i * 4 + 3]);
    }
    fprintf(stderr, "[fps tc stats] passes 1..%d: %.1f rounds / pass (%.1f lanes each), busiest warp %.2f rounds on average, %.0f %% of the rounds have <= 4 lanes, %d passes without any round\n",
            s.m - 2, double(tr) / (s.m - 2), tr ? double(tl) / tr : 0.0, double(tmax) / (s.m - 2), tr ? 100.0 * tsm / tr : 0.0, none);
  }
  if (p.trace) {
    std::vector<long long> h(size_t(6) * s.m);
    UOC_CUDA(cudaStreamSynchronize(stream));
    UOC_CUDA(cudaMemcpy(h.data(), p.trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    double sum[4] = {0, 0, 0, 0};
    for (int i = 0; i + 1 < s.m; ++i) {
      const long long* q = h.data() + size_t(i) * 6;
      if (i < 6 || i % 10 == 0)
        fprintf(stderr, "[fps tc trace] cta %d pass %d: screen %lld  fp32 rounds+arg-max (warp 0) %lld  wait for all warps %lld  exchange+seed %lld clk; fp32 rounds (warp 0) %lld\n",
                p.trace_cta, i, q[1] - q[0], q[2] - q[1], q[5] - q[2], q[3] - q[5], q[4]);
      if (i >= 10) { sum[0] += q[1] - q[0]; sum[1] += q[2] - q[1]; sum[2] += q[5] - q[2]; sum[3] += q[3] - q[5]; }
    }
    const double cnt = s.m - 11;
    fprintf(stderr, "[fps tc trace] mean over passes >= 10: screen %.0f  fp32+arg-max (warp 0) %.0f  wait for all warps %.0f  exchange+seed %.0f clk  (T %d, TA %d)\n",
            sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt, int(T), TA);
  }
  *used = true;
  return UOC_OK;
}

}  // namespace uoc
