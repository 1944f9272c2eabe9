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
  unsigned long long* passlog;  // debug (UOC_FPS_STATS): [m] per pass of CTA 0: (exchange ? 1 << 63 : 0) | clock64 at the pass end
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

// top-2 keys of two (first, second) pairs
__device__ __forceinline__ void merge_top2(unsigned long long& a1, unsigned long long& a2, unsigned long long b1,
                                           unsigned long long b2) {
  const unsigned long long hi = a1 > b1 ? a1 : b1;
  const unsigned long long lo = a1 > b1 ? b1 : a1;
  const unsigned long long s2 = a2 > b2 ? a2 : b2;
  a1 = hi;
  a2 = lo > s2 ? lo : s2;
}
__device__ __forceinline__ float key_r(unsigned long long key) {      // inverse of orderable4 on the high word
  const unsigned int u = static_cast<unsigned int>(key >> 32);
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// Speculative seed chain (round 2).  A pass used to end with an all-to-all exchange of every CTA's arg-max key (~3.2k clk
// plus the wait for the slowest CTA) although the answer is usually known in advance: r[] only ever DECREASES, so the
// exchange of pass i also tells every CTA the runners-up.  Every CTA publishes its TOP-2 keys; all CTAs derive the same
//   list  = the first keys k1 that exceed   bound = max over CTAs of the second key k2   (<= 32 entries), and
//   bound = an upper bound, for ever, of every point that is not in the list.
// After a new seed s every CTA re-evaluates the list entries exactly (canonical fp32 chain on x_q, s -- any CTA can do
// that, it is 64 loads per entry) and takes the largest; if it is still above `bound` it IS the global arg-max (same key
// order, same tie-break) and becomes the next seed WITHOUT any exchange: the CTAs run such passes uncoupled.  Only when the
// list runs dry, or its best entry drops to the bound, a real exchange rebuilds it.  Indices are bit-identical by
// construction: every decision is taken on exact keys.
// canonical fp32 chain acc = fmaf(x_k, s_k, acc), k = 0 .. D-1 from +0 (bit-identical to fps_kernel / fps2_kernel / the
// oracle) for ONE point of the planar field.  (An fp32 pixel-major copy for these rows was built and measured: faster on
// clustered fields, slower on the bench frame, and 21 us per frame in the head kernel: profiles/r02_fps_pixel_major.txt.)
template <int D>
__device__ __forceinline__ float exact_chain4(const float* __restrict__ xp, long long sd, const float* seed, bool want_sq,
                                              float& sq) {
  float acc = 0.f;
  sq = 0.f;
  constexpr int XB = 32;                      // loads in flight per lane (64 were measured slower, also for the sparse
                                              // rounds of the later passes: profiles/r02_fps_speculation.txt)
#pragma unroll 1
  for (int k0 = 0; k0 < D; k0 += XB) {
    float x[XB];
#pragma unroll
    for (int k = 0; k < XB; ++k) x[k] = __ldg(xp + (k0 + k) * sd);
#pragma unroll
    for (int k = 0; k < XB; ++k) {
      acc = fmaf(x[k], seed[k0 + k], acc);
      if (want_sq) sq = fmaf(x[k], x[k], sq);
    }
  }
  return acc;
}

template <int D>
__global__ void __launch_bounds__(kThreads4, 1) fps4_kernel(Fps4Params p) {
  constexpr int KB = D / 64;                 // 64-channel blocks: one 128-byte swizzled row per point and block
  constexpr int NL = D / 8;                  // 16-byte units per point
  constexpr int ACOLS = D / 2;               // tensor-memory columns of one A tile (two bf16 per column)
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_seed[2][D];             // seed i lives in buffer i & 1: the next seed is staged while pass i runs
  __shared__ float s_ns[2];
  __shared__ unsigned long long s_red[2 * kWarps4];
  __shared__ unsigned long long s_list[32];
  __shared__ unsigned long long s_drop, s_bound;
  __shared__ int s_cnt, s_newlist;
  __shared__ long long s_next[2];            // [pass parity] next seed index decided by the speculative chain, -1: exchange needed
  __shared__ uint64_t s_bar[2];              // the screen is committed in two halves (tile slots 0-1 / 2-4)
  __shared__ uint32_t s_tmem;
  __shared__ int s_fail;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bsm = smem;                       // B operand, two buffers (seed parity): KB blocks of [16 rows][128 B] each
  uint8_t* atiles = smem + 2 * KB * 2048;    // shared-memory A tiles: [tile][KB][128 rows][128 B]

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
    if (lane == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); s_fail = 0; s_next[0] = s_next[1] = -1; s_newlist = 0; }
  }
  for (int e = tid; e < 2 * KB * 2048 / 16; e += kThreads4) reinterpret_cast<uint4*>(bsm)[e] = make_uint4(0u, 0u, 0u, 0u);
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

  float r[kMaxSlots], nx[kMaxSlots];
#pragma unroll
  for (int s = 0; s < kMaxSlots; ++s) { r[s] = 0.f; nx[s] = 0.f; }
  // warp 1 (idle while warp 0 issues the MMAs of the screen), lane j: entry j of the candidate list (exact current key,
  // 0 = empty); `bound` is warp-uniform
  unsigned long long cand = 0ull, bound = ~0ull;
  int nex = 0;                                // exchanges so far (the same in every CTA: all take the same decisions)

  // one warp: fetch seed `idx` = seed number i (fp32), its norm, and the B operand {bf16(s), bf16(s - bf16(s))} of the
  // screen into the buffers of parity i & 1
  auto stage_seed = [&](long long idx, int i) {
    float* seedbuf = s_seed[i & 1];
    uint8_t* bbuf = bsm + (i & 1) * KB * 2048;
    float sq = 0.f;
#pragma unroll
    for (int h = 0; h < D / 64; ++h) {
      const int c0 = h * 64 + 2 * lane;
      const float v0 = __ldg(Xb + c0 * p.sd + idx), v1 = __ldg(Xb + (c0 + 1) * p.sd + idx);
      seedbuf[c0] = v0; seedbuf[c0 + 1] = v1;
      sq = fmaf(v0, v0, fmaf(v1, v1, sq));
      const float h0 = __bfloat162float(__float2bfloat16_rn(v0)), h1 = __bfloat162float(__float2bfloat16_rn(v1));
      const int chunk = lane >> 2, off = (lane & 3) * 4;          // channels 2*lane, 2*lane+1 of block h
      // row r (= column of B), chunk c of block h at h * 2048 + r * 128 + ((c ^ (r & 7)) << 4)
      *reinterpret_cast<uint32_t*>(bbuf + h * 2048 + 0 * 128 + ((chunk ^ 0) << 4) + off) = pack_bf16x2(h0, h1);
      *reinterpret_cast<uint32_t*>(bbuf + h * 2048 + 1 * 128 + ((chunk ^ 1) << 4) + off) = pack_bf16x2(v0 - h0, v1 - h1);
      if (rank == 0) {
        float* so = p.seeds_out + (size_t(b) * p.m + i) * D;
        so[c0] = v0; so[c0 + 1] = v1;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) {
      s_ns[i & 1] = sqrtf(sq) * 1.000001f;
      if (rank == 0) p.selected_out[size_t(b) * p.m + i] = idx;
    }
    fence_proxy_async();
  };
  if (warp == 0) stage_seed(p.first[b], 0);
  tc_fence_before();
  __syncthreads();

  constexpr uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
  const uint32_t bsm_addr = smem_u32(bsm), at_addr = smem_u32(atiles);

  for (int i = 0; i + 1 < p.m; ++i) {
    // ---- screen: all tiles of the CTA against seed i
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        // two halves, one commit each: the warps start on the tiles of slots 0-1 while the tensor core still works on the rest
        const uint32_t bcur = bsm_addr + uint32_t(i & 1) * KB * 2048;
        const int T0 = T < 8 ? T : 8;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int tb = half ? T0 : 0, te = half ? T : T0;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {         // k-step major: consecutive MMAs accumulate into different tiles
              const uint64_t bd = make_smem_desc_sw128(bcur + kb * 2048 + ks * 32, 16, 1024);
              const uint32_t accumulate = (kb | ks) ? 1u : 0u;
              for (int t = tb; t < te && t < TA; ++t)
                umma_ts_f16(tmem_base + dcol0 + 16u * t, tmem_base + acol0 + uint32_t(t) * ACOLS + kb * 32 + ks * 8, bd, idesc, accumulate);
              for (int t = (tb > TA ? tb : TA); t < te; ++t) {
                const uint64_t ad = make_smem_desc_sw128(at_addr + uint32_t((t - TA) * KB + kb) * 16384u + ks * 32, 16, 1024);
                umma_ss_f16(tmem_base + dcol0 + 16u * t, ad, bd, idesc, accumulate);
              }
            }
          }
          umma_commit(&s_bar[half]);
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ---- speculative chain (while warp 0 issues the screen): exact keys of the candidates after seed i
      if (s_newlist) {                            // the last pass ended with an exchange: take the rebuilt list
        const int cnt = s_cnt < 32 ? s_cnt : 32;
        cand = lane < cnt ? s_list[lane] : 0ull;
        bound = s_bound;
        __syncwarp();
        if (lane == 0) s_newlist = 0;
      }
      if (__any_sync(0xffffffffu, cand != 0ull)) {
        if (cand != 0ull) {
          const unsigned int idx = 0xFFFFFFFFu - static_cast<unsigned int>(cand & 0xFFFFFFFFull);
          float sq_unused;
          const float acc = exact_chain4<D>(Xb + idx, p.sd, s_seed[i & 1], false, sq_unused);
          const float dist = 0.5f * (1.0f - acc);
          const float rq = key_r(cand);
          cand = pack_key4(dist < rq ? dist : rq, idx);
        }
        unsigned long long best = cand;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
          best = other > best ? other : best;
        }
        if (best > bound) {                       // above everything that is not in the list: the global arg-max
          if (cand == best) cand = 0ull;          // consumed (keys are unique: they carry the point index)
          const long long nidx = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(best & 0xFFFFFFFFull));
          if (lane == 0) s_next[i & 1] = nidx;
          stage_seed(nidx, i + 1);                // into the other buffers, while this pass still runs
        } else {
          cand = 0ull;                            // the list is stale beyond use: rebuilt by the exchange
          if (lane == 0) s_next[i & 1] = -1;
        }
      } else if (lane == 0) {
        s_next[i & 1] = -1;
      }
    }
    const float ns = s_ns[i & 1];
    const float* seed = s_seed[i & 1];
    bool ok = true;
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s) {
      if (s == 0 || s == 2) {                 // slots 0-1 = tiles 0..7 (first commit), slots 2-4 = the rest (second commit)
        ok = mbar_wait(&s_bar[s == 0 ? 0 : 1], uint32_t(i) & 1u, p.err) && ok;
        tc_fence_after();
      }
      const int t = gi + 4 * s;
      if (t < T) {                            // warp-uniform
        uint32_t c0, c1;
        tmem_ld_32x32b_x2(lane_addr + dcol0 + 16u * uint32_t(t), c0, c1);
        tmem_wait_ld();
        const long long gp = tile_base(t) + q * 32 + lane;
        const bool valid = gp < p.n;
        bool need = valid;
        if (i > 0 && ok) {
          const float dapprox = 0.5f * (1.0f - (__uint_as_float(c0) + __uint_as_float(c1)));
          need = valid && !((dapprox - fmaf(p.rel_margin * nx[s], ns, kAbsMargin4)) >= r[s]);
        }
        if (__any_sync(0xffffffffu, need)) {
          if (need) {
            float sq;
            const float acc = exact_chain4<D>(Xb + gp, p.sd, seed, i == 0, sq);
            const float dist = 0.5f * (1.0f - acc);
            if (i == 0) {
              r[s] = dist;
              nx[s] = sqrtf(sq) * 1.000001f;
            } else {
              r[s] = dist < r[s] ? dist : r[s];
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();                          // everybody is done with the accumulators of this pass; s_next is final
    const long long next = s_next[i & 1];     // same decision in every CTA (deterministic)
    // next >= 0: speculative pass end -- no exchange, the CTAs stay uncoupled; warp 1 has staged the seed already
    if (next < 0) {
      // ---- exchange: every CTA publishes its two best keys; everybody rebuilds the candidate list
      // (four keys per CTA were measured: 66 % instead of 50 % of the passes stay speculative on the bench frame, but the
      // wider exchange costs more than that returns: 0.74 vs 0.61 ms, profiles/r02_fps_speculation.txt)
      unsigned long long k1 = 1ull, k2 = 1ull;          // non-zero sentinels: an empty slice still signals arrival
#pragma unroll
      for (int s = 0; s < kMaxSlots; ++s) {
        const int t = gi + 4 * s;
        const long long gp = tile_base(t) + q * 32 + lane;
        if (t < T && gp < p.n) merge_top2(k1, k2, pack_key4(r[s], static_cast<unsigned int>(gp)), 1ull);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        merge_top2(k1, k2, o1, o2);
      }
      if (lane == 0) { s_red[2 * warp] = k1; s_red[2 * warp + 1] = k2; }
      if (tid == 0) { s_cnt = 0; s_drop = 0ull; }
      __syncthreads();
      if (warp == 0) {
        k1 = (lane < kWarps4) ? s_red[2 * lane] : 1ull;
        k2 = (lane < kWarps4) ? s_red[2 * lane + 1] : 1ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
          merge_top2(k1, k2, o1, o2);
        }
        // all-to-all: this CTA's pair goes into column `rank` of EVERY CTA's private row; then poll the own row
        // The key matrices form a RING of two (exchange number parity), not one per pass: a CTA can be at most one
        // exchange ahead of another (it needs the other's keys of exchange c + 1 to get past it, and those are published
        // after the other has read exchange c), so exchange c + 2 may reuse the matrix of exchange c.  Arrival is a TAG:
        // bits 20..31 of a key's low word are all ones (index < 2^20), they carry the pass number on the wire.  700 KB
        // per field instead of 35 MB to clear per launch and to push through L2 next to the field.
        unsigned long long* mat = p.slots + (size_t(b) * 2 + (nex & 1)) * nb * nb * 2;
        const unsigned long long tag = static_cast<unsigned long long>((i + 1) & 0xFFF) << 20;
        const unsigned long long w1 = (k1 & ~0xFFF00000ull) | tag, w2 = (k2 & ~0xFFF00000ull) | tag;
#pragma unroll
        for (int c5 = 0; c5 < 5; ++c5) {
          const int c = lane + 32 * c5;
          if (c < nb) {
            unsigned long long* dst = mat + (size_t(c) * nb + rank) * 2;
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst), "l"(w1) : "memory");
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(w2) : "memory");
          }
        }
        const unsigned long long* row = mat + size_t(rank) * nb * 2;
        unsigned long long kv1[5], kv2[5];
        bool done = false;
        for (unsigned int it = 0; it < (1u << 22) && !done; ++it) {
          bool all = true;
#pragma unroll
          for (int c5 = 0; c5 < 5; ++c5) {
            const int c = lane + 32 * c5;
            kv1[c5] = tag; kv2[c5] = tag;
            if (c < nb) {
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(kv1[c5]) : "l"(row + 2 * c));
              asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(kv2[c5]) : "l"(row + 2 * c + 1));
            }
            all = all && ((kv1[c5] & 0xFFF00000ull) == tag) && ((kv2[c5] & 0xFFF00000ull) == tag);
          }
          done = __all_sync(0xffffffffu, all);
        }
#pragma unroll
        for (int c5 = 0; c5 < 5; ++c5) { kv1[c5] |= 0xFFF00000ull; kv2[c5] |= 0xFFF00000ull; }   // back to plain keys
        if (done && ok) {
          // bound = the largest SECOND key: every point outside the CTAs' first keys is below it
          unsigned long long bk = 0ull, gmax = 0ull;
#pragma unroll
          for (int c5 = 0; c5 < 5; ++c5) { bk = kv2[c5] > bk ? kv2[c5] : bk; gmax = kv1[c5] > gmax ? kv1[c5] : gmax; }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(0xffffffffu, bk, o), og = __shfl_xor_sync(0xffffffffu, gmax, o);
            bk = ob > bk ? ob : bk;
            gmax = og > gmax ? og : gmax;
          }
          // list = the first keys above the bound, except the winner itself (at most 32; the overflow raises the bound)
#pragma unroll
          for (int c5 = 0; c5 < 5; ++c5) {
            if (kv1[c5] > bk && kv1[c5] != gmax) {
              const int pos = atomicAdd(&s_cnt, 1);
              if (pos < 32) s_list[pos] = kv1[c5];
              else atomicMax(&s_drop, kv1[c5]);
            }
          }
          __syncwarp();
          if (lane == 0) { s_bound = s_drop > bk ? s_drop : bk; s_newlist = 1; }
          stage_seed(static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(gmax & 0xFFFFFFFFull)), i + 1);
        } else if (lane == 0) {
          s_fail = 1;
          atomicOr(p.err, ERR_GRID_BARRIER_TIMEOUT);
        }
      }
      tc_fence_before();
      __syncthreads();                        // the exchanged seed is staged
      ++nex;
    }
    if (p.passlog && blockIdx.x == 0 && tid == 0)
      p.passlog[i] = (next < 0 ? (1ull << 63) : 0ull) | (static_cast<unsigned long long>(clock64()) & ~(1ull << 63));
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
constexpr int kMaxSlots5 = 12;             // 32-point tiles per warp: n <= 148 * 16 * 32 * 12 = 909k points

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
  if (knobs().fps_tc == 0) return UOC_OK;
  const int sms = sm_count();
  if (sms <= 0 || s.batch > sms) return UOC_OK;
  // a batch of fields that do not fit on chip together: the fields take turns in the resident-slice kernel
  // (launch_select_seeds); UOC_FPS_BATCH_STREAM=1 streams them side by side instead (A/B knob: slower)
  if (s.batch > 1) {
    const bool side_by_side = knobs().fps_batch_stream != 0;   // measured (profiles/r02_fps_batch.txt): turns 0.72 ms / field, side by side 1.25 - 1.40
    const long long tiles_per_cta = ((s.n + 127) / 128 + (sms / s.batch) - 1) / (sms / s.batch);
    if (tiles_per_cta > 4 * kMaxSlots && !side_by_side) return UOC_OK;
  }
  const int nb = sms / s.batch;
  if (nb > 160) return UOC_OK;                      // the poll loop reads at most 5 keys per lane
  float rel_margin = kRelMargin4;
  if (knobs().fps_rn_margin != 0) rel_margin = 0.00205f;   // A/B: copy known to be round-to-nearest
  const long long total_tiles = (s.n + 127) / 128;
  const long long T = (total_tiles + nb - 1) / nb;  // tiles of the busiest CTA
  bool streaming = false;                           // the field does not fit on chip: stream the bf16 copy in every pass
  if (knobs().fps_stream != 0) streaming = true;            // test knob: force the streaming variant
  if (T > 4 * kMaxSlots) streaming = true;
  {
    const int acols0 = s.d / 2, kb0 = s.d / 64;
    int TA0 = int((512 - 16 * T) / acols0);
    if (TA0 > T) TA0 = int(T);
    if (TA0 < 0) TA0 = 0;
    if (1024 + 2 * size_t(kb0) * 2048 + size_t(T - TA0) * kb0 * 16384 > 225 * 1024) streaming = true;
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
    p5.passlog = nullptr;
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
  { const int v = knobs().fps_tmem_tiles; if (v >= 0 && v < TA) TA = v; }   // test / A-B knob
  const size_t smem = 1024 + 2 * size_t(kb) * 2048 + size_t(T - TA) * kb * 16384;
  if (smem > 225 * 1024) return UOC_OK;
  if (s.n > (1ll << 20) || s.m > 4095) return UOC_OK;                 // key tags (pass number in the index word's free bits)
  const size_t slot_need = size_t(s.batch) * 2 * nb * nb * 16;       // ring of two matrices of (first, second) keys per CTA pair
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
  p.passlog = knobs().fps_stats ? w.keys : nullptr;      // the key slots of the fp32 kernels are unused here
  UOC_CUDA(cudaMemsetAsync(w.slots, 0, slot_need, stream));
  void* args[] = {&p};
  UOC_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nb * s.batch), dim3(kThreads4), args, smem, stream));
  count_launch();
  if (p.passlog) {
    std::vector<unsigned long long> h(s.m);
    UOC_CUDA(cudaStreamSynchronize(stream));
    UOC_CUDA(cudaMemcpy(h.data(), p.passlog, sizeof(unsigned long long) * s.m, cudaMemcpyDeviceToHost));
    int nx = 0, nsp = 0;
    double cx = 0, csp = 0;
    for (int i = 2; i + 1 < s.m; ++i) {
      const bool ex = (h[i] >> 63) != 0;
      const double dt = double((h[i] & ~(1ull << 63)) - (h[i - 1] & ~(1ull << 63)));
      if (ex) { ++nx; cx += dt; } else { ++nsp; csp += dt; }
    }
    fprintf(stderr, "[fps stats] passes 2..%d of CTA 0: %d ended with an exchange (%.0f clk per pass), %d speculative (%.0f clk per pass)\n",
            s.m - 2, nx, nx ? cx / nx : 0.0, nsp, nsp ? csp / nsp : 0.0);
  }
  *used = true;
  return UOC_OK;
}

}  // namespace uoc
