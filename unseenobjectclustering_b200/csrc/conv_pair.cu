// K1 (second generation): convolution as an implicit GEMM on CTA PAIRS (tcgen05 cta_group::2), persistent, stream-K.
// Replaces the cuDNN convolutions + BatchNorm + ReLU + residual add of lib/networks/resnet.py:57-73, :236-270
// (BN folded into weights / bias on the host, see backbone.cu).  NHWC bf16 activations, fp32 accumulation.
//
// Why (measured on B200, profiles/r02_conv_layers_before.txt): the first generation ran one 128 x 128 tile per CTA, two
// CTAs per SM.  At 32 KB of operands per 64-channel K block it sits on the ~48 B/clk an SM can ingest through TMA (603
// clk per K block against 256 clk of MMA), its grids are 1.08 - 2.03 waves (half-empty last wave), and every tile pays
// TMEM allocation, barrier set-up, a cold pipeline and an epilogue nothing overlaps.
//
//   * CTA pair, M = 256 x N = BLOCK_N (64 / 128 / 256): each CTA stages its own 128 pixels of A and only HALF of the weight
//     rows; the tensor core reads the other half from the peer's shared memory: half the bytes per MAC at BLOCK_N = 256.
//   * persistent: <= one pair per TPC for the whole launch; the work is the list of (tile, K block) UNITS cut into equal
//     contiguous ranges (stream-K), so the grid is always exactly one "wave".  A tile cut by a range boundary is
//     finished WITHOUT ANY WAITING: every pair that holds a part of it writes its fp32 partial accumulator to a scratch
//     slot and bumps the tile's counter; whoever arrives LAST sums the parts (in part order: the result does not depend
//     on the arrival order) and runs the epilogue.  No CTA ever waits for another pair, so the kernel needs no
//     co-residency guarantee (frames in flight on other streams run cooperative kernels beside it).  Layers with few K
//     blocks (1x1) deal whole tiles.
//   * two TMEM accumulators: the epilogue of segment i (TMEM -> registers -> bias / residual / ReLU -> bf16 -> global)
//     overlaps the main loop of segment i + 1; 6 - 8 stage TMA ring.
//   warp 0: TMA producer (both CTAs; loads signal the LEADER's barrier), warp 1: MMA issuer (leader only; commits are
//   multicast to both CTAs), warps 2-5: epilogue of the own 128 rows.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv.cuh"

namespace uoc {

namespace {

constexpr int kThreads = 320;      // warp 0 producer, warp 1 MMA issuer, warps 2-9 epilogue (two per tensor-memory lane quadrant)
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 ch bf16

struct ConvPairParams {
  CUtensorMap tmap_x[2];
  CUtensorMap tmap_w[2];
  CUtensorMap tmap_y[2];  // bf16 output [N][Ho][Wo][Cout], box [64 ch][16][2][1] = one epilogue warp's 32 pixels, 128B swizzle
                          // (TMA store clips ragged tiles)
  CUtensorMap tmap_r[2];  // residual, same shape (TMA load, zero fill)
  const float* bias[2];
  const void* residual[2];
  void* y[2];
  int N, Ho, Wo, Cin, Cout;
  int tiles_x, tiles_y, m_pairs, n_tiles, groups;
  int ksize, stride, dil, pad;
  int relu, out_fp32;
  int KB;                 // K blocks per tile
  int tiles;              // m_pairs * n_tiles * groups
  int pairs;              // CTA pairs in the grid
  int streamk;            // 1: contiguous unit ranges (tiles may be split), 0: whole tiles dealt round-robin
  float* partials;        // [pairs][2 slots][2 ranks][128][BLOCK_N] fp32 scratch: parts of tiles cut by a range boundary
  unsigned int* counters; // [tiles][2 ranks] parts arrived (zero outside a launch)
  unsigned int* err;
  int dbg;                // measurement knobs (UOC_CONV_DEBUG): 1 no A loads, 2 no B loads, 4 no MMA, 8 no epilogue stores
  long long* trace;       // [16] clock sums of pair 0 (UOC_CONV_TRACE): see launch_conv_pair
};

template <int BLOCK_N>
struct PairCfg {
  static constexpr int kBBytes = (BLOCK_N / 2) * 128;       // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kChunks = BLOCK_N / 64;              // 64-channel output chunks of a tile: one 16 KB staging buffer each
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 7 : 8);
  static constexpr int kOutBytes = kChunks * 16384;
  static constexpr int kBiasBytes = 8 * BLOCK_N * 4;        // one copy of the tile's bias per epilogue warp
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kOutBytes + 256 + kBiasBytes;   // + 16 B static: <= 227 KB
  static constexpr uint32_t kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;   // two accumulators
};

struct Segment {
  int tile, kb0, kb1;
};

// The i-th segment of pair `pair` (false when the pair's work is exhausted).  Every role walks the same list.
struct SegmentWalker {
  long long u, u1;        // stream-K: next unit / end of the range
  int t, tiles, step, KB; // whole tiles: next tile, stride
  bool streamk;
  __device__ SegmentWalker(const ConvPairParams& p, int pair) {
    KB = p.KB; tiles = p.tiles; step = p.pairs; streamk = p.streamk != 0;
    const long long total = (long long)p.tiles * p.KB;
    u = total * pair / p.pairs;
    u1 = total * (pair + 1) / p.pairs;
    t = pair;
  }
  __device__ bool next(Segment* s) {
    if (streamk) {
      if (u >= u1) return false;
      s->tile = int(u / KB);
      s->kb0 = int(u - (long long)s->tile * KB);
      const long long left = u1 - u;
      s->kb1 = (KB - s->kb0 < left) ? KB : int(s->kb0 + left);
      u += s->kb1 - s->kb0;
      return true;
    }
    if (t >= tiles) return false;
    s->tile = t; s->kb0 = 0; s->kb1 = KB;
    t += step;
    return true;
  }
};

__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// range of pair q = [total * q / P, total * (q + 1) / P);  pair_of(u) = the pair whose range holds unit u
__device__ __forceinline__ long long range_start(long long total, int P, int q) { return total * q / P; }
__device__ __forceinline__ int pair_of(long long total, int P, long long u) { return int(((u + 1) * P + total - 1) / total) - 1; }

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
conv_pair_kernel(const __grid_constant__ ConvPairParams p) {
  using Cfg = PairCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* outbuf = smem + Cfg::kStages * Cfg::kStageBytes;     // [chunks][128 rows][128 B] output staging (128B swizzle)
  uint64_t* bars = reinterpret_cast<uint64_t*>(outbuf + Cfg::kOutBytes);
  uint64_t* full = bars;                               // [stages]  leader: TMA bytes of BOTH CTAs landed
  uint64_t* empty = bars + Cfg::kStages;               // [stages]  both: the MMAs that read the stage have completed
  uint64_t* acc_full = bars + 2 * Cfg::kStages;        // [2]       both: accumulator a is complete
  uint64_t* acc_empty = acc_full + 2;                  // [2]       leader: both CTAs' epilogues have drained accumulator a
  uint64_t* res_full = acc_empty + 2;                  // [chunks]  own: the residual chunk has landed in its staging buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 4);
  float* s_bias = reinterpret_cast<float*>(bars + 32);  // [8 warps][BLOCK_N] bias of the current tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = int(cluster_ctarank());               // 0 = leader
  const int pair = int(blockIdx.x) >> 1;
  const int cblocks = p.Cin >> 6;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x[0]);
    tma_prefetch_desc(&p.tmap_w[0]);
    if (p.groups > 1) { tma_prefetch_desc(&p.tmap_x[1]); tma_prefetch_desc(&p.tmap_w[1]); }
    tma_prefetch_desc(&p.tmap_y[0]);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 16); }   // 8 epilogue warps x 2 CTAs
    for (int c = 0; c < 4; ++c) mbar_init(&res_full[c], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // both CTAs' barriers are initialised and both allocations done before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above overlapped the previous layer's tail; its activations (and the
  // stream-K counters it reset) are complete and visible after the wait.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // tile index -> (group, n tile, M pair): the M pair runs fastest so that neighbouring pairs share a weight tile in L2
  auto decode = [&](int tile, int* g, int* n0, int* tx, int* ty, int* img) {
    const int mp = tile % p.m_pairs;
    int r = tile / p.m_pairs;
    *n0 = (r % p.n_tiles) * BLOCK_N;
    *g = r / p.n_tiles;
    int m_tile = mp * 2 + rank;                          // may exceed the real tile count (padding CTA): zero A, no stores
    *tx = m_tile % p.tiles_x; m_tile /= p.tiles_x;
    *ty = m_tile % p.tiles_y;
    *img = m_tile / p.tiles_y;
  };

  if (warp == 0) {
    if (elect_one()) {
      SegmentWalker walk(p, pair);
      Segment sg;
      uint32_t it = 0;                                   // K blocks issued so far (ring position)
      bool ok = true;
      const bool trc = p.trace && blockIdx.x == 0;
      long long t_wait = 0, t0 = clock64();
      const uint32_t stage_tx = ((p.dbg & 1) ? 0u : uint32_t(kABytes)) + ((p.dbg & 2) ? 0u : uint32_t(Cfg::kBBytes));
      while (ok && walk.next(&sg)) {
        int g, n0, tx, ty, img;
        decode(sg.tile, &g, &n0, &tx, &ty, &img);
        const int x_base = tx * 16 * p.stride - p.pad;
        const int y_base = ty * 8 * p.stride - p.pad;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++it) {
          const int s = int(it % Cfg::kStages);
          const long long w0 = trc ? clock64() : 0;
          if (!mbar_wait(&empty[s], ((it / Cfg::kStages) & 1) ^ 1u, p.err)) { ok = false; break; }
          if (trc) t_wait += clock64() - w0;
          const int tap = kb / cblocks, cb = kb - tap * cblocks;
          const int r = tap / p.ksize, sx = tap - r * p.ksize;
          uint8_t* st = smem + s * Cfg::kStageBytes;
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * stage_tx);     // the bytes of BOTH CTAs
          const uint32_t lbar = leader_bar_addr(&full[s]);
          if (!(p.dbg & 1)) tma_load_4d_2sm(st, &p.tmap_x[g], lbar, cb * 64, x_base + sx * p.dil, y_base + r * p.dil, img);
          if (!(p.dbg & 2)) tma_load_2d_2sm(st + kABytes, &p.tmap_w[g], lbar, kb * 64, n0 + rank * (BLOCK_N / 2));
        }
      }
      if (trc) { p.trace[0] = t_wait; p.trace[1] = clock64() - t0; p.trace[2] = it; }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, 0, 0);
      SegmentWalker walk(p, pair);
      Segment sg;
      uint32_t it = 0, seg = 0;
      bool ok = true;
      const bool trc = p.trace && blockIdx.x == 0;
      long long t_full = 0, t_acc = 0, t0 = clock64();
      while (ok && walk.next(&sg)) {
        const uint32_t a = seg & 1u;
        const long long w1 = trc ? clock64() : 0;
        if (!mbar_wait(&acc_empty[a], ((seg >> 1) & 1u) ^ 1u, p.err)) break;       // epilogue of segment seg - 2 has drained it
        if (trc) t_acc += clock64() - w1;
        tc_fence_after();
        const uint32_t tacc = tmem_base + a * BLOCK_N;
        for (int kb = sg.kb0; kb < sg.kb1; ++kb, ++it) {
          const int s = int(it % Cfg::kStages);
          const long long w0 = trc ? clock64() : 0;
          if (!mbar_wait(&full[s], (it / Cfg::kStages) & 1, p.err)) { ok = false; break; }
          if (trc) t_full += clock64() - w0;
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + ks * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + ks * 32, 16, 1024);
            if (!(p.dbg & 4)) umma_ss_f16_2sm(tacc, ad, bd, idesc, (kb > sg.kb0 || ks > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty[s], 3);
        }
        if (ok) umma_commit_2sm(&acc_full[a], 3);
        ++seg;
      }
      if (trc) { p.trace[3] = t_full; p.trace[4] = t_acc; p.trace[5] = clock64() - t0; p.trace[6] = seg; }
    }
  } else {
    // Eight epilogue warps: the epilogue arithmetic is ISSUE bound (one warp per scheduler issues an fp32 instruction
    // every other clock: ~800 clk per 64-column chunk, profiles/r02_conv_trace.txt), so every tensor-memory lane quadrant
    // is served by two warps that take the 64-column chunks of a tile in turn.
    const int q = warp & 3;                              // TMEM lane quadrant of this warp
    const int half = (warp - 2) >> 2;                    // takes the chunks c with (c & 1) == half
    const int row = q * 32 + lane;
    const bool leader_thread = (warp == 2 && lane == 0); // issues the residual loads, updates the stream-K counters
    const uint32_t acc_empty_leader[2] = {map_to_cta(smem_u32(&acc_empty[0]), 0), map_to_cta(smem_u32(&acc_empty[1]), 0)};
    __shared__ int s_last;
    SegmentWalker walk(p, pair);
    Segment sg;
    uint32_t seg = 0, fin = 0;                           // segments seen / tiles finished (phase of the residual barriers)
    const long long total_units = (long long)p.tiles * p.KB;
    const bool trc = p.trace && blockIdx.x == 0 && warp == 2 && lane == 0;
    long long t_epi_wait = 0, t_split = 0, t_fin = 0, t0 = clock64();
    // partial accumulators in scratch: float4 index ((c * 8 + e) * 128 + row) -- consecutive lanes, consecutive 16 bytes
    auto part_ptr = [&](int qp, long long tile_u0) {
      const int which = (range_start(total_units, p.pairs, qp) >= tile_u0) ? 0 : 1;
      return reinterpret_cast<float4*>(p.partials + ((size_t(qp) * 2 + which) * 2 + rank) * 128 * BLOCK_N) + row;
    };
    while (walk.next(&sg)) {
      const uint32_t a = seg & 1u;
      int g, n0, tx, ty, img;
      decode(sg.tile, &g, &n0, &tx, &ty, &img);
      const bool split = sg.kb0 > 0 || sg.kb1 < p.KB;    // another pair holds the rest of this tile
      // parts of a split tile: the consecutive pairs whose ranges intersect it; slot 0 of a pair = a tile that its range
      // starts in, slot 1 = a tile that began in an earlier pair's range
      const long long tile_u0 = (long long)sg.tile * p.KB;
      const int first_part = split ? pair_of(total_units, p.pairs, tile_u0) : pair;
      const int nparts = split ? pair_of(total_units, p.pairs, tile_u0 + p.KB - 1) - first_part + 1 : 1;
      const long long w0 = trc ? clock64() : 0;
      if (!mbar_wait(&acc_full[a], (seg >> 1) & 1u, p.err)) break;
      const long long w1 = trc ? clock64() : 0;
      if (trc) t_epi_wait += w1 - w0;
      tc_fence_after();
      const uint32_t ta = tmem_base + a * BLOCK_N + (uint32_t(q * 32) << 16);
      bool finish = !split;                              // whole tile: the epilogue runs straight from tensor memory
      bool direct = !split;                              // this CTA's own part is still in tensor memory while it finishes
      const int me = pair - first_part;                  // index of the own part
      if (split) {
        // Has every other part arrived already?  (the usual case for the pair that holds the tile's FIRST K blocks: they
        // are the end of its range, while the other parts are the beginning of their pairs' ranges).  Then this CTA
        // finishes the tile straight from tensor memory plus the others' partials, without publishing its own part.
        if (leader_thread) {
          unsigned int cnt;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cnt) : "l"(p.counters + size_t(sg.tile) * 2 + rank) : "memory");
          const int all_there = (cnt == (unsigned int)(nparts - 1)) ? 1 : 0;
          if (all_there) p.counters[size_t(sg.tile) * 2 + rank] = 0u;   // nobody else touches it any more: zero for the next launch
          s_last = all_there ? 2 : 0;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        direct = s_last == 2;
        finish = direct;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // s_last is re-written below
      }
      if (split && !direct) {
        float4* mine = part_ptr(pair, tile_u0);
#pragma unroll 1
        for (int c = half; c < BLOCK_N / 64; c += 2) {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32b_x32(ta + c * 64, v0);
          tmem_ld_32x32b_x32(ta + c * 64 + 32, v1);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e)
            __stcg(mine + size_t((2 * c) * 8 + e) * 128, make_float4(__uint_as_float(v0[4 * e]), __uint_as_float(v0[4 * e + 1]),
                                                                       __uint_as_float(v0[4 * e + 2]), __uint_as_float(v0[4 * e + 3])));
#pragma unroll
          for (int e = 0; e < 8; ++e)
            __stcg(mine + size_t((2 * c + 1) * 8 + e) * 128, make_float4(__uint_as_float(v1[4 * e]), __uint_as_float(v1[4 * e + 1]),
                                                                           __uint_as_float(v1[4 * e + 2]), __uint_as_float(v1[4 * e + 3])));
        }
        // the accumulator is drained: hand it back to the MMA issuer before the (possibly long) finishing pass
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader[a]);
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps: all partial rows are written ...
        if (leader_thread) {
          __threadfence();                               // ... and (cumulatively) ordered before the counter update
          const unsigned int old = atomicAdd(p.counters + size_t(sg.tile) * 2 + rank, 1u);
          __threadfence();
          const int last = (old == (unsigned int)(nparts - 1)) ? 1 : 0;
          if (last) p.counters[size_t(sg.tile) * 2 + rank] = 0u;
          s_last = last;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        finish = s_last != 0;
        asm volatile("bar.sync 1, 256;" ::: "memory");   // s_last is re-written by the next split segment
      }
      const long long w2 = trc ? clock64() : 0;
      if (trc) t_split += w2 - w1;
      if (finish) {
        const bool has_res = p.residual[g] != nullptr && !p.out_fp32;
        const int x0 = tx * 16, y0 = ty * 8;
        // this warp's staging blocks are free again once the TMA engine has read its previous stores (bulk groups are per
        // thread: lane 0 of every warp issues and tracks its own); the warp's own copy of the tile's bias
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        float* wbias = s_bias + (warp - 2) * BLOCK_N;
        for (int e = lane; e < BLOCK_N; e += 32) wbias[e] = __ldg(p.bias[g] + n0 + e);
        __syncwarp();
        if (has_res) {
          asm volatile("bar.sync 1, 256;" ::: "memory");  // every warp's blocks are free before the residual lands in them
          if (leader_thread) {
#pragma unroll
            for (int c = 0; c < Cfg::kChunks; ++c) {
              mbar_arrive_expect_tx(&res_full[c], 16384);
              tma_load_4d(outbuf + c * 16384, &p.tmap_r[g], &res_full[c], n0 + c * 64, x0, y0, img);
            }
          }
        }
#pragma unroll 1
        for (int c = half; c < Cfg::kChunks; c += 2) {
          // own pixel row of the staging buffer: 16-byte unit j of row r sits at ((j ^ (r & 7)) << 4) (128B swizzle)
          uint8_t* myrow = outbuf + c * 16384 + row * 128;
          if (has_res) mbar_wait(&res_full[c], fin & 1u, p.err);
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {                   // 32 columns at a time (register budget of 10 warps)
            float f[32];
            uint32_t v[32];
            if (direct) {                                 // tcgen05.ld is warp-aligned: every lane executes it
              tmem_ld_32x32b_x32(ta + c * 64 + h * 32, v);
              tmem_wait_ld();
            }
            if (!split) {
#pragma unroll
              for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
            } else {
              // parts summed in part order (0 + part 0 + part 1 ...): the bits do not depend on who finishes, nor on
              // whether the own part comes from tensor memory or from scratch
#pragma unroll
              for (int e = 0; e < 32; ++e) f[e] = 0.f;
              for (int j = 0; j < nparts; ++j) {
                if (direct && j == me) {
#pragma unroll
                  for (int e = 0; e < 32; ++e) f[e] += __uint_as_float(v[e]);
                  continue;
                }
                const float4* pp = part_ptr(first_part + j, tile_u0) + size_t(2 * c + h) * 8 * 128;
                float4 pv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) pv[e] = __ldcg(pp + size_t(e) * 128);
#pragma unroll
                for (int e = 0; e < 8; ++e) { f[4 * e + 0] += pv[e].x; f[4 * e + 1] += pv[e].y; f[4 * e + 2] += pv[e].z; f[4 * e + 3] += pv[e].w; }
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 bv = *reinterpret_cast<const float4*>(wbias + c * 64 + h * 32 + 4 * e);      // broadcast
              f[4 * e + 0] += bv.x; f[4 * e + 1] += bv.y; f[4 * e + 2] += bv.z; f[4 * e + 3] += bv.w;
            }
            if (p.out_fp32) {
              // fp32 output (the 1x1 `fc` layer only): direct stores of the own pixel row
              const int oy = y0 + (row >> 4), ox = x0 + (row & 15);
              if ((oy < p.Ho) && (ox < p.Wo) && (img < p.N) && !(p.dbg & 8)) {
                const size_t pix = (size_t(img) * p.Ho + oy) * p.Wo + ox;
                if (p.relu) {
#pragma unroll
                  for (int e = 0; e < 32; ++e) f[e] = fmaxf(f[e], 0.f);
                }
                float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.y[g]) + pix * p.Cout + n0 + c * 64 + h * 32);
#pragma unroll
                for (int e = 0; e < 8; ++e) op[e] = make_float4(f[4 * e], f[4 * e + 1], f[4 * e + 2], f[4 * e + 3]);
              }
              continue;
            }
            if (has_res) {
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
          if (p.out_fp32) continue;
          fence_proxy_async();                            // generic-proxy writes -> visible to the TMA engine
          __syncwarp();
          if (lane == 0 && !(p.dbg & 8)) {                // this warp's 32 pixels (2 rows of 16) x 64 channels
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&p.tmap_y[g])), "r"(smem_u32(outbuf + c * 16384 + q * 4096)),
                           "r"(n0 + c * 64), "r"(x0), "r"(y0 + 2 * q), "r"(img)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        __syncwarp();                                     // wbias is rewritten by the next tile
        ++fin;
      }
      if (trc) t_fin += clock64() - w2;
      if (direct) {
        // the accumulator is drained: hand it back to the MMA issuer (one arrival per warp, on the LEADER's barrier)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader[a]);
      }
      ++seg;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // this warp's output stores have completed
    if (trc) { p.trace[7] = t_epi_wait; p.trace[8] = t_split; p.trace[9] = t_fin; p.trace[10] = clock64() - t0; p.trace[15] = fin; }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();        // the peer may still be reading this CTA's weight half / arriving on its barriers
  if (warp == 1) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

template <int BLOCK_N>
int launch_pair(const ConvPairParams& prm, cudaStream_t stream) {
  using Cfg = PairCfg<BLOCK_N>;
  int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(&conv_pair_kernel<BLOCK_N>), Cfg::kSmemBytes);
  if (rc != UOC_OK) return rc;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(prm.pairs * 2, 1, 1);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute lattr[2];
  lattr[0].id = cudaLaunchAttributeClusterDimension;
  lattr[0].val.clusterDim.x = 2;
  lattr[0].val.clusterDim.y = 1;
  lattr[0].val.clusterDim.z = 1;
  lattr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue overlaps the previous layer's tail
  lattr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = 2;
  UOC_CUDA(cudaLaunchKernelEx(&cfg, conv_pair_kernel<BLOCK_N>, prm));
  count_launch();
  return UOC_OK;
}

}  // namespace

size_t conv_pair_scratch_bytes() {
  // the counters, then [pairs][2 slots][2 CTAs][128 rows][256 columns] fp32 partials for the widest tile
  return kConvCounterBytes + size_t(kConvMaxPairs) * 2 * 2 * 128 * 256 * 4;
}

bool conv_pair_supported(const ConvProblem& p) {
  return p.Cin % 64 == 0 && p.Cout % 64 == 0 && (p.ksize == 1 || p.ksize == 3) && (p.stride == 1 || p.stride == 2) &&
         p.groups >= 1 && p.groups <= 2;
}

int launch_conv_pair(const ConvProblem& p, void* scratch, size_t scratch_bytes, cudaStream_t stream) {
  if (!conv_pair_supported(p)) return fail(UOC_ERR_UNSUPPORTED, "conv_pair: Cin, Cout multiples of 64; 1x1 or 3x3; stride 1 or 2");
  if (!scratch || scratch_bytes < conv_pair_scratch_bytes()) return fail(UOC_ERR_WORKSPACE, "conv_pair: stream-K scratch too small");
  ConvPairParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.Ho = conv_out_dim(p.H, p.ksize, p.stride, p.dilation);
  prm.Wo = conv_out_dim(p.W, p.ksize, p.stride, p.dilation);
  prm.Cin = p.Cin; prm.Cout = p.Cout; prm.N = p.N; prm.groups = p.groups;
  prm.tiles_x = (prm.Wo + 15) / 16;
  prm.tiles_y = (prm.Ho + 7) / 8;
  const int block_n = (p.Cout % 256 == 0) ? 256 : (p.Cout % 128 == 0 ? 128 : 64);
  prm.n_tiles = p.Cout / block_n;
  prm.ksize = p.ksize; prm.stride = p.stride; prm.dil = p.dilation; prm.pad = (p.ksize == 3) ? p.dilation : 0;
  prm.relu = p.relu; prm.out_fp32 = p.out_fp32;
  prm.err = device_error_word();
  if (!prm.err) return fail(UOC_ERR_CUDA, "no device error word");
  const int taps = p.ksize * p.ksize;
  prm.KB = taps * (p.Cin / 64);
  const int m_tiles = p.N * prm.tiles_y * prm.tiles_x;
  prm.m_pairs = (m_tiles + 1) / 2;
  prm.tiles = prm.m_pairs * prm.n_tiles * p.groups;
  int max_pairs = sm_count() / 2;
  if (max_pairs > kConvMaxPairs) max_pairs = kConvMaxPairs;
  if (max_pairs < 1) return fail(UOC_ERR_CUDA, "no SM count");
  prm.pairs = prm.tiles < max_pairs ? prm.tiles : max_pairs;
  // stream-K when splitting a tile costs less than the idle tail it removes: tiles with >= 8 K blocks (a contribution
  // writes and re-reads 128 x BLOCK_N fp32, about two K blocks' worth of operand bytes), and an uneven tile count
  prm.streamk = (prm.KB >= 8 && prm.tiles % max_pairs != 0) ? 1 : 0;
  if (prm.streamk) {
    // never cut ranges shorter than 4 K blocks
    const long long units = (long long)prm.tiles * prm.KB;
    long long pr = units / 4;
    if (pr < 1) pr = 1;
    prm.pairs = int(pr < max_pairs ? pr : max_pairs);
  }
  char* sc = static_cast<char*>(scratch);
  prm.counters = reinterpret_cast<unsigned int*>(sc);
  prm.partials = reinterpret_cast<float*>(sc + kConvCounterBytes);
  if (prm.streamk && size_t(prm.tiles) * 2 * sizeof(unsigned int) > kConvCounterBytes) prm.streamk = 0;   // (huge batches) deal whole tiles
  for (int g = 0; g < p.groups; ++g) {
    const uint64_t xd[4] = {uint64_t(p.Cin), uint64_t(p.W), uint64_t(p.H), uint64_t(p.N)};
    const uint64_t xs[3] = {uint64_t(p.Cin) * 2, uint64_t(p.W) * p.Cin * 2, uint64_t(p.H) * p.W * p.Cin * 2};
    const uint32_t xb[4] = {64, uint32_t(16 * p.stride), uint32_t(8 * p.stride), 1};
    const uint32_t xe[4] = {1, uint32_t(p.stride), uint32_t(p.stride), 1};
    int rc = make_tmap_bf16(&prm.tmap_x[g], p.g[g].x, 4, xd, xs, xb, xe);
    if (rc != UOC_OK) return rc;
    const uint64_t wd[2] = {uint64_t(taps) * p.Cin, uint64_t(p.Cout)};
    const uint64_t wsb[1] = {uint64_t(taps) * p.Cin * 2};
    const uint32_t wb[2] = {64, uint32_t(block_n / 2)};            // each CTA of the pair fetches one half of the rows
    rc = make_tmap_bf16(&prm.tmap_w[g], p.g[g].w, 2, wd, wsb, wb, nullptr);
    if (rc != UOC_OK) return rc;
    prm.bias[g] = p.g[g].bias;
    prm.residual[g] = p.g[g].residual;
    prm.y[g] = p.g[g].y;
    if (!p.out_fp32) {
      const uint64_t yd[4] = {uint64_t(p.Cout), uint64_t(prm.Wo), uint64_t(prm.Ho), uint64_t(p.N)};
      const uint64_t ys[3] = {uint64_t(p.Cout) * 2, uint64_t(prm.Wo) * p.Cout * 2, uint64_t(prm.Ho) * prm.Wo * p.Cout * 2};
      const uint32_t yb[4] = {64, 16, 2, 1};          // one epilogue warp's rows per store
      const uint32_t rb[4] = {64, 16, 8, 1};          // the whole 128-pixel tile per residual load
      rc = make_tmap_bf16(&prm.tmap_y[g], p.g[g].y, 4, yd, ys, yb, nullptr);
      if (rc != UOC_OK) return rc;
      if (p.g[g].residual) {
        rc = make_tmap_bf16(&prm.tmap_r[g], p.g[g].residual, 4, yd, ys, rb, nullptr);
        if (rc != UOC_OK) return rc;
      }
    }
  }
  prm.dbg = 0;
  prm.trace = nullptr;
  prm.dbg = knobs().conv_debug;
  const bool want_trace = knobs().conv_trace != 0;
  if (want_trace) {
    UOC_CUDA(cudaMalloc(&prm.trace, 16 * sizeof(long long)));
    UOC_CUDA(cudaMemsetAsync(prm.trace, 0, 16 * sizeof(long long), stream));
  }
  int rc = (block_n == 256) ? launch_pair<256>(prm, stream) : (block_n == 128 ? launch_pair<128>(prm, stream) : launch_pair<64>(prm, stream));
  if (want_trace && rc == UOC_OK) {
    long long h[16];
    UOC_CUDA(cudaStreamSynchronize(stream));
    UOC_CUDA(cudaMemcpy(h, prm.trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(prm.trace);
    fprintf(stderr, "[conv trace] Cin %d Cout %d k %d N %d: tiles %d KB %d pairs %d streamk %d block_n %d | pair 0: producer %lld units, "
            "waits for a free stage %lld of %lld clk | MMA issuer: waits for data %lld, for a free accumulator %lld of %lld clk, %lld segments | "
            "epilogue: waits for an accumulator %lld, split hand-over %lld, finishing %lld of %lld clk (%lld tiles finished)\n",
            p.Cin, p.Cout, p.ksize, p.N, prm.tiles, prm.KB, prm.pairs, prm.streamk, block_n, h[2], h[0], h[1], h[3], h[4], h[5], h[6],
            h[7], h[8], h[9], h[10], h[15]);
  }
  return rc;
}

}  // namespace uoc
