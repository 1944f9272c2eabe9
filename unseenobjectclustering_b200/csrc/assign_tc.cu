// K5b on the tensor cores: nearest-seed labels with an exactness certificate.
// (lib/utils/mean_shift.py:206-215: labels = seed_labels[argmin_j 0.5 (1 - x . z_j)].)
//
// The output must equal the canonical fp32 arg-min bit for bit, but a pixel's LABEL only depends on which
// cluster of seeds wins, and clusters are separated by far more than bf16 rounding.  So:
//   pass 1 (this kernel): S^T[128 points x 128 seeds] = Xtile . Zs^T on tcgen05 (bf16 operands, fp32 accumulate) from the
//           bf16 pixel-major copy of X; every point finds the best dot v1 (label l1) and the best dot v2 among seeds whose
//           label differs from l1.  The seed COLUMNS are sorted by label first, so a label is a run of consecutive columns:
//           per run one max-reduction straight from tensor memory (runtime column address), then a top-2 over the runs --
//           no label logic per element (the element-wise sweep was ~14 instructions per seed and point: 44 us per field).  |bf16 dot - canonical fp32 dot| <= E = 2^-8 (|x||z| = 1)
//           plus accumulation noise, so v1 - v2 > 2E + slack PROVES that the canonical arg-min carries label l1.
//   pass 2 (assign_fix_kernel): the few uncertified points (cluster borders) are recomputed with the canonical fp32 chain.
// Result: identical labels to the fp32 kernel (tests/test_gpu_clustering.py::test_assign_bit_exact), one pass over
// 39 MB of bf16 instead of 3.9 GFLOP of fp32 FMA.
#include <cstring>

#include "cluster.cuh"

namespace uoc {

namespace {

constexpr int kTile = 128;
constexpr int kThreads = 320;           // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9: two epilogue groups (even / odd tiles)
constexpr int kBoxBytes = 128 * 128;
constexpr float kCertGap = 0.0085f;     // > 2 * (2^-8 + accumulation slack)

template <int D>
struct AsCfg {
  static constexpr int kKBlocks = D / 64;
  static constexpr int kStageBytes = kKBlocks * kBoxBytes;
  static constexpr int kStages = (D == 64) ? 6 : 4;
  static constexpr int kZBytes = kKBlocks * kBoxBytes;
  static constexpr int kSmemBytes = 1024 + kZBytes + kStages * kStageBytes + 256;
  static constexpr uint32_t kTmemCols = 256;
};

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const float* __restrict__ Z, const int* __restrict__ seed_labels,
                 int m, long long n, int P, int* __restrict__ labels_tmp, int* __restrict__ hist,
                 unsigned int* __restrict__ uncertain_count, int* __restrict__ uncertain_list, unsigned int* err) {
  using Cfg = AsCfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* zs = smem;
  uint8_t* stages = smem + Cfg::kZBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* x_full = bars;
  uint64_t* x_empty = bars + Cfg::kStages;
  uint64_t* s_full = bars + 2 * Cfg::kStages;
  uint64_t* s_empty = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + 2);
  __shared__ int s_lab[UOC_MAX_SEEDS];         // label of seed i
  __shared__ int s_hist[UOC_MAX_SEEDS];
  __shared__ int s_perm[UOC_MAX_SEEDS];        // column (sorted position) -> seed
  __shared__ int s_sorted[UOC_MAX_SEEDS];      // label of column
  __shared__ int s_run_start[UOC_MAX_SEEDS + 1];   // first column of run r; [nruns] = m
  __shared__ int s_run_lab[UOC_MAX_SEEDS];
  __shared__ int s_nruns;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, b = blockIdx.y;
  const long long tiles_total = (n + kTile - 1) / kTile;
  const int T = (cta < tiles_total) ? int((tiles_total - cta + P - 1) / P) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(&s_empty[0], 128); mbar_init(&s_empty[1], 128);
    fence_mbar_init();
  }
  // (8-column loads of a run that ends near column 128 may read up to 7 columns past the buffer: more columns then)
  const uint32_t tmem_cols = (m > 120) ? 512u : Cfg::kTmemCols;
  if (warp == 1) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  if (threadIdx.x < UOC_MAX_SEEDS) {
    s_lab[threadIdx.x] = threadIdx.x < m ? seed_labels[size_t(b) * m + threadIdx.x] : -1;
    s_hist[threadIdx.x] = 0;
  }
  __syncthreads();
  if (warp >= 2 && warp < 6) {
    // stable sort of the seeds by label: column of seed t = number of seeds that sort before it
    const int t = threadIdx.x - 64;
    if (t < m) {
      const int lt = s_lab[t];
      int pos = 0;
      for (int i = 0; i < m; ++i) {
        const int li = s_lab[i];
        pos += (li < lt || (li == lt && i < t)) ? 1 : 0;
      }
      s_perm[pos] = t;
      s_sorted[pos] = lt;
    }
  }
  __syncthreads();
  if (warp == 2) {
    // runs of equal labels
    int count = 0;
    for (int base = 0; base < UOC_MAX_SEEDS; base += 32) {
      const int pos = base + lane;
      const bool st = pos < m && (pos == 0 || s_sorted[pos] != s_sorted[pos - 1]);
      const unsigned int bal = __ballot_sync(0xffffffffu, st);
      if (st) {
        const int r = count + __popc(bal & ((1u << lane) - 1u));
        s_run_start[r] = pos;
        s_run_lab[r] = s_sorted[pos];
      }
      count += __popc(bal);
    }
    if (lane == 0) { s_nruns = count; s_run_start[count] = m; }
  }
  if (warp >= 2) {
    const int r = (threadIdx.x - 64) & 127;
    const int half = (threadIdx.x - 64) >> 7;
    const float* zr = Z + (size_t(b) * m + (r < m ? s_perm[r] : 0)) * D;     // column r holds seed s_perm[r]
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

  if (warp == 0) {
    if (elect_one()) {
      for (int j = 0; j < T; ++j) {
        const int s = j % Cfg::kStages;
        if (!mbar_wait(&x_empty[s], ((j / Cfg::kStages) & 1) ^ 1u, err)) break;
        mbar_arrive_expect_tx(&x_full[s], Cfg::kStageBytes);
        const int row0 = int((cta + (long long)j * P) * kTile);
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb)
          tma_load_3d(stages + s * Cfg::kStageBytes + kb * kBoxBytes, &tmap_x, &x_full[s], kb * 64, row0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t zs_addr = smem_u32(zs), st_addr = smem_u32(stages);
      for (int j = 0; j < T; ++j) {
        const int s = j % Cfg::kStages, buf = j & 1;
        if (!mbar_wait(&s_empty[buf], ((j >> 1) & 1) ^ 1u, err)) break;     // epilogue drained this TMEM buffer
        if (!mbar_wait(&x_full[s], (j / Cfg::kStages) & 1, err)) break;
        tc_fence_after();
        const uint32_t xa = st_addr + s * Cfg::kStageBytes;
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = make_smem_desc_sw128(xa + kb * kBoxBytes + ks * 32, 16, 1024);        // A = points
            const uint64_t bd = make_smem_desc_sw128(zs_addr + kb * kBoxBytes + ks * 32, 16, 1024);  // B = seeds
            umma_ss_f16(tmem_base + buf * 128, ad, bd, idesc, (kb | ks) ? 1u : 0u);
          }
        }
        umma_commit(&x_empty[s]);
        umma_commit(&s_full[buf]);
      }
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    for (int j = grp; j < T; j += 2) {
      if (!mbar_wait(&s_full[grp], (j >> 1) & 1, err)) break;
      tc_fence_after();
      float v1 = -INFINITY, v2 = -INFINITY;
      int l1 = -1;
      const uint32_t sbase = lane_addr + grp * 128;
      const int nruns = s_nruns;
      for (int r = 0; r < nruns; ++r) {
        const int c0 = s_run_start[r], len = s_run_start[r + 1] - c0;
        float cur = -INFINITY;
        int g = 0;
        for (; len - g >= 32; g += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(sbase + c0 + g, v);
          tmem_wait_ld();
          float t8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            t8[e] = fmaxf(fmaxf(__uint_as_float(v[e]), __uint_as_float(v[e + 8])), fmaxf(__uint_as_float(v[e + 16]), __uint_as_float(v[e + 24])));
          cur = fmaxf(cur, fmaxf(fmaxf(fmaxf(t8[0], t8[1]), fmaxf(t8[2], t8[3])), fmaxf(fmaxf(t8[4], t8[5]), fmaxf(t8[6], t8[7]))));
        }
        for (; g < len; g += 8) {
          uint32_t v[8];
          tmem_ld_32x32b_x8(sbase + c0 + g, v);
          tmem_wait_ld();
          const int left = len - g;                       // columns of this load that belong to the run
          float t8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) t8[e] = (e < left) ? __uint_as_float(v[e]) : -INFINITY;
          cur = fmaxf(cur, fmaxf(fmaxf(fmaxf(t8[0], t8[1]), fmaxf(t8[2], t8[3])), fmaxf(fmaxf(t8[4], t8[5]), fmaxf(t8[6], t8[7]))));
        }
        // top-2 over the runs (their labels are distinct); an exact tie of two labels leaves v1 == v2: not certified
        const bool gt = cur > v1;
        v2 = gt ? v1 : fmaxf(v2, cur);
        l1 = gt ? s_run_lab[r] : l1;
        v1 = gt ? cur : v1;
      }
      tc_fence_before();
      mbar_arrive(&s_empty[grp]);
      const long long pnt = (cta + (long long)j * P) * kTile + row;
      int label = -1;
      if (pnt < n) {
        if (v1 - v2 > kCertGap) {
          label = l1;
          labels_tmp[size_t(b) * n + pnt] = l1;
        } else {
          const unsigned int slot = atomicAdd(uncertain_count + b, 1u);
          uncertain_list[size_t(b) * n + slot] = int(pnt);
        }
      }
      const unsigned int active = __ballot_sync(0xffffffffu, label >= 0);
      if (label >= 0) {
        const unsigned int peers = __match_any_sync(active, label);
        if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[label], __popc(peers));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < UOC_MAX_SEEDS && s_hist[threadIdx.x] != 0) atomicAdd(hist + size_t(b) * m + threadIdx.x, s_hist[threadIdx.x]);
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// pass 2: canonical fp32 evaluation of the uncertified points (same arithmetic as assign_kernel)
__global__ void __launch_bounds__(128) assign_fix_kernel(const float* __restrict__ X, long long sb, long long sd, long long n,
                                                         int d, int m, const float* __restrict__ Z,
                                                         const int* __restrict__ seed_labels,
                                                         const unsigned int* __restrict__ uncertain_count,
                                                         const int* __restrict__ uncertain_list, int* __restrict__ hist,
                                                         int* __restrict__ labels_tmp) {
  extern __shared__ float zsf[];  // [m][d]
  __shared__ int s_lab[UOC_MAX_SEEDS];
  const int b = blockIdx.y, tid = threadIdx.x;
  const unsigned int count = uncertain_count[b];
  if (count == 0) return;
  const float* Zb = Z + size_t(b) * m * d;
  for (int e = tid; e < m * d; e += blockDim.x) zsf[e] = Zb[e];
  if (tid < UOC_MAX_SEEDS) s_lab[tid] = tid < m ? seed_labels[size_t(b) * m + tid] : 0;
  __syncthreads();
  const float* Xb = X + b * sb;
  for (unsigned int u = blockIdx.x * blockDim.x + tid; u < count; u += gridDim.x * blockDim.x) {
    const long long pnt = uncertain_list[size_t(b) * n + u];
    float best = 0.f;
    int bj = 0;
    for (int j = 0; j < m; ++j) {
      float acc = 0.f;
      for (int k = 0; k < d; ++k) acc = fmaf(__ldg(Xb + k * sd + pnt), zsf[j * d + k], acc);
      const float dist = 0.5f * (1.0f - acc);
      if (j == 0 || dist < best) { best = dist; bj = j; }
    }
    const int label = s_lab[bj];
    labels_tmp[size_t(b) * n + pnt] = label;
    atomicAdd(hist + size_t(b) * m + label, 1);
  }
}

template <int D>
int launch_tc(const CUtensorMap& tmap, const ClusterShape& s, const float* Z, const int* seed_labels, int P, int* labels_tmp,
              int* hist, unsigned int* ucount, int* ulist, cudaStream_t stream) {
  using Cfg = AsCfg<D>;
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&assign_tc_kernel<D>), int(Cfg::kSmemBytes));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  assign_tc_kernel<D><<<dim3(P, s.batch), kThreads, Cfg::kSmemBytes, stream>>>(tmap, Z, seed_labels, s.m, s.n, P, labels_tmp,
                                                                               hist, ucount, ulist, device_error_word());
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

}  // namespace

int launch_assign_tc(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w, const float* Z,
                     const int* seed_labels, int* hist, int* labels_tmp, cudaStream_t stream) {
  if (s.d != 64 && s.d != 128) return fail(UOC_ERR_UNSUPPORTED, "tcgen05 assignment supports d = 64 or 128");
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
  unsigned int* ucount = reinterpret_cast<unsigned int*>(w.keys);       // sampling scratch is free again at this point
  int* ulist = reinterpret_cast<int*>(w.r);
  UOC_CUDA(cudaMemsetAsync(ucount, 0, sizeof(unsigned int) * size_t(s.batch), stream));
  rc = (s.d == 64) ? launch_tc<64>(tmap, s, Z, seed_labels, P, labels_tmp, hist, ucount, ulist, stream)
                   : launch_tc<128>(tmap, s, Z, seed_labels, P, labels_tmp, hist, ucount, ulist, stream);
  if (rc != UOC_OK) return rc;
  const size_t smem = sizeof(float) * size_t(s.m) * s.d;
  {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&assign_fix_kernel), int(100 * 1024));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  assign_fix_kernel<<<dim3(sm_count() > 0 ? sm_count() : 148, s.batch), 128, smem, stream>>>(
      X, s.stride_b, s.stride_d, s.n, s.d, s.m, Z, seed_labels, ucount, ulist, hist, labels_tmp);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

}  // namespace uoc
