// K3 (third generation): farthest point sampling with a bf16 screening pass.
// (lib/utils/mean_shift.py:128-189: seed 0 is given; seed i+1 = argmax_p min_{j<=i} 0.5 (1 - x_p . x_{s_j}).)
//
// The second-generation kernel (cluster_kernels.cu, fps2_kernel) re-reads the whole fp32 field in every one of the
// m-1 passes: 78.6 MB per pass at 640x480x64, served by L2 at the ~48 B/clk/SM ingest ceiling -> ~5.6 us per pass.
// Here the running minimum r[p] is only ever LOWERED by a new seed, so a pass needs the exact distance only for the
// points the new seed can lower -- essentially its Voronoi cell, ~n/i points in pass i.  Everything else is screened
// with the bf16 pixel-major copy of the field (the one the mean-shift loop streams), most of which stays resident in
// shared memory for all passes (2 bytes/channel: 75% of a CTA's 2076 points fit in 200 KB):
//     d~ = 0.5 (1 - bf16(x_p) . s)             (fp32 accumulate, lanes split the channels)
//     |d~ - d| <= 2^-8 |x_p| |s| (+ fp32 slop)  (bf16 rounding or truncation of x_p, Cauchy-Schwarz)
//     d~ - margin >= r[p]  =>  d >= r[p]  =>  r[p] is unchanged: no fp32 read
// and only the remaining points evaluate the canonical fp32 chain (one fmaf chain over the channels from +0, exactly
// as fps_kernel / fps2_kernel / oracle), so r[], the arg-max keys and the selected indices are bit-identical.
// Pass 0 evaluates every point exactly and records |x_p|.
//
// Work split: one CTA per SM, 16 warps; the 32-point rounds of an item are dealt round-robin to its CTAs (the points
// a seed can lower form a compact blob of the image) and a CTA's rounds round-robin to its warps; a round is NL = d/8 warp-wide 16-byte loads (512 contiguous bytes each), lane l holds 8
// channels of point (32/NL)*q + l/NL of load q; a transposing butterfly leaves lane l with the full dot product of
// point (32/NL)*(l % NL) + l / NL, whose r[] and |x| it keeps in registers.
// Exchange between the CTAs: the all-to-all key matrix of fps2_kernel (SYNC == 2).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cluster.cuh"

namespace uoc {

namespace {

constexpr int kThreads3 = 512;
constexpr int kWarps3 = kThreads3 / 32;
constexpr float kRelMargin = 0.0040f;   // 2^-8 (bf16 truncation, halved by the 0.5 of the distance) + slack for the fp32 sum order
constexpr float kAbsMargin = 2e-6f;

struct Fps3Params {
  const float* X;
  const uint4* xb;          // bf16 pixel-major copy [batch][n][d], as 16-byte units
  long long sb, sd, n;
  int d, m, batch;
  const long long* first;
  long long* selected_out;
  float* seeds_out;
  unsigned int* err;
  unsigned long long* slots;
  int nb, res_rounds;
  long long* trace;         // debug (UOC_FPS3_TRACE=<cta>): per pass {start, warp 0 done, all warps done, exchanged, seed loaded, exact rounds of warp 0}
  int trace_cta;
};

__device__ __forceinline__ unsigned int orderable3(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_key3(float r, unsigned int idx) {
  return (static_cast<unsigned long long>(orderable3(r)) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float dot8_bf16(const uint4& u, const float (&s)[8]) {
  float a = __uint_as_float(u.x << 16) * s[0];
  a = fmaf(__uint_as_float(u.x & 0xFFFF0000u), s[1], a);
  a = fmaf(__uint_as_float(u.y << 16), s[2], a);
  a = fmaf(__uint_as_float(u.y & 0xFFFF0000u), s[3], a);
  a = fmaf(__uint_as_float(u.z << 16), s[4], a);
  a = fmaf(__uint_as_float(u.z & 0xFFFF0000u), s[5], a);
  a = fmaf(__uint_as_float(u.w << 16), s[6], a);
  a = fmaf(__uint_as_float(u.w & 0xFFFF0000u), s[7], a);
  return a;
}

template <int D, int MAXR>
__global__ void __launch_bounds__(kThreads3, 1) fps3_kernel(Fps3Params p) {
  constexpr int NL = D / 8;        // 16-byte loads per 32-point round == lanes per point
  constexpr int PPL = 32 / NL;     // points per warp-wide load
  extern __shared__ uint4 xres[];  // resident rounds [res_rounds][NL][32]
  __shared__ float s_seed[D];
  __shared__ unsigned long long s_red[kWarps3];
  __shared__ long long s_idx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nb = p.nb;
  const int b = blockIdx.x / nb, rank = blockIdx.x % nb;
  // 32-point rounds are dealt round-robin to the CTAs of an item (local round rho <-> global round rho * nb + rank):
  // the points a new seed can lower form a compact blob of the image, which a contiguous split would hand to a few CTAs
  const int total_rounds = int((p.n + 31) >> 5);
  const int nrounds = rank < total_rounds ? (total_rounds - rank + nb - 1) / nb : 0;
  const int res_rounds = p.res_rounds < nrounds ? p.res_rounds : nrounds;
  const float* Xb = p.X + b * p.sb;
  const uint4* xbb = p.xb + size_t(b) * p.n * NL;
  auto round_base = [&](int rho) { return ((long long)rho * nb + rank) * 32; };   // first point of a local round

  // resident rounds (straight copies of 32 * d * 2 contiguous bytes each)
  for (int e = tid; e < res_rounds * 32 * NL; e += kThreads3) {
    const int rho = e / (32 * NL), o = e % (32 * NL);
    const long long pnt = round_base(rho) + o / NL;
    xres[e] = (pnt < p.n) ? __ldg(xbb + round_base(rho) * NL + o) : make_uint4(0u, 0u, 0u, 0u);
  }

  const int g = lane / NL, j = lane % NL;
  const int pt = PPL * j + g;                 // the point of a round this lane owns after the butterfly
  float r[MAXR], nx[MAXR];
#pragma unroll
  for (int s = 0; s < MAXR; ++s) { r[s] = 0.f; nx[s] = 0.f; }
  long long idx = p.first[b];
  if (tid < D) s_seed[tid] = __ldg(Xb + tid * p.sd + idx);
  __syncthreads();

  for (int i = 0; i < p.m; ++i) {
    if (rank == 0) {
      if (tid == 0) p.selected_out[size_t(b) * p.m + i] = idx;
      if (tid < D) p.seeds_out[(size_t(b) * p.m + i) * D + tid] = s_seed[tid];
    }
    if (i + 1 == p.m) break;
    const bool tr = p.trace && int(blockIdx.x) == p.trace_cta && tid == 0;
    int n_exact = 0;
    if (tr) p.trace[i * 6 + 0] = clock64();
    float sreg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sreg[e] = s_seed[j * 8 + e];
    float ns = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ns = fmaf(sreg[e], sreg[e], ns);
#pragma unroll
    for (int h = NL / 2; h > 0; h >>= 1) ns += __shfl_xor_sync(0xffffffffu, ns, h);
    ns = sqrtf(ns) * 1.000001f;

    unsigned long long best = 1ull;           // non-zero sentinel: an empty slice still signals arrival
#pragma unroll
    for (int s = 0; s < MAXR; ++s) {
      const int rho = warp + s * kWarps3;
      if (rho < nrounds) {                    // warp-uniform
        const long long rb = round_base(rho);
        const long long gp = rb + pt;
        const bool valid = gp < p.n;
        bool need = valid;
        if (i > 0) {
          float v[NL];
          if (rho < res_rounds) {
            const uint4* src = xres + (size_t(rho) * NL) * 32 + lane;
#pragma unroll
            for (int q = 0; q < NL; ++q) v[q] = dot8_bf16(src[q * 32], sreg);
          } else {
            const uint4* src = xbb + rb * NL + lane;
            uint4 u[NL];
#pragma unroll
            for (int q = 0; q < NL; ++q)
              u[q] = (rb + q * PPL + g < p.n) ? __ldg(src + q * 32) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int q = 0; q < NL; ++q) v[q] = dot8_bf16(u[q], sreg);
          }
          // transposing butterfly over the NL lanes of a point group: lane j ends with the dot product of load q == j
#pragma unroll
          for (int h = NL / 2; h > 0; h >>= 1) {
            const bool up = (j & h) != 0;
#pragma unroll
            for (int t = 0; t < h; ++t) {
              const float keep = up ? v[t + h] : v[t];
              const float send = up ? v[t] : v[t + h];
              v[t] = keep + __shfl_xor_sync(0xffffffffu, send, h);
            }
          }
          const float dapprox = 0.5f * (1.0f - v[0]);
          need = valid && !((dapprox - fmaf(kRelMargin * nx[s], ns, kAbsMargin)) >= r[s]);
        }
        if (__any_sync(0xffffffffu, need)) {
          ++n_exact;
          if (need) {
            // canonical fp32 chain (bit-identical to fps_kernel / fps2_kernel / the oracle)
            const float* xp = Xb + gp;
            float acc = 0.f, sq = 0.f;
#pragma unroll 1
            for (int k0 = 0; k0 < D; k0 += 32) {      // 32 loads in flight per lane
              float x[32];
#pragma unroll
              for (int k = 0; k < 32; ++k) x[k] = __ldg(xp + (k0 + k) * p.sd);
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                acc = fmaf(x[k], s_seed[k0 + k], acc);
                if (i == 0) sq = fmaf(x[k], x[k], sq);
              }
            }
            const float dist = 0.5f * (1.0f - acc);
            if (i == 0) {
              r[s] = dist;
              nx[s] = sqrtf(sq) * 1.000001f;
            } else {
              r[s] = dist < r[s] ? dist : r[s];
            }
          }
        }
        if (valid) {
          const unsigned long long key = pack_key3(r[s], static_cast<unsigned int>(gp));
          best = key > best ? key : best;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) s_red[warp] = best;
    if (tr) { p.trace[i * 6 + 1] = clock64(); p.trace[i * 6 + 5] = n_exact; }
    __syncthreads();                          // also: everybody is done reading s_seed of this pass
    if (tr) p.trace[i * 6 + 2] = clock64();
    if (warp == 0) {
      unsigned long long v = (lane < kWarps3) ? s_red[lane] : 0ull;
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
      if (lane == 0) {
        s_idx = done ? static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(gmax & 0xFFFFFFFFull)) : -1;
        if (!done) atomicOr(p.err, ERR_GRID_BARRIER_TIMEOUT);
      }
    }
    __syncthreads();
    if (tr) p.trace[i * 6 + 3] = clock64();
    idx = s_idx;
    if (idx < 0) return;
    if (tid < D) s_seed[tid] = __ldg(Xb + tid * p.sd + idx);
    __syncthreads();
    if (tr) p.trace[i * 6 + 4] = clock64();
  }
}

}  // namespace

int launch_select_seeds_pruned(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w,
                               int64_t* selected_out, float* seeds_out, cudaStream_t stream, bool* used) {
  *used = false;
  if (!xb || (s.d != 64 && s.d != 128)) return UOC_OK;
  if (reinterpret_cast<uintptr_t>(xb) % 16 != 0) return UOC_OK;
  bool enabled = false;   // opt-in until it beats the second-generation kernel (profiles/r01_fps3_trace.txt)
  if (const char* e = getenv("UOC_FPS_PRUNED")) enabled = atoi(e) != 0;
  if (!enabled) return UOC_OK;
  const int sms = sm_count();
  if (sms <= 0 || s.batch > sms) return UOC_OK;
  const int nb = sms / s.batch;
  if (nb > 160) return UOC_OK;                      // the poll loop reads at most 5 keys per lane
  const long long total_rounds = (s.n + 31) / 32;
  const long long rounds = (total_rounds + nb - 1) / nb;     // local rounds of the busiest CTA
  if (rounds > 12 * kWarps3) return UOC_OK;         // > 6144 points per CTA: second-generation kernel
  const size_t slot_need = size_t(s.batch) * s.m * nb * nb * 8;
  if (slot_need > w.slot_bytes) return UOC_OK;
  unsigned int* err = device_error_word();
  if (!err) return fail(UOC_ERR_CUDA, "no device error word");

  size_t budget = 200 * 1024;
  if (const char* e = getenv("UOC_FPS_SMEM_KB")) budget = size_t(atoi(e)) * 1024;
  if (budget > 220 * 1024) budget = 220 * 1024;
  const size_t round_bytes = size_t(32) * s.d * 2;
  int res_rounds = int(budget / round_bytes);
  if (res_rounds > rounds) res_rounds = int(rounds);
  const size_t smem = size_t(res_rounds) * round_bytes;

  void* kern;
  if (s.d == 64) kern = rounds <= 6 * kWarps3 ? reinterpret_cast<void*>(&fps3_kernel<64, 6>) : reinterpret_cast<void*>(&fps3_kernel<64, 12>);
  else kern = rounds <= 6 * kWarps3 ? reinterpret_cast<void*>(&fps3_kernel<128, 6>) : reinterpret_cast<void*>(&fps3_kernel<128, 12>);
  UOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));

  Fps3Params p;
  p.X = X; p.xb = reinterpret_cast<const uint4*>(xb);
  p.sb = s.stride_b; p.sd = s.stride_d; p.n = s.n; p.d = s.d; p.m = s.m; p.batch = s.batch;
  p.first = w.first;
  p.selected_out = reinterpret_cast<long long*>(selected_out);
  p.seeds_out = seeds_out;
  p.err = err;
  p.slots = w.slots;
  p.nb = nb; p.res_rounds = res_rounds;
  p.trace = nullptr;
  p.trace_cta = 0;
  if (const char* e = getenv("UOC_FPS3_TRACE")) {
    p.trace_cta = atoi(e);
    UOC_CUDA(cudaMalloc(&p.trace, sizeof(long long) * 6 * s.m));
    UOC_CUDA(cudaMemsetAsync(p.trace, 0, sizeof(long long) * 6 * s.m, stream));
  }
  UOC_CUDA(cudaMemsetAsync(w.slots, 0, slot_need, stream));
  void* args[] = {&p};
  UOC_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nb * s.batch), dim3(kThreads3), args, smem, stream));
  count_launch();
  if (p.trace) {
    std::vector<long long> h(size_t(6) * s.m);
    UOC_CUDA(cudaStreamSynchronize(stream));
    UOC_CUDA(cudaMemcpy(h.data(), p.trace, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    long long sum[4] = {0, 0, 0, 0};
    for (int i = 0; i + 1 < s.m; ++i) {
      const long long* q = h.data() + size_t(i) * 6;
      if (i < 12 || i % 10 == 0)
        fprintf(stderr, "[fps3 trace] cta %d pass %d: warp0 %lld  all warps %lld  exchange %lld  seed %lld clk; exact rounds (warp 0) %lld\n",
                p.trace_cta, i, q[1] - q[0], q[2] - q[0], q[3] - q[2], q[4] - q[3], q[5]);
      if (i >= 10) { sum[0] += q[1] - q[0]; sum[1] += q[2] - q[0]; sum[2] += q[3] - q[2]; sum[3] += q[4] - q[3]; }
    }
    const double cnt = s.m - 11;
    fprintf(stderr, "[fps3 trace] mean over passes >= 10: warp0 %.0f  all warps %.0f  exchange %.0f  seed %.0f clk\n",
            sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt);
  }
  *used = true;
  return UOC_OK;
}

}  // namespace uoc
