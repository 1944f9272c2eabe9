// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (one CTA per SM, 4 or 8 warps, 32x32b.x32 loads).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_ld32(uint32_t a, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(a));
}
__device__ __forceinline__ void tmem_st16(uint32_t a, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]));
}

// MODE 0: loads only (wait after every load)   1: loads, two in flight   2: load + store (mean-shift pattern)
// MODE 3: load + 32 MUFU.EX2 + store           4: 32 MUFU.EX2 only per iteration
// MODE 5: load + 24 MUFU.EX2 + 8 polynomial ex2 (FMA pipe) + store      6: load + 20 MUFU + 12 polynomial + store
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -120.f);
  const float r = x + 12582912.f;
  const float f = x - (r - 12582912.f);
  const float p = fmaf(fmaf(fmaf(0.055268917f, f, 0.24221092f), f, 0.6932298f), f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23));
}
template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
  uint32_t va[32], vb[32], pk[16];
  for (int i = 0; i < 16; ++i) pk[i] = threadIdx.x + i;
  for (int c = 0; c < 8; ++c) { tmem_st16(base + c * 16, pk); }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t a = base + (it & 3) * 32;
    if (MODE == 0) {
      tmem_ld32(a, va);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      acc += __uint_as_float(va[it & 31]);
    } else if (MODE == 1) {
      tmem_ld32(a, va);
      tmem_ld32(a ^ 64, vb);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      acc += __uint_as_float(va[it & 31]) + __uint_as_float(vb[it & 31]);
    } else if (MODE == 5 || MODE == 6) {
      constexpr int POLY = (MODE == 5) ? 8 : 12;
      tmem_ld32(a, va);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float p0 = fmaf(__uint_as_float(va[2 * e]), 1e-30f, -0.5f), p1 = fmaf(__uint_as_float(va[2 * e + 1]), 1e-30f, -0.5f);
        const bool poly0 = ((2 * e) * POLY) / 32 != ((2 * e + 1) * POLY) / 32;
        const bool poly1 = ((2 * e + 1) * POLY) / 32 != ((2 * e + 2) * POLY) / 32;
        if (poly0) p0 = ex2_poly(p0); else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(p0));
        if (poly1) p1 = ex2_poly(p1); else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(p1));
        asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[e]) : "r"(__float_as_uint(p0)), "r"(__float_as_uint(p1)));
      }
      tmem_st16(base + (it & 7) * 16, pk);
      if ((it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;");
    } else if (MODE == 2 || MODE == 3) {
      tmem_ld32(a, va);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float p0 = __uint_as_float(va[2 * e]) * 1e-30f, p1 = __uint_as_float(va[2 * e + 1]) * 1e-30f;
        if (MODE == 3) {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(p0));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(p1));
        }
        asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk[e]) : "r"(__float_as_uint(p0)), "r"(__float_as_uint(p1)));
      }
      tmem_st16(base + (it & 7) * 16, pk);
      if ((it & 3) == 3) asm volatile("tcgen05.wait::st.sync.aligned;");
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) { float p = acc * 1e-3f - e; asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(p)); va[e] = __float_as_uint(p); }
#pragma unroll
      for (int e = 0; e < 32; ++e) acc += __uint_as_float(va[e]);
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

template <int MODE>
void run(const char* name, int warps, double loads_per_iter) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 8192;
  k<MODE><<<148, warps * 32>>>(out, clk, iters);
  k<MODE><<<148, warps * 32>>>(out, clk, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += double(h[i]) / 148;
  const double bytes = double(iters) * loads_per_iter * warps * 32 * 32 * 4;   // TMEM bytes read per SM
  printf("%-44s %d warps: %8.1f clk/iter  TMEM read %6.1f B/clk/SM  (%s)\n", name, warps, avg / iters, bytes / avg,
         cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {4, 8}) {
    run<0>("ld.32x32b.x32 + wait", w, 1);
    run<1>("2 x ld.32x32b.x32 + wait", w, 2);
    run<2>("ld + pack + st.x16 (no MUFU)", w, 1);
    run<3>("ld + 32 ex2 + pack + st.x16 (mean-shift)", w, 1);
    run<4>("32 ex2 only", w, 0);
    run<5>("ld + 24 ex2 + 8 poly + pack + st.x16", w, 1);
    run<6>("ld + 20 ex2 + 12 poly + pack + st.x16", w, 1);
  }
  return 0;
}
