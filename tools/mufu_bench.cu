// Micro-benchmark: MUFU.EX2 / F2FP / PRMT issue rates per SM (one CTA per SM, W warps per CTA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mufu_bench tools/mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = -0.001f * (threadIdx.x + i + 1);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {            // MUFU.EX2 only
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      } else if (MODE == 1) {     // FFMA + MUFU.EX2
        v[i] = fmaf(v[i], 0.999f, -0.0001f);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      }
    }
    if (MODE == 2) {              // F2FP only (8 packs of the 16 values)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[2 * i]), "f"(v[2 * i + 1]));
        acc ^= r;
        v[2 * i] = __uint_as_float(r | 0x3f000000u);
      }
    }
    if (MODE == 3) {              // MUFU + F2FP at the mean-shift ratio (2 : 1)
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v[2 * i]), "f"(v[2 * i + 1]));
        acc ^= r;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, double ops_per_iter_per_thread) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 4096;
  k<MODE><<<148, warps * 32>>>(out, clk, iters);
  k<MODE><<<148, warps * 32>>>(out, clk, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-28s warps/SM %2d : %.1f clk/iter, %.2f thread-ops/clk/SM\n", name, warps, c / iters,
         ops_per_iter_per_thread * warps * 32 * iters / c);
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("MUFU.EX2", w, 16);
    run<1>("FFMA+MUFU.EX2 (per ex2)", w, 16);
    run<2>("F2FP.BF16 pack (per pack)", w, 8);
    run<3>("16 MUFU + 8 F2FP (per ex2)", w, 16);
  }
  return 0;
}
