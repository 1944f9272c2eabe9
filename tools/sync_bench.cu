// Micro-benchmark of inter-CTA signalling latency through L2 on B200:
//  (a) ping-pong between two CTAs (store / poll), (b) grid-wide exchange of 148 CTAs with three protocols.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cg(unsigned long long* p, unsigned long long v) {
  asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_weak(unsigned long long* p, unsigned long long v) {
  asm volatile("st.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__global__ void pingpong(unsigned long long* flags, int iters, long long* out, int mode) {
  if (threadIdx.x != 0) return;
  unsigned long long* mine = flags + blockIdx.x * 32;       // separate lines
  unsigned long long* other = flags + (1 - blockIdx.x) * 32;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (blockIdx.x == 0) {
      st_relaxed(other, i);
      while ((mode ? ld_volatile(mine) : ld_relaxed(mine)) < (unsigned long long)i) {}
    } else {
      while ((mode ? ld_volatile(mine) : ld_relaxed(mine)) < (unsigned long long)i) {}
      st_relaxed(other, i);
    }
  }
  long long t1 = clock64();
  if (blockIdx.x == 0) out[0] = t1 - t0;
}

// proto 0: atomic counter barrier; proto 1: all-to-all matrix [dst][src]; proto 2: one slot per CTA (own line), everybody polls all slots
__global__ void exchange(unsigned long long* buf, unsigned int* counter, int iters, long long* out, int proto, int stmode) {
  const int nb = gridDim.x, rank = blockIdx.x, tid = threadIdx.x;
  long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (proto == 0) {
      if (tid == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < (unsigned int)i * nb);
      }
    } else if (proto == 1) {
      if (tid < 32) {
        for (int c = tid; c < nb; c += 32) {
          if (stmode == 0) st_relaxed(buf + (size_t)c * nb + rank, i);
          else if (stmode == 1) st_cg(buf + (size_t)c * nb + rank, i);
          else st_weak(buf + (size_t)c * nb + rank, i);
        }
        bool done = false;
        while (!done) {
          bool all = true;
          unsigned long long kv[5];
#pragma unroll
          for (int s = 0; s < 5; ++s) { int c = tid + 32 * s; kv[s] = ~0ull; if (c < nb) kv[s] = ld_relaxed(buf + (size_t)rank * nb + c); }
#pragma unroll
          for (int s = 0; s < 5; ++s) all = all && kv[s] >= (unsigned long long)i;
          done = __all_sync(0xffffffffu, all);
        }
      }
    } else {
      if (tid < 32) {
        if (tid == 0) st_relaxed(buf + (size_t)rank * 16, i);
        bool done = false;
        while (!done) {
          bool all = true;
          unsigned long long kv[5];
#pragma unroll
          for (int s = 0; s < 5; ++s) { int c = tid + 32 * s; kv[s] = ~0ull; if (c < nb) kv[s] = ld_relaxed(buf + (size_t)c * 16); }
#pragma unroll
          for (int s = 0; s < 5; ++s) all = all && kv[s] >= (unsigned long long)i;
          done = __all_sync(0xffffffffu, all);
        }
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
}

int main() {
  unsigned long long* buf; unsigned int* counter; long long* out;
  cudaMalloc(&buf, 148 * 148 * 8 + 4096); cudaMalloc(&counter, 4); cudaMalloc(&out, 8);
  long long h;
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(buf, 0, 4096);
    pingpong<<<2, 32>>>(buf, 2000, out, mode);
    cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("ping-pong (%s poll): %.0f clk per round trip (2 hops)\n", mode ? "ld.volatile" : "ld.relaxed.gpu", h / 2000.0);
  }
  const char* names[3] = {"atomic counter barrier", "all-to-all matrix", "slot per CTA, all poll all"};
  const char* stn[3] = {"st.relaxed.gpu", "st.global.cg", "st.global (weak)"};
  for (int grid : {148}) {
    for (int stmode = 0; stmode < 3; ++stmode) {
      int proto = 1;
      cudaMemset(buf, 0, 148 * 148 * 8 + 4096); cudaMemset(counter, 0, 4);
      void* args[] = {&buf, &counter, nullptr, &out, &proto, &stmode};
      int iters = 2000; args[2] = &iters;
      cudaLaunchCooperativeKernel((void*)exchange, dim3(grid), dim3(128), args, 0, 0);
      cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("grid %3d all-to-all with %-18s: %.0f clk per exchange\n", grid, stn[stmode], h / 2000.0);
    }
  }
  for (int grid : {8, 32, 148}) {
    for (int proto = 0; proto < 3; ++proto) {
      int stmode = 0;
      cudaMemset(buf, 0, 148 * 148 * 8 + 4096); cudaMemset(counter, 0, 4);
      void* args[] = {&buf, &counter, nullptr, &out, &proto, &stmode};
      int iters = 2000; args[2] = &iters;
      cudaError_t e = cudaLaunchCooperativeKernel((void*)exchange, dim3(grid), dim3(128), args, 0, 0);
      cudaError_t e2 = cudaDeviceSynchronize(); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("grid %3d %-28s: %.0f clk per exchange %s %s\n", grid, names[proto], h / 2000.0, e == cudaSuccess ? "" : cudaGetErrorString(e),
             e2 == cudaSuccess ? "" : cudaGetErrorString(e2));
    }
  }
  return 0;
}
