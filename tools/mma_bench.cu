// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (bf16) for several tile shapes with operands
// resident in shared memory (no TMA traffic), optionally with a concurrent TMA stream into other
// shared-memory buffers.  Prints cycles per MMA and MAC/clk/SM.  Build: nvcc -arch=sm_100a ... ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../unseenobjectclustering_b200/csrc/uoc_common.cuh"

using namespace uoc;

template <int N, bool TS, int NACC = 1>
__global__ void __launch_bounds__(192, 1) mma_rate_kernel(int iters, int kper, const __grid_constant__ CUtensorMap tmap,
                                                          int tma_bytes_per_iter, unsigned long long* out, unsigned int* err,
                                                          int a_sbo = 1024, int a_off = 0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a = smem;                 // 16 KB: 128 x 64 bf16 (K-major SW128)
  uint8_t* b = smem + 16384;         // 32 KB: 256 x 64
  uint8_t* sink = smem + 49152;      // TMA sink, 4 x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152 + 65536);
  uint64_t* done = bars;
  uint64_t* tfull = bars + 1;        // 4
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(done, 1); for (int i = 0; i < 4; ++i) mbar_init(&tfull[i], 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (warp == 0 && lane == 0 && tma_bytes_per_iter > 0) {
    // stream boxes of 16 KB round-robin into the sink while the MMAs run
    int n = 0;
    const int boxes = iters * kper * tma_bytes_per_iter / 16384 / 4;   // per-iteration bytes, issued in groups of 4
    for (int g = 0; g < boxes; ++g) {
      for (int i = 0; i < 4; ++i) {
        if (g > 0) mbar_wait(&tfull[i], (g - 1) & 1, err);
        mbar_arrive_expect_tx(&tfull[i], 16384);
        tma_load_2d(sink + i * 16384, &tmap, &tfull[i], 0, ((n++) * 128) % 32768);
      }
    }
  }
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t aa = smem_u32(a), ba = smem_u32(b);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < kper; ++k) {
        const int ks = k & 3;
        const uint64_t bd = make_smem_desc_sw128(ba + ks * 32, 16, 1024);
        const uint32_t dcol = tm + uint32_t((k % NACC) * N);          // NACC independent accumulators, round-robin
        if (TS) umma_ts_f16(dcol, tm + 256 + ks * 8, bd, idesc, 1u);
        else { const uint64_t ad = make_smem_desc_sw128(aa + a_off + ks * 32, 16, a_sbo); umma_ss_f16(dcol, ad, bd, idesc, 1u); }
      }
    }
    umma_commit(done);
    mbar_wait(done, 0, err);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}

template <int N, bool TS, int NACC = 1>
void run(const char* name, int grid, int tma_bytes_per_iter, const CUtensorMap& tmap, unsigned long long* dout, unsigned int* derr,
         int a_sbo = 1024, int a_off = 0) {
  const int iters = 2000, kper = 4;
  const int smem = 1024 + 49152 + 65536 + 256;
  cudaFuncSetAttribute(mma_rate_kernel<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma_rate_kernel<N, TS, NACC><<<grid, 192, smem>>>(100, kper, tmap, tma_bytes_per_iter, dout, derr, a_sbo, a_off);
  cudaEventRecord(e0);
  mma_rate_kernel<N, TS, NACC><<<grid, 192, smem>>>(iters, kper, tmap, tma_bytes_per_iter, dout, derr, a_sbo, a_off);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long cyc = 0; unsigned int herr = 0;
  cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  const double mmas = double(iters) * kper;
  const double macs = mmas * 128.0 * N * 16.0;
  printf("%-34s grid %3d tma %6d B/iter: %.1f clk/MMA, %.0f MAC/clk/SM, %.2f ms, %.1f MHz eff, %.1f TFLOP/s chip  err=%u %s\n", name, grid,
         tma_bytes_per_iter, cyc / mmas, macs / cyc, ms, cyc / (ms * 1e3), 2.0 * macs * grid / (ms * 1e-3) / 1e12, herr,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

// TMA-only: DEPTH boxes of 16 KB in flight per CTA, round-robin; reports L2->SMEM bytes/clk/SM
template <int DEPTH>
__global__ void __launch_bounds__(64, 1) tma_rate_kernel(int boxes, const __grid_constant__ CUtensorMap tmap, int mcast,
                                                         unsigned long long* out, unsigned int* err) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DEPTH * 16384);
  if (threadIdx.x == 0) { for (int i = 0; i < DEPTH; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
  __syncthreads();
  if (mcast > 1) cluster_sync_all();
  if (threadIdx.x == 0) {
    const uint32_t rank = mcast > 1 ? cluster_ctarank() : 0;
    long long t0 = clock64();
    for (int g = 0; g < boxes; ++g) {
      const int i = g % DEPTH;
      if (g >= DEPTH) mbar_wait(&bars[i], ((g / DEPTH) - 1) & 1, err);
      mbar_arrive_expect_tx(&bars[i], 16384);
      const int row = ((g * 148 + blockIdx.x) * 128) % 32768;
      if (mcast <= 1) tma_load_2d(smem + i * 16384, &tmap, &bars[i], 0, row);
      else {
        // every CTA of the cluster fetches a 1/mcast slice (128/mcast rows) and multicasts it to all
        const int rows = 128 / mcast;
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
            " [%0], [%1, {%4, %5}], [%2], %3;"
            ::"r"(smem_u32(smem + i * 16384 + rank * rows * 128)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bars[i])),
            "h"(uint16_t((1u << mcast) - 1u)), "r"(0), "r"(row + int(rank) * rows)
            : "memory");
      }
    }
    for (int i = 0; i < DEPTH && i < boxes; ++i) {
      const int uses = (boxes - i + DEPTH - 1) / DEPTH;
      mbar_wait(&bars[i], (uses - 1) & 1, err);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  __syncthreads();
  if (mcast > 1) cluster_sync_all();
}

template <int DEPTH>
void run_tma(int grid, int mcast, const CUtensorMap& tmap, const CUtensorMap& tmap_slice, unsigned long long* dout, unsigned int* derr) {
  const int boxes = 4000;
  const int smem = 1024 + DEPTH * 16384 + 256;
  cudaFuncSetAttribute(tma_rate_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = mcast > 1 ? mcast : 1;
  at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, tma_rate_kernel<DEPTH>, boxes, mcast > 1 ? tmap_slice : tmap, mcast, dout, derr);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long cyc = 0; unsigned int herr = 0;
  cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  printf("TMA only: %2d x 16KB in flight, grid %3d, multicast %d: %.1f B/clk/SM ingest, %.2f TB/s chip ingest, %.3f ms err=%u %s\n", DEPTH, grid,
         mcast, double(boxes) * 16384 / cyc, double(boxes) * 16384 * grid / (ms * 1e-3) / 1e12, ms, herr, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  unsigned long long* dout; unsigned int* derr; void* src;
  cudaMalloc(&dout, 8); cudaMalloc(&derr, 4); cudaMemset(derr, 0, 4);
  cudaMalloc(&src, 32768 * 128); cudaMemset(src, 0, 32768 * 128);
  CUtensorMap tmap;
  const uint64_t dims[2] = {64, 32768}; const uint64_t str[1] = {128}; const uint32_t box[2] = {64, 128};
  if (make_tmap_bf16(&tmap, src, 2, dims, str, box, nullptr) != 0) return 1;
  CUtensorMap tmap_s2, tmap_s4;
  { const uint32_t b2[2] = {64, 64}; const uint32_t b4[2] = {64, 32};
    make_tmap_bf16(&tmap_s2, src, 2, dims, str, b2, nullptr); make_tmap_bf16(&tmap_s4, src, 2, dims, str, b4, nullptr); }
  for (int grid : {1, 148}) {
    run_tma<8>(grid, 1, tmap, tmap, dout, derr);
  }
  run_tma<8>(148, 2, tmap, tmap_s2, dout, derr);
  run_tma<8>(148, 4, tmap, tmap_s4, dout, derr);
  run_tma<12>(148, 4, tmap, tmap_s4, dout, derr);
  for (int grid : {148}) {
    run<128, false>("SS 128x128x16", grid, 0, tmap, dout, derr);
    run<128, false>("SS 128x128x16 A SBO 2048", grid, 0, tmap, dout, derr, 2048, 0);
    run<128, false>("SS 128x128x16 A SBO 3072", grid, 0, tmap, dout, derr, 3072, 0);
    run<128, false>("SS 128x128x16 A SBO 3072 off 512", grid, 0, tmap, dout, derr, 3072, 512);
    run<128, false>("SS 128x128x16 A SBO 1024 off 512", grid, 0, tmap, dout, derr, 1024, 512);
    run<128, false>("SS 128x128x16 A SBO 2048 off 128", grid, 0, tmap, dout, derr, 2048, 128);
    run<256, false>("SS 128x256x16", grid, 0, tmap, dout, derr);
    run<64, false>("SS 128x64x16", grid, 0, tmap, dout, derr);
    run<128, true>("TS 128x128x16 (A in TMEM)", grid, 0, tmap, dout, derr);
    run<256, true>("TS 128x256x16 (A in TMEM)", grid, 0, tmap, dout, derr);
    run<64, true>("TS 128x64x16 (A in TMEM)", grid, 0, tmap, dout, derr);
    run<64, true, 2>("TS 128x64x16, 2 accumulators", grid, 0, tmap, dout, derr);
    run<64, true, 4>("TS 128x64x16, 4 accumulators", grid, 0, tmap, dout, derr);
    run<64, false, 4>("SS 128x64x16, 4 accumulators", grid, 0, tmap, dout, derr);
    run<128, false, 2>("SS 128x128x16, 2 accumulators", grid, 0, tmap, dout, derr);
    run<128, true, 2>("TS 128x128x16, 2 accumulators", grid, 0, tmap, dout, derr);
  }
  return 0;
}
