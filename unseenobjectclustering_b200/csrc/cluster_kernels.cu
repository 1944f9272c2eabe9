// SIMT kernels of the clustering path (everything except the tcgen05 mean-shift loop):
//   K3  farthest point sampling       -- persistent cooperative kernel, one grid barrier per seed
//   K4' mean-shift iteration, fp32    -- validation kernel (UOC_FLAG_LOOP_SIMT), not the product path
//   K4r partial-sum reduce + L2 normalise
//   K5a greedy seed labelling         -- one CTA per field
//   K5b nearest-seed assignment, label histogram, label-0 swap
//   pack fp32 planar -> bf16 pixel-major
//
// Canonical fp32 arithmetic (mirrored bit for bit by oracle/uoc_oracle_c.c): every dot product that
// feeds a discrete decision (arg-max of the sampling, epsilon threshold of the labelling, arg-min
// of the assignment) is one sequential fused-multiply-add chain over the channels k = 0..d-1
// starting from +0, and the cosine distance is 0.5f * (1.0f - dot).  Euclidean metric (METRIC_EUCLIDEAN): the chain is
// acc = fmaf(t, t, acc) with t = x_k - z_k, and the distance is sqrtf(acc) (IEEE round-to-nearest).
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "cluster.cuh"

namespace uoc {

// ----------------------------------------------------------------------------------------------
// workspace
// ----------------------------------------------------------------------------------------------
static int partial_capacity(int batch) {
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int P = (sms + batch - 1) / batch;
  return P < 1 ? 1 : P;
}

static size_t carve(size_t& off, size_t bytes) {
  size_t at = off;
  off = align_up(off + bytes, 256);
  return at;
}

struct WsLayout {
  size_t r, keys, slots, barrier, first, Z, partials, seed_labels, num_unique, hist, labels_tmp, xb, wsum, total;
  size_t slot_bytes;
  int P;
};

static WsLayout ws_layout(int batch, int64_t n, int d, int m) {
  WsLayout L;
  size_t off = 0;
  L.P = partial_capacity(batch);
  L.r = carve(off, sizeof(float) * size_t(batch) * n);
  L.keys = carve(off, sizeof(unsigned long long) * size_t(batch) * m);
  {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int nb = batch <= sms ? sms / batch : 1;
    size_t per_pass = size_t(nb) * 128;                               // leader protocol: one line per CTA
    if (size_t(nb) * nb * 16 > per_pass) per_pass = size_t(nb) * nb * 16;   // all-to-all protocol: nb x nb matrix of (first, second) keys
    L.slot_bytes = size_t(batch) * m * per_pass + size_t(batch) * m * d * 8;   // key slots + seed mailboxes
    // a batch whose fields do not fit on chip together is sampled one field at a time at full width (nb = all SMs)
    const size_t single = size_t(m) * size_t(sms) * sms * 16 + size_t(m) * d * 8;
    if (batch > 1 && single > L.slot_bytes) L.slot_bytes = single;
    L.slots = carve(off, L.slot_bytes);
  }
  L.barrier = carve(off, 256);
  L.first = carve(off, sizeof(long long) * size_t(batch));
  L.Z = carve(off, sizeof(float) * size_t(batch) * m * d);
  L.partials = carve(off, sizeof(float) * size_t(batch) * L.P * 128 * d);
  L.seed_labels = carve(off, sizeof(int) * size_t(batch) * m);
  L.num_unique = carve(off, sizeof(int) * size_t(batch));
  L.hist = carve(off, sizeof(int) * size_t(batch) * m);
  L.labels_tmp = carve(off, sizeof(int) * size_t(batch) * n);
  L.xb = carve(off, sizeof(__nv_bfloat16) * size_t(batch) * n * d);
  L.wsum = carve(off, sizeof(float) * size_t(batch) * L.P * 128);
  L.total = off;
  return L;
}

size_t cluster_workspace_bytes(int batch, int64_t n, int d, int m) { return ws_layout(batch, n, d, m).total; }

int carve_cluster_workspace(void* ws, size_t ws_bytes, int batch, int64_t n, int d, int m, ClusterWorkspace* out) {
  if (!ws) return fail(UOC_ERR_INVALID, "workspace pointer is null");
  if (reinterpret_cast<uintptr_t>(ws) % 256 != 0) return fail(UOC_ERR_INVALID, "workspace must be 256-byte aligned");
  WsLayout L = ws_layout(batch, n, d, m);
  if (ws_bytes < L.total) {
    char buf[160];
    snprintf(buf, sizeof(buf), "workspace too small: need %zu bytes, got %zu", L.total, ws_bytes);
    return fail(UOC_ERR_WORKSPACE, buf);
  }
  char* base = static_cast<char*>(ws);
  out->r = reinterpret_cast<float*>(base + L.r);
  out->keys = reinterpret_cast<unsigned long long*>(base + L.keys);
  out->slots = reinterpret_cast<unsigned long long*>(base + L.slots);
  out->slot_bytes = L.slot_bytes;
  out->barrier = reinterpret_cast<unsigned int*>(base + L.barrier);
  out->first = reinterpret_cast<long long*>(base + L.first);
  out->Z = reinterpret_cast<float*>(base + L.Z);
  out->partials = reinterpret_cast<float*>(base + L.partials);
  out->seed_labels = reinterpret_cast<int*>(base + L.seed_labels);
  out->num_unique = reinterpret_cast<int*>(base + L.num_unique);
  out->hist = reinterpret_cast<int*>(base + L.hist);
  out->labels_tmp = reinterpret_cast<int*>(base + L.labels_tmp);
  out->xb = reinterpret_cast<__nv_bfloat16*>(base + L.xb);
  out->wsum = reinterpret_cast<float*>(base + L.wsum);
  out->max_partials = L.P;
  return UOC_OK;
}

// ----------------------------------------------------------------------------------------------
// K3: farthest point sampling
// ----------------------------------------------------------------------------------------------
struct FpsParams {
  const float* X;
  long long sb, sd;
  long long n;
  int d, m, batch;
  const long long* first;
  float* r;
  unsigned long long* keys;
  unsigned int* barrier;
  long long* selected_out;
  float* seeds_out;
  unsigned int* err;
  long long* trace;       // optional (debug): per pass {start, after local arg-max, after exchange} clock64 of CTA `trace_cta`
  int trace_cta;
  int debug_mode;         // 0 normal; 1 = skip the inter-CTA exchange (each CTA follows its own arg-max); 2 = skip the streaming loop
  const float* init_seeds = nullptr;   // [batch][num_init][d]: seeds chosen already (mean_shift.py:144-149, :164-169); fps_kernel only
  int num_init = 0;
};

__device__ __forceinline__ unsigned int orderable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long pack_key(float r, unsigned int idx) {
  return (static_cast<unsigned long long>(orderable(r)) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - idx);
}

__device__ __forceinline__ bool grid_barrier(unsigned int* counter, unsigned int target, unsigned int* err) {
  __syncthreads();
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    int ok = 0;
    for (unsigned int it = 0; it < (1u << 27); ++it) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) { ok = 1; break; }
    }
    if (!ok) atomicOr(err, ERR_GRID_BARRIER_TIMEOUT);
    s_ok = ok;
    __threadfence();
  }
  __syncthreads();
  return s_ok != 0;
}

template <int VEC, int METRIC>
__global__ void __launch_bounds__(256) fps_kernel(FpsParams p) {
  extern __shared__ float s_seed[];  // d floats
  __shared__ unsigned long long s_red[8];
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  // block -> item mapping: with G >= batch every item owns nb = G / batch blocks
  const int nb = (G >= p.batch) ? (G / p.batch) : 1;
  const int my_first_item = (G >= p.batch) ? (int(blockIdx.x) / nb) : int(blockIdx.x);
  const int item_step = (G >= p.batch) ? p.batch : G;  // one item per block when G >= batch
  const int rank = (G >= p.batch) ? (int(blockIdx.x) % nb) : 0;
  const long long ngroups = p.n / VEC;
  unsigned int target = 0;

  for (int i = 0; i < p.m; ++i) {
    for (int b = my_first_item; b < p.batch; b += item_step) {
      long long idx;
      const float* Xb = p.X + b * p.sb;
      if (i < p.num_init) {
        // a seed the caller chose before (mean_shift.py:164-169): its distances enter the running minimum like any
        // other column; there is no index for it (selected_indices stays -1, :140)
        idx = -1;
        for (int k = tid; k < p.d; k += blockDim.x) s_seed[k] = __ldg(p.init_seeds + (size_t(b) * p.num_init + i) * p.d + k);
      } else {
        if (i == 0) {
          idx = p.first[b];
        } else {
          unsigned long long key = __ldcg(p.keys + size_t(b) * p.m + i);
          idx = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(key & 0xFFFFFFFFull));
        }
        for (int k = tid; k < p.d; k += blockDim.x) s_seed[k] = __ldg(Xb + k * p.sd + idx);
      }
      __syncthreads();
      if (rank == 0) {
        if (tid == 0) p.selected_out[size_t(b) * p.m + i] = idx;
        for (int k = tid; k < p.d; k += blockDim.x) p.seeds_out[(size_t(b) * p.m + i) * p.d + k] = s_seed[k];
      }
      if (i + 1 < p.m) {
        unsigned long long best = 0ull;
        float* rb = p.r + size_t(b) * p.n;
        for (long long g = (long long)rank * blockDim.x + tid; g < ngroups; g += (long long)nb * blockDim.x) {
          const float* xp = Xb + g * VEC;
          float acc[VEC];
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll 8
          for (int k = 0; k < p.d; ++k) {
            const float sk = s_seed[k];
            if (VEC == 4) {
              float4 v = __ldg(reinterpret_cast<const float4*>(xp + k * p.sd));
              if (METRIC == METRIC_EUCLIDEAN) {
                v.x -= sk; v.y -= sk; v.z -= sk; v.w -= sk;
                acc[0] = fmaf(v.x, v.x, acc[0]);
                acc[1 % VEC] = fmaf(v.y, v.y, acc[1 % VEC]);
                acc[2 % VEC] = fmaf(v.z, v.z, acc[2 % VEC]);
                acc[3 % VEC] = fmaf(v.w, v.w, acc[3 % VEC]);
              } else {
                acc[0] = fmaf(v.x, sk, acc[0]);
                acc[1 % VEC] = fmaf(v.y, sk, acc[1 % VEC]);
                acc[2 % VEC] = fmaf(v.z, sk, acc[2 % VEC]);
                acc[3 % VEC] = fmaf(v.w, sk, acc[3 % VEC]);
              }
            } else if (METRIC == METRIC_EUCLIDEAN) {
              const float t = __ldg(xp + k * p.sd) - sk;
              acc[0] = fmaf(t, t, acc[0]);
            } else {
              acc[0] = fmaf(__ldg(xp + k * p.sd), sk, acc[0]);
            }
          }
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] = (METRIC == METRIC_EUCLIDEAN) ? sqrtf(acc[j]) : 0.5f * (1.0f - acc[j]);
          float rn[VEC];
          if (i == 0) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) rn[j] = acc[j];
          } else {
            float ro[VEC];
            if (VEC == 4) {
              const float4 o = *reinterpret_cast<const float4*>(rb + g * VEC);
              ro[0] = o.x; ro[1 % VEC] = o.y; ro[2 % VEC] = o.z; ro[3 % VEC] = o.w;
            } else {
              ro[0] = rb[g];
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              const float dj = acc[j];
              rn[j] = dj < ro[j] ? dj : ro[j];
            }
          }
          if (VEC == 4) {
            *reinterpret_cast<float4*>(rb + g * VEC) = make_float4(rn[0], rn[1 % VEC], rn[2 % VEC], rn[3 % VEC]);
          } else {
            rb[g] = rn[0];
          }
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const unsigned long long key = pack_key(rn[j], static_cast<unsigned int>(g * VEC + j));
            best = key > best ? key : best;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
          best = other > best ? other : best;
        }
        if ((tid & 31) == 0) s_red[tid >> 5] = best;
        __syncthreads();
        if (tid < 32) {
          unsigned long long v = (tid < (blockDim.x >> 5)) ? s_red[tid] : 0ull;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
            v = other > v ? other : v;
          }
          if (tid == 0 && v != 0ull) atomicMax(p.keys + size_t(b) * p.m + i + 1, v);
        }
      }
      __syncthreads();  // s_seed / s_red reuse
    }
    if (i + 1 < p.m) {
      target += gridDim.x;
      if (!grid_barrier(p.barrier, target, p.err)) return;
    }
  }
}

// Second-generation sampling kernel: one CTA per SM, every CTA owns a FIXED contiguous slice of points for
// all m passes.
//  (1) the running min r[] of a thread's points lives in registers;
//  (2) as much of the slice as fits (~200 KB) is copied once into shared memory and re-read from there in
//      every pass -- only the rest is streamed from L2;
//  (3) the slice is split evenly (no thread runs a second trip), UNROLL channel loads in flight per thread;
//  (4) SYNC == 0: contention-free exchange without atomics or fences.  Each CTA stores its packed
//      (distance, index) arg-max key into its OWN 128-byte slot; warp 0 of the item's rank-0 CTA (the leader)
//      polls the nb slots, takes the maximum, gathers the winning point's channels and publishes them as d
//      self-validating 64-bit words {pass tag, float bits} in a per-pass mailbox; all CTAs poll the mailbox
//      words they need.  Every word carries its own validity (non-zero key / matching tag), so no ordering
//      between stores is required, nothing is ever reset inside the kernel and batch items never wait for
//      each other.   SYNC == 1: atomicMax + counting grid barrier (first-generation protocol).
// Same canonical arithmetic as fps_kernel (bit-identical results).
template <int GPT, int MAXT, int UNROLL, int SYNC>
__global__ void __launch_bounds__(MAXT, (MAXT <= 320 ? 2 : 1)) fps2_kernel(FpsParams p, int nb, int chunk, int rg, unsigned long long* slots,
                                                       unsigned long long* mail) {
  extern __shared__ float4 xs4[];                       // [d][rg] resident slice, then d floats of the current seed
  float* s_seed = reinterpret_cast<float*>(xs4 + size_t(p.d) * rg);
  __shared__ unsigned long long s_red[32];
  __shared__ long long s_idx;
  __shared__ int s_fail;
  const int tid = threadIdx.x, T = blockDim.x;
  const int b = blockIdx.x / nb, rank = blockIdx.x % nb;
  const long long ngroups = p.n / 4;
  const long long g0 = (long long)rank * chunk;
  const long long g1 = (g0 + chunk < ngroups) ? g0 + chunk : ngroups;
  const float* Xb = p.X + b * p.sb;
  if (tid == 0) s_fail = 0;
  for (int lg = tid; lg < rg && g0 + lg < g1; lg += T) {
    const float* xp = Xb + (g0 + lg) * 4;
#pragma unroll 8
    for (int k = 0; k < p.d; ++k) xs4[size_t(k) * rg + lg] = __ldg(reinterpret_cast<const float4*>(xp + k * p.sd));
  }
  float r[GPT][4];
#pragma unroll
  for (int q = 0; q < GPT; ++q) { r[q][0] = r[q][1] = r[q][2] = r[q][3] = 0.f; }
  long long idx = p.first[b];
  for (int k = tid; k < p.d; k += T) s_seed[k] = __ldg(Xb + k * p.sd + idx);
  __syncthreads();
  for (int i = 0; i < p.m; ++i) {
    if (rank == 0) {
      if (tid == 0) p.selected_out[size_t(b) * p.m + i] = idx;
      for (int k = tid; k < p.d; k += T) p.seeds_out[(size_t(b) * p.m + i) * p.d + k] = s_seed[k];
    }
    if (i + 1 == p.m) break;
    if (p.trace && tid == 0) { if (int(blockIdx.x) == p.trace_cta) p.trace[i * 3 + 0] = clock64(); else if (p.trace_cta < 0 && i == 50) { unsigned int smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); p.trace[blockIdx.x * 8 + 0] = clock64(); p.trace[blockIdx.x * 8 + 3] = smid; } }
    unsigned long long best = 1ull;                      // non-zero sentinel: an empty slice still signals arrival
#pragma unroll
    for (int q = 0; q < GPT; ++q) {
      const int lg = tid + q * T;
      const long long g = g0 + lg;
      if (g < g1 && p.debug_mode != 2) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (lg < rg) {
          const float4* xp = xs4 + lg;
#pragma unroll UNROLL
          for (int k = 0; k < p.d; ++k) {
            const float4 v = xp[size_t(k) * rg];
            const float sk = s_seed[k];
            a0 = fmaf(v.x, sk, a0); a1 = fmaf(v.y, sk, a1); a2 = fmaf(v.z, sk, a2); a3 = fmaf(v.w, sk, a3);
          }
        } else {
          const float* xp = Xb + g * 4;
#pragma unroll UNROLL
          for (int k = 0; k < p.d; ++k) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xp + k * p.sd));
            const float sk = s_seed[k];
            a0 = fmaf(v.x, sk, a0); a1 = fmaf(v.y, sk, a1); a2 = fmaf(v.z, sk, a2); a3 = fmaf(v.w, sk, a3);
          }
        }
        const float d0 = 0.5f * (1.0f - a0), d1 = 0.5f * (1.0f - a1), d2 = 0.5f * (1.0f - a2), d3 = 0.5f * (1.0f - a3);
        if (i == 0) {
          r[q][0] = d0; r[q][1] = d1; r[q][2] = d2; r[q][3] = d3;
        } else {
          r[q][0] = d0 < r[q][0] ? d0 : r[q][0];
          r[q][1] = d1 < r[q][1] ? d1 : r[q][1];
          r[q][2] = d2 < r[q][2] ? d2 : r[q][2];
          r[q][3] = d3 < r[q][3] ? d3 : r[q][3];
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const unsigned long long key = pack_key(r[q][jj], static_cast<unsigned int>(g * 4 + jj));
          best = key > best ? key : best;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if ((tid & 31) == 0) s_red[tid >> 5] = best;
    __syncthreads();                                     // also: everybody is done reading s_seed of this pass
    if (p.trace && tid == 0) { if (int(blockIdx.x) == p.trace_cta) p.trace[i * 3 + 1] = clock64(); else if (p.trace_cta < 0 && i == 50) p.trace[blockIdx.x * 8 + 1] = clock64(); }
    if (tid < 32) {
      unsigned long long v = (tid < (T >> 5)) ? s_red[tid] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
      }
      if (SYNC == 0) {
        unsigned long long* row = slots + (size_t(b) * p.m + (i + 1)) * nb * 16;     // one 128-byte line per CTA
        if (tid == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(row + size_t(rank) * 16), "l"(v) : "memory");
        if (rank == 0) {
          // leader: wait for every CTA of the item, arg-max, gather the winner's channels, publish
          unsigned long long gmax = 0ull;
          bool done = false;
          for (unsigned int it = 0; it < (1u << 22) && !done; ++it) {
            unsigned long long kv[5];
#pragma unroll
            for (int sidx = 0; sidx < 5; ++sidx) {
              const int c = tid + 32 * sidx;
              kv[sidx] = 1ull;
              if (c < nb) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(kv[sidx]) : "l"(row + size_t(c) * 16));
            }
            bool all = true;
            gmax = 0ull;
#pragma unroll
            for (int sidx = 0; sidx < 5; ++sidx) {
              all = all && (kv[sidx] != 0ull);
              gmax = kv[sidx] > gmax ? kv[sidx] : gmax;
            }
            done = __all_sync(0xffffffffu, all);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, gmax, o);
            gmax = other > gmax ? other : gmax;
          }
          const long long widx = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(gmax & 0xFFFFFFFFull));
          if (tid == 0) s_idx = widx;
          unsigned long long* mrow = mail + (size_t(b) * p.m + (i + 1)) * p.d;
          const unsigned long long tag = done ? (static_cast<unsigned long long>(i + 1) << 32)
                                              : (0xFFFFFFFFull << 32);            // poison tag: readers give up too
          for (int k = tid; k < p.d; k += 32) {
            const float val = done ? __ldg(Xb + k * p.sd + widx) : 0.f;
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mrow + k), "l"(tag | __float_as_uint(val)) : "memory");
          }
          if (!done && tid == 0) atomicOr(p.err, ERR_GRID_BARRIER_TIMEOUT);
        }
      } else if (SYNC == 2 && p.debug_mode == 1) {
        if (tid == 0) s_idx = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(v & 0xFFFFFFFFull)) % p.n;
      } else if (SYNC == 2) {
        // all-to-all: this CTA's key goes into column `rank` of EVERY CTA's private row; then poll the own row
        unsigned long long* mat = slots + (size_t(b) * p.m + (i + 1)) * nb * nb;
#pragma unroll
        for (int sidx = 0; sidx < 5; ++sidx) {
          const int c = tid + 32 * sidx;
          if (c < nb) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(mat + size_t(c) * nb + rank), "l"(v) : "memory");
        }
        const unsigned long long* row = mat + size_t(rank) * nb;
        if (p.trace && p.trace_cta < 0 && i == 50 && tid == 0) p.trace[blockIdx.x * 8 + 4] = clock64();
        unsigned long long gmax = 0ull;
        bool done = false;
        for (unsigned int it = 0; it < (1u << 22) && !done; ++it) {
          unsigned long long kv[5];
#pragma unroll
          for (int sidx = 0; sidx < 5; ++sidx) {
            const int c = tid + 32 * sidx;
            kv[sidx] = 1ull;
            if (c < nb) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(kv[sidx]) : "l"(row + c));
          }
          bool all = true;
          gmax = 0ull;
#pragma unroll
          for (int sidx = 0; sidx < 5; ++sidx) {
            all = all && (kv[sidx] != 0ull);
            gmax = kv[sidx] > gmax ? kv[sidx] : gmax;
          }
          done = __all_sync(0xffffffffu, all);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, gmax, o);
          gmax = other > gmax ? other : gmax;
        }
        if (tid == 0) {
          s_idx = done ? static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(gmax & 0xFFFFFFFFull)) : -1;
          if (!done) atomicOr(p.err, ERR_GRID_BARRIER_TIMEOUT);
          if (p.trace && p.trace_cta < 0 && i == 50) p.trace[blockIdx.x * 8 + 5] = clock64();
        }
      } else {
        if (tid == 0) atomicMax(p.keys + size_t(b) * p.m + i + 1, v);
      }
    }
    if (SYNC == 2) {
      __syncthreads();
      idx = s_idx;
      if (idx < 0) return;
      for (int k = tid; k < p.d; k += T) s_seed[k] = __ldg(Xb + k * p.sd + idx);
      __syncthreads();
      if (p.trace && tid == 0) { if (int(blockIdx.x) == p.trace_cta) p.trace[i * 3 + 2] = clock64(); else if (p.trace_cta < 0 && i == 50) p.trace[blockIdx.x * 8 + 2] = clock64(); }
    } else if (SYNC == 0) {
      // every CTA: fetch the new seed from the mailbox (each word validates itself by its tag)
      const unsigned long long* mrow = mail + (size_t(b) * p.m + (i + 1)) * p.d;
      for (int k = tid; k < p.d; k += T) {
        unsigned long long w = 0ull;
        unsigned int it = 0;
        for (; it < (1u << 22); ++it) {
          asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(mrow + k));
          if (static_cast<unsigned int>(w >> 32) == static_cast<unsigned int>(i + 1)) break;
          if (static_cast<unsigned int>(w >> 32) == 0xFFFFFFFFu) { it = 1u << 22; break; }
        }
        if (it >= (1u << 22)) s_fail = 1;
        s_seed[k] = __uint_as_float(static_cast<unsigned int>(w & 0xFFFFFFFFull));
      }
      __syncthreads();
      if (s_fail) { if (tid == 0) atomicOr(p.err, ERR_GRID_BARRIER_TIMEOUT); return; }
      idx = s_idx;                                       // meaningful on the leader only (it alone writes the outputs)
      if (p.trace && tid == 0) { if (int(blockIdx.x) == p.trace_cta) p.trace[i * 3 + 2] = clock64(); else if (p.trace_cta < 0 && i == 50) p.trace[blockIdx.x * 8 + 2] = clock64(); }
    } else {
      if (!grid_barrier(p.barrier, (unsigned int)(i + 1) * gridDim.x, p.err)) return;
      if (tid == 0) {
        const unsigned long long key = __ldcg(p.keys + size_t(b) * p.m + i + 1);
        s_idx = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned int>(key & 0xFFFFFFFFull));
      }
      __syncthreads();
      idx = s_idx;
      for (int k = tid; k < p.d; k += T) s_seed[k] = __ldg(Xb + k * p.sd + idx);
      __syncthreads();
      if (p.trace && tid == 0) { if (int(blockIdx.x) == p.trace_cta) p.trace[i * 3 + 2] = clock64(); else if (p.trace_cta < 0 && i == 50) p.trace[blockIdx.x * 8 + 2] = clock64(); }
    }
  }
}

static int launch_select_seeds_v2(FpsParams p, const ClusterShape& s, unsigned long long* slots, size_t slot_bytes,
                                  cudaStream_t stream, bool* used) {
  *used = false;
  const int sms = sm_count();
  if (sms <= 0 || s.batch > sms) return UOC_OK;
  const int nb = sms / s.batch;
  if (nb > 160) return UOC_OK;                      // the leader's poll loop reads at most 5 slots per lane
  const long long ngroups = s.n / 4;
  const long long chunk_ll = (ngroups + nb - 1) / nb;
  if (chunk_ll > 4096) return UOC_OK;
  const int chunk = int(chunk_ll);
  constexpr int variant = 2;   // all-to-all key exchange (0: leader + mailbox and 1: atomicMax + grid barrier were measured slower)
  const size_t slot_need = (variant == 2) ? size_t(s.batch) * s.m * nb * nb * 8 : size_t(s.batch) * s.m * nb * 128;
  const size_t mail_need = size_t(s.batch) * s.m * s.d * 8;
  if (slot_need + mail_need > slot_bytes) return UOC_OK;
  unsigned long long* mail = slots + slot_need / 8;
  // Footprint: the pass time is dominated by the inter-CTA exchange latency, not by the streaming loop, so the default
  // is a LIGHT CTA (<= 320 threads, <= 96 KB shared memory) that leaves room on every SM for the kernels of another
  // frame running on a second stream (see pipeline.py).  UOC_FPS_HEAVY=1 selects the widest configuration instead.
  int gpt_t, maxt;
  if (chunk <= 576) { gpt_t = 1; maxt = 576; }
  else if (chunk <= 1024) { gpt_t = 1; maxt = 1024; }
  else if (chunk <= 2048) { gpt_t = 2; maxt = 1024; }
  else { gpt_t = 4; maxt = 1024; }
  int threads = ((chunk + gpt_t - 1) / gpt_t + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  const size_t budget = 96 * 1024;
  int rg = int(budget / (16 * size_t(s.d)));
  if (rg > chunk) rg = chunk;
  const size_t smem = size_t(rg) * s.d * 16 + size_t(s.d) * 4 + 16;
  void* kern;
#define UOC_FPS_PICK(G, MT, U) reinterpret_cast<void*>(&fps2_kernel<G, MT, U, 2>)
  if (maxt == 576) kern = UOC_FPS_PICK(1, 576, 16);
  else if (gpt_t == 1) kern = UOC_FPS_PICK(1, 1024, 8);
  else if (gpt_t == 2) kern = UOC_FPS_PICK(2, 1024, 8);
  else kern = UOC_FPS_PICK(4, 1024, 8);
#undef UOC_FPS_PICK
  UOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));   // per device, cheap
  if (variant == 0 || variant == 2) {
    UOC_CUDA(cudaMemsetAsync(slots, 0, slot_need + (variant == 0 ? mail_need : 0), stream));
  } else {
    UOC_CUDA(cudaMemsetAsync(p.keys, 0, sizeof(unsigned long long) * size_t(s.batch) * s.m, stream));
    UOC_CUDA(cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), stream));
  }
  int nb_i = nb, chunk_i = chunk, rg_i = rg;
  void* args[] = {&p, &nb_i, &chunk_i, &rg_i, &slots, &mail};
  UOC_CUDA(cudaLaunchCooperativeKernel(kern, dim3(nb * s.batch), dim3(threads), args, smem, stream));
  count_launch();
  *used = true;
  return UOC_OK;
}

// the generic cooperative fp32 sampler (any d, both metrics, optional seeds chosen before)
static int launch_fps_generic(FpsParams p, const ClusterShape& s, bool vec4, int metric, cudaStream_t stream) {
  void* kern = vec4 ? reinterpret_cast<void*>(&fps_kernel<4, METRIC_COSINE>) : reinterpret_cast<void*>(&fps_kernel<1, METRIC_COSINE>);
  if (metric == METRIC_EUCLIDEAN)
    kern = vec4 ? reinterpret_cast<void*>(&fps_kernel<4, METRIC_EUCLIDEAN>) : reinterpret_cast<void*>(&fps_kernel<1, METRIC_EUCLIDEAN>);
  const int threads = 256;
  const size_t smem = sizeof(float) * size_t(s.d);
  int per_sm = 0;
  UOC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return fail(UOC_ERR_CUDA, "fps kernel does not fit on an SM");
  int want = 2;
  if (want > per_sm) want = per_sm;
  const int sms = sm_count();
  long long groups_total = (s.n / (vec4 ? 4 : 1)) * (long long)s.batch;
  int grid = sms * want;
  long long needed = (groups_total + threads - 1) / threads;
  if (needed < grid) grid = int(needed < 1 ? 1 : needed);
  if (grid < s.batch && s.batch <= sms * per_sm) grid = s.batch;
  void* args[] = {&p};
  UOC_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(threads), args, smem, stream));
  count_launch();
  return UOC_OK;
}

int launch_select_seeds(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w,
                        int64_t* selected_out, float* seeds_out, cudaStream_t stream, int metric) {
  if (xb && metric == METRIC_COSINE) {
    bool used = false;
    int rc = launch_select_seeds_tc(X, xb, s, w, selected_out, seeds_out, stream, &used);
    if (rc != UOC_OK || used) return rc;
    if (s.batch > 1) {
      // The resident-slice sampler keeps a whole field on chip (tensor + shared memory of all SMs); a batch of large
      // fields does not fit at once, so the fields take turns, each at full width (same kernel, same indices).
      ClusterShape s1 = s;
      s1.batch = 1;
      int done = 0;
      for (int b = 0; b < s.batch; ++b, ++done) {
        ClusterWorkspace w1 = w;
        w1.first = w.first + b;
        rc = launch_select_seeds_tc(X + b * s.stride_b, xb + size_t(b) * s.n * s.d, s1, w1, selected_out + size_t(b) * s.m,
                                    seeds_out + size_t(b) * s.m * s.d, stream, &used);
        if (rc != UOC_OK) return rc;
        if (!used) break;
      }
      if (done == s.batch) return UOC_OK;
    }
  }
  FpsParams p;
  p.X = X; p.sb = s.stride_b; p.sd = s.stride_d; p.n = s.n; p.d = s.d; p.m = s.m; p.batch = s.batch;
  p.first = w.first; p.r = w.r; p.keys = w.keys; p.barrier = w.barrier;
  p.selected_out = reinterpret_cast<long long*>(selected_out);
  p.seeds_out = seeds_out;
  p.err = device_error_word();
  if (!p.err) return fail(UOC_ERR_CUDA, "no device error word");
  p.trace = nullptr;
  p.trace_cta = 0;
  p.debug_mode = 0;
  UOC_CUDA(cudaMemsetAsync(w.keys, 0, sizeof(unsigned long long) * size_t(s.batch) * s.m, stream));
  UOC_CUDA(cudaMemsetAsync(w.barrier, 0, sizeof(unsigned int), stream));
  const bool vec4 = (s.n % 4 == 0) && (s.stride_d % 4 == 0) && (s.stride_b % 4 == 0) &&
                    (reinterpret_cast<uintptr_t>(X) % 16 == 0);
  if (vec4 && metric == METRIC_COSINE) {
    bool used = false;
    int rc = launch_select_seeds_v2(p, s, w.slots, w.slot_bytes, stream, &used);
    if (rc != UOC_OK || used) return rc;
  }
  return launch_fps_generic(p, s, vec4, metric, stream);
}

// select_smart_seeds with init_seeds (mean_shift.py:144-149, :164-169): the first num_init seeds are given as vectors
int launch_select_seeds_init(const float* X, const ClusterShape& s, const ClusterWorkspace& w, const float* init_seeds,
                             int num_init, int64_t* selected_out, float* seeds_out, cudaStream_t stream, int metric) {
  if (num_init < 1 || num_init > s.m || !init_seeds) return fail(UOC_ERR_INVALID, "init_seeds: 1 <= num_init_seeds <= num_seeds");
  FpsParams p;
  p.X = X; p.sb = s.stride_b; p.sd = s.stride_d; p.n = s.n; p.d = s.d; p.m = s.m; p.batch = s.batch;
  p.first = w.first; p.r = w.r; p.keys = w.keys; p.barrier = w.barrier;
  p.selected_out = reinterpret_cast<long long*>(selected_out);
  p.seeds_out = seeds_out;
  p.err = device_error_word();
  if (!p.err) return fail(UOC_ERR_CUDA, "no device error word");
  p.trace = nullptr;
  p.trace_cta = 0;
  p.debug_mode = 0;
  p.init_seeds = init_seeds;
  p.num_init = num_init;
  UOC_CUDA(cudaMemsetAsync(w.keys, 0, sizeof(unsigned long long) * size_t(s.batch) * s.m, stream));
  UOC_CUDA(cudaMemsetAsync(w.barrier, 0, sizeof(unsigned int), stream));
  const bool vec4 = (s.n % 4 == 0) && (s.stride_d % 4 == 0) && (s.stride_b % 4 == 0) &&
                    (reinterpret_cast<uintptr_t>(X) % 16 == 0);
  return launch_fps_generic(p, s, vec4, metric, stream);
}

// ----------------------------------------------------------------------------------------------
// K4': fp32 SIMT mean-shift iteration (validation path)
// ----------------------------------------------------------------------------------------------
// grid (P, m, batch); each block accumulates sum_p exp(kappa * x_p.z_j) * x_p over its point slice.
// METRIC_EUCLIDEAN: weights exp(-kappa ||x_p - z_j||^2) (mean_shift.py:21-24), their sum goes to wsum[b][part][j].
template <int METRIC>
__global__ void __launch_bounds__(256) meanshift_simt_kernel(const float* __restrict__ X, long long sb, long long sd,
                                                             long long n, int d, int m, const float* __restrict__ Z,
                                                             float kappa, float* __restrict__ partials, int P,
                                                             float* __restrict__ wsum) {
  extern __shared__ float sm[];  // z[d] + red[8][32]
  float* z = sm;
  float* red = sm + d;
  const int part = blockIdx.x, j = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const float* Xb = X + b * sb;
  for (int k = tid; k < d; k += blockDim.x) z[k] = Z[(size_t(b) * m + j) * d + k];
  __syncthreads();
  const long long chunk = (n + P - 1) / P;
  const long long p0 = part * chunk;
  const long long p1 = (p0 + chunk < n) ? p0 + chunk : n;
  float* out = partials + ((size_t(b) * P + part) * 128 + j) * d;
  for (int kc = 0; kc < d; kc += 32) {
    float acc[32];
    float wtot = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) acc[q] = 0.f;
    for (long long pnt = p0 + tid; pnt < p1; pnt += blockDim.x) {
      float s = 0.f;
      float wgt;
      if (METRIC == METRIC_EUCLIDEAN) {
        for (int k = 0; k < d; ++k) { const float t = __ldg(Xb + k * sd + pnt) - z[k]; s = fmaf(t, t, s); }
        wgt = expf(-kappa * s);
        wtot += wgt;
      } else {
        for (int k = 0; k < d; ++k) s = fmaf(__ldg(Xb + k * sd + pnt), z[k], s);
        wgt = expf(kappa * s);
      }
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (kc + q < d) acc[q] = fmaf(wgt, __ldg(Xb + (kc + q) * sd + pnt), acc[q]);
    }
    if (METRIC == METRIC_EUCLIDEAN && kc == 0) {      // block sum of the weights, fixed order
      __shared__ float s_w[8];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wtot += __shfl_xor_sync(0xffffffffu, wtot, o);
      if ((tid & 31) == 0) s_w[tid >> 5] = wtot;
      __syncthreads();
      if (tid == 0) {
        float v = 0.f;
        for (int wq = 0; wq < int(blockDim.x >> 5); ++wq) v += s_w[wq];
        wsum[(size_t(b) * P + part) * 128 + j] = v;
      }
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) {
      float v = acc[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((tid & 31) == 0) red[(tid >> 5) * 32 + q] = v;
    }
    __syncthreads();
    if (tid < 32 && kc + tid < d) {
      float v = 0.f;
      for (int wq = 0; wq < int(blockDim.x >> 5); ++wq) v += red[wq * 32 + tid];
      out[kc + tid] = v;
    }
    __syncthreads();
  }
}

// grid (m, batch), block 256 (8 warps): Z[b][j][:] = normalize(sum_part partials[b][part][j][:]).
// Warp w adds parts w, w+8, w+16, ... (each a coalesced row read), the 8 warp sums are combined in a
// fixed order -> deterministic.  Lane l owns channels l, l+32, ... (d <= 256).
__global__ void __launch_bounds__(256) reduce_normalize_kernel(const float* __restrict__ partials, int P, int m, int d,
                                                               int row_stride, float* __restrict__ Z,
                                                               const float* __restrict__ wsum) {
  __shared__ float s_part[8][256];
  __shared__ float s_sq[8];
  asm volatile("griddepcontrol.wait;" ::: "memory");   // programmatic dependent launch: partials are complete
  const int j = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float v[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = 0.f;
  for (int part = warp; part < P; part += 8) {
    const float* src = partials + ((size_t(b) * P + part) * row_stride + j) * d;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (lane + 32 * q < d) v[q] += __ldcg(src + lane + 32 * q);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) s_part[warp][lane + 32 * q] = v[q];
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  float tot = 0.f;
  if (tid < d) {
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_part[w][tid];
  }
  float sq = tot * tot;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0) s_sq[warp] = sq;
  __syncthreads();
  float all = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) all += s_sq[w];
  float denom = fmaxf(sqrtf(all), 1e-12f);  // F.normalize eps (lib/utils/mean_shift.py:107)
  if (wsum) {                                // euclidean: Z = new_Z / clamp(sum of weights, min=1)  (mean_shift.py:101-105)
    float sw = 0.f;
    for (int part = 0; part < P; ++part) sw += __ldcg(wsum + (size_t(b) * P + part) * row_stride + j);
    denom = fmaxf(sw, 1.0f);
  }
  if (tid < d) Z[(size_t(b) * m + j) * d + tid] = tot / denom;
}

int launch_reduce_normalize(const float* partials, int batch, int P, int m, int d, int row_stride, float* Z,
                            cudaStream_t stream, const float* wsum) {
  if (d > 256) return fail(UOC_ERR_UNSUPPORTED, "d > 256 is not supported");
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(m, batch);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  UOC_CUDA(cudaLaunchKernelEx(&cfg, reduce_normalize_kernel, partials, P, m, d, row_stride, Z, wsum));
  count_launch();
  return UOC_OK;
}

int launch_hill_climb_simt(const float* X, const ClusterShape& s, const ClusterWorkspace& w, float* Z, float kappa,
                           int iters, cudaStream_t stream, int metric) {
  int P = w.max_partials;
  const long long min_pts = 2048;
  long long maxP = (s.n + min_pts - 1) / min_pts;
  if (P > maxP) P = int(maxP);
  if (P < 1) P = 1;
  const size_t smem = sizeof(float) * (size_t(s.d) + 8 * 32);
  for (int it = 0; it < iters; ++it) {
    if (metric == METRIC_EUCLIDEAN)
      meanshift_simt_kernel<METRIC_EUCLIDEAN><<<dim3(P, s.m, s.batch), 256, smem, stream>>>(
          X, s.stride_b, s.stride_d, s.n, s.d, s.m, Z, kappa, w.partials, P, w.wsum);
    else
      meanshift_simt_kernel<METRIC_COSINE><<<dim3(P, s.m, s.batch), 256, smem, stream>>>(
          X, s.stride_b, s.stride_d, s.n, s.d, s.m, Z, kappa, w.partials, P, nullptr);
    UOC_CHECK_LAUNCH();
    int rc = launch_reduce_normalize(w.partials, s.batch, P, s.m, s.d, 128, Z, stream,
                                     metric == METRIC_EUCLIDEAN ? w.wsum : nullptr);
    if (rc != UOC_OK) return rc;
  }
  return UOC_OK;
}

// ----------------------------------------------------------------------------------------------
// K5a: greedy seed labelling, one CTA (512 threads) per field
// ----------------------------------------------------------------------------------------------
template <int DREG>   // channels held in registers (0 = generic d)
__global__ void __launch_bounds__(512) label_seeds_kernel(const float* __restrict__ Z, int m, int d, float eps, int metric,
                                                          int* __restrict__ seed_labels, int* __restrict__ num_unique) {
  extern __shared__ __align__(16) float zs[];  // [m][d+4]: rows 16-byte aligned (float4 broadcast reads), 4-way bank conflicts at most on the per-lane row reads
  __shared__ unsigned int adj[UOC_MAX_SEEDS][4];
  __shared__ int labels[UOC_MAX_SEEDS];
  __shared__ int cnt[UOC_MAX_SEEDS];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int ld = d + 4;
  const float* Zb = Z + size_t(b) * m * d;
  for (int e = tid; e < m * d; e += blockDim.x) zs[(e / d) * ld + (e % d)] = Zb[e];
  if (tid < UOC_MAX_SEEDS) { labels[tid] = -1; cnt[tid] = 0; }
  __syncthreads();
  // adjacency: bit j of adj[i] <=> 0.5 * (1 - z_j . z_i) <= eps      (mean_shift.py:62-63)
  // thread (j, iq): row j held in registers (DREG channels at a time), z_i read as smem broadcasts
  {
    const int j = tid & 127, iq = tid >> 7, w4 = (tid >> 5) & 3;
    if (DREG > 0) {
      float zj[DREG > 0 ? DREG : 1];
#pragma unroll
      for (int k = 0; k < DREG; ++k) zj[k] = (j < m) ? zs[j * ld + k] : 0.f;
      if (metric == METRIC_EUCLIDEAN) {        // ||z_j - z_i|| <= eps  (mean_shift.py:58-60)
        for (int i = iq; i < m; i += 4) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < DREG; ++k) { const float t = zj[k] - zs[i * ld + k]; acc = fmaf(t, t, acc); }
          const bool in = (j < m) && (sqrtf(acc) <= eps);
          const unsigned int bits = __ballot_sync(0xffffffffu, in);
          if ((tid & 31) == 0) adj[i][w4] = bits;
        }
      } else {
        for (int i = iq; i < m; i += 4) {
          float acc = 0.f;
          const float4* zi = reinterpret_cast<const float4*>(zs + i * ld);     // broadcast reads, same fmaf order as before
#pragma unroll
          for (int k4 = 0; k4 < DREG / 4; ++k4) {
            const float4 z4 = zi[k4];
            acc = fmaf(zj[4 * k4 + 0], z4.x, acc);
            acc = fmaf(zj[4 * k4 + 1], z4.y, acc);
            acc = fmaf(zj[4 * k4 + 2], z4.z, acc);
            acc = fmaf(zj[4 * k4 + 3], z4.w, acc);
          }
          const bool in = (j < m) && ((0.5f * (1.0f - acc)) <= eps);
          const unsigned int bits = __ballot_sync(0xffffffffu, in);
          if ((tid & 31) == 0) adj[i][w4] = bits;
        }
      }
    } else {
      for (int i = iq; i < m; i += 4) {
        bool in = false;
        if (j < m) {
          float acc = 0.f;
          if (metric == METRIC_EUCLIDEAN) {
            for (int k = 0; k < d; ++k) { const float t = zs[j * ld + k] - zs[i * ld + k]; acc = fmaf(t, t, acc); }
            in = sqrtf(acc) <= eps;
          } else {
            for (int k = 0; k < d; ++k) acc = fmaf(zs[j * ld + k], zs[i * ld + k], acc);
            in = (0.5f * (1.0f - acc)) <= eps;
          }
        }
        const unsigned int bits = __ballot_sync(0xffffffffu, in);
        if ((tid & 31) == 0) adj[i][w4] = bits;
      }
    }
  }
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    int K = 0;
    for (int i = 0; i < m; ++i) {
      if (labels[i] != -1) continue;
      for (int l = lane; l < K; l += 32) cnt[l] = 0;
      __syncwarp();
      bool any = false;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        const bool in = (j < m) && ((adj[i][q] >> lane) & 1u);
        if (in && labels[j] != -1) { atomicAdd(&cnt[labels[j]], 1); any = true; }
      }
      any = __any_sync(0xffffffffu, any);
      __syncwarp();
      int lab;
      if (any) {
        // mode of the labelled members, ties -> smallest label (mean_shift.py:30-38)
        unsigned int bestkey = 0;
        for (int l = lane; l < K; l += 32) {
          const unsigned int key = (static_cast<unsigned int>(cnt[l]) << 8) | static_cast<unsigned int>(255 - l);
          bestkey = key > bestkey ? key : bestkey;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned int other = __shfl_xor_sync(0xffffffffu, bestkey, o);
          bestkey = other > bestkey ? other : bestkey;
        }
        lab = 255 - int(bestkey & 0xFFu);
      } else {
        lab = K;
        K += 1;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = lane + 32 * q;
        if ((j < m) && ((adj[i][q] >> lane) & 1u)) labels[j] = lab;   // overwrite, labelled or not (:74)
      }
      __syncwarp();
    }
    // len(unique(labels))  (mean_shift.py:218)
    for (int l = lane; l < UOC_MAX_SEEDS; l += 32) cnt[l] = 0;
    __syncwarp();
    for (int j = lane; j < m; j += 32) cnt[labels[j]] = 1;
    __syncwarp();
    int c = 0;
    for (int l = lane; l < UOC_MAX_SEEDS; l += 32) c += cnt[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) num_unique[b] = c;
    for (int j = lane; j < m; j += 32) seed_labels[size_t(b) * m + j] = labels[j];
  }
}

int launch_label_seeds(const float* Z, int batch, int m, int d, float epsilon, int* seed_labels, int* num_unique,
                       cudaStream_t stream, int metric) {
  const size_t smem = sizeof(float) * size_t(m) * (d + 4);
  {
    const void* kern = d == 64 ? reinterpret_cast<const void*>(&label_seeds_kernel<64>)
                     : d == 128 ? reinterpret_cast<const void*>(&label_seeds_kernel<128>)
                                : reinterpret_cast<const void*>(&label_seeds_kernel<0>);
    const int rc_attr = ensure_dynamic_smem(kern, 160 * 1024);
    if (rc_attr != UOC_OK) return rc_attr;
  }
  if (d == 64) label_seeds_kernel<64><<<batch, 512, smem, stream>>>(Z, m, d, epsilon, metric, seed_labels, num_unique);
  else if (d == 128) label_seeds_kernel<128><<<batch, 512, smem, stream>>>(Z, m, d, epsilon, metric, seed_labels, num_unique);
  else label_seeds_kernel<0><<<batch, 512, smem, stream>>>(Z, m, d, epsilon, metric, seed_labels, num_unique);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

// ----------------------------------------------------------------------------------------------
// K5b: nearest-seed assignment + histogram, then the label-0 swap
// ----------------------------------------------------------------------------------------------
template <int D>  // D > 0: channels held in registers; D == 0: generic (re-reads x through L1)
__global__ void __launch_bounds__(128) assign_kernel(const float* __restrict__ X, long long sb, long long sd, long long n,
                                                     int d, int m, int metric, const float* __restrict__ Z,
                                                     const int* __restrict__ seed_labels, int* __restrict__ hist,
                                                     int* __restrict__ labels_tmp) {
  extern __shared__ float zs[];  // [m][d]
  __shared__ int s_lab[UOC_MAX_SEEDS];
  __shared__ int s_hist[UOC_MAX_SEEDS];
  const int b = blockIdx.y, tid = threadIdx.x;
  const float* Zb = Z + size_t(b) * m * d;
  for (int e = tid; e < m * d; e += blockDim.x) zs[e] = Zb[e];
  if (tid < UOC_MAX_SEEDS) { s_lab[tid] = tid < m ? seed_labels[size_t(b) * m + tid] : 0; s_hist[tid] = 0; }
  __syncthreads();
  const float* Xb = X + b * sb;
  const long long pnt = (long long)blockIdx.x * blockDim.x + tid;
  int label = -1;
  if (pnt < n) {
    float best = 0.f;
    int bj = 0;
    if (D > 0) {
      float x[D > 0 ? D : 1];
#pragma unroll
      for (int k = 0; k < D; ++k) x[k] = __ldg(Xb + k * sd + pnt);
      for (int j = 0; j < m; ++j) {
        const float4* zj = reinterpret_cast<const float4*>(zs + j * D);
        float acc = 0.f;
        if (metric == METRIC_EUCLIDEAN) {      // ||x - z_j||  (mean_shift.py:207-209)
#pragma unroll
          for (int k4 = 0; k4 < D / 4; ++k4) {
            const float4 zv = zj[k4];
            const float t0 = x[4 * k4 + 0] - zv.x, t1 = x[4 * k4 + 1] - zv.y, t2 = x[4 * k4 + 2] - zv.z, t3 = x[4 * k4 + 3] - zv.w;
            acc = fmaf(t0, t0, acc);
            acc = fmaf(t1, t1, acc);
            acc = fmaf(t2, t2, acc);
            acc = fmaf(t3, t3, acc);
          }
        } else {
#pragma unroll
          for (int k4 = 0; k4 < D / 4; ++k4) {
            const float4 zv = zj[k4];
            acc = fmaf(x[4 * k4 + 0], zv.x, acc);
            acc = fmaf(x[4 * k4 + 1], zv.y, acc);
            acc = fmaf(x[4 * k4 + 2], zv.z, acc);
            acc = fmaf(x[4 * k4 + 3], zv.w, acc);
          }
        }
        const float dist = (metric == METRIC_EUCLIDEAN) ? sqrtf(acc) : 0.5f * (1.0f - acc);
        if (j == 0 || dist < best) { best = dist; bj = j; }   // first minimum (torch.argmin)
      }
    } else {
      for (int j = 0; j < m; ++j) {
        float acc = 0.f;
        if (metric == METRIC_EUCLIDEAN) {
          for (int k = 0; k < d; ++k) { const float t = __ldg(Xb + k * sd + pnt) - zs[j * d + k]; acc = fmaf(t, t, acc); }
        } else {
          for (int k = 0; k < d; ++k) acc = fmaf(__ldg(Xb + k * sd + pnt), zs[j * d + k], acc);
        }
        const float dist = (metric == METRIC_EUCLIDEAN) ? sqrtf(acc) : 0.5f * (1.0f - acc);
        if (j == 0 || dist < best) { best = dist; bj = j; }
      }
    }
    label = s_lab[bj];
    labels_tmp[size_t(b) * n + pnt] = label;
  }
  // warp-aggregated histogram
  const unsigned int active = __ballot_sync(0xffffffffu, label >= 0);
  if (label >= 0) {
    const unsigned int peers = __match_any_sync(active, label);
    if ((tid & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[label], __popc(peers));
  }
  __syncthreads();
  if (tid < UOC_MAX_SEEDS && s_hist[tid] != 0) atomicAdd(hist + size_t(b) * m + tid, s_hist[tid]);
}

// swap label 0 <-> argmax_i count[i], i in range(num_unique)   (mean_shift.py:217-227)
__global__ void __launch_bounds__(256) relabel_kernel(const int* __restrict__ labels_tmp, const int* __restrict__ hist,
                                                      const int* __restrict__ num_unique, long long n, int m,
                                                      int* __restrict__ labels_out, float* __restrict__ labels_f32_out,
                                                      unsigned char* __restrict__ labels_u8_out) {
  __shared__ int s_max;
  const int b = blockIdx.y;
  if (threadIdx.x == 0) {
    int num = num_unique[b];
    if (num > m) num = m;
    int best = 0, bi = 0;
    for (int i = 0; i < num; ++i) {
      const int c = hist[size_t(b) * m + i];
      if (i == 0 || c > best) { best = c; bi = i; }
    }
    s_max = bi;
  }
  __syncthreads();
  const int lm = s_max;
  const long long pnt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pnt < n) {
    int l = labels_tmp[size_t(b) * n + pnt];
    if (lm != 0) {
      if (l == 0) l = lm;
      else if (l == lm) l = 0;
    }
    labels_out[size_t(b) * n + pnt] = l;
    // the reference's API type (float32 label maps, test_dataset.py:48) and the wire type of the label all-gather
    // (uint8: ids < num_seeds <= 128) written by the same pass instead of separate conversion kernels
    if (labels_f32_out) labels_f32_out[size_t(b) * n + pnt] = float(l);
    if (labels_u8_out) labels_u8_out[size_t(b) * n + pnt] = static_cast<unsigned char>(l);
  }
}

int launch_assign(const float* X, const __nv_bfloat16* xb, const ClusterShape& s, const ClusterWorkspace& w, const float* Z,
                  const int* seed_labels, const int* num_unique, int* hist, int* labels_tmp, int* labels_out,
                  cudaStream_t stream, int metric, float* labels_f32_out, unsigned char* labels_u8_out) {
  UOC_CUDA(cudaMemsetAsync(hist, 0, sizeof(int) * size_t(s.batch) * s.m, stream));
  bool use_tc = xb != nullptr && (s.d == 64 || s.d == 128) && metric == METRIC_COSINE;
  if (knobs().assign_simt != 0) use_tc = false;
  if (use_tc) {
    int rc = launch_assign_tc(X, xb, s, w, Z, seed_labels, hist, labels_tmp, stream);
    if (rc != UOC_OK) return rc;
    const dim3 grid2(static_cast<unsigned int>((s.n + 255) / 256), s.batch);
    relabel_kernel<<<grid2, 256, 0, stream>>>(labels_tmp, hist, num_unique, s.n, s.m, labels_out, labels_f32_out, labels_u8_out);
    UOC_CHECK_LAUNCH();
    return UOC_OK;
  }
  const size_t smem = sizeof(float) * size_t(s.m) * s.d;
  const dim3 grid(static_cast<unsigned int>((s.n + 127) / 128), s.batch);
  {
    const void* kern = s.d == 64 ? reinterpret_cast<const void*>(&assign_kernel<64>)
                     : s.d == 128 ? reinterpret_cast<const void*>(&assign_kernel<128>)
                                  : reinterpret_cast<const void*>(&assign_kernel<0>);
    const int rc_attr = ensure_dynamic_smem(kern, (s.d == 64 || s.d == 128) ? 100 * 1024 : 160 * 1024);
    if (rc_attr != UOC_OK) return rc_attr;
  }
  if (s.d == 64)
    assign_kernel<64><<<grid, 128, smem, stream>>>(X, s.stride_b, s.stride_d, s.n, s.d, s.m, metric, Z, seed_labels, hist, labels_tmp);
  else if (s.d == 128)
    assign_kernel<128><<<grid, 128, smem, stream>>>(X, s.stride_b, s.stride_d, s.n, s.d, s.m, metric, Z, seed_labels, hist, labels_tmp);
  else
    assign_kernel<0><<<grid, 128, smem, stream>>>(X, s.stride_b, s.stride_d, s.n, s.d, s.m, metric, Z, seed_labels, hist, labels_tmp);
  UOC_CHECK_LAUNCH();
  const dim3 grid2(static_cast<unsigned int>((s.n + 255) / 256), s.batch);
  relabel_kernel<<<grid2, 256, 0, stream>>>(labels_tmp, hist, num_unique, s.n, s.m, labels_out, labels_f32_out, labels_u8_out);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

// ----------------------------------------------------------------------------------------------
// pack: fp32 planar [d][n] -> bf16 pixel-major [n][d]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_bf16_kernel(const float* __restrict__ X, long long sb, long long sd,
                                                        long long n, int d, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float tile[];  // [d][65]
  const int b = blockIdx.y, tid = threadIdx.x;
  const long long p0 = (long long)blockIdx.x * 64;
  const float* Xb = X + b * sb;
  for (int e = tid; e < d * 64; e += blockDim.x) {
    const int k = e >> 6, pp = e & 63;
    tile[k * 65 + pp] = (p0 + pp < n) ? __ldg(Xb + k * sd + p0 + pp) : 0.f;
  }
  __syncthreads();
  const int half = d >> 1;
  uint32_t* o32 = reinterpret_cast<uint32_t*>(out + (size_t(b) * n + p0) * d);
  for (int e = tid; e < half * 64; e += blockDim.x) {
    const int pp = e / half, k2 = e % half;
    if (p0 + pp < n) o32[size_t(pp) * half + k2] = pack_bf16x2(tile[(2 * k2) * 65 + pp], tile[(2 * k2 + 1) * 65 + pp]);
  }
}

int launch_pack_bf16(const float* X, const ClusterShape& s, __nv_bfloat16* xb, cudaStream_t stream) {
  if (s.d % 2 != 0) return fail(UOC_ERR_UNSUPPORTED, "d must be even");
  const size_t smem = sizeof(float) * size_t(s.d) * 65;
  if (smem > 48 * 1024) {
    const int rc_attr = ensure_dynamic_smem(reinterpret_cast<const void*>(&pack_bf16_kernel), int(smem));
    if (rc_attr != UOC_OK) return rc_attr;
  }
  const dim3 grid(static_cast<unsigned int>((s.n + 63) / 64), s.batch);
  pack_bf16_kernel<<<grid, 256, smem, stream>>>(X, s.stride_b, s.stride_d, s.n, s.d, xb);
  UOC_CHECK_LAUNCH();
  return UOC_OK;
}

}  // namespace uoc
