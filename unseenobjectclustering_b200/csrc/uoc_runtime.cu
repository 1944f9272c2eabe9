// Host-side runtime helpers shared by the UOC entry points: thread-local error string, device
// capability gate (sm_100 only -- there is no fallback), device error word, tensor-map encoding.
#include "uoc_common.cuh"

#include <atomic>
#include <cstdio>
#include <mutex>

namespace uoc {

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const std::string& msg) { g_last_error = msg; }

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[1024];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d in `%s`", int(e), cudaGetErrorString(e), file, line, what);
  g_last_error = buf;
  cudaGetLastError();  // clear the non-sticky part
  return UOC_ERR_CUDA;
}

struct DevInfo {
  int ok = -1;  // -1 unknown
  int device = -1;
  int sms = 0, major = 0, minor = 0;
};
static DevInfo g_dev;
static std::mutex g_mu;

static int query_device() {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
  if (g_dev.ok >= 0 && g_dev.device == dev) return UOC_OK;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__);
  g_dev.device = dev;
  g_dev.sms = prop.multiProcessorCount;
  g_dev.major = prop.major;
  g_dev.minor = prop.minor;
  g_dev.ok = (prop.major == 10) ? 1 : 0;
  return UOC_OK;
}

int require_sm100() {
  int rc = query_device();
  if (rc != UOC_OK) return rc;
  if (g_dev.ok != 1) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "this library is built for sm_100a (B200) only; current device is sm_%d%d and there is no fallback path",
             g_dev.major, g_dev.minor);
    return fail(UOC_ERR_UNSUPPORTED, buf);
  }
  return UOC_OK;
}

int sm_count() {
  if (query_device() != UOC_OK) return 0;
  return g_dev.sms;
}

static unsigned int* g_err_word = nullptr;
static int g_err_dev = -1;

unsigned int* device_error_word() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_err_word == nullptr || g_err_dev != dev) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, sizeof(unsigned int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(unsigned int));
    g_err_word = p;  // one small allocation per device change; intentionally never freed
    g_err_dev = dev;
  }
  return g_err_word;
}

int check_device_error(cudaStream_t stream) {
  unsigned int* w = device_error_word();
  if (!w) return fail(UOC_ERR_CUDA, "could not allocate the device error word");
  unsigned int h = 0;
  UOC_CUDA(cudaMemcpyAsync(&h, w, sizeof(h), cudaMemcpyDeviceToHost, stream));
  UOC_CUDA(cudaStreamSynchronize(stream));
  if (h != 0) {
    cudaMemsetAsync(w, 0, sizeof(unsigned int), stream);
    char buf[160];
    snprintf(buf, sizeof(buf), "device-side pipeline error word 0x%x (1=mbarrier time-out, 2=grid-barrier time-out, 4=bad config)", h);
    return fail(UOC_ERR_DEVICE, buf);
  }
  return UOC_OK;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    else cudaGetLastError();
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(UOC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1u;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu .., box %u %u ..)",
             int(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
             rank > 1 ? box[1] : 0);
    return fail(UOC_ERR_CUDA, buf);
  }
  return UOC_OK;
}

}  // namespace uoc

extern "C" {

const char* uoc_last_error(void) { return uoc::g_last_error.c_str(); }

int uoc_version(void) { return 100; }

unsigned long long uoc_launch_count(void) { return uoc::g_launches.load(); }

int uoc_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int rc = uoc::query_device();
  if (rc != UOC_OK) return rc;
  if (sm_count_out) *sm_count_out = uoc::g_dev.sms;
  if (cc_major) *cc_major = uoc::g_dev.major;
  if (cc_minor) *cc_minor = uoc::g_dev.minor;
  return uoc::require_sm100();
}

}  // extern "C"
