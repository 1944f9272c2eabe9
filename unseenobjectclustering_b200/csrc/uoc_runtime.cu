// Host-side runtime helpers shared by the UOC entry points: thread-local error string, device
// capability gate (sm_100 only -- there is no fallback), device error word, tensor-map encoding.
#include "uoc_common.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace uoc {

static thread_local std::string g_last_error;
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct KnobEntry { const char* name; const char* env; int Knobs::*field; };
static const KnobEntry kKnobTable[] = {
    {"conv_pair", "UOC_CONV_PAIR", &Knobs::conv_pair},          {"conv_wres", "UOC_CONV_WRES", &Knobs::conv_wres},
    {"conv_debug", "UOC_CONV_DEBUG", &Knobs::conv_debug},
    {"conv_trace", "UOC_CONV_TRACE", &Knobs::conv_trace},       {"fps_tc", "UOC_FPS_TC", &Knobs::fps_tc},
    {"fps_stream", "UOC_FPS_STREAM", &Knobs::fps_stream},       {"fps_tmem_tiles", "UOC_FPS_TC_TMEM_TILES", &Knobs::fps_tmem_tiles},
    {"fps_batch_stream", "UOC_FPS_BATCH_STREAM", &Knobs::fps_batch_stream},
    {"fps_rn_margin", "UOC_FPS_RN_MARGIN", &Knobs::fps_rn_margin}, {"fps_stats", "UOC_FPS_STATS", &Knobs::fps_stats},
    {"loop_trace", "UOC_LOOP_TRACE", &Knobs::loop_trace},       {"assign_simt", "UOC_ASSIGN_SIMT", &Knobs::assign_simt},
};
static Knobs& mutable_knobs() {
  static Knobs k = [] {
    Knobs v;
    for (const KnobEntry& e : kKnobTable)
      if (const char* s = getenv(e.env)) v.*(e.field) = atoi(s);
    return v;
  }();
  return k;
}
const Knobs& knobs() { return mutable_knobs(); }
int set_knob(const char* name, int value) {
  if (!name) return fail(UOC_ERR_INVALID, "null knob name");
  for (const KnobEntry& e : kKnobTable)
    if (std::string(name) == e.name || std::string(name) == e.env) { mutable_knobs().*(e.field) = value; return UOC_OK; }
  return fail(UOC_ERR_INVALID, std::string("unknown knob: ") + name);
}

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
  int sms = 0, major = 0, minor = 0;
};
static DevInfo g_devs[64];
static std::mutex g_mu;

static int query_device(DevInfo* out = nullptr) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
  if (dev < 0 || dev >= 64) return fail(UOC_ERR_UNSUPPORTED, "device ordinal out of range");
  std::lock_guard<std::mutex> lk(g_mu);
  DevInfo& d = g_devs[dev];
  if (d.ok < 0) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    d.sms = prop.multiProcessorCount;
    d.major = prop.major;
    d.minor = prop.minor;
    d.ok = (prop.major == 10) ? 1 : 0;
  }
  if (out) *out = d;
  return UOC_OK;
}

int require_sm100() {
  DevInfo g_dev;
  int rc = query_device(&g_dev);
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
  DevInfo d;
  if (query_device(&d) != UOC_OK) return 0;
  return d.sms;
}

// one error word per device (multi-GPU processes: DataParallel replicas, one handle per device)
static const int kMaxDevices = 64;
static unsigned int* g_err_word[kMaxDevices] = {nullptr};

unsigned int* device_error_word() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_err_word[dev] == nullptr) {
    unsigned int* p = nullptr;
    if (cudaMalloc(&p, sizeof(unsigned int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(unsigned int));
    g_err_word[dev] = p;  // one small allocation per device; intentionally never freed
  }
  return g_err_word[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember (device, kernel) pairs already set
int ensure_dynamic_smem(const void* func, int bytes) {
  int dev = -1;
  UOC_CUDA(cudaGetDevice(&dev));
  static std::map<std::pair<int, const void*>, int> done;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = done.find(std::make_pair(dev, func));
    if (it != done.end() && it->second >= bytes) return UOC_OK;
  }
  UOC_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  std::lock_guard<std::mutex> lk(g_mu);
  done[std::make_pair(dev, func)] = bytes;
  return UOC_OK;
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

// fp32 tensor map without swizzle (the stem's input patches: box rows land linearly in shared memory, OOB -> zero)
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail(UOC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1u; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, cuuint32_t(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rank %d, dims %llu %llu ..)", int(r), rank,
             (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0));
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
  uoc::DevInfo d;
  int rc = uoc::query_device(&d);
  if (rc != UOC_OK) return rc;
  if (sm_count_out) *sm_count_out = d.sms;
  if (cc_major) *cc_major = d.major;
  if (cc_minor) *cc_minor = d.minor;
  return uoc::require_sm100();
}


int uoc_set_knob(const char* name, int value) { return uoc::set_knob(name, value); }

int uoc_check_device_error(uoc_stream_t stream) {
  int rc = uoc::require_sm100();
  if (rc != UOC_OK) return rc;
  return uoc::check_device_error(static_cast<cudaStream_t>(stream));
}

int uoc_peek_device_error_async(uint32_t* word_host, uoc_stream_t stream) {
  int rc = uoc::require_sm100();
  if (rc != UOC_OK) return rc;
  if (!word_host) return uoc::fail(UOC_ERR_INVALID, "word_host is null");
  unsigned int* w = uoc::device_error_word();
  if (!w) return uoc::fail(UOC_ERR_CUDA, "could not allocate the device error word");
  UOC_CUDA(cudaMemcpyAsync(word_host, w, sizeof(uint32_t), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream)));
  return UOC_OK;
}

}  // extern "C"
