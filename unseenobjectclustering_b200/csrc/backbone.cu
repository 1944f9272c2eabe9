// Backbone handle + forward schedule + C-ABI (include/uoc.h): the two-branch ResNet34-8s of
// lib/networks/SEG.py:69-71,:88-119 / lib/networks/resnet_dilated.py:287-327 / lib/networks/resnet.py.
// Eval-mode BatchNorm is folded into the preceding convolution on the host at create time; weights
// are re-packed to bf16 [Cout][tap][Cin] (K-major rows for the implicit GEMM).
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "conv.cuh"

namespace uoc {

struct ConvLayer {
  int Cin = 0, Cout = 0, ksize = 1, stride = 1, dil = 1;
  size_t w_off = 0;     // byte offset of the bf16 weights in the device blob
  size_t b_off = 0;     // byte offset of the fp32 bias
};

struct Block {
  ConvLayer conv1, conv2, down;
  bool has_down = false;
};

struct Branch {
  size_t stem_w_off = 0, stem_b_off = 0, stem_wbf_off = 0;
  std::vector<Block> blocks;
  ConvLayer fc;
};

}  // namespace uoc

struct uoc_backbone {
  int num_units = 64;
  int input_type = UOC_INPUT_RGBD, fusion = UOC_FUSION_ADD, normalize = 1;
  int groups = 2;         // trunks evaluated side by side: 2 for RGBD add / cat, 1 for COLOR / DEPTH / early fusion
  int cin = 3;            // stem input channels: 6 for early fusion (SEG.py:101-103, :178-181)
  int feat_dim = 64;      // channels of the returned field: 2 * num_units for cat fusion (SEG.py:110)
  int device = -1;
  uoc::Branch br[2];
  char* blob = nullptr;   // device
  size_t blob_bytes = 0;
};

namespace uoc {

static uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return uint16_t(u >> 16);  // inf / nan
  u += 0x7FFFu + ((u >> 16) & 1u);
  return uint16_t(u >> 16);
}

typedef std::map<std::string, const uoc_weight_desc*> WeightMap;

static const float* find(const WeightMap& wm, const std::string& name, int64_t numel, std::string* err) {
  auto it = wm.find(name);
  if (it == wm.end()) { *err = "state_dict tensor missing: " + name; return nullptr; }
  if (it->second->numel != numel) {
    *err = "state_dict tensor " + name + " has " + std::to_string(it->second->numel) + " elements, expected " +
           std::to_string(numel);
    return nullptr;
  }
  return it->second->data_host;
}

// scale/shift of eval-mode BN:  y = x * scale + shift
static bool bn_fold(const WeightMap& wm, const std::string& p, int C, std::vector<float>* scale, std::vector<float>* shift,
                    std::string* err) {
  const float* g = find(wm, p + ".weight", C, err);
  const float* b = g ? find(wm, p + ".bias", C, err) : nullptr;
  const float* mu = b ? find(wm, p + ".running_mean", C, err) : nullptr;
  const float* var = mu ? find(wm, p + ".running_var", C, err) : nullptr;
  if (!var) return false;
  scale->resize(C);
  shift->resize(C);
  for (int c = 0; c < C; ++c) {
    const double s = double(g[c]) / std::sqrt(double(var[c]) + 1e-5);
    (*scale)[c] = float(s);
    (*shift)[c] = float(double(b[c]) - double(mu[c]) * s);
  }
  return true;
}

struct BlobBuilder {
  std::vector<char> host;
  size_t reserve(size_t bytes) {
    size_t at = align_up(host.size(), 256);
    host.resize(at + bytes);
    return at;
  }
};

// conv weight [Cout][Cin][k][k] (+ optional BN, + optional bias) -> bf16 [Cout][k*k][Cin], fp32 bias
static bool pack_conv(const WeightMap& wm, const std::string& wname, const std::string& bnname, const std::string& biasname,
                      ConvLayer* L, BlobBuilder* bb, std::string* err) {
  const int taps = L->ksize * L->ksize;
  const float* w = find(wm, wname, int64_t(L->Cout) * L->Cin * taps, err);
  if (!w) return false;
  std::vector<float> scale(L->Cout, 1.f), shift(L->Cout, 0.f);
  if (!bnname.empty() && !bn_fold(wm, bnname, L->Cout, &scale, &shift, err)) return false;
  if (!biasname.empty()) {
    const float* b = find(wm, biasname, L->Cout, err);
    if (!b) return false;
    for (int c = 0; c < L->Cout; ++c) shift[c] += b[c] * scale[c];
  }
  L->w_off = bb->reserve(size_t(L->Cout) * taps * L->Cin * 2);
  L->b_off = bb->reserve(size_t(L->Cout) * 4);
  uint16_t* wd = reinterpret_cast<uint16_t*>(bb->host.data() + L->w_off);
  float* bd = reinterpret_cast<float*>(bb->host.data() + L->b_off);
  for (int co = 0; co < L->Cout; ++co) {
    for (int t = 0; t < taps; ++t)
      for (int ci = 0; ci < L->Cin; ++ci)
        wd[(size_t(co) * taps + t) * L->Cin + ci] = f32_to_bf16_rne(w[(size_t(co) * L->Cin + ci) * taps + t] * scale[co]);
    bd[co] = shift[co];
  }
  return true;
}

static const int kPlanes[4] = {64, 128, 256, 512};
static const int kBlocks[4] = {3, 4, 6, 3};
static const int kStride[4] = {1, 2, 1, 1};   // after the output_stride = 8 conversion (resnet.py:203-214)
static const int kDil[4] = {1, 1, 2, 4};

static bool build_branch(const WeightMap& wm, const std::string& prefix, int num_units, int cin, Branch* br,
                         BlobBuilder* bb, std::string* err) {
  // stem: [64][cin][7][7] -> fp32 [64][(r*7+s)*cin + c], BN folded
  const int K = 49 * cin, Kpad = (K + 63) / 64 * 64;
  const float* w = find(wm, prefix + "conv1.weight", int64_t(64) * cin * 49, err);
  if (!w) return false;
  std::vector<float> scale, shift;
  if (!bn_fold(wm, prefix + "bn1", 64, &scale, &shift, err)) return false;
  br->stem_w_off = bb->reserve(size_t(64) * K * 4);
  br->stem_b_off = bb->reserve(64 * 4);
  float* sw = reinterpret_cast<float*>(bb->host.data() + br->stem_w_off);
  float* sb = reinterpret_cast<float*>(bb->host.data() + br->stem_b_off);
  for (int co = 0; co < 64; ++co) {
    for (int c = 0; c < cin; ++c)
      for (int t = 0; t < 49; ++t) sw[co * K + t * cin + c] = w[(co * cin + c) * 49 + t] * scale[co];
    sb[co] = shift[co];
  }
  // bf16 copy for the tensor-core stem: [64][Kpad], k = tap*cin + c, zero padded beyond K
  br->stem_wbf_off = bb->reserve(size_t(64) * Kpad * 2);
  {
    uint16_t* swb = reinterpret_cast<uint16_t*>(bb->host.data() + br->stem_wbf_off);
    const float* swf = reinterpret_cast<const float*>(bb->host.data() + br->stem_w_off);
    for (int co = 0; co < 64; ++co)
      for (int k = 0; k < Kpad; ++k) swb[co * Kpad + k] = (k < K) ? f32_to_bf16_rne(swf[co * K + k]) : uint16_t(0);
  }
  int inplanes = 64;
  for (int li = 0; li < 4; ++li) {
    for (int b = 0; b < kBlocks[li]; ++b) {
      Block blk;
      const std::string q = prefix + "layer" + std::to_string(li + 1) + "." + std::to_string(b) + ".";
      const int stride = (b == 0) ? kStride[li] : 1;
      blk.conv1.Cin = inplanes; blk.conv1.Cout = kPlanes[li]; blk.conv1.ksize = 3; blk.conv1.stride = stride; blk.conv1.dil = kDil[li];
      blk.conv2.Cin = kPlanes[li]; blk.conv2.Cout = kPlanes[li]; blk.conv2.ksize = 3; blk.conv2.stride = 1; blk.conv2.dil = kDil[li];
      if (!pack_conv(wm, q + "conv1.weight", q + "bn1", "", &blk.conv1, bb, err)) return false;
      if (!pack_conv(wm, q + "conv2.weight", q + "bn2", "", &blk.conv2, bb, err)) return false;
      // resnet.py:198 `if stride != 1 or self.inplanes != planes`: true for the first block of layers 2-4
      // (evaluated with the un-converted stride 2); the 1x1 downsample is never dilated (:216-221).
      if (b == 0 && li >= 1) {
        blk.has_down = true;
        blk.down.Cin = inplanes; blk.down.Cout = kPlanes[li]; blk.down.ksize = 1; blk.down.stride = stride; blk.down.dil = 1;
        if (!pack_conv(wm, q + "downsample.0.weight", q + "downsample.1", "", &blk.down, bb, err)) return false;
      }
      br->blocks.push_back(blk);
      inplanes = kPlanes[li];
    }
  }
  br->fc.Cin = 512; br->fc.Cout = num_units; br->fc.ksize = 1; br->fc.stride = 1; br->fc.dil = 1;
  return pack_conv(wm, prefix + "fc.weight", "", prefix + "fc.bias", &br->fc, bb, err);
}

struct Dims {
  int H1, W1;   // stem output
  int H2, W2;   // after max-pool (layer1)
  int H3, W3;   // layer2..4 and trunk output
};

static Dims trunk_dims(int H, int W) {
  Dims d;
  d.H1 = (H + 6 - 7) / 2 + 1; d.W1 = (W + 6 - 7) / 2 + 1;
  d.H2 = (d.H1 + 2 - 3) / 2 + 1; d.W2 = (d.W1 + 2 - 3) / 2 + 1;
  d.H3 = conv_out_dim(d.H2, 3, 2, 1); d.W3 = conv_out_dim(d.W2, 3, 2, 1);
  return d;
}

struct WsPlan {
  size_t stem[2], buf[2][4], trunk[2], scratch, total;
};

static WsPlan plan_ws(int num_units, int N, int H, int W) {
  const Dims d = trunk_dims(H, W);
  WsPlan p;
  size_t off = 0;
  const size_t stem_b = size_t(N) * d.H1 * d.W1 * 64 * 2;
  size_t act = size_t(N) * d.H2 * d.W2 * 64 * 2;
  const size_t deep = size_t(N) * d.H3 * d.W3 * 512 * 2;
  if (deep > act) act = deep;
  for (int g = 0; g < 2; ++g) {
    p.stem[g] = off; off = align_up(off + stem_b, 1024);
    for (int i = 0; i < 4; ++i) { p.buf[g][i] = off; off = align_up(off + act, 1024); }
    p.trunk[g] = off; off = align_up(off + size_t(N) * d.H3 * d.W3 * num_units * 4, 1024);
  }
  p.scratch = off;            // stream-K partials + flags of the convolution kernel (private to this workspace = stream)
  off = align_up(off + conv_pair_scratch_bytes(), 1024);
  p.total = off;
  return p;
}

static int run_conv(const ConvLayer* L[2], const uoc_backbone* bb, const void* x[2], const void* res[2], void* y[2], int N,
                    int H, int W, int relu, int out_fp32, int flags, cudaStream_t st, void* scratch) {
  ConvProblem p;
  memset(&p, 0, sizeof(p));
  p.groups = bb->groups;
  for (int g = 0; g < bb->groups; ++g) {
    p.g[g].x = x[g];
    p.g[g].w = bb->blob + L[g]->w_off;
    p.g[g].bias = reinterpret_cast<const float*>(bb->blob + L[g]->b_off);
    p.g[g].residual = res ? res[g] : nullptr;
    p.g[g].y = y[g];
  }
  p.N = N; p.H = H; p.W = W; p.Cin = L[0]->Cin; p.Cout = L[0]->Cout;
  p.ksize = L[0]->ksize; p.stride = L[0]->stride; p.dilation = L[0]->dil; p.relu = relu; p.out_fp32 = out_fp32;
  return (flags & UOC_FLAG_CONV_SIMT) ? launch_conv_simt(p, st) : launch_conv_auto(p, scratch, conv_pair_scratch_bytes(), st);
}

}  // namespace uoc

using namespace uoc;

extern "C" {

int uoc_backbone_create(uoc_backbone** out, const uoc_weight_desc* tensors, int n_tensors, int num_units) {
  return uoc_backbone_create_ex(out, tensors, n_tensors, num_units, UOC_INPUT_RGBD, UOC_FUSION_ADD, 1);
}

int uoc_backbone_feature_dim(const uoc_backbone* bb) { return bb ? bb->feat_dim : 0; }

int uoc_backbone_create_ex(uoc_backbone** out, const uoc_weight_desc* tensors, int n_tensors, int num_units, int input_type,
                           int fusion_type, int normalize) {
  if (!out || !tensors || n_tensors < 1) return fail(UOC_ERR_INVALID, "null argument");
  *out = nullptr;
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (num_units != 64 && num_units != 128) return fail(UOC_ERR_UNSUPPORTED, "num_units must be 64 or 128");
  if (input_type < UOC_INPUT_RGBD || input_type > UOC_INPUT_DEPTH) return fail(UOC_ERR_INVALID, "bad input_type");
  if (fusion_type < UOC_FUSION_ADD || fusion_type > UOC_FUSION_EARLY) return fail(UOC_ERR_INVALID, "bad fusion_type");
  const bool rgbd = input_type == UOC_INPUT_RGBD;
  if (rgbd && fusion_type == UOC_FUSION_CAT && num_units != 64)
    return fail(UOC_ERR_UNSUPPORTED, "cat fusion is supported for num_units = 64 (128-channel field)");
  WeightMap wm;
  for (int i = 0; i < n_tensors; ++i) {
    if (!tensors[i].name || !tensors[i].data_host) return fail(UOC_ERR_INVALID, "weight descriptor with null name / data");
    std::string name = tensors[i].name;
    if (name.rfind("module.", 0) == 0) name = name.substr(7);   // lib/networks/SEG.py:145-146
    wm[name] = &tensors[i];
  }
  uoc_backbone* bb = new uoc_backbone();
  bb->num_units = num_units;
  bb->input_type = input_type;
  bb->fusion = rgbd ? fusion_type : UOC_FUSION_ADD;
  bb->normalize = normalize ? 1 : 0;
  // SEG.py:69-71: the depth trunk exists only for INPUT == 'RGBD' with a late fusion
  bb->groups = (rgbd && fusion_type != UOC_FUSION_EARLY) ? 2 : 1;
  bb->cin = (rgbd && fusion_type == UOC_FUSION_EARLY) ? 6 : 3;
  bb->feat_dim = (rgbd && fusion_type == UOC_FUSION_CAT) ? 2 * num_units : num_units;
  BlobBuilder builder;
  std::string err;
  if (!build_branch(wm, "fcn.resnet34_8s.", num_units, bb->cin, &bb->br[0], &builder, &err) ||
      (bb->groups == 2 && !build_branch(wm, "fcn_depth.resnet34_8s.", num_units, 3, &bb->br[1], &builder, &err))) {
    delete bb;
    return fail(UOC_ERR_INVALID, err);
  }
  cudaGetDevice(&bb->device);
  bb->blob_bytes = builder.host.size();
  cudaError_t e = cudaMalloc(&bb->blob, bb->blob_bytes);
  if (e != cudaSuccess) { delete bb; return cuda_fail(e, "cudaMalloc(weights)", __FILE__, __LINE__); }
  e = cudaMemcpy(bb->blob, builder.host.data(), bb->blob_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(bb->blob); delete bb; return cuda_fail(e, "cudaMemcpy(weights)", __FILE__, __LINE__); }
  *out = bb;
  return UOC_OK;
}

void uoc_backbone_destroy(uoc_backbone* bb) {
  if (!bb) return;
  if (bb->blob) cudaFree(bb->blob);
  delete bb;
}

size_t uoc_backbone_workspace_bytes(const uoc_backbone* bb, int N, int H, int W) {
  if (!bb || N < 1 || H < 16 || W < 16) return 0;
  return plan_ws(bb->num_units, N, H, W).total;
}

int uoc_backbone_forward(uoc_backbone* bb, const float* rgb, const float* xyz, int N, int H, int W, float* features_out,
                         void* features_bf16_out, void* workspace, size_t workspace_bytes, int flags,
                         uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (!bb || !features_out || !workspace) return fail(UOC_ERR_INVALID, "null argument");
  if (bb->input_type != UOC_INPUT_DEPTH && !rgb) return fail(UOC_ERR_INVALID, "rgb is required for this input type");
  if (bb->input_type != UOC_INPUT_COLOR && !xyz) return fail(UOC_ERR_INVALID, "xyz is required for this input type");
  if (N < 1 || H < 16 || W < 16) return fail(UOC_ERR_INVALID, "bad N / H / W");
  if (reinterpret_cast<uintptr_t>(workspace) % 1024 != 0) return fail(UOC_ERR_INVALID, "workspace must be 1024-byte aligned");
  const WsPlan wp = plan_ws(bb->num_units, N, H, W);
  if (workspace_bytes < wp.total) return fail(UOC_ERR_WORKSPACE, "backbone workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const Dims d = trunk_dims(H, W);
  const int G = bb->groups;
  void* scratch = ws + wp.scratch;
  // a caller may hand the same memory to another tensor between calls (torch's caching allocator): the stream-K counters are
  // re-zeroed at the start of every forward pass (4 KB memset, stream ordered)
  UOC_CUDA(cudaMemsetAsync(scratch, 0, kConvCounterBytes, st));

  // SEG.py:97-108: DEPTH -> fcn(depth); COLOR -> fcn(img); early -> fcn(cat(img, depth)); else fcn(img), fcn_depth(depth)
  StemGroup sg[2];
  const float* inputs[2] = {bb->input_type == UOC_INPUT_DEPTH ? xyz : rgb, xyz};
  for (int g = 0; g < 2; ++g) {
    const int b = g < G ? g : 0;
    sg[g].x = inputs[b];
    sg[g].x2 = xyz;
    sg[g].w = reinterpret_cast<const float*>(bb->blob + bb->br[b].stem_w_off);
    sg[g].bias = reinterpret_cast<const float*>(bb->blob + bb->br[b].stem_b_off);
    sg[g].y = ws + wp.stem[b];
  }
  if (flags & UOC_FLAG_CONV_SIMT) {
    rc = launch_stem(sg, G, bb->cin, N, H, W, st);
  } else {
    const void* wbf[2] = {bb->blob + bb->br[0].stem_wbf_off, bb->blob + bb->br[G - 1].stem_wbf_off};
    rc = launch_stem_tc(sg, wbf, G, bb->cin, N, H, W, st);
  }
  if (rc != UOC_OK) return rc;
  const void* px[2] = {ws + wp.stem[0], ws + wp.stem[1]};
  void* py[2] = {ws + wp.buf[0][0], ws + wp.buf[1][0]};
  rc = launch_maxpool(px, py, G, N, d.H1, d.W1, 64, st);
  if (rc != UOC_OK) return rc;

  int cur = 0, curH = d.H2, curW = d.W2;
  for (size_t bi = 0; bi < bb->br[0].blocks.size(); ++bi) {
    const Block& b0 = bb->br[0].blocks[bi];
    const Block& b1 = bb->br[G - 1].blocks[bi];
    const int t = (cur + 1) & 3, r = (cur + 2) & 3, y = (cur + 3) & 3;
    const void* xin[2] = {ws + wp.buf[0][cur], ws + wp.buf[1][cur]};
    void* tout[2] = {ws + wp.buf[0][t], ws + wp.buf[1][t]};
    const ConvLayer* L1[2] = {&b0.conv1, &b1.conv1};
    rc = run_conv(L1, bb, xin, nullptr, tout, N, curH, curW, 1, 0, flags, st, scratch);
    if (rc != UOC_OK) return rc;
    const int oH = conv_out_dim(curH, 3, b0.conv1.stride, b0.conv1.dil);
    const int oW = conv_out_dim(curW, 3, b0.conv1.stride, b0.conv1.dil);
    const void* res[2] = {xin[0], xin[1]};
    if (b0.has_down) {
      void* rout[2] = {ws + wp.buf[0][r], ws + wp.buf[1][r]};
      const ConvLayer* LD[2] = {&b0.down, &b1.down};
      rc = run_conv(LD, bb, xin, nullptr, rout, N, curH, curW, 0, 0, flags, st, scratch);
      if (rc != UOC_OK) return rc;
      res[0] = rout[0]; res[1] = rout[1];
    }
    const void* tin[2] = {tout[0], tout[1]};
    void* yout[2] = {ws + wp.buf[0][y], ws + wp.buf[1][y]};
    const ConvLayer* L2[2] = {&b0.conv2, &b1.conv2};
    rc = run_conv(L2, bb, tin, res, yout, N, oH, oW, 1, 0, flags, st, scratch);
    if (rc != UOC_OK) return rc;
    cur = y; curH = oH; curW = oW;
  }
  {
    const void* xin[2] = {ws + wp.buf[0][cur], ws + wp.buf[1][cur]};
    void* tr[2] = {ws + wp.trunk[0], ws + wp.trunk[1]};
    const ConvLayer* LF[2] = {&bb->br[0].fc, &bb->br[G - 1].fc};
    rc = run_conv(LF, bb, xin, nullptr, tr, N, curH, curW, 0, 1, flags, st, scratch);
    if (rc != UOC_OK) return rc;
  }
  const int mode = (G == 1) ? HEAD_SINGLE : (bb->fusion == UOC_FUSION_CAT ? HEAD_CAT : HEAD_ADD);
  rc = launch_head(reinterpret_cast<const float*>(ws + wp.trunk[0]), reinterpret_cast<const float*>(ws + wp.trunk[1]), mode,
                   bb->normalize, N, curH, curW, bb->feat_dim, H, W, features_out, features_bf16_out, st);
  if (rc != UOC_OK) return rc;
  if (flags & UOC_FLAG_SYNC_CHECK) return check_device_error(st);
  return UOC_OK;
}

int uoc_backbone_read_trunk(uoc_backbone* bb, int branch, int N, int H, int W, const void* workspace, float* out,
                            uoc_stream_t stream) {
  if (!bb || !workspace || !out || branch < 0 || branch >= bb->groups) return fail(UOC_ERR_INVALID, "bad argument");
  const WsPlan wp = plan_ws(bb->num_units, N, H, W);
  const Dims d = trunk_dims(H, W);
  return launch_nhwc_to_nchw(reinterpret_cast<const float*>(static_cast<const char*>(workspace) + wp.trunk[branch]), N, d.H3,
                             d.W3, bb->num_units, out, static_cast<cudaStream_t>(stream));
}

int uoc_conv2d_bf16(const void* x, const void* w, const float* bias, const void* residual, void* y, int N, int H, int W,
                    int Cin, int Cout, int ksize, int stride, int dilation, int relu, int flags, uoc_stream_t stream) {
  int rc = require_sm100();
  if (rc != UOC_OK) return rc;
  if (!x || !w || !bias || !y) return fail(UOC_ERR_INVALID, "null argument");
  ConvProblem p;
  memset(&p, 0, sizeof(p));
  p.groups = 1;
  p.g[0].x = x; p.g[0].w = w; p.g[0].bias = bias; p.g[0].residual = residual; p.g[0].y = y;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ksize = ksize; p.stride = stride; p.dilation = dilation;
  p.relu = relu; p.out_fp32 = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // test hook: one scratch region per device, allocated on first use (the product path takes it from the caller's workspace)
  static void* hook_scratch[64] = {nullptr};
  int devid = 0;
  UOC_CUDA(cudaGetDevice(&devid));
  if (devid < 0 || devid >= 64) return fail(UOC_ERR_UNSUPPORTED, "device ordinal out of range");
  if (!hook_scratch[devid]) {
    UOC_CUDA(cudaMalloc(&hook_scratch[devid], conv_pair_scratch_bytes()));
    UOC_CUDA(cudaMemset(hook_scratch[devid], 0, kConvCounterBytes));
  }
  rc = (flags & UOC_FLAG_CONV_SIMT) ? launch_conv_simt(p, st)
                                    : launch_conv_auto(p, hook_scratch[devid], conv_pair_scratch_bytes(), st);
  if (rc != UOC_OK) return rc;
  if (flags & UOC_FLAG_SYNC_CHECK) return check_device_error(st);
  return UOC_OK;
}

}  // extern "C"
