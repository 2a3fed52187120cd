// embed.cu — the EfficientNet-B0 embedding tower as a native runtime object (C ABI kws_embed_*).
//
// Replaces `base_model = tf.keras.models.load_model(...)` + `Model(inputs, get_layer("dense_2").output)`
// + `.predict(...)` of the reference (multilingual_kws/embedding/transfer_learning.py:36-43,
// distance_filtering.py:12-27; architecture train_multilingual_embedding.py:66-83).
//
// At create time the Keras-named fp32 weights are folded for inference (BatchNorm -> scale/shift of the
// preceding conv, eps 1e-3), converted to the layouts the kernels want (K-major 16-bit — fp16 by default, bf16 on request —
// for every tensor-core contraction, fp32 for depthwise / SE / biases) and uploaded once.  A forward pass is a fixed list of
// launches on the caller's stream: stem conv, then per MBConv block
//   [expand 1x1 GEMM + BN + swish] -> [depthwise + BN + swish + SE, fused] -> [project 1x1 GEMM + BN (+ residual)]
// then top 1x1 GEMM + BN + swish + 2x2 average pool (fused epilogue) and the three dense GEMMs.
// Activations are NHWC 16-bit in three ping-pong workspace buffers; the batch is walked in chunks so
// one chunk's inter-layer activations stay resident in the 126 MB L2.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <new>
#include <string>
#include <vector>

#include "common.h"
#include "conv_kernels.h"
#include "gemm_tcgen05.cuh"
#include "mbconv_fused.h"

using namespace kws;

namespace {

struct HostTensor {
  std::vector<uint32_t> dims;
  const float* data = nullptr;
  size_t size = 0;
};
typedef std::map<std::string, HostTensor> WeightMap;

// container: "KWSW0001", u32 n; per entry: u32 name_len, name (padded to 4), u32 ndim, u32 dims[], f32 data[]
bool parse_blob(const void* blob, size_t bytes, WeightMap* out, std::string* err) {
  const uint8_t* p = static_cast<const uint8_t*>(blob);
  const uint8_t* end = p + bytes;
  if (bytes < 12 || memcmp(p, "KWSW0001", 8) != 0) { *err = "bad magic"; return false; }
  p += 8;
  uint32_t n;
  memcpy(&n, p, 4); p += 4;
  for (uint32_t i = 0; i < n; ++i) {
    if (p + 4 > end) { *err = "truncated"; return false; }
    uint32_t nl; memcpy(&nl, p, 4); p += 4;
    const uint32_t nlp = (nl + 3) & ~3u;
    if (p + nlp + 4 > end) { *err = "truncated"; return false; }
    std::string name(reinterpret_cast<const char*>(p), nl); p += nlp;
    uint32_t nd; memcpy(&nd, p, 4); p += 4;
    if (nd > 8 || p + 4 * nd > end) { *err = "bad ndim for " + name; return false; }
    HostTensor t;
    t.size = 1;
    for (uint32_t d = 0; d < nd; ++d) { uint32_t v; memcpy(&v, p, 4); p += 4; t.dims.push_back(v); t.size *= v; }
    if (p + 4 * t.size > end) { *err = "truncated data for " + name; return false; }
    t.data = reinterpret_cast<const float*>(p);
    p += 4 * t.size;
    (*out)[name] = t;
  }
  return true;
}

struct BlockCfg { int k, reps, fin, fout, e, s; };
const BlockCfg kStages[7] = {{3, 1, 32, 16, 1, 1}, {3, 2, 16, 24, 6, 2}, {5, 2, 24, 40, 6, 2}, {3, 3, 40, 80, 6, 2},
                             {5, 3, 80, 112, 6, 1}, {5, 4, 112, 192, 6, 2}, {3, 1, 192, 320, 6, 1}};
const float kBnEps = 1e-3f;

enum OpKind { kOpStem, kOpGemm, kOpDwse };

struct Op {
  OpKind kind;
  std::string name;            // Keras layer whose output this op produces (for taps)
  int in_buf, out_buf;         // 0 = X, 1 = E, 2 = D, -1 = external (features / embedding)
  int res_buf = -1;
  // geometry
  int rows_per_clip = 0;       // GEMM: M = rows_per_clip * batch
  int N = 0, K = 0;
  int act = 0, gap4 = 0, out_f32 = 0;
  const uint16_t* w = nullptr;        // GEMM weights [N][K], fp16 or bf16 bits
  const float* bias = nullptr;
  StemParams stem{};
  DwseParams dw{};
  // external SE (wide layers): 16-bit GEMM operands, se padded to a multiple of 16
  int se_pad = 0;
  const uint16_t* se_w1 = nullptr;   // [se_pad][C]
  const float* se_b1 = nullptr;       // [se_pad]
  const uint16_t* se_w2 = nullptr;   // [C][se_pad]
  int dw_group = 1;
  int block_id = -1;                   // MBConv block this op belongs to (index into kws_embed::fblocks), -1: none
  size_t out_elems_per_clip = 0;
  double flops_per_clip = 0;           // 2 * MACs
  double bytes_per_clip = 0;           // algorithmic: activation read + write (weights excluded)
};

}  // namespace

// One MBConv block in the form the fused tail kernel wants (mbconv_fused.cu); ops [op_lo, op_hi] are its layer-wise ops.
struct FusedBlock {
  FusedBlockInfo info{};
  int op_lo = 0, op_hi = 0;
  bool fusable = false;
  // host-side pieces the device descriptor is built from
  const uint16_t *w_exp = nullptr, *w_se1 = nullptr, *w_se2 = nullptr, *w_proj = nullptr;
  const float *b_exp = nullptr, *w_dw = nullptr, *b_dw = nullptr, *b_se1 = nullptr, *b_se2 = nullptr, *b_proj = nullptr;
  int se = 0;
};
// A run of consecutive fusable blocks executed by one launch
struct FusedSegment {
  int blk_lo = 0, nblocks = 0;         // blocks [blk_lo, blk_lo + nblocks)
  int op_lo = 0, op_hi = 0;            // layer-wise ops it replaces
  int in_place = 0;                    // every block has a skip connection: the output may overwrite the input rows
};

struct GraphEntry {
  const float* feats; float* emb; void* ws; int batch; long long chunk;   // chunk: schedule key (chunks + SM budgets)
  cudaGraphExec_t exec;                                                   // nullptr: key seen once, not captured yet
};

struct kws_embed {
  std::vector<GraphEntry> graphs;      // captured forward passes, keyed by (buffers, batch, chunk)
  std::vector<GraphEntry> seen;        // keys met once (plain launches); a key is captured when it comes back
  cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the caller's may be the legacy stream)
  int use_graph = 1;
  int H = 49, W = 40, out_dim = 0;
  float in_scale = 1.0f / 255.0f, in_shift = 0.0f;
  std::vector<Op> ops;
  std::vector<void*> dev_allocs;
  // Two schedule segments.  "early" ops (stem .. the project conv of the block whose expanded map is the last big
  // one) have up to 96 KB of activations per clip and are walked in small chunks so they stay L2-resident; "late"
  // ops have <= 37 KB per clip and run over large chunks so GEMM tiles / depthwise groups fill the 148 SMs.
  int split_op = 0;                    // first late op
  size_t max_se_channels = 0;          // widest externally-gated layer (scratch: pooled means, squeeze, gates)
  size_t buf_elems[2][3] = {{0, 0, 0}, {0, 0, 0}};   // [segment][X, E, D] per clip, 16-bit elements
  int sm_count = 0, max_smem = 0;
  // Throughput schedule (kws_embed_forward_budget): when several forward passes are in flight on different streams, the
  // launch/latency-bound tail of the network (everything after block3b: maps of <= 4x3 pixels) is sized for a subset
  // of the SMs, so its persistent CTAs leave room for the throughput-bound head of the neighbouring pass.
  int tail_op = 0;                     // first op of the tail
  int chunk = 1024;                    // early-segment clips per pass
  int chunk_late = 4096;               // late-segment clips per pass
  int se_via_gemm = 1;                 // wide layers: SE FCs as batched tcgen05 GEMMs (0: inside the depthwise kernel)
  int bf16 = 0;                        // 16-bit storage / tensor-core operand type: 0 fp16 (default), 1 bf16
  double flops_per_clip = 0;
  // Fused tail (mbconv_fused.cu): 0 = layer by layer, 1 = one launch per MBConv block, 2 = runs of blocks per launch
  int fuse = 0;                        // (default stays layer-wise until the fused kernel wins in the graph; see kws_embed_set_fuse)
  int stop_after_tap = 0;              // kws_embed_forward_until: run only the ops up to (and including) the tapped one
  std::vector<FusedBlock> fblocks;
  std::vector<FusedBlockInfo> finfos;  // fblocks[i].info, contiguous (the launcher takes an array)
  FusedBlockDev* d_fblocks = nullptr;  // device array, same order as fblocks (+ the top-conv pseudo-block at the end)
  std::vector<FusedSegment> segs[4];   // per fuse mode
};

namespace {

template <typename T>
T* upload(kws_embed* m, const std::vector<T>& h, cudaError_t* err) {
  T* d = nullptr;
  *err = cudaMalloc(&d, h.size() * sizeof(T));
  if (*err != cudaSuccess) return nullptr;
  m->dev_allocs.push_back(d);
  *err = cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

struct Builder {
  kws_embed* m;
  const WeightMap& w;
  std::string err;
  cudaError_t cerr = cudaSuccess;

  const HostTensor* get(const std::string& name, size_t expect) {
    auto it = w.find(name);
    if (it == w.end()) { err = "missing weight " + name; return nullptr; }
    if (it->second.size != expect) { err = "wrong size for " + name; return nullptr; }
    return &it->second;
  }
  // BN inference fold: y = x*scale + shift
  bool bn(const std::string& name, int c, std::vector<float>* scale, std::vector<float>* shift) {
    const HostTensor *g = get(name + "/gamma", c), *b = get(name + "/beta", c), *mu = get(name + "/moving_mean", c),
                     *var = get(name + "/moving_variance", c);
    if (!g || !b || !mu || !var) return false;
    scale->resize(c); shift->resize(c);
    for (int i = 0; i < c; ++i) {
      const float s = g->data[i] / sqrtf(var->data[i] + kBnEps);
      (*scale)[i] = s;
      (*shift)[i] = b->data[i] - mu->data[i] * s;
    }
    return true;
  }
  // 1x1 conv / dense kernel [K][N] (Keras) -> 16-bit [N][K] with per-output scale
  const uint16_t* gemm_weight(const std::string& name, int K, int N, const std::vector<float>* scale) {
    const HostTensor* t = get(name, (size_t)K * N);
    if (!t) return nullptr;
    std::vector<uint16_t> h((size_t)N * K);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const float v = t->data[(size_t)k * N + n] * (scale ? (*scale)[n] : 1.0f);
        uint16_t bits;
        if (m->bf16) { const __nv_bfloat16 b = __float2bfloat16(v); memcpy(&bits, &b, 2); }
        else { const __half b = __float2half_rn(v); memcpy(&bits, &b, 2); }
        h[(size_t)n * K + k] = bits;
      }
    return upload(m, h, &cerr);
  }
  // same, from a host fp32 [N][K] matrix (used for folded projection -> expansion pairs)
  const uint16_t* gemm_weight_host(const std::vector<double>& wnk) {
    std::vector<uint16_t> h(wnk.size());
    for (size_t i = 0; i < wnk.size(); ++i) h[i] = h16((float)wnk[i]);
    return upload(m, h, &cerr);
  }
  const float* vec(const std::vector<float>& v) { return upload(m, v, &cerr); }
  uint16_t h16(float v) const {
    uint16_t bits;
    if (m->bf16) { const __nv_bfloat16 b = __float2bfloat16(v); memcpy(&bits, &b, 2); }
    else { const __half b = __float2half_rn(v); memcpy(&bits, &b, 2); }
    return bits;
  }
  const uint16_t* vec16(const std::vector<uint16_t>& v) { return upload(m, v, &cerr); }
};

}  // namespace

// Device descriptors of the fusable blocks (weight tensor maps are encoded once, here) and the launch plans of the two
// fused modes: per block, and runs of consecutive blocks that fit one CTA's shared memory / TMEM with >= kMinGroup clips.
static int build_fused_plan(kws_embed* m) {
  const int n = (int)m->fblocks.size();
  std::vector<FusedBlockDev> dev(n);
  for (int i = 0; i < n; ++i) {
    FusedBlock& fb = m->fblocks[i];
    // fusion applies to the network's tail only (tiny maps); a block must also fit on its own
    if (fb.fusable && fb.op_lo < m->tail_op) fb.fusable = false;
    if (fb.fusable && fused_max_group(&fb.info, 1, m->max_smem) < 1) fb.fusable = false;
    FusedBlockDev& d = dev[i];
    memset(&d, 0, sizeof(d));
    if (!fb.fusable) continue;
    const FusedBlockInfo& I = fb.info;
    int rc = make_tmap_h16(&d.tm_exp, fb.w_exp, (uint64_t)I.cexp, (uint64_t)I.cin, kFusedBoxRows, m->bf16, 64);
    if (rc == KWS_OK && !I.pool_out) rc = make_tmap_h16(&d.tm_se1, fb.w_se1, (uint64_t)I.se_pad, (uint64_t)I.cexp, (uint32_t)I.se_pad, m->bf16, 64);
    if (rc == KWS_OK && !I.pool_out) rc = make_tmap_h16(&d.tm_se2, fb.w_se2, (uint64_t)I.cexp, (uint64_t)I.se_pad, kFusedBoxRows, m->bf16, 64);
    if (rc == KWS_OK && !I.pool_out) rc = make_tmap_h16(&d.tm_proj, fb.w_proj, (uint64_t)I.cout, (uint64_t)I.cexp, kFusedBoxRows, m->bf16, 64);
    if (rc != KWS_OK) return rc;
    d.b_exp = fb.b_exp; d.w_dw = fb.w_dw; d.b_dw = fb.b_dw; d.b_se1 = fb.b_se1; d.b_se2 = fb.b_se2; d.b_proj = fb.b_proj;
    d.cin = I.cin; d.cexp = I.cexp; d.cout = I.cout; d.se = fb.se; d.se_pad = I.se_pad; d.geom = I.geom;
    d.pin = I.pin; d.pout = I.pout; d.residual = I.residual; d.pool_out = I.pool_out;
  }
  KWS_CUDA_CHECK(cudaMalloc(&m->d_fblocks, sizeof(FusedBlockDev) * (size_t)n));
  m->dev_allocs.push_back(m->d_fblocks);
  KWS_CUDA_CHECK(cudaMemcpy(m->d_fblocks, dev.data(), sizeof(FusedBlockDev) * (size_t)n, cudaMemcpyHostToDevice));
  // mode 1: one launch per fusable block (the top conv stays a GEMM: it has its own well-fed tile shape)
  for (int i = 0; i < n; ++i) {
    const FusedBlock& fb = m->fblocks[i];
    if (!fb.fusable || fb.info.pool_out) continue;
    FusedSegment sg;
    sg.blk_lo = i; sg.nblocks = 1; sg.op_lo = fb.op_lo; sg.op_hi = fb.op_hi; sg.in_place = fb.info.residual;
    m->segs[1].push_back(sg);
  }
  // modes 2 and 3: greedy runs.  A run keeps growing while the merged launch still fits at least kMinGroup clips per CTA.
  // Mode 3 fuses only the blocks with <= 672 expanded channels (4a .. 6a: 12 pixels per clip, where one fused launch beats
  // the six layer-wise ones); the 1152-channel blocks on 2x2 maps (6b .. 7a) give a fused CTA only 28 rows per step
  // (measured slower than their six layer-wise launches) and stay layer-wise.
  const int kMinGroup = 7;
  m->finfos.resize(n);
  for (int i = 0; i < n; ++i) m->finfos[i] = m->fblocks[i].info;
  const std::vector<FusedBlockInfo>& infos = m->finfos;
  for (int mode = 2; mode <= 3; ++mode) {
    auto ok = [&](int i) {
      const FusedBlock& fb = m->fblocks[i];
      return fb.fusable && (mode == 2 || (!fb.info.pool_out && fb.info.cexp <= 672));
    };
    for (int i = 0; i < n;) {
      if (!ok(i) || m->fblocks[i].info.pool_out) { ++i; continue; }
      int cnt = 1;
      while (i + cnt < n && ok(i + cnt) && m->fblocks[i + cnt].op_lo == m->fblocks[i + cnt - 1].op_hi + 1 &&
             fused_max_group(&infos[i], cnt + 1, m->max_smem) >= kMinGroup)
        ++cnt;
      FusedSegment sg;
      sg.blk_lo = i; sg.nblocks = cnt; sg.op_lo = m->fblocks[i].op_lo; sg.op_hi = m->fblocks[i + cnt - 1].op_hi;
      sg.in_place = 1;
      for (int q = 0; q < cnt; ++q) if (!m->fblocks[i + q].info.residual) sg.in_place = 0;
      m->segs[mode].push_back(sg);
      i += cnt;
    }
  }
  return KWS_OK;
}

extern "C" int kws_embed_create(kws_embed_t** out, const void* blob, size_t bytes, int act_dtype) {
  KWS_REQUIRE(out && blob, "kws_embed_create: NULL argument");
  *out = nullptr;
  KWS_REQUIRE(act_dtype == 0 || act_dtype == 1, "kws_embed_create: act_dtype must be 0 (fp16) or 1 (bf16)");
  WeightMap wm;
  std::string perr;
  if (!parse_blob(blob, bytes, &wm, &perr)) {
    set_error("kws_embed_create: weight container: %s", perr.c_str());
    return KWS_ERR_ARG;
  }
  int dev = 0;
  kws_embed* m = new (std::nothrow) kws_embed();
  KWS_REQUIRE(m != nullptr, "out of host memory");
  m->bf16 = act_dtype;
  if (const char* env = getenv("KWS_SE_IN_KERNEL")) m->se_via_gemm = atoi(env) ? 0 : 1;   // tuning knob (A/B measurements)
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) {
    set_error("kws_embed_create: CUDA device required (%s); there is no CPU fallback", cudaGetErrorString(e));
    delete m;
    return KWS_ERR_CUDA;
  }
  Builder B{m, wm};
  auto fail = [&](int code) {
    if (B.cerr != cudaSuccess) set_error("kws_embed_create: %s", cudaGetErrorString(B.cerr));
    else set_error("kws_embed_create: %s", B.err.c_str());
    for (void* p : m->dev_allocs) cudaFree(p);
    delete m;
    return code;
  };
#define CK(x) do { if (!(x) || B.cerr != cudaSuccess) return fail(B.cerr != cudaSuccess ? KWS_ERR_CUDA : KWS_ERR_ARG); } while (0)

  // Rescaling(1/255) + Normalization(mean, var)
  {
    auto itm = wm.find("normalization/mean"), itv = wm.find("normalization/variance");
    float mean = 0.f, var = 1.f;
    if (itm != wm.end() && itm->second.size >= 1) mean = itm->second.data[0];
    if (itv != wm.end() && itv->second.size >= 1) var = itv->second.data[0];
    const float sd = fmaxf(sqrtf(var), 1e-7f);
    m->in_scale = 1.0f / (255.0f * sd);
    m->in_shift = -mean / sd;
  }
  int h = m->H, w = m->W;
  double macs = 0;
  // ---- stem
  {
    std::vector<float> sc, sh;
    CK(B.bn("stem_bn", 32, &sc, &sh));
    const HostTensor* k = B.get("stem_conv/kernel", 9 * 32);
    CK(k);
    std::vector<float> wf(9 * 32);
    for (int t = 0; t < 9; ++t)
      for (int c = 0; c < 32; ++c) wf[t * 32 + c] = k->data[t * 32 + c] * sc[c];
    Op op;
    op.kind = kOpStem; op.name = "stem_activation"; op.in_buf = -1; op.out_buf = 0;
    op.stem.H = h; op.stem.W = w;
    op.stem.pad_top = 1 - (1 - h % 2); op.stem.pad_left = 1 - (1 - w % 2);
    op.stem.Ho = (h + op.stem.pad_top + 1 - 3) / 2 + 1;
    op.stem.Wo = (w + op.stem.pad_left + 1 - 3) / 2 + 1;
    op.stem.in_scale = m->in_scale; op.stem.in_shift = m->in_shift; op.stem.bf16 = m->bf16;
    op.stem.w = B.vec(wf); op.stem.bias = B.vec(sh);
    CK(op.stem.w && op.stem.bias);
    h = op.stem.Ho; w = op.stem.Wo;
    op.out_elems_per_clip = (size_t)h * w * 32;
    macs += (double)h * w * 32 * 9;
    m->ops.push_back(op);
  }
  // ---- MBConv blocks
  // A projection without a skip connection whose only consumer is the next block's expansion (also without a
  // skip from that tensor) is a linear map followed by a linear map: W_e (W_p d + b_p) + b_e.  The pair is folded
  // into one GEMM (K = the projection's input channels) and the narrow tensor in between never exists.  With the
  // Keras B0 configuration this is block1a's projection into block2a's expansion (32 -> 16 -> 96 becomes 32 -> 96).
  const bool fold_linear_pairs = !(getenv("KWS_NO_FOLD") && atoi(getenv("KWS_NO_FOLD")));
  std::vector<double> pend_w;          // pending projection [cout][cexp], BN scale applied
  std::vector<double> pend_b;          // its shift [cout]
  int pend_k = 0;                      // its input channels (0: nothing pending)
  for (int si = 0; si < 7; ++si) {
    for (int r = 0; r < kStages[si].reps; ++r) {
      const BlockCfg& c = kStages[si];
      const std::string n = "block" + std::to_string(si + 1) + std::string(1, (char)('a' + r));
      const int stride = r == 0 ? c.s : 1;
      const int cin = r == 0 ? c.fin : c.fout;
      const int cexp = cin * c.e, cout = c.fout;
      const int se = cin / 4 > 1 ? cin / 4 : 1;
      FusedBlock fb;
      fb.op_lo = (int)m->ops.size();
      fb.info.cin = cin; fb.info.cexp = cexp; fb.info.cout = cout; fb.info.pin = h * w; fb.info.pool_out = 0;
      fb.info.residual = (stride == 1 && cin == cout) ? 1 : 0;
      fb.se = se;
      fb.fusable = c.e != 1 && !pend_k;          // needs its own expansion (a folded pair feeds the expand GEMM from D)
      const int block_id = (int)m->fblocks.size();
      if (c.e != 1) {
        std::vector<float> sc, sh;
        CK(B.bn(n + "_expand_bn", cexp, &sc, &sh));
        Op op;
        op.kind = kOpGemm; op.name = n + "_expand_activation"; op.in_buf = 0; op.out_buf = 1;
        op.rows_per_clip = h * w; op.N = cexp; op.K = cin; op.act = kActSwish;
        if (pend_k) {
          const HostTensor* ke = B.get(n + "_expand_conv/kernel", (size_t)cin * cexp);
          CK(ke);
          std::vector<double> wf((size_t)cexp * pend_k, 0.0);
          for (int o = 0; o < cexp; ++o) {
            double bacc = sh[o];
            for (int j = 0; j < cin; ++j) {
              const double we = (double)ke->data[(size_t)j * cexp + o] * sc[o];
              bacc += we * pend_b[j];
              for (int q = 0; q < pend_k; ++q) wf[(size_t)o * pend_k + q] += we * pend_w[(size_t)j * pend_k + q];
            }
            sh[o] = (float)bacc;
          }
          op.in_buf = 2; op.K = pend_k;
          op.w = B.gemm_weight_host(wf);
          pend_k = 0;
        } else {
          op.w = B.gemm_weight(n + "_expand_conv/kernel", cin, cexp, &sc);
        }
        op.bias = B.vec(sh);
        CK(op.w && op.bias);
        op.out_elems_per_clip = (size_t)h * w * cexp;
        macs += (double)h * w * op.K * cexp;
        op.block_id = block_id;
        fb.w_exp = op.w; fb.b_exp = op.bias;
        m->ops.push_back(op);
      }
      {
        std::vector<float> sc, sh;
        CK(B.bn(n + "_bn", cexp, &sc, &sh));
        const int k = c.k;
        const HostTensor* dk = B.get(n + "_dwconv/depthwise_kernel", (size_t)k * k * cexp);
        const HostTensor* w1 = B.get(n + "_se_reduce/kernel", (size_t)cexp * se);
        const HostTensor* b1 = B.get(n + "_se_reduce/bias", se);
        const HostTensor* w2 = B.get(n + "_se_expand/kernel", (size_t)se * cexp);
        const HostTensor* b2 = B.get(n + "_se_expand/bias", cexp);
        CK(dk && w1 && b1 && w2 && b2);
        std::vector<float> wd((size_t)k * k * cexp), w1t((size_t)se * cexp);
        for (int t = 0; t < k * k; ++t)
          for (int ch = 0; ch < cexp; ++ch) wd[(size_t)t * cexp + ch] = dk->data[(size_t)t * cexp + ch] * sc[ch];
        for (int j = 0; j < se; ++j)
          for (int ch = 0; ch < cexp; ++ch) w1t[(size_t)j * cexp + ch] = w1->data[(size_t)ch * se + j];
        Op op;
        op.kind = kOpDwse; op.name = n + "_se_excite"; op.in_buf = c.e != 1 ? 1 : 0; op.out_buf = 2;
        DwseParams& P = op.dw;
        P.H = h; P.W = w; P.C = cexp; P.K = k; P.S = stride; P.se = se; P.bf16 = m->bf16;
        if (stride == 2) {
          P.pad_top = k / 2 - (1 - h % 2); P.pad_left = k / 2 - (1 - w % 2);
          P.Ho = (h + P.pad_top + k / 2 - k) / 2 + 1; P.Wo = (w + P.pad_left + k / 2 - k) / 2 + 1;
        } else {
          P.pad_top = P.pad_left = k / 2; P.Ho = h; P.Wo = w;
        }
        P.w_dw = B.vec(wd); P.b_dw = B.vec(sh);
        P.w_se1 = B.vec(w1t); P.b_se1 = B.vec(std::vector<float>(b1->data, b1->data + se));
        P.w_se2 = B.vec(std::vector<float>(w2->data, w2->data + (size_t)se * cexp));
        P.b_se2 = B.vec(std::vector<float>(b2->data, b2->data + cexp));
        CK(P.w_dw && P.b_dw && P.w_se1 && P.b_se1 && P.w_se2 && P.b_se2);
        P.se_external = 0; P.pooled_out = nullptr;
        fb.info.geom = fused_geom_id(k, stride, h, w, P.pad_top, P.pad_left);
        fb.info.pout = P.Ho * P.Wo;
        if (!fb.info.geom) fb.fusable = false;
        const bool wide_se = cexp > 256 && m->se_via_gemm;   // C <= 256 runs the in-kernel (narrow) squeeze-excite
        if (wide_se || fb.fusable) {
          // wide layer: SE as two batched tensor-core GEMMs.  FC1: [B,C] x W1[se_pad,C]^T (+b1, swish);
          // FC2: [B,se_pad] x W2[C,se_pad]^T (+b2, sigmoid).  Padding rows / columns are zero.
          const int sp = (se + 15) & ~15;
          std::vector<uint16_t> w1h((size_t)sp * cexp, B.h16(0.0f)), w2h((size_t)cexp * sp, B.h16(0.0f));
          std::vector<float> b1p(sp, 0.0f);
          for (int j = 0; j < se; ++j) {
            b1p[j] = b1->data[j];
            for (int ch = 0; ch < cexp; ++ch) {
              w1h[(size_t)j * cexp + ch] = B.h16(w1->data[(size_t)ch * se + j]);
              w2h[(size_t)ch * sp + j] = B.h16(w2->data[(size_t)j * cexp + ch]);
            }
          }
          P.se_external = wide_se ? 1 : 0;
          op.se_pad = sp;
          op.se_w1 = B.vec16(w1h); op.se_b1 = B.vec(b1p); op.se_w2 = B.vec16(w2h);
          CK(op.se_w1 && op.se_b1 && op.se_w2);
          if (wide_se && (size_t)cexp > m->max_se_channels) m->max_se_channels = (size_t)cexp;
          fb.info.se_pad = sp;
          fb.w_se1 = op.se_w1; fb.b_se1 = op.se_b1; fb.w_se2 = op.se_w2;
        }
        fb.w_dw = P.w_dw; fb.b_dw = P.b_dw; fb.b_se2 = P.b_se2;
        op.block_id = block_id;
        op.dw_group = dwse_pick_group(P, m->max_smem, 1 << 20, m->sm_count);
        if (op.dw_group < 1) { B.err = "depthwise layer " + n + " does not fit shared memory"; return fail(KWS_ERR_UNSUPPORTED); }
        macs += (double)P.Ho * P.Wo * cexp * k * k + 2.0 * cexp * se;
        h = P.Ho; w = P.Wo;
        op.out_elems_per_clip = (size_t)h * w * cexp;
        m->ops.push_back(op);
      }
      {
        std::vector<float> sc, sh;
        CK(B.bn(n + "_project_bn", cout, &sc, &sh));
        const bool has_skip = stride == 1 && cin == cout;
        const bool last_rep = r + 1 == kStages[si].reps;
        if (fold_linear_pairs && !has_skip && last_rep && si + 1 < 7 && kStages[si + 1].e != 1 &&
            !(kStages[si + 1].s == 1 && kStages[si + 1].fin == kStages[si + 1].fout)) {
          const HostTensor* kp = B.get(n + "_project_conv/kernel", (size_t)cexp * cout);
          CK(kp);
          pend_w.assign((size_t)cout * cexp, 0.0); pend_b.assign(cout, 0.0);
          for (int o = 0; o < cout; ++o) {
            pend_b[o] = sh[o];
            for (int q = 0; q < cexp; ++q) pend_w[(size_t)o * cexp + q] = (double)kp->data[(size_t)q * cout + o] * sc[o];
          }
          pend_k = cexp;
          fb.fusable = false;
          fb.op_hi = (int)m->ops.size() - 1;
          m->fblocks.push_back(fb);
          continue;
        }
        Op op;
        op.kind = kOpGemm; op.name = n + "_out"; op.in_buf = 2; op.out_buf = 0;
        op.rows_per_clip = h * w; op.N = cout; op.K = cexp; op.act = kActNone;
        op.res_buf = (stride == 1 && cin == cout) ? 0 : -1;
        op.w = B.gemm_weight(n + "_project_conv/kernel", cexp, cout, &sc);
        op.bias = B.vec(sh);
        CK(op.w && op.bias);
        op.out_elems_per_clip = (size_t)h * w * cout;
        macs += (double)h * w * cexp * cout;
        op.block_id = block_id;
        fb.w_proj = op.w; fb.b_proj = op.bias;
        fb.op_hi = (int)m->ops.size();
        m->ops.push_back(op);
        m->fblocks.push_back(fb);
      }
    }
  }
  // ---- top conv + BN + swish + GAP (needs the 2x2 final map so the pool is a 4-row mean)
  if (h * w != 4) { B.err = "top GAP epilogue expects a 2x2 final feature map"; return fail(KWS_ERR_UNSUPPORTED); }
  {
    std::vector<float> sc, sh;
    CK(B.bn("top_bn", 1280, &sc, &sh));
    Op op;
    op.kind = kOpGemm; op.name = "top_gap"; op.in_buf = 0; op.out_buf = 1;
    op.rows_per_clip = 4; op.N = 1280; op.K = 320; op.act = kActSwish; op.gap4 = 1;
    op.w = B.gemm_weight("top_conv/kernel", 320, 1280, &sc);
    op.bias = B.vec(sh);
    CK(op.w && op.bias);
    op.out_elems_per_clip = 1280;
    macs += 4.0 * 320 * 1280;
    FusedBlock fb;
    fb.op_lo = fb.op_hi = (int)m->ops.size();
    fb.info.cin = 320; fb.info.cexp = 1280; fb.info.cout = 1280; fb.info.pin = 4; fb.info.pout = 4; fb.info.pool_out = 1;
    fb.info.geom = 7; fb.info.residual = 0; fb.info.se_pad = 0;
    fb.w_exp = op.w; fb.b_exp = op.bias;
    fb.fusable = true;
    op.block_id = (int)m->fblocks.size();
    m->ops.push_back(op);
    m->fblocks.push_back(fb);
  }
  // ---- dense tower: dense, dense_1, ... (relu ... relu, last = selu), cut at the last one present
  {
    int fan = 1280, idx = 0, in_buf = 1;
    std::vector<std::string> names;
    for (;; ++idx) {
      const std::string nm = idx == 0 ? "dense" : "dense_" + std::to_string(idx);
      if (wm.find(nm + "/kernel") == wm.end()) break;
      names.push_back(nm);
    }
    if (names.empty()) { B.err = "no dense layers in the weight container"; return fail(KWS_ERR_ARG); }
    // Each Dense layer keeps ITS OWN activation wherever the tower is cut (base_model.get_layer(name).output in the
    // reference, transfer_learning.py:38-43): the models of train_*_embedding.py:81-100 have relu, relu, selu and then the
    // linear classifier layer, so dense / dense_1 -> ReLU, dense_2 -> SELU, anything after it -> none.
    for (size_t i = 0; i < names.size(); ++i) {
      const HostTensor& kt = wm[names[i] + "/kernel"];
      if (kt.dims.size() != 2 || (int)kt.dims[0] != fan) { B.err = "bad shape for " + names[i]; return fail(KWS_ERR_ARG); }
      const int units = (int)kt.dims[1];
      const HostTensor* bt = B.get(names[i] + "/bias", units);
      CK(bt);
      const bool last = i + 1 == names.size();
      Op op;
      op.kind = kOpGemm; op.name = names[i]; op.in_buf = in_buf; op.out_buf = last ? -1 : (in_buf == 1 ? 2 : 1);
      op.rows_per_clip = 1; op.N = units; op.K = fan; op.act = i < 2 ? kActRelu : (i == 2 ? kActSelu : kActNone);
      op.out_f32 = last ? 1 : 0;
      op.w = B.gemm_weight(names[i] + "/kernel", fan, units, nullptr);
      op.bias = B.vec(std::vector<float>(bt->data, bt->data + units));
      CK(op.w && op.bias);
      op.out_elems_per_clip = units;
      macs += (double)fan * units;
      m->ops.push_back(op);
      in_buf = op.out_buf;
      fan = units;
    }
    m->out_dim = fan;
  }
#undef CK
  for (Op& op : m->ops) {
    if (op.kind == kOpStem) {
      op.flops_per_clip = 2.0 * op.stem.Ho * op.stem.Wo * 32 * 9;
      op.bytes_per_clip = 4.0 * op.stem.H * op.stem.W + 2.0 * op.out_elems_per_clip;
    } else if (op.kind == kOpDwse) {
      const DwseParams& P = op.dw;
      op.flops_per_clip = 2.0 * ((double)P.Ho * P.Wo * P.C * P.K * P.K + 2.0 * P.C * P.se);
      op.bytes_per_clip = 2.0 * ((double)P.H * P.W * P.C + (double)P.Ho * P.Wo * P.C);
    } else {
      op.flops_per_clip = 2.0 * op.rows_per_clip * (double)op.N * op.K;
      op.bytes_per_clip = 2.0 * op.rows_per_clip * op.K + (op.out_f32 ? 4.0 : 2.0) * op.out_elems_per_clip +
                          (op.res_buf >= 0 ? 2.0 * op.out_elems_per_clip : 0.0);
    }
  }
  // split after the last op that produces or consumes more than 20 000 elements per clip
  m->split_op = 0;
  for (size_t i = 0; i < m->ops.size(); ++i)
    if (m->ops[i].out_elems_per_clip > 20000) m->split_op = (int)i + 3 < (int)m->ops.size() ? (int)i + 3 : (int)m->ops.size();
  // (expand op i -> dwse i+1 -> project i+2; the late segment starts at the next block)
  for (size_t i = 0; i < m->ops.size(); ++i) {
    const Op& op = m->ops[i];
    const int seg = (int)i < m->split_op ? 0 : 1;
    if (op.out_buf >= 0 && op.out_elems_per_clip > m->buf_elems[seg][op.out_buf]) m->buf_elems[seg][op.out_buf] = op.out_elems_per_clip;
    // the hand-off tensor (output of the last early op) lives in the late X buffer
    if ((int)i == m->split_op - 1 && op.out_buf == 0 && op.out_elems_per_clip > m->buf_elems[1][0])
      m->buf_elems[1][0] = op.out_elems_per_clip;
  }
  // the late X region is addressed compactly at [clip * hand-off elems]: no rounding of buf_elems[1][0], and every
  // late X tensor must fit the hand-off row
  for (int sgi = 0; sgi < 2; ++sgi)
    for (int i = 0; i < 3; ++i)
      if (!(sgi == 1 && i == 0)) m->buf_elems[sgi][i] = round_up(m->buf_elems[sgi][i], 64);
  if (m->split_op > 0 && m->split_op < (int)m->ops.size() &&
      m->buf_elems[1][0] != m->ops[m->split_op - 1].out_elems_per_clip) {
    set_error("kws_embed_create: schedule split needs the hand-off tensor to be the largest late trunk tensor");
    for (void* p : m->dev_allocs) cudaFree(p);
    delete m;
    return KWS_ERR_UNSUPPORTED;
  }
  m->tail_op = m->split_op;
  for (size_t i = 0; i < m->ops.size(); ++i)
    if (m->ops[i].name == "block3b_out") m->tail_op = (int)i + 1;
  m->flops_per_clip = 2.0 * macs;
  if (const char* env = getenv("KWS_FUSE")) m->fuse = atoi(env) < 0 ? 0 : (atoi(env) > 3 ? 3 : atoi(env));
  {
    const int rc = build_fused_plan(m);
    if (rc != KWS_OK) {
      for (void* p : m->dev_allocs) cudaFree(p);
      delete m;
      return rc;
    }
  }
  *out = m;
  return KWS_OK;
}

extern "C" void kws_embed_destroy(kws_embed_t* m) {
  if (!m) return;
  for (auto& g : m->graphs) cudaGraphExecDestroy(g.exec);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  for (void* p : m->dev_allocs) cudaFree(p);
  delete m;
}

extern "C" int kws_embed_info(const kws_embed_t* m, int* in_h, int* in_w, int* out_dim, int* n_ops, double* flops_per_clip) {
  KWS_REQUIRE(m != nullptr, "kws_embed_info: NULL handle");
  if (in_h) *in_h = m->H;
  if (in_w) *in_w = m->W;
  if (out_dim) *out_dim = m->out_dim;
  if (n_ops) *n_ops = (int)m->ops.size();
  if (flops_per_clip) *flops_per_clip = m->flops_per_clip;
  return KWS_OK;
}

extern "C" int kws_embed_op_name(const kws_embed_t* m, int op, char* buf, size_t buf_bytes, int64_t* out_elems_per_clip) {
  KWS_REQUIRE(m && op >= 0 && op < (int)m->ops.size(), "kws_embed_op_name: bad op index");
  if (buf && buf_bytes) {
    strncpy(buf, m->ops[op].name.c_str(), buf_bytes - 1);
    buf[buf_bytes - 1] = 0;
  }
  if (out_elems_per_clip) *out_elems_per_clip = (int64_t)m->ops[op].out_elems_per_clip;
  return KWS_OK;
}

// kind: 0 stem, 1 gemm (tcgen05), 2 depthwise+SE.  flops / bytes are per clip (bytes: activations only).
extern "C" int kws_embed_op_info(const kws_embed_t* m, int op, int* kind, double* flops_per_clip, double* bytes_per_clip,
                                 int* gemm_n, int* gemm_k, int* rows_per_clip) {
  KWS_REQUIRE(m && op >= 0 && op < (int)m->ops.size(), "kws_embed_op_info: bad op index");
  const Op& o = m->ops[op];
  if (kind) *kind = o.kind == kOpStem ? 0 : (o.kind == kOpGemm ? 1 : 2);
  if (flops_per_clip) *flops_per_clip = o.flops_per_clip;
  if (bytes_per_clip) *bytes_per_clip = o.bytes_per_clip;
  if (gemm_n) *gemm_n = o.N;
  if (gemm_k) *gemm_k = o.K;
  if (rows_per_clip) *rows_per_clip = o.rows_per_clip;
  return KWS_OK;
}

extern "C" int kws_embed_set_chunk(kws_embed_t* m, int chunk) {
  KWS_REQUIRE(m && chunk >= 1, "kws_embed_set_chunk: bad argument");
  m->chunk = chunk;
  return KWS_OK;
}

extern "C" int kws_embed_set_chunk_late(kws_embed_t* m, int chunk) {
  KWS_REQUIRE(m && chunk >= 1, "kws_embed_set_chunk_late: bad argument");
  m->chunk_late = chunk;
  return KWS_OK;
}

// Number of kernel launches one forward pass of `batch` clips issues (both schedule segments).
extern "C" int kws_embed_launches(const kws_embed_t* m, int batch) {
  if (!m || batch <= 0) return 0;
  const int ce = batch < m->chunk ? batch : m->chunk, cl = batch < m->chunk_late ? batch : m->chunk_late;
  const int n_ops = (int)m->ops.size();
  int launches = 0;
  for (int i = 0; i < n_ops; ++i) {
    int per = (m->ops[i].kind == kOpDwse && m->ops[i].dw.se_external) ? 4 : 1;   // dw+pool, 2 SE GEMMs, gating
    for (const FusedSegment& c : m->segs[m->fuse])
      if (i >= c.op_lo && i <= c.op_hi) per = i == c.op_hi ? 1 : 0;               // one launch for the whole run
    launches += per * (i < m->split_op ? (batch + ce - 1) / ce : (batch + cl - 1) / cl);
  }
  return launches;
}

// 0: layer-by-layer schedule; 1: one fused launch per MBConv block of the tail; 2 (default): runs of blocks per launch
extern "C" int kws_embed_set_fuse(kws_embed_t* m, int mode) {
  KWS_REQUIRE(m != nullptr && mode >= 0 && mode <= 3, "kws_embed_set_fuse: bad argument");
  m->fuse = mode;
  return KWS_OK;
}

extern "C" int kws_embed_set_graph(kws_embed_t* m, int enable) {
  KWS_REQUIRE(m != nullptr, "kws_embed_set_graph: NULL handle");
  m->use_graph = enable ? 1 : 0;
  return KWS_OK;
}

extern "C" size_t kws_embed_workspace_bytes(const kws_embed_t* m, int batch) {
  if (!m || batch < 0) return 0;
  const size_t ce = (size_t)(batch < m->chunk ? batch : m->chunk);
  const size_t cl = (size_t)(batch < m->chunk_late ? batch : m->chunk_late);
  const size_t early = (m->buf_elems[0][0] + m->buf_elems[0][1] + m->buf_elems[0][2]) * ce;
  const size_t late = m->buf_elems[1][0] * (size_t)batch + (m->buf_elems[1][1] + m->buf_elems[1][2]) * cl;
  const size_t se_scratch = (ce > cl ? ce : cl) * (2 * m->max_se_channels + 64);   // pooled means, gates, squeeze (16-bit)
  return (early + late + se_scratch) * 2 + 4096;
}

// tap_op >= 0: additionally copy the output of op `tap_op` (16-bit NHWC, or fp32 for the last op) to d_tap.
static int run_ops(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace, int tap_op,
                   void* d_tap, float* host_op_ms, cudaStream_t st, int sm_head, int sm_tail);

// The launch list of one forward pass is fixed for given buffers / batch, so it is captured once into a CUDA
// graph (35 tensor-map encodes + ~55 launches per chunk collapse into one cudaGraphLaunch) and replayed.
static int embed_forward_impl(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                              size_t ws_bytes, int tap_op, void* d_tap, float* host_op_ms, void* stream, int sm_head = 0,
                              int sm_tail = 0) {
  KWS_REQUIRE(m != nullptr, "kws_embed_forward: NULL handle");
  KWS_REQUIRE(batch >= 0, "kws_embed_forward: negative batch");
  if (batch == 0) return KWS_OK;
  KWS_REQUIRE(d_feats && d_emb && d_workspace, "kws_embed_forward: NULL device buffer");
  KWS_REQUIRE(ws_bytes >= kws_embed_workspace_bytes(m, batch), "kws_embed_forward: workspace too small");
  KWS_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "kws_embed_forward: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (m->use_graph && tap_op < 0 && !host_op_ms && cudaStreamIsCapturing(st, &cap) == cudaSuccess &&
      cap == cudaStreamCaptureStatusNone) {
    const long long sched_key = ((((long long)m->chunk * 100003 + m->chunk_late) * 1024 + sm_head) * 1024 + sm_tail) * 8 + m->fuse;
    for (auto& g : m->graphs)
      if (g.feats == d_feats && g.emb == d_emb && g.ws == d_workspace && g.batch == batch && g.chunk == sched_key) {
        KWS_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
        return KWS_OK;
      }
    // A key is captured the second time it is met: callers that pass fresh buffers on every call (a new output tensor
    // per predict() chunk) would otherwise pay a stream capture + cudaGraphInstantiate per call and churn the cache.
    {
      bool met = false;
      for (size_t i = 0; i < m->seen.size() && !met; ++i) {
        const GraphEntry& g = m->seen[i];
        if (g.feats == d_feats && g.emb == d_emb && g.ws == d_workspace && g.batch == batch && g.chunk == sched_key) {
          m->seen.erase(m->seen.begin() + (long)i);
          met = true;
        }
      }
      if (!met) {
        if (m->seen.size() >= 64) m->seen.erase(m->seen.begin());
        m->seen.push_back(GraphEntry{d_feats, d_emb, d_workspace, batch, sched_key, nullptr});
        return run_ops(m, d_feats, batch, d_emb, d_workspace, -1, nullptr, nullptr, st, sm_head, sm_tail);
      }
    }
    if (!m->cap_stream) KWS_CUDA_CHECK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
    KWS_CUDA_CHECK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = run_ops(m, d_feats, batch, d_emb, d_workspace, -1, nullptr, nullptr, m->cap_stream, sm_head, sm_tail);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(m->cap_stream, &graph);
    if (rc != KWS_OK) {
      if (graph) cudaGraphDestroy(graph);
      return rc;
    }
    KWS_CUDA_CHECK(ce);
    GraphEntry e{d_feats, d_emb, d_workspace, batch, sched_key, nullptr};
    const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
    cudaGraphDestroy(graph);
    KWS_CUDA_CHECK(ie);
    if (m->graphs.size() >= 32) {
      cudaGraphExecDestroy(m->graphs.front().exec);
      m->graphs.erase(m->graphs.begin());
    }
    m->graphs.push_back(e);
    KWS_CUDA_CHECK(cudaGraphLaunch(e.exec, st));
    return KWS_OK;
  }
  return run_ops(m, d_feats, batch, d_emb, d_workspace, tap_op, d_tap, host_op_ms, st, sm_head, sm_tail);
}

static int run_ops(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace, int tap_op,
                   void* d_tap, float* host_op_ms, cudaStream_t st, int sm_head, int sm_tail) {
  const int sms_head = (sm_head > 0 && sm_head < m->sm_count) ? sm_head : m->sm_count;
  const int sms_tail = (sm_tail > 0 && sm_tail < m->sm_count) ? sm_tail : m->sm_count;
  const int n_ops = (int)m->ops.size();
  const int chunk_seg[2] = {batch < m->chunk ? batch : m->chunk, batch < m->chunk_late ? batch : m->chunk_late};
  // a tap inside a fused run (not its last op) needs the layer-wise schedule
  int fuse_mode = m->fuse;
  if (tap_op >= 0)
    for (const FusedSegment& c : m->segs[fuse_mode])
      if (tap_op >= c.op_lo && tap_op < c.op_hi) fuse_mode = 0;
  // workspace carve-up: [Xe | Ee | De] (early chunk) [H = late X, whole batch] [El | Dl] (late chunk)
  uint16_t* base = static_cast<uint16_t*>(d_workspace);
  uint16_t* early[3];
  early[0] = base;
  early[1] = early[0] + m->buf_elems[0][0] * chunk_seg[0];
  early[2] = early[1] + m->buf_elems[0][1] * chunk_seg[0];
  uint16_t* H = early[2] + m->buf_elems[0][2] * chunk_seg[0];
  uint16_t* late_e = H + m->buf_elems[1][0] * (size_t)batch;
  uint16_t* late_d = late_e + m->buf_elems[1][1] * chunk_seg[1];
  const size_t se_rows = (size_t)(chunk_seg[0] > chunk_seg[1] ? chunk_seg[0] : chunk_seg[1]);
  uint16_t* se_pooled = reinterpret_cast<uint16_t*>(
      (reinterpret_cast<uintptr_t>(late_d + m->buf_elems[1][2] * chunk_seg[1]) + 255) & ~(uintptr_t)255);
  uint16_t* se_gates = se_pooled + se_rows * m->max_se_channels;
  uint16_t* se_squeeze = se_gates + se_rows * m->max_se_channels;

  std::vector<cudaEvent_t> evs;
  size_t ev_i = 0;
  if (host_op_ms) {
    size_t n = 0;
    for (int sgi = 0; sgi < 2; ++sgi) {
      const int ops_in = sgi == 0 ? m->split_op : n_ops - m->split_op;
      if (ops_in > 0) n += (size_t)((batch + chunk_seg[sgi] - 1) / chunk_seg[sgi]) * (ops_in + 1);
    }
    evs.resize(n);
    for (auto& e : evs) KWS_CUDA_CHECK(cudaEventCreate(&e));
    for (int i = 0; i < n_ops; ++i) host_op_ms[i] = 0.f;
  }
  std::vector<int> ev_op;   // op index ending at event i (-1 = chunk start marker)

  for (int sgi = 0; sgi < 2; ++sgi) {
    const int op_lo = sgi == 0 ? 0 : m->split_op, op_hi = sgi == 0 ? m->split_op : n_ops;
    if (op_lo >= op_hi) continue;
    const int chunk = chunk_seg[sgi];
    for (int b0 = 0; b0 < batch; b0 += chunk) {
      const int nb = batch - b0 < chunk ? batch - b0 : chunk;
      uint16_t* bufs[3];
      if (sgi == 0) { bufs[0] = early[0]; bufs[1] = early[1]; bufs[2] = early[2]; }
      else { bufs[0] = H + m->buf_elems[1][0] * (size_t)b0; bufs[1] = late_e; bufs[2] = late_d; }
      if (host_op_ms) { KWS_CUDA_CHECK(cudaEventRecord(evs[ev_i++], st)); ev_op.push_back(-1); }
      int trunk = 0;                    // buffer holding the current block input: 0 = X, 2 = D (after a fused launch)
      for (int oi = op_lo; oi < op_hi; ++oi) {
        if (m->stop_after_tap && tap_op >= 0 && oi > tap_op) break;
        const Op& op = m->ops[oi];
        const int sms = oi >= m->tail_op ? sms_tail : sms_head;
        const FusedSegment* sg = nullptr;
        if (sgi == 1)
          for (const FusedSegment& c : m->segs[fuse_mode])
            if (c.op_lo == oi) sg = &c;
        if (sg) {
          // a run of MBConv blocks (and possibly the top conv) in one launch: mbconv_fused.cu
          const Op& lastop = m->ops[sg->op_hi];
          void* in_ptr = bufs[trunk];
          void* outp;
          if (lastop.out_buf != 0) outp = bufs[lastop.out_buf];
          else if (sg->in_place) outp = in_ptr;
          else { trunk = trunk == 0 ? 2 : 0; outp = bufs[trunk]; }
          const int rcf = launch_mbconv_fused(in_ptr, nb, m->d_fblocks + sg->blk_lo, m->finfos.data() + sg->blk_lo,
                                              sg->nblocks, outp, m->bf16, sms, m->max_smem, st);
          if (rcf != KWS_OK) return rcf;
          if (host_op_ms) { KWS_CUDA_CHECK(cudaEventRecord(evs[ev_i++], st)); ev_op.push_back(sg->op_hi); }
          if (sg->op_hi == tap_op && d_tap)
            KWS_CUDA_CHECK(cudaMemcpyAsync(static_cast<uint8_t*>(d_tap) + (size_t)b0 * lastop.out_elems_per_clip * 2, outp,
                                           (size_t)nb * lastop.out_elems_per_clip * 2, cudaMemcpyDeviceToDevice, st));
          oi = sg->op_hi;
          continue;
        }
        if (trunk != 0 && (op.in_buf == 0 || op.res_buf == 0)) {
          // a layer-wise op follows a fused launch that left the trunk in the D buffer: bring it back to X
          const size_t el = m->ops[oi - 1].out_elems_per_clip;
          KWS_CUDA_CHECK(cudaMemcpyAsync(bufs[0], bufs[2], (size_t)nb * el * 2, cudaMemcpyDeviceToDevice, st));
          trunk = 0;
        }
        void* out_ptr = op.out_buf >= 0 ? (void*)bufs[op.out_buf] : (void*)(d_emb + (size_t)b0 * m->out_dim);
        if (sgi == 0 && oi == m->split_op - 1 && op.out_buf == 0)      // hand-off: early chunk -> late X (whole batch)
          out_ptr = H + m->buf_elems[1][0] * (size_t)b0;
        int rc = KWS_OK;
        if (op.kind == kOpStem) {
          rc = launch_stem(d_feats + (size_t)b0 * m->H * m->W, nb, op.stem, out_ptr, sms, st);
        } else if (op.kind == kOpDwse) {
          DwseParams P = op.dw;
          P.pooled_out = se_pooled;
          rc = launch_dwse(bufs[op.in_buf], nb, P, out_ptr, dwse_pick_group(P, m->max_smem, nb, sms), sms, st);
          if (rc == KWS_OK && P.se_external) {
            GemmEpilogue e1;
            e1.bias = op.se_b1; e1.residual = nullptr; e1.out = se_squeeze; e1.ldo = op.se_pad; e1.ldr = op.se_pad;
            e1.act = kActSwish; e1.out_f32 = 0; e1.gap4 = 0; e1.bf16 = m->bf16;
            rc = gemm_h16(se_pooled, op.se_w1, nb, op.se_pad, P.C, 0, e1, sms, st);
            if (rc == KWS_OK) {
              GemmEpilogue e2 = e1;
              // (applying the gates inside this epilogue measured slower than the separate coalesced gating pass:
              //  one thread per clip would walk the pixels with a 2*C-byte stride)
              e2.bias = P.b_se2; e2.out = se_gates; e2.ldo = P.C; e2.ldr = P.C; e2.act = kActSigmoid;
              rc = gemm_h16(se_squeeze, op.se_w2, nb, P.C, op.se_pad, 0, e2, sms, st);
            }
            if (rc == KWS_OK) rc = launch_se_scale(out_ptr, se_gates, nb, P.Ho * P.Wo, P.C, m->bf16, sms, st);
          }
        } else {
          GemmEpilogue ep;
          ep.bias = op.bias;
          ep.residual = op.res_buf >= 0 ? bufs[op.res_buf] : nullptr;
          ep.out = out_ptr; ep.ldo = op.N; ep.ldr = op.N;
          ep.act = op.act; ep.out_f32 = op.out_f32; ep.gap4 = op.gap4; ep.bf16 = m->bf16;
          rc = gemm_h16(bufs[op.in_buf], op.w, op.rows_per_clip * nb, op.N, op.K, 0, ep, sms, st);
        }
        if (rc != KWS_OK) return rc;
        if (host_op_ms) { KWS_CUDA_CHECK(cudaEventRecord(evs[ev_i++], st)); ev_op.push_back(oi); }
        if (oi == tap_op && d_tap) {
          const size_t esz = op.out_f32 ? 4 : 2;
          KWS_CUDA_CHECK(cudaMemcpyAsync(static_cast<uint8_t*>(d_tap) + (size_t)b0 * op.out_elems_per_clip * esz, out_ptr,
                                         (size_t)nb * op.out_elems_per_clip * esz, cudaMemcpyDeviceToDevice, st));
        }
      }
    }
  }
  if (host_op_ms) {
    KWS_CUDA_CHECK(cudaStreamSynchronize(st));
    for (size_t i = 1; i < ev_i; ++i)
      if (ev_op[i] >= 0) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, evs[i - 1], evs[i]);
        host_op_ms[ev_op[i]] += ms;
      }
    for (auto& e : evs) cudaEventDestroy(e);
  }
  return KWS_OK;
}

extern "C" int kws_embed_forward_tap(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                                     size_t ws_bytes, int tap_op, void* d_tap, void* stream) {
  return embed_forward_impl(m, d_feats, batch, d_emb, d_workspace, ws_bytes, tap_op, d_tap, nullptr, stream);
}

// The frozen part of the network only: runs ops 0 .. tap_op and copies that op's output (16-bit NHWC) to d_tap; nothing
// after it is executed and no embedding is written.  Used by the phase-2 fine-tune step, whose trainable tail starts at the
// tapped tensor (finetune.TailTrainer).
extern "C" int kws_embed_forward_until(kws_embed_t* m, const float* d_feats, int batch, void* d_workspace, size_t ws_bytes,
                                       int tap_op, void* d_tap, void* stream) {
  KWS_REQUIRE(m && d_tap && tap_op >= 0 && tap_op + 1 < (int)m->ops.size(), "kws_embed_forward_until: bad argument");
  m->stop_after_tap = 1;
  float dummy;
  const int rc = embed_forward_impl(m, d_feats, batch, &dummy, d_workspace, ws_bytes, tap_op, d_tap, nullptr, stream);
  m->stop_after_tap = 0;
  return rc;
}

extern "C" int kws_embed_forward(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                                 size_t ws_bytes, void* stream) {
  return embed_forward_impl(m, d_feats, batch, d_emb, d_workspace, ws_bytes, -1, nullptr, nullptr, stream);
}

// Throughput schedule: as kws_embed_forward, with the grids of the network's head (stem ... block3b) sized for sm_head
// SMs and those of its latency-bound tail for sm_tail SMs (0 = all).  Meant for callers that keep several forward
// passes in flight on different streams (EmbedPipeline): results are identical, a single pass gets slower, the
// overlapped throughput higher (profiles/README.md item 11).
extern "C" int kws_embed_forward_budget(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                                        size_t ws_bytes, int sm_head, int sm_tail, void* stream) {
  KWS_REQUIRE(sm_head >= 0 && sm_tail >= 0 && sm_head < 1024 && sm_tail < 1024, "kws_embed_forward_budget: bad SM budget");
  return embed_forward_impl(m, d_feats, batch, d_emb, d_workspace, ws_bytes, -1, nullptr, nullptr, stream, sm_head, sm_tail);
}

// Profiling variant: CUDA events around every op on `stream`; host_op_ms[n_ops] receives the per-op device time
// (summed over chunks).  Synchronises the stream.
extern "C" int kws_embed_forward_timed(kws_embed_t* m, const float* d_feats, int batch, float* d_emb, void* d_workspace,
                                       size_t ws_bytes, float* host_op_ms, void* stream) {
  KWS_REQUIRE(host_op_ms != nullptr, "kws_embed_forward_timed: NULL timing buffer");
  return embed_forward_impl(m, d_feats, batch, d_emb, d_workspace, ws_bytes, -1, nullptr, host_op_ms, stream);
}

// Standalone tensor-core contraction with the fused epilogue (the "pointwise conv / dense" operator):
// out[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual); A, W, residual fp16 (dtype 0) or bf16 (1); out same or fp32.
extern "C" int kws_gemm_h16(const void* d_a, const void* d_w, int M, int N, int K, const float* d_bias, int act,
                            const void* d_residual, void* d_out, int out_f32, int gap4, int block_n, int dtype,
                            void* stream) {
  KWS_REQUIRE(d_a && d_w && d_out && M >= 0 && N > 0 && K > 0 && (dtype == 0 || dtype == 1), "kws_gemm_h16: bad argument");
  if (M == 0) return KWS_OK;
  const int sm = device_sm_count();
  KWS_REQUIRE(sm > 0, "kws_gemm_h16: CUDA device required; there is no CPU fallback");
  GemmEpilogue ep;
  ep.bias = d_bias; ep.residual = d_residual;
  ep.out = d_out; ep.ldo = N; ep.ldr = N; ep.act = act; ep.out_f32 = out_f32; ep.gap4 = gap4; ep.bf16 = dtype;
  return gemm_h16(d_a, d_w, M, N, K, block_n, ep, sm, (cudaStream_t)stream);
}
