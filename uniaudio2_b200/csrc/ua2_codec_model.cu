// Handle API of the SEANet + transformer + residual-VQ codec ("Mimi twin"): encode (wav -> codes) and decode
// (codes -> wav), sequencing the kernels of ua2_codec.cu, the skinny-linear family and the attention kernel.
//
// Replaces tools/tokenizer/MimiCodec/model/models/MimiCodec.py:93-110 (MimiCodec.encode / .decode), whose modules are
// the byte-identical, importable twins of llm_modules/{seanet,conv,resample,transformer,rope}.py.
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"

extern "C" {
int ua2_conv1d_causal_f32(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int,
                          int, int, void*);
int ua2_conv1d_causal_gemm_f32(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int,
                               int, int, void*);
int ua2_convtr1d_causal_f32(const float*, const float*, const float*, float*, int, int, int, int, int, int, void*);
int ua2_convtr1d_depthwise_f32(const float*, const float*, float*, int, int, int, int, void*);
int ua2_convtr1d_repack_phase_f32(const float*, float*, int, int, int, void*);
int ua2_convtr1d_causal_gemm_f32(const float*, const float*, const float*, float*, int, int, int, int, int, int, void*);
int ua2_rvq_encode_f32(const float*, const float*, const float*, int64_t*, int, int, int, int, int, int, int, void*);
int ua2_rvq_decode_f32(const int64_t*, const float*, float*, int, int, int, int, int, int, int, void*);
int ua2_rvq_encode_gemm_f32(float*, const float*, const float*, float*, int64_t*, int, int, int, int, int, int, int, void*);
}

namespace ua2 {
namespace {

// (Cout, Cin, K) -> (Cin, K, Cout)   [conv]   or   (Cin, Cout, K) -> (Cin, K, Cout)   [transposed conv]
__global__ void repack_conv_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int Cin, int K,
                                   int transposed) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)Cout * Cin * K;
  if (i >= total) return;
  const int co = (int)(i % Cout);
  const int k = (int)((i / Cout) % K);
  const int ci = (int)(i / ((size_t)Cout * K));
  const size_t s = transposed ? ((size_t)ci * Cout + co) * K + k : ((size_t)co * Cin + ci) * K + k;
  dst[i] = src[s];
}

// EuclideanCodebook.embedding (core_vq.py:142-150) and its row squared norms
__global__ void codebook_kernel(const float* __restrict__ esum, const float* __restrict__ usage, float* __restrict__ emb,
                                float* __restrict__ sq, int K, int D, float eps) {
  const int j = blockIdx.x;
  const float u = fmaxf(usage[j], eps);
  float acc = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float e = esum[(size_t)j * D + d] / u;
    emb[(size_t)j * D + d] = e;
    acc = fmaf(e, e, acc);
  }
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    sq[j] = t;
  }
}

// (B, C, T) <-> (B, T, C)
__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int Cc) {
  // src (batch, R, Cc) -> dst (batch, Cc, R)
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* s = src + (size_t)b * R * Cc;
  float* d = dst + (size_t)b * R * Cc;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? s[(size_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) d[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

__global__ void posidx_kernel(int32_t* pos, int32_t* bidx, int M, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) {
    pos[i] = i % T;
    bidx[i] = i / T;
  }
}

struct ConvW {
  const float* w_src = nullptr;  // torch layout
  const float* bias = nullptr;
  float* w = nullptr;  // repacked (Cin, K, Cout) [direct kernels] or (stride, Cout, Cin, 2) [transposed conv, phase GEMM]
  int cout = 0, cin = 0, k = 0, transposed = 0;
};

struct TLayer {
  const float *in_proj = nullptr, *out_proj = nullptr, *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr,
              *lin1 = nullptr, *lin2 = nullptr, *ls1 = nullptr, *ls2 = nullptr;
};

struct Rvq {
  const float *in_proj = nullptr, *out_proj = nullptr;  // (Dq, D, 1) / (D, Dq, 1): 1x1 convs, repacked like convs
  float *in_w = nullptr, *out_w = nullptr;
  std::vector<const float*> esum, usage;
  float* emb = nullptr;  // (n_q, K, Dq)
  float* sq = nullptr;   // (n_q, K)
  int n_q = 0;
};

}  // namespace
}  // namespace ua2

using namespace ua2;

struct ua2_codec {
  ua2_codec_cfg cfg{};
  std::map<std::string, ConvW> convs;  // keyed by the reference prefix, e.g. "encoder.model.0.conv.conv"
  const float *down_w_src = nullptr, *up_w = nullptr;
  float* down_w = nullptr;
  std::vector<TLayer> tl[2];  // encoder_transformer, decoder_transformer
  Rvq rvq[2];                 // rvq_first, rvq_rest
  bool ready = false;
  std::vector<void*> owned;
  // workspace
  float* ws = nullptr;
  size_t ws_floats = 0;
  int32_t *pos = nullptr, *bidx = nullptr;
  size_t posidx_cap = 0;
  TcWorkspace tc;  // scratch of the tcgen05 3xTF32 path for the transformer linears (grown on demand)
  size_t tc_rows = 0;
  int hop = 1, rs = 1;
};

namespace {

int calloc_dev(ua2_codec* h, void** p, size_t bytes) {
  UA2_CHECK_CUDA(cudaMalloc(p, bytes));
  h->owned.push_back(*p);
  return UA2_OK;
}

int reserve_ws(ua2_codec* h, size_t floats) {
  if (floats <= h->ws_floats) return UA2_OK;
  if (h->ws) {
    UA2_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(h->ws);
    h->ws = nullptr;
  }
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->ws, floats * sizeof(float)));
  h->ws_floats = floats;
  return UA2_OK;
}

int reserve_posidx(ua2_codec* h, size_t M) {
  if (M <= h->posidx_cap) return UA2_OK;
  if (h->pos) {
    UA2_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(h->pos);
    cudaFree(h->bidx);
  }
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->pos, M * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->bidx, M * 4));
  h->posidx_cap = M;
  return UA2_OK;
}

// split operands + raw product of the largest transformer linear at M rows (ua2_tcgemm.cu)
int reserve_tc(ua2_codec* h, size_t M) {
  if (!tc_gemm_available() || M > 65536) return UA2_OK;  // larger batches stay on the fp32 SIMT tiles
  if (M <= h->tc_rows) return UA2_OK;
  const size_t C = h->cfg.latent_dim, F = h->cfg.dim_feedforward;
  if (h->tc.a) {
    UA2_CHECK_CUDA(cudaDeviceSynchronize());
    cudaFree(h->tc.a);
    cudaFree(h->tc.slots);
    cudaFree(h->tc.c);
    h->tc = TcWorkspace();
    h->tc_rows = 0;
  }
  const size_t kmax = std::max(C, F), nmax = std::max(3 * C, F);
  h->tc.a_floats = M * 2 * kmax;
  h->tc.slots_floats = tc_slots_max_floats();
  h->tc.c_floats = M * nmax;
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->tc.a, h->tc.a_floats * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->tc.slots, h->tc.slots_floats * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->tc.c, h->tc.c_floats * 4));
  h->tc_rows = M;
  return UA2_OK;
}

const ConvW* find_conv(ua2_codec* h, const std::string& key) {
  auto it = h->convs.find(key);
  return it == h->convs.end() ? nullptr : &it->second;
}

#define RUN(expr)              \
  do {                         \
    int _rc = (expr);          \
    if (_rc != UA2_OK) return _rc; \
  } while (0)

int conv(ua2_codec* h, const std::string& key, const float* x, const float* res, float* y, int B, int T, int stride, int pre_elu,
         void* st) {
  const ConvW* c = find_conv(h, key);
  UA2_REQUIRE(c && c->w, key + ": conv weight missing");
  return ua2_conv1d_causal_gemm_f32(x, c->w_src, c->bias, res, y, B, c->cin, c->cout, T, c->k, stride, 1, pre_elu, 0, st);
}

// SEANetResnetBlock (modules/seanet.py:21-94): y = x + conv_k1(ELU(conv_k3(ELU(x)))); `tmp` holds the hidden activation of the
// two-launch form.  Option "resblock_fused": one kernel for the 64 -> 32 -> 64 blocks (ua2_resblock.cu).
int resblock(ua2_codec* h, const std::string& p, const float* x, float* tmp, float* y, int B, int T, void* st) {
  // the fused fp32 kernel serves builds / options without the tensor-core convolution (two conv_umma launches: 3.1 ms against 4.1 ms
  // for the 64-channel block at batch 16 x 10 s)
  if (get_resblock_fused() && !get_conv_umma()) {
    const ConvW *c1 = find_conv(h, p + "1.conv.conv"), *c3 = find_conv(h, p + "3.conv.conv");
    if (c1 && c3 && c1->w_src && c3->w_src && c1->k == 3 && c3->k == 1 && c1->cout == c3->cin && c1->cin == c3->cout) {
      LaunchCtx lc;
      lc.stream = (cudaStream_t)st;
      const cudaError_t e = launch_resblock_fused(lc, x, c1->w_src, c1->bias, c3->w_src, c3->bias, y, B, c1->cin, c1->cout, T);
      if (e == cudaSuccess) return UA2_OK;
      if (e != cudaErrorNotSupported) UA2_CHECK_CUDA(e);
    }
  }
  RUN(conv(h, p + "1.conv.conv", x, nullptr, tmp, B, T, 1, 1, st));
  return conv(h, p + "3.conv.conv", tmp, x, y, B, T, 1, 1, st);  // x + block(x)
}

// ProjectedTransformer(conv_layout=True) over x (B, C, T) in place; tmp buffers carved from the workspace by the caller
int run_transformer(ua2_codec* h, int which, float* x_bct, float* xt, float* qbuf, float* kbuf, float* vbuf, float* hbuf,
                    float* o_part, float* ml_part, float* sg_ws, size_t sg_ws_floats, int B, int T, void* st) {
  const int C = h->cfg.latent_dim, H = h->cfg.num_heads, hs = C / H, F = h->cfg.dim_feedforward;
  const int M = B * T;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)st;
  RUN(reserve_posidx(h, M));
  RUN(reserve_tc(h, M));
  const TcWorkspace* tcw = (h->tc.a != nullptr && (size_t)M <= h->tc_rows) ? &h->tc : nullptr;
  UA2_CHECK_CUDA(launch(lc, posidx_kernel, dim3((M + 255) / 256), dim3(256), 0, h->pos, h->bidx, M, T));
  // (B, C, T) -> (B, T, C)
  UA2_CHECK_CUDA(launch(lc, transpose_kernel, dim3((T + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, (const float*)x_bct, xt, C, T));
  const int max_splits = (T + ATTN_CHUNK - 1) / ATTN_CHUNK;
  const int CH = 32768;  // rows per attention launch (grid.z limit)
  for (size_t l = 0; l < h->tl[which].size(); ++l) {
    const TLayer& w = h->tl[which][l];
    {  // norm1 -> in_proj -> interleaved RoPE -> K/V (B, H, T, hs)
      GemvParams p;
      p.W = w.in_proj;
      p.N = 3 * C;
      p.K = C;
      p.M = M;
      p.ws = sg_ws;
      p.ws_floats = sg_ws_floats;
      p.tc = tcw;
      p.X = xt;
      p.ldx = C;
      p.norm_w = w.n1w;
      p.norm_b = w.n1b;
      p.eps = 1e-5f;
      p.pos = h->pos;
      p.bidx = h->bidx;
      p.n_head = H;
      p.n_groups = H;
      p.hs = hs;
      p.q_out = qbuf;
      p.k_cache = kbuf;
      p.v_cache = vbuf;
      p.S_max = T;
      p.rope_max_period = h->cfg.max_period;
      UA2_CHECK_CUDA(launch_gemv(lc, PRO_LAYERNORM, EPI_QKV_IL, p));
    }
    for (int r0 = 0; r0 < M; r0 += CH) {
      const int Mc = std::min(CH, M - r0);
      AttnParams a;
      a.q = qbuf + (size_t)r0 * C;
      a.k_cache = kbuf;
      a.v_cache = vbuf;
      a.pos = h->pos + r0;
      a.bidx = h->bidx + r0;
      a.o_part = o_part;
      a.ml_part = ml_part;
      a.M = Mc;
      a.n_head = H;
      a.n_groups = H;
      a.hs = hs;
      a.S_max = T;
      a.max_splits = max_splits;
      a.n_splits_launch = max_splits;
      a.window = h->cfg.context;
      UA2_CHECK_CUDA(launch_attn(lc, a));
      GemvParams p;  // combine -> out_proj -> x + layer_scale_1 * update
      p.W = w.out_proj;
      p.N = C;
      p.K = C;
      p.M = Mc;
      p.ws = sg_ws;
      p.ws_floats = sg_ws_floats;
      p.tc = tcw;
      p.o_part = o_part;
      p.ml_part = ml_part;
      p.max_splits = max_splits;
      p.n_splits = max_splits;
      p.pos = h->pos + r0;
      p.n_head = H;
      p.hs = hs;
      p.Y = xt + (size_t)r0 * C;
      p.ldy = C;
      p.R = xt + (size_t)r0 * C;
      p.ldr = C;
      p.scale = w.ls1;
      UA2_CHECK_CUDA(launch_gemv(lc, PRO_ATTN, EPI_SCALE_RESADD, p));
    }
    {  // norm2 -> linear1 -> gelu
      GemvParams p;
      p.W = w.lin1;
      p.N = F;
      p.K = C;
      p.M = M;
      p.ws = sg_ws;
      p.ws_floats = sg_ws_floats;
      p.tc = tcw;
      p.X = xt;
      p.ldx = C;
      p.norm_w = w.n2w;
      p.norm_b = w.n2b;
      p.eps = 1e-5f;
      p.Y = hbuf;
      p.ldy = F;
      UA2_CHECK_CUDA(launch_gemv(lc, PRO_LAYERNORM, EPI_GELU, p));
    }
    {  // linear2 -> x + layer_scale_2 * update
      GemvParams p;
      p.W = w.lin2;
      p.N = C;
      p.K = F;
      p.M = M;
      p.ws = sg_ws;
      p.ws_floats = sg_ws_floats;
      p.tc = tcw;
      p.X = hbuf;
      p.ldx = F;
      p.Y = xt;
      p.ldy = C;
      p.R = xt;
      p.ldr = C;
      p.scale = w.ls2;
      UA2_CHECK_CUDA(launch_gemv(lc, PRO_PLAIN, EPI_SCALE_RESADD, p));
    }
  }
  UA2_CHECK_CUDA(launch(lc, transpose_kernel, dim3((C + 31) / 32, (T + 31) / 32, B), dim3(32, 8), 0, (const float*)xt, x_bct, T, C));
  return UA2_OK;
}

size_t transformer_ws_floats(const ua2_codec_cfg& c, int B, int T) {
  const size_t M = (size_t)B * T, C = c.latent_dim;
  const size_t Mc = std::min<size_t>(M, 32768);
  const size_t splits = (T + ATTN_CHUNK - 1) / ATTN_CHUNK;
  // xt, q, k, v (M*C each), hbuf (M*F), o_part (Mc*C*splits), ml_part (Mc*H*splits*2)
  return 4 * M * C + M * c.dim_feedforward + Mc * C * splits + Mc * c.num_heads * splits * 2 + (2 * M + Mc * C) + 1024;
}

}  // namespace

extern "C" {

int ua2_codec_create(const ua2_codec_cfg* cfg, ua2_codec** out) {
  UA2_REQUIRE(cfg && out, "null cfg/out");
  UA2_REQUIRE(cfg->n_ratios >= 1 && cfg->n_ratios <= 8, "1..8 ratios");
  for (int i = 0; i < cfg->n_ratios; ++i) UA2_REQUIRE(cfg->ratios[i] >= 1 && cfg->ratios[i] <= 8, "ratios must be in 1..8");
  UA2_REQUIRE(cfg->latent_dim % cfg->num_heads == 0, "latent_dim % num_heads");
  const int hs = cfg->latent_dim / cfg->num_heads;
  UA2_REQUIRE(hs == 32 || hs == 64 || hs == 128, "transformer head size must be 32/64/128");
  UA2_REQUIRE(cfg->rvq_layers >= 2 && cfg->codebook_dim % 4 == 0, "rvq_layers >= 2, codebook_dim % 4 == 0");
  UA2_REQUIRE(cfg->resample_stride >= 1 && cfg->resample_stride <= 8, "resample stride 1..8");
  ua2_codec* h = new ua2_codec();
  h->cfg = *cfg;
  h->hop = 1;
  for (int i = 0; i < cfg->n_ratios; ++i) h->hop *= cfg->ratios[i];
  h->rs = cfg->resample_stride;
  h->tl[0].resize(cfg->num_layers);
  h->tl[1].resize(cfg->num_layers);
  h->rvq[0].n_q = 1;
  h->rvq[1].n_q = cfg->rvq_layers - 1;
  for (int g = 0; g < 2; ++g) {
    h->rvq[g].esum.assign(h->rvq[g].n_q, nullptr);
    h->rvq[g].usage.assign(h->rvq[g].n_q, nullptr);
  }
  *out = h;
  return UA2_OK;
}

int ua2_codec_destroy(ua2_codec* h) {
  if (!h) return UA2_OK;
  for (void* p : h->owned) cudaFree(p);
  if (h->ws) cudaFree(h->ws);
  if (h->pos) cudaFree(h->pos);
  if (h->bidx) cudaFree(h->bidx);
  if (h->tc.a) {
    cudaFree(h->tc.a);
    cudaFree(h->tc.slots);
    cudaFree(h->tc.c);
  }
  delete h;
  return UA2_OK;
}

int ua2_codec_load_weight(ua2_codec* h, const char* key_c, const float* dptr, const int64_t* shape, int ndim) {
  UA2_REQUIRE(h && key_c && dptr && shape, "null argument");
  const std::string key(key_c);
  auto ends = [&](const std::string& suf) { return key.size() >= suf.size() && key.compare(key.size() - suf.size(), suf.size(), suf) == 0; };
  auto starts = [&](const std::string& pre) { return key.compare(0, pre.size(), pre) == 0; };
  if (key == "downsample.conv.conv.conv.weight") {
    UA2_REQUIRE(ndim == 3, key + ": bad shape");
    ConvW& c = h->convs["downsample"];
    c.w_src = dptr;
    c.cout = (int)shape[0];
    c.cin = (int)shape[1];
    c.k = (int)shape[2];
    return UA2_OK;
  }
  if (key == "upsample.convtr.convtr.convtr.weight") {
    UA2_REQUIRE(ndim == 3 && shape[1] == 1, key + ": bad shape");
    h->up_w = dptr;
    return UA2_OK;
  }
  if (starts("encoder.model.") || starts("decoder.model.")) {
    const bool is_w = ends(".weight");
    UA2_REQUIRE(is_w || ends(".bias"), key + ": unknown parameter");
    const std::string base = key.substr(0, key.rfind('.'));
    ConvW& c = h->convs[base];
    if (is_w) {
      UA2_REQUIRE(ndim == 3, key + ": bad shape");
      c.w_src = dptr;
      c.transposed = base.find("convtr") != std::string::npos ? 1 : 0;
      if (c.transposed) {
        c.cin = (int)shape[0];
        c.cout = (int)shape[1];
      } else {
        c.cout = (int)shape[0];
        c.cin = (int)shape[1];
      }
      c.k = (int)shape[2];
    } else {
      c.bias = dptr;
    }
    return UA2_OK;
  }
  for (int t = 0; t < 2; ++t) {
    const std::string pre = std::string(t == 0 ? "encoder_transformer" : "decoder_transformer") + ".transformer.layers.";
    if (!starts(pre)) continue;
    const size_t dot = key.find('.', pre.size());
    UA2_REQUIRE(dot != std::string::npos, key + ": unknown parameter");
    const int l = atoi(key.substr(pre.size(), dot - pre.size()).c_str());
    UA2_REQUIRE(l >= 0 && l < (int)h->tl[t].size(), key + ": layer out of range");
    const std::string leaf = key.substr(dot + 1);
    TLayer& w = h->tl[t][l];
    const int C = h->cfg.latent_dim, F = h->cfg.dim_feedforward;
    auto is2 = [&](int64_t a, int64_t b) { return ndim == 2 && shape[0] == a && shape[1] == b; };
    auto is1 = [&](int64_t a) { return ndim == 1 && shape[0] == a; };
    if (leaf == "self_attn.in_proj_weight") {
      UA2_REQUIRE(is2(3 * C, C), key + ": bad shape");
      w.in_proj = dptr;
    } else if (leaf == "self_attn.out_proj.weight") {
      UA2_REQUIRE(is2(C, C), key + ": bad shape");
      w.out_proj = dptr;
    } else if (leaf == "norm1.weight") {
      UA2_REQUIRE(is1(C), key + ": bad shape");
      w.n1w = dptr;
    } else if (leaf == "norm1.bias") {
      UA2_REQUIRE(is1(C), key + ": bad shape");
      w.n1b = dptr;
    } else if (leaf == "norm2.weight") {
      UA2_REQUIRE(is1(C), key + ": bad shape");
      w.n2w = dptr;
    } else if (leaf == "norm2.bias") {
      UA2_REQUIRE(is1(C), key + ": bad shape");
      w.n2b = dptr;
    } else if (leaf == "linear1.weight") {
      UA2_REQUIRE(is2(F, C), key + ": bad shape");
      w.lin1 = dptr;
    } else if (leaf == "linear2.weight") {
      UA2_REQUIRE(is2(C, F), key + ": bad shape");
      w.lin2 = dptr;
    } else if (leaf == "layer_scale_1.scale") {
      UA2_REQUIRE(is1(C), key + ": bad shape");
      w.ls1 = dptr;
    } else if (leaf == "layer_scale_2.scale") {
      UA2_REQUIRE(is1(C), key + ": bad shape");
      w.ls2 = dptr;
    } else {
      UA2_REQUIRE(false, key + ": unknown parameter");
    }
    return UA2_OK;
  }
  for (int g = 0; g < 2; ++g) {
    const std::string pre = g == 0 ? "quantizer.rvq_first." : "quantizer.rvq_rest.";
    if (!starts(pre)) continue;
    Rvq& r = h->rvq[g];
    const std::string rest = key.substr(pre.size());
    if (rest == "input_proj.weight") {
      UA2_REQUIRE(ndim == 3 && shape[0] == h->cfg.codebook_dim && shape[1] == h->cfg.latent_dim && shape[2] == 1, key + ": bad shape");
      r.in_proj = dptr;
      return UA2_OK;
    }
    if (rest == "output_proj.weight") {
      UA2_REQUIRE(ndim == 3 && shape[1] == h->cfg.codebook_dim && shape[0] == h->cfg.latent_dim && shape[2] == 1, key + ": bad shape");
      r.out_proj = dptr;
      return UA2_OK;
    }
    const std::string lp = "vq.layers.";
    UA2_REQUIRE(rest.compare(0, lp.size(), lp) == 0, key + ": unknown parameter");
    const size_t dot = rest.find('.', lp.size());
    const int i = atoi(rest.substr(lp.size(), dot - lp.size()).c_str());
    UA2_REQUIRE(i >= 0 && i < r.n_q, key + ": quantizer index out of range");
    const std::string leaf = rest.substr(dot + 1);
    if (leaf == "_codebook.embedding_sum") {
      UA2_REQUIRE(ndim == 2 && shape[0] == h->cfg.codebook_size && shape[1] == h->cfg.codebook_dim, key + ": bad shape");
      r.esum[i] = dptr;
    } else if (leaf == "_codebook.cluster_usage") {
      UA2_REQUIRE(ndim == 1 && shape[0] == h->cfg.codebook_size, key + ": bad shape");
      r.usage[i] = dptr;
    } else if (leaf == "_codebook._initialized") {
      // training-time flag, unused at inference
    } else {
      UA2_REQUIRE(false, key + ": unknown parameter");
    }
    return UA2_OK;
  }
  if (key.compare(0, 23, "semantic_mapping_layer.") == 0) return UA2_OK;  // training-time distillation head (MimiCodec.py:15-23)
  UA2_REQUIRE(false, key + ": unknown parameter");
}

int ua2_codec_finalize(ua2_codec* h, void* stream) {
  UA2_REQUIRE(h, "null handle");
  if (h->ready) return UA2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LaunchCtx lc;
  lc.stream = st;
  int rc;
  UA2_REQUIRE(h->up_w, "upsample weight missing");
  for (auto& kv : h->convs) {
    ConvW& c = kv.second;
    UA2_REQUIRE(c.w_src, kv.first + ": weight missing");
    const size_t n = (size_t)c.cout * c.cin * c.k;
    if ((rc = calloc_dev(h, (void**)&c.w, n * 4))) return rc;
    if (c.transposed) {
      if ((rc = ua2_convtr1d_repack_phase_f32(c.w_src, c.w, c.cin, c.cout, c.k / 2, st))) return rc;
    } else {
      UA2_CHECK_CUDA(launch(lc, repack_conv_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, c.w_src, c.w, c.cout, c.cin, c.k, 0));
    }
  }
  for (int t = 0; t < 2; ++t)
    for (auto& w : h->tl[t])
      UA2_REQUIRE(w.in_proj && w.out_proj && w.n1w && w.n1b && w.n2w && w.n2b && w.lin1 && w.lin2 && w.ls1 && w.ls2,
                  "transformer layer weights missing");
  const int K = h->cfg.codebook_size, Dq = h->cfg.codebook_dim, D = h->cfg.latent_dim;
  for (int g = 0; g < 2; ++g) {
    Rvq& r = h->rvq[g];
    UA2_REQUIRE(r.in_proj && r.out_proj, "rvq projections missing");
    if ((rc = calloc_dev(h, (void**)&r.emb, (size_t)r.n_q * K * Dq * 4))) return rc;
    if ((rc = calloc_dev(h, (void**)&r.sq, (size_t)r.n_q * K * 4))) return rc;
    for (int i = 0; i < r.n_q; ++i) {
      UA2_REQUIRE(r.esum[i] && r.usage[i], "codebook buffers missing");
      UA2_CHECK_CUDA(launch(lc, codebook_kernel, dim3(K), dim3(128), 0, r.esum[i], r.usage[i], r.emb + (size_t)i * K * Dq,
                            r.sq + (size_t)i * K, K, Dq, 1e-5f));
    }
    // 1x1 convs: (Dq, D, 1) -> (D, 1, Dq) ; (D, Dq, 1) -> (Dq, 1, D)
    if ((rc = calloc_dev(h, (void**)&r.in_w, (size_t)Dq * D * 4))) return rc;
    if ((rc = calloc_dev(h, (void**)&r.out_w, (size_t)Dq * D * 4))) return rc;
    const size_t n = (size_t)Dq * D;
    UA2_CHECK_CUDA(launch(lc, repack_conv_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, r.in_proj, r.in_w, Dq, D, 1, 0));
    UA2_CHECK_CUDA(launch(lc, repack_conv_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, r.out_proj, r.out_w, D, Dq, 1, 0));
  }
  UA2_CHECK_CUDA(cudaStreamSynchronize(st));
  h->ready = true;
  return UA2_OK;
}

int64_t ua2_codec_frames(ua2_codec* h, int64_t T_samples) {
  if (!h) return -1;
  int64_t T = T_samples;
  for (int i = h->cfg.n_ratios - 1; i >= 0; --i) T = (T + h->cfg.ratios[i] - 1) / h->cfg.ratios[i];
  return (T + h->rs - 1) / h->rs;
}

int ua2_codec_encode(ua2_codec* h, const float* wav, int B, int T, int64_t* codes, void* st) {
  UA2_REQUIRE(h && h->ready && wav && codes, "bad argument (finalize the handle first)");
  UA2_REQUIRE(B >= 1 && T >= 1, "bad shape");
  const ua2_codec_cfg& c = h->cfg;
  const int nf = c.n_filters, D = c.latent_dim, Dq = c.codebook_dim;
  // frame counts down the stack
  std::vector<int> Ts(1, T);
  for (int i = c.n_ratios - 1; i >= 0; --i) Ts.push_back((Ts.back() + c.ratios[i] - 1) / c.ratios[i]);
  const int Tz = Ts.back(), Tq = (Tz + h->rs - 1) / h->rs;
  // workspace: three activation buffers of the largest layer + the transformer scratch
  size_t act = 0;
  {
    int mult = 1;
    for (int i = 0; i <= c.n_ratios; ++i) {
      act = std::max(act, (size_t)B * mult * nf * Ts[i]);
      mult *= 2;
    }
    act = std::max(act, (size_t)B * D * Tz);
  }
  const size_t tws = transformer_ws_floats(c, B, Tz);
  const size_t rvq_ws = (size_t)B * Tq * (2 * Dq + c.codebook_size) + 64;  // projected latent, frame-major residual, scores
  RUN(reserve_ws(h, 3 * act + std::max(tws, rvq_ws) + 64));
  float *a = h->ws, *b = a + act, *v = b + act, *tw = v + act;
  RUN(conv(h, "encoder.model.0.conv.conv", wav, nullptr, a, B, T, 1, 0, st));
  int idx = 1;
  for (int i = 0; i < c.n_ratios; ++i) {
    const int ratio = c.ratios[c.n_ratios - 1 - i];
    const std::string p = "encoder.model." + std::to_string(idx) + ".block.";
    RUN(resblock(h, p, a, v, b, B, Ts[i], st));
    RUN(conv(h, "encoder.model." + std::to_string(idx + 2) + ".conv.conv", b, nullptr, a, B, Ts[i], ratio, 1, st));
    idx += 3;
  }
  RUN(conv(h, "encoder.model." + std::to_string(idx + 1) + ".conv.conv", a, nullptr, b, B, Tz, 1, 1, st));
  {
    const size_t MC = (size_t)B * Tz * D, MF = (size_t)B * Tz * c.dim_feedforward;
    float *xt = tw, *q = xt + MC, *k = q + MC, *vv = k + MC, *hb = vv + MC, *op = hb + MF;
    const size_t Mc = std::min<size_t>((size_t)B * Tz, 32768), splits = (Tz + ATTN_CHUNK - 1) / ATTN_CHUNK;
    float* ml = op + Mc * D * splits;
    float* sg = ml + Mc * c.num_heads * splits * 2;
    RUN(run_transformer(h, 0, b, xt, q, k, vv, hb, op, ml, sg, 2 * (size_t)B * Tz + Mc * D, B, Tz, st));
  }
  // ConvDownsample1d: kernel 2*stride, replicate padding, no bias (modules/resample.py:14-65)
  const ConvW* dw = find_conv(h, "downsample");
  RUN(ua2_conv1d_causal_gemm_f32(b, dw->w_src, nullptr, nullptr, a, B, D, D, Tz, dw->k, h->rs, 1, 0, 1, st));
  // SplitResidualVectorQuantizer.encode (quantization/vq.py:305-315): both quantizers see the same latent
  float* xq = tw;
  for (int g = 0; g < 2; ++g) {
    Rvq& r = h->rvq[g];
    RUN(ua2_conv1d_causal_gemm_f32(a, r.in_proj, nullptr, nullptr, xq, B, D, Dq, Tq, 1, 1, 1, 0, 0, st));
    if ((size_t)B * Tq >= 128 && (Dq % 8) == 0) {
      // many frames: frame-major residual + per-quantizer tiled GEMM (scores) + argmin/update
      float* r_md = xq + (size_t)B * Dq * Tq;
      float* S = r_md + (size_t)B * Dq * Tq;
      LaunchCtx lc;
      lc.stream = (cudaStream_t)st;
      UA2_CHECK_CUDA(launch(lc, transpose_kernel, dim3((Tq + 31) / 32, (Dq + 31) / 32, B), dim3(32, 8), 0, (const float*)xq, r_md, Dq, Tq));
      RUN(ua2_rvq_encode_gemm_f32(r_md, r.emb, r.sq, S, codes, B, Dq, Tq, c.codebook_size, r.n_q, c.rvq_layers, g == 0 ? 0 : 1, st));
    } else {
      RUN(ua2_rvq_encode_f32(xq, r.emb, r.sq, codes, B, Dq, Tq, c.codebook_size, r.n_q, c.rvq_layers, g == 0 ? 0 : 1, st));
    }
  }
  return UA2_OK;
}

int ua2_codec_decode(ua2_codec* h, const int64_t* codes, int B, int Tq, float* wav, void* st) {
  UA2_REQUIRE(h && h->ready && wav && codes, "bad argument (finalize the handle first)");
  UA2_REQUIRE(B >= 1 && Tq >= 1, "bad shape");
  const ua2_codec_cfg& c = h->cfg;
  const int nf = c.n_filters, D = c.latent_dim, Dq = c.codebook_dim;
  const int Tz = Tq * h->rs;
  std::vector<int> Ts(1, Tz);
  for (int i = 0; i < c.n_ratios; ++i) Ts.push_back(Ts.back() * c.ratios[i]);
  size_t act = (size_t)B * D * Tz;
  {
    int mult = 1 << c.n_ratios;
    for (int i = 0; i <= c.n_ratios; ++i) {
      act = std::max(act, (size_t)B * mult * nf * Ts[i]);
      if (i < c.n_ratios) act = std::max(act, (size_t)B * (mult / 2) * nf * Ts[i + 1]);
      mult /= 2;
    }
  }
  const size_t tws = transformer_ws_floats(c, B, Tz);
  RUN(reserve_ws(h, 3 * act + tws + 64));
  float *a = h->ws, *b = a + act, *v = b + act, *tw = v + act;
  // SplitRVQ.decode (vq.py:317-323): first.decode + rest.decode, each = sum of codebook rows -> 1x1 output_proj
  RUN(ua2_rvq_decode_f32(codes, h->rvq[0].emb, v, B, Dq, Tq, c.codebook_size, 1, c.rvq_layers, 0, st));
  RUN(ua2_conv1d_causal_gemm_f32(v, h->rvq[0].out_proj, nullptr, nullptr, a, B, Dq, D, Tq, 1, 1, 1, 0, 0, st));
  RUN(ua2_rvq_decode_f32(codes, h->rvq[1].emb, v, B, Dq, Tq, c.codebook_size, c.rvq_layers - 1, c.rvq_layers, 1, st));
  RUN(ua2_conv1d_causal_gemm_f32(v, h->rvq[1].out_proj, nullptr, a, b, B, Dq, D, Tq, 1, 1, 1, 0, 0, st));
  RUN(ua2_convtr1d_depthwise_f32(b, h->up_w, a, B, D, Tq, h->rs, st));
  {
    const size_t MC = (size_t)B * Tz * D, MF = (size_t)B * Tz * c.dim_feedforward;
    float *xt = tw, *q = xt + MC, *k = q + MC, *vv = k + MC, *hb = vv + MC, *op = hb + MF;
    const size_t Mc = std::min<size_t>((size_t)B * Tz, 32768), splits = (Tz + ATTN_CHUNK - 1) / ATTN_CHUNK;
    float* ml = op + Mc * D * splits;
    float* sg = ml + Mc * c.num_heads * splits * 2;
    RUN(run_transformer(h, 1, a, xt, q, k, vv, hb, op, ml, sg, 2 * (size_t)B * Tz + Mc * D, B, Tz, st));
  }
  RUN(conv(h, "decoder.model.0.conv.conv", a, nullptr, b, B, Tz, 1, 0, st));
  int idx = 1;
  float *x = b, *y = a;
  for (int i = 0; i < c.n_ratios; ++i) {
    const int ratio = c.ratios[i];
    const ConvW* ct = find_conv(h, "decoder.model." + std::to_string(idx + 1) + ".convtr.convtr");
    UA2_REQUIRE(ct && ct->w, "decoder convtr weight missing");
    RUN(ua2_convtr1d_causal_gemm_f32(x, ct->w, ct->bias, y, B, ct->cin, ct->cout, Ts[i], ratio, 1, st));
    const std::string p = "decoder.model." + std::to_string(idx + 2) + ".block.";
    RUN(resblock(h, p, y, v, x, B, Ts[i + 1], st));
    idx += 3;
  }
  RUN(conv(h, "decoder.model." + std::to_string(idx + 1) + ".conv.conv", x, nullptr, wav, B, Ts.back(), 1, 1, st));
  return UA2_OK;
}

}  // extern "C"
