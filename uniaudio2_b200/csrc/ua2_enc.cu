// Whisper encoder of ReasoningCodec_film's tokenize (the first of the three SSL front-ends):
//   tools/tokenizer/ReasoningCodec_film/models/modeling_whisper.py
//     WhisperEncoder.forward       :766-867  conv1 k3 p1 + GELU, conv2 k3 s2 p1 + GELU, + embed_positions, layers, layer_norm
//     WhisperEncoderLayer.forward  :394-443  x + attn(LN(x)); x + fc2(gelu(fc1(LN(x))))     (erf GELU, LayerNorm eps 1e-5)
//     WhisperAttention.forward     :255-374  q = (x Wq^T + bq) hs^-0.5, k = x Wk^T (no bias), v = x Wv^T + bv, softmax(q k^T) v, out_proj
//   called as `self.whisper_encoder(mels).last_hidden_state` (AudioDiffusion1D.py:334-343) under torch.autocast(bfloat16)
//   (reason_tokenizer.py:114-118).
//
// Two arithmetic modes, like the flow decoder (ua2_dit.cu):
//   fp32 class (default; what the 1e-4 parity tests against the fp32 oracle run): every linear on the tcgen05 3xTF32 GEMM
//       (ua2_umma.cu), attention on the fp32 SIMT kernel of ua2_dit.cu
//   bf16 (option "bf16" = the reference's autocast arithmetic): linears on tcgen05 kind::f16 with fp32 accumulation, both
//       contractions of the attention on tcgen05 (ua2_flash.cu, head size 64), activations handed from kernel to kernel as bf16,
//       residual stream and LayerNorm statistics fp32; 9 launches per layer
// The two convolutions of the stem are GEMMs over [x[t-1] | x[t] | x[t+1]] rows (weights repacked once at finalize).
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include <cuda_bf16.h>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"
#include "ua2_umma.cuh"
#include "ua2_enc_dev.cuh"

namespace ua2 {
namespace {

// torch Conv1d weight (Cout, Cin, 3) -> (Cout, 3 * Cin) with column k * Cin + c
__global__ void enc_repack_conv3_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin) {
  const long long n = (long long)Cout * Cin * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / (3 * Cin)), r = (int)(i - (long long)co * 3 * Cin), k = r / Cin, c = r - k * Cin;
    out[i] = w[((size_t)co * Cin + c) * 3 + k];
  }
}

// stem conv1 input rows: col[(b, t)][k * C + c] = mel[b, c, t + k - 1] (zero outside), mel channels-first (B, C, T)
template <typename OUT>
__global__ void enc_im2col_cf_kernel(const float* __restrict__ mel, OUT* __restrict__ col, int B, int C, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = (long long)B * T * 3 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // thread order (b, k, c, t): reads of mel coalesced along t; the (3 C)-wide rows are short (240 columns), writes stay in L2
    const int t = (int)(i % T);
    long long r = i / T;
    const int c = (int)(r % C);
    r /= C;
    const int k = (int)(r % 3), b = (int)(r / 3);
    const int ts = t + k - 1;
    const float v = (ts >= 0 && ts < T) ? mel[((size_t)b * C + c) * T + ts] : 0.f;
    col[((size_t)b * T + t) * (3 * C) + k * C + c] = (OUT)v;
  }
}

// stem conv2 input rows (k 3, stride 2, padding 1): col[(b, t2)][k * D + c] = y[b, 2 t2 + k - 1, c], y channels-last (B, T1, D)
template <typename OUT>
__global__ void enc_im2col_s2_kernel(const float* __restrict__ y, OUT* __restrict__ col, int B, int T1, int T2, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n4 = (long long)B * T2 * 3 * (D / 4);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % (D / 4));
    long long r = i / (D / 4);
    const int k = (int)(r % 3);
    r /= 3;
    const int t2 = (int)(r % T2), b = (int)(r / T2);
    const int ts = 2 * t2 + k - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ts >= 0 && ts < T1) v = *reinterpret_cast<const float4*>(y + ((size_t)b * T1 + ts) * D + 4 * c4);
    OUT* dst = col + ((size_t)b * T2 + t2) * (3 * (size_t)D) + (size_t)k * D + 4 * c4;
    if (sizeof(OUT) == 2) {
      *reinterpret_cast<uint2*>(dst) = pack4_bf16(v.x, v.y, v.z, v.w);
    } else {
      *reinterpret_cast<float4*>(dst) = v;
    }
  }
}

struct EncLayer {
  Lin q, k, v, o, fc1, fc2;
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  float *wqkv = nullptr, *bqkv = nullptr;  // owned: [q | k | v] rows, k's bias zero
};

}  // namespace
}  // namespace ua2

using namespace ua2;

struct ua2_whisper {
  ua2_whisper_cfg cfg{};
  Lin conv1, conv2;
  const float *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
  float *conv1_w = nullptr, *conv2_w = nullptr;  // owned GEMM-form stem weights
  std::vector<EncLayer> layers;
  std::vector<void*> owned;
  bool ready = false;
  int max_batch = 0;
  float *col = nullptr, *y1 = nullptr, *h = nullptr, *n = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *ff = nullptr,
        *stats = nullptr;
  TcWorkspace tc;
  int opt_bf16 = 0;
  __nv_bfloat16* a16 = nullptr;
  std::map<const float*, __nv_bfloat16*> w16;
  int last_launches = 0;
};

namespace {

#define RUN(expr)                  \
  do {                             \
    int _rc = (expr);              \
    if (_rc != UA2_OK) return _rc; \
  } while (0)
#define CU(expr)                                                     \
  do {                                                               \
    cudaError_t _e = (expr);                                         \
    if (_e != cudaSuccess) {                                         \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); \
      return UA2_ERR_CUDA;                                           \
    }                                                                \
  } while (0)

void free_ws(ua2_whisper* h) {
  for (float** p : {&h->col, &h->y1, &h->h, &h->n, &h->q, &h->k, &h->v, &h->att, &h->ff, &h->stats, &h->tc.a, &h->tc.slots, &h->tc.c}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (h->a16) cudaFree(h->a16);
  h->a16 = nullptr;
}

int dmalloc(float** p, size_t floats) {
  UA2_CHECK_CUDA(cudaMalloc((void**)p, std::max<size_t>(floats, 4) * sizeof(float)));
  return UA2_OK;
}

int reserve(ua2_whisper* h, int B) {
  if (B <= h->max_batch) return UA2_OK;
  const ua2_whisper_cfg& c = h->cfg;
  const size_t P = c.max_source_positions, D = c.d_model, F = c.encoder_ffn_dim, C = c.num_mel_bins;
  const size_t M = (size_t)B * P, M1 = 2 * M;
  if (h->max_batch) UA2_CHECK_CUDA(cudaDeviceSynchronize());
  free_ws(h);
  const size_t kmax = std::max({3 * D, F, 3 * C});  // widest operand row
  RUN(dmalloc(&h->col, std::max(M1 * 3 * C, M * 3 * D)));
  RUN(dmalloc(&h->y1, M1 * D));
  RUN(dmalloc(&h->h, M * D));
  RUN(dmalloc(&h->n, M * D));
  RUN(dmalloc(&h->q, M * D));
  RUN(dmalloc(&h->k, M * D));
  RUN(dmalloc(&h->v, M * D));
  RUN(dmalloc(&h->att, M * D));
  RUN(dmalloc(&h->ff, M * F));
  RUN(dmalloc(&h->stats, 2 * M1 + 8));
  h->tc.a_floats = 2 * std::max(M1 * 3 * C, M * kmax);
  h->tc.slots_floats = tc_slots_max_floats();
  h->tc.c_floats = std::max(M1 * D, M * std::max(3 * D, F));
  RUN(dmalloc(&h->tc.a, h->tc.a_floats));
  RUN(dmalloc(&h->tc.slots, h->tc.slots_floats));
  RUN(dmalloc(&h->tc.c, h->tc.c_floats));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->a16, std::max(M1 * 3 * C, M * kmax) * sizeof(__nv_bfloat16)));
  h->max_batch = B;
  return UA2_OK;
}

// *fresh: the copy was made by a kernel just launched on this stream.  The GEMM's weight producer starts its TMA loads BEFORE
// griddepcontrol.wait (weights are normally static), so the launch that follows must not be a programmatic dependent launch.
int weight16(ua2_whisper* h, const LaunchCtx& lc, const float* W, int N, int K, const __nv_bfloat16** out, bool* fresh) {
  auto it = h->w16.find(W);
  *fresh = it == h->w16.end();
  if (it == h->w16.end()) {
    __nv_bfloat16* wb = nullptr;
    UA2_CHECK_CUDA(cudaMalloc((void**)&wb, (size_t)N * K * sizeof(__nv_bfloat16)));
    it = h->w16.emplace(W, wb).first;
    const long long n4 = (long long)N * K / 4;
    CU(launch(lc, enc_to_bf16_kernel, dim3(grid_for(n4)), dim3(256), 0, W, wb, n4));
  }
  *out = it->second;
  return UA2_OK;
}

// x @ W^T, raw (no bias) into the handle's product buffer; the epilogue description comes back filled in.
//   bf16 mode: operand rows are already bf16 in h->a16; the stream-K side slots are left for the epilogue to sum
//   fp32 class: operand rows fp32 at x32; launch_gemv's tensor-core path (activation split + 3xTF32 GEMM + slot fix-up)
int enc_linear(ua2_whisper* h, const LaunchCtx& lc, const float* x32, const float* W, const float* bias, int M, int N, int K, int T, EncEpi* e) {
  *e = EncEpi{};
  e->bias = bias;
  e->M = M;
  e->N = N;
  e->T = T;
  e->eps = 1e-5f;
  if (h->opt_bf16) {
    const __nv_bfloat16* w16 = nullptr;
    bool fresh = false;
    RUN(weight16(h, lc, W, N, K, &w16, &fresh));
    const UmmaPlan pl = umma_plan(M, N, 1, K, true);
    UA2_REQUIRE(pl.slot_floats <= h->tc.slots_floats && (size_t)M * N <= h->tc.c_floats && (K % 8) == 0 && (N % 4) == 0,
                "encoder linear outside the tensor-core path's shapes");
    LaunchCtx lg = lc;
    if (fresh) lg.pdl = false;  // full stream order behind the conversion kernel (see weight16)
    CU(run_umma_bf16(lg, h->a16, w16, h->tc.c, N, h->tc.slots, M, N, K, pl));
    e->src = h->tc.c;
    e->slots = h->tc.slots;
    e->pl = pl;
    e->has_split = umma_has_split_tiles(pl) ? 1 : 0;
    return UA2_OK;
  }
  GemvParams p;
  p.W = W;
  p.N = N;
  p.K = K;
  p.M = M;
  p.X = x32;
  p.ldx = K;
  p.Y = h->tc.c;  // a many-row call is served by the tensor-core path and leaves the product in its own buffer (raw_out)
  p.ldy = N;
  p.ws = h->stats;
  p.ws_floats = 2 * (size_t)h->max_batch * h->cfg.max_source_positions * 2 + 8;
  p.tc = &h->tc;
  const float* raw = nullptr;
  p.raw_out = &raw;
  CU(launch_gemv(lc, PRO_PLAIN, EPI_STORE, p));
  e->src = raw ? raw : h->tc.c;
  return UA2_OK;
}

int enc_forward(ua2_whisper* h, const LaunchCtx& lc, const float* mel, float* out, int B) {
  const ua2_whisper_cfg& c = h->cfg;
  const int P = c.max_source_positions, D = c.d_model, F = c.encoder_ffn_dim, C = c.num_mel_bins, H = c.encoder_attention_heads, hs = D / H;
  const int T1 = 2 * P, M1 = B * T1, M = B * P;
  const bool bf = h->opt_bf16 != 0;
  const unsigned ln_threads = (unsigned)(((D / 4) + 31) / 32 * 32);
  EncEpi e;
  // ---- stem: conv1 (k 3, p 1) + GELU -> y1 (B, 2P, D) fp32
  if (bf) {
    CU(launch(lc, enc_im2col_cf_kernel<__nv_bfloat16>, dim3(grid_for((long long)M1 * 3 * C)), dim3(256), 0, mel, h->a16, B, C, T1));
  } else {
    CU(launch(lc, enc_im2col_cf_kernel<float>, dim3(grid_for((long long)M1 * 3 * C)), dim3(256), 0, mel, h->col, B, C, T1));
  }
  RUN(enc_linear(h, lc, h->col, h->conv1_w, h->conv1.b, M1, D, 3 * C, T1, &e));
  e.y32 = h->y1;
  CU(launch_enc_epi<EE_GELU>(lc, e));
  // ---- conv2 (k 3, s 2, p 1) + GELU + embed_positions -> h (B, P, D) fp32, the residual stream
  if (bf) {
    CU(launch(lc, enc_im2col_s2_kernel<__nv_bfloat16>, dim3(grid_for((long long)M * 3 * (D / 4))), dim3(256), 0, (const float*)h->y1, h->a16, B, T1, P, D));
  } else {
    CU(launch(lc, enc_im2col_s2_kernel<float>, dim3(grid_for((long long)M * 3 * (D / 4))), dim3(256), 0, (const float*)h->y1, h->col, B, T1, P, D));
  }
  RUN(enc_linear(h, lc, h->col, h->conv2_w, h->conv2.b, M, D, 3 * D, P, &e));
  e.y32 = h->h;
  e.pos = h->pos;
  CU(launch_enc_epi<EE_GELU_POS>(lc, e));
  // ---- layers
  auto layer_norm = [&](const float* g, const float* b) -> cudaError_t {  // LN(h) -> bf16 operand rows or h->n
    EncEpi l{};
    l.M = M;
    l.N = D;
    l.res = h->h;
    l.ln_g = g;
    l.ln_b = b;
    l.eps = 1e-5f;
    l.y16 = bf ? h->a16 : nullptr;
    l.y32 = h->n;
    return launch(lc, enc_res_ln_kernel<false>, dim3(M), dim3(ln_threads), 0, l);
  };
  for (size_t li = 0; li < h->layers.size(); ++li) {
    const EncLayer& L = h->layers[li];
    if (li == 0) CU(layer_norm(L.ln1_g, L.ln1_b));
    // q / k / v in one GEMM (k's bias rows are zero)
    RUN(enc_linear(h, lc, h->n, L.wqkv, L.bqkv, M, 3 * D, D, P, &e));
    e.H = H;
    e.hs = hs;
    if (bf) {
      e.q16 = reinterpret_cast<__nv_bfloat16*>(h->q);
      e.k16 = reinterpret_cast<__nv_bfloat16*>(h->k);
      e.v16 = reinterpret_cast<__nv_bfloat16*>(h->v);
      CU(launch_enc_epi<EE_QKV16>(lc, e));
      CU(launch_flash_bf16(lc, e.q16, e.k16, e.v16, nullptr, h->a16, B, P, H, hs));
    } else {
      e.q = h->q;
      e.k = h->k;
      e.v = h->v;
      CU(launch_enc_epi<EE_QKV>(lc, e));
      CU(launch_dense_attn_f32(lc, h->q, h->k, h->v, h->att, B, P, H, hs));
    }
    // out_proj + residual, then final_layer_norm of the updated row in the same kernel
    RUN(enc_linear(h, lc, h->att, L.o.w, L.o.b, M, D, D, P, &e));
    e.res = h->h;
    e.ln_g = L.ln2_g;
    e.ln_b = L.ln2_b;
    e.y16 = bf ? h->a16 : nullptr;
    e.y32 = h->n;
    CU(launch(lc, enc_res_ln_kernel<true>, dim3(M), dim3(ln_threads), 0, e));
    // fc1 + GELU
    RUN(enc_linear(h, lc, h->n, L.fc1.w, L.fc1.b, M, F, D, P, &e));
    e.y16 = bf ? h->a16 : nullptr;
    e.y32 = h->ff;
    CU(launch_enc_epi<EE_GELU>(lc, e));
    // fc2 + residual, then the NEXT LayerNorm (next layer's self_attn_layer_norm, or the encoder's layer_norm into `out`)
    RUN(enc_linear(h, lc, h->ff, L.fc2.w, L.fc2.b, M, D, F, P, &e));
    e.res = h->h;
    const bool last = li + 1 == h->layers.size();
    e.ln_g = last ? h->lnf_g : h->layers[li + 1].ln1_g;
    e.ln_b = last ? h->lnf_b : h->layers[li + 1].ln1_b;
    e.y16 = (bf && !last) ? h->a16 : nullptr;
    e.y32 = last ? out : h->n;
    CU(launch(lc, enc_res_ln_kernel<true>, dim3(M), dim3(ln_threads), 0, e));
  }
  return UA2_OK;
}

bool parse_layer_key(const std::string& key, int& idx, std::string& rest) {
  const std::string pre = "layers.";
  if (key.compare(0, pre.size(), pre) != 0) return false;
  size_t p = pre.size(), q = p;
  while (q < key.size() && key[q] >= '0' && key[q] <= '9') ++q;
  if (q == p || q >= key.size() || key[q] != '.') return false;
  idx = std::stoi(key.substr(p, q - p));
  rest = key.substr(q + 1);
  return true;
}

}  // namespace

extern "C" {

int ua2_whisper_create(const ua2_whisper_cfg* cfg, ua2_whisper** out) {
  UA2_REQUIRE(cfg && out, "null argument");
  const ua2_whisper_cfg& c = *cfg;
  UA2_REQUIRE(c.d_model >= 32 && c.encoder_attention_heads >= 1 && c.encoder_layers >= 1 && c.encoder_ffn_dim >= 8 && c.max_source_positions >= 16 &&
                  c.num_mel_bins >= 1,
              "bad dimensions");
  UA2_REQUIRE(c.d_model % c.encoder_attention_heads == 0, "d_model must be divisible by encoder_attention_heads");
  const int hs = c.d_model / c.encoder_attention_heads;
  UA2_REQUIRE(hs == 32 || hs == 64 || hs == 128, "head size must be 32 / 64 / 128");
  UA2_REQUIRE(c.d_model % 8 == 0 && c.d_model <= 4096 && c.encoder_ffn_dim % 8 == 0 && (3 * c.num_mel_bins) % 8 == 0,
              "d_model (<= 4096), encoder_ffn_dim and 3 * num_mel_bins must be multiples of 8");
  ua2_whisper* h = new ua2_whisper();
  h->cfg = c;
  h->layers.resize(c.encoder_layers);
  *out = h;
  return UA2_OK;
}

int ua2_whisper_destroy(ua2_whisper* h) {
  if (!h) return UA2_OK;
  cudaDeviceSynchronize();
  free_ws(h);
  for (void* p : h->owned) cudaFree(p);
  for (auto& kv : h->w16) cudaFree(kv.second);
  delete h;
  return UA2_OK;
}

int ua2_whisper_load_weight(ua2_whisper* h, const char* key_c, const float* dptr, const int64_t* shape, int ndim) {
  UA2_REQUIRE(h && key_c && dptr && shape && ndim >= 1, "null argument");
  const std::string key(key_c);
  const ua2_whisper_cfg& c = h->cfg;
  const int64_t D = c.d_model, F = c.encoder_ffn_dim, C = c.num_mel_bins, P = c.max_source_positions;
  auto is = [&](std::initializer_list<int64_t> want) {
    if ((int)want.size() != ndim) return false;
    int i = 0;
    for (int64_t w : want)
      if (shape[i++] != w) return false;
    return true;
  };
  struct Ent {
    const char* name;
    const float** dst;
    std::initializer_list<int64_t> shp;
  };
  const Ent tops[] = {{"conv1.weight", &h->conv1.w, {D, C, 3}}, {"conv1.bias", &h->conv1.b, {D}},     {"conv2.weight", &h->conv2.w, {D, D, 3}},
                      {"conv2.bias", &h->conv2.b, {D}},         {"embed_positions.weight", &h->pos, {P, D}}, {"layer_norm.weight", &h->lnf_g, {D}},
                      {"layer_norm.bias", &h->lnf_b, {D}}};
  for (const Ent& t : tops)
    if (key == t.name) {
      UA2_REQUIRE(is(t.shp), key + ": shape mismatch");
      *t.dst = dptr;
      return UA2_OK;
    }
  int li = -1;
  std::string rest;
  UA2_REQUIRE(parse_layer_key(key, li, rest) && li >= 0 && li < c.encoder_layers, "unexpected key " + key);
  EncLayer& L = h->layers[li];
  const Ent ents[] = {{"self_attn.q_proj.weight", &L.q.w, {D, D}},     {"self_attn.q_proj.bias", &L.q.b, {D}},
                      {"self_attn.k_proj.weight", &L.k.w, {D, D}},     {"self_attn.v_proj.weight", &L.v.w, {D, D}},
                      {"self_attn.v_proj.bias", &L.v.b, {D}},          {"self_attn.out_proj.weight", &L.o.w, {D, D}},
                      {"self_attn.out_proj.bias", &L.o.b, {D}},        {"self_attn_layer_norm.weight", &L.ln1_g, {D}},
                      {"self_attn_layer_norm.bias", &L.ln1_b, {D}},    {"fc1.weight", &L.fc1.w, {F, D}},
                      {"fc1.bias", &L.fc1.b, {F}},                     {"fc2.weight", &L.fc2.w, {D, F}},
                      {"fc2.bias", &L.fc2.b, {D}},                     {"final_layer_norm.weight", &L.ln2_g, {D}},
                      {"final_layer_norm.bias", &L.ln2_b, {D}}};
  for (const Ent& t : ents)
    if (rest == t.name) {
      UA2_REQUIRE(is(t.shp), key + ": shape mismatch");
      *t.dst = dptr;
      return UA2_OK;
    }
  UA2_REQUIRE(false, "unexpected key " + key);
}

int ua2_whisper_finalize(ua2_whisper* h, void* stream) {
  UA2_REQUIRE(h, "null handle");
  const ua2_whisper_cfg& c = h->cfg;
  const size_t D = c.d_model, C = c.num_mel_bins;
  UA2_REQUIRE(h->conv1.w && h->conv1.b && h->conv2.w && h->conv2.b && h->pos && h->lnf_g && h->lnf_b,
              "missing top-level parameters (conv1.*, conv2.*, embed_positions.weight, layer_norm.*)");
  for (int i = 0; i < c.encoder_layers; ++i) {
    const EncLayer& L = h->layers[i];
    UA2_REQUIRE(L.q.w && L.q.b && L.k.w && L.v.w && L.v.b && L.o.w && L.o.b && L.ln1_g && L.ln1_b && L.fc1.w && L.fc1.b && L.fc2.w && L.fc2.b &&
                    L.ln2_g && L.ln2_b,
                "missing parameters of layers." + std::to_string(i));
  }
  cudaStream_t st = (cudaStream_t)stream;
  LaunchCtx lc;
  lc.stream = st;
  auto own = [&](float** p, size_t floats) -> int {
    UA2_CHECK_CUDA(cudaMalloc((void**)p, floats * sizeof(float)));
    h->owned.push_back(*p);
    return UA2_OK;
  };
  if (!h->ready) {
    RUN(own(&h->conv1_w, D * 3 * C));
    RUN(own(&h->conv2_w, D * 3 * D));
    for (EncLayer& L : h->layers) {
      RUN(own(&L.wqkv, 3 * D * D));
      RUN(own(&L.bqkv, 3 * D));
    }
  }
  CU(launch(lc, enc_repack_conv3_kernel, dim3(grid_for((long long)D * 3 * C)), dim3(256), 0, h->conv1.w, h->conv1_w, (int)D, (int)C));
  CU(launch(lc, enc_repack_conv3_kernel, dim3(grid_for((long long)D * 3 * D)), dim3(256), 0, h->conv2.w, h->conv2_w, (int)D, (int)D));
  for (EncLayer& L : h->layers) {
    const size_t wb = D * D * sizeof(float), bb = D * sizeof(float);
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.wqkv, L.q.w, wb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.wqkv + D * D, L.k.w, wb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.wqkv + 2 * D * D, L.v.w, wb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.bqkv, L.q.b, bb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemsetAsync(L.bqkv + D, 0, bb, st));  // k_proj has no bias (:240)
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.bqkv + 2 * D, L.v.b, bb, cudaMemcpyDeviceToDevice, st));
  }
  for (auto& kv : h->w16) cudaFree(kv.second);  // weights may have changed: bf16 copies are rebuilt at next use
  h->w16.clear();
  h->ready = true;
  return UA2_OK;
}

int ua2_whisper_forward(ua2_whisper* h, const float* input_features, float* out, int B, void* stream) {
  UA2_REQUIRE(h && input_features && out, "null argument");
  UA2_REQUIRE(h->ready, "ua2_whisper_finalize has not run");
  UA2_REQUIRE(B >= 1 && B <= 4096, "batch out of range");
  UA2_REQUIRE((long long)B * h->cfg.max_source_positions >= 32, "fewer than 32 rows");
  UA2_REQUIRE(!h->opt_bf16 || h->cfg.d_model / h->cfg.encoder_attention_heads == 64, "bf16 mode serves head size 64");
  RUN(reserve(h, B));
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.pdl = true;
  int launches = 0;
  lc.launch_counter = &launches;
  const int rc = enc_forward(h, lc, input_features, out, B);
  h->last_launches = launches;
  return rc;
}

int ua2_whisper_set_option(ua2_whisper* h, const char* name, int value) {
  UA2_REQUIRE(h && name, "null argument");
  const std::string n(name);
  if (n == "bf16") {
    h->opt_bf16 = value ? 1 : 0;
    return UA2_OK;
  }
  UA2_REQUIRE(false, "unknown option " + n);
}

int ua2_whisper_last_launch_count(ua2_whisper* h) { return h ? h->last_launches : 0; }

}  // extern "C"
