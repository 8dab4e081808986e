// Reasoning encoder (`AudioThinking`) of ReasoningCodec_film's tokenize direction, up to the query tokens
// that go into reasoning_vq:
//   tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py
//     :169-188  AudioThinking: cls_token, 5 x TransformerBlock(dim 768, dim_heads 128, power_normalized, layer_scale, add_rope, qk_norm,
//               ff mult 4), semantic_merge_proj Linear(whisper_dim + 1024 -> dim), down_sampling_layer_whisper Conv1d(k 2, s 2)
//     :372-390  encode_reasoning_part; :458-486 set_masking (a query token after every 5 frames) / extract_mask_positions
//   tools/tokenizer/ReasoningCodec_film/modules/transformer.py
//     :645-783  TransformerBlock (power_normalized removes the pre-norms): x += scale1 * to_out(attn(to_qkv(x))); x += scale2 * ff(x)
//     :293-598  Attention: fused to_qkv (no bias), LayerNorm(128) of q and k per head, rotary embedding on the first 64 head dims
//               (pairs (i, i + 32), angle pos * inv_freq[i]), softmax(q k^T / sqrt(128)) v, to_out (no bias)
//     :206-291  FeedForward: GLU(Linear 768 -> 2 x 3072, x * sigmoid(gate)), Linear 3072 -> 768; :197-202 LayerScale
// Weight-normed linears arrive with the normalisation applied (the Python drop-in folds g * v / ||v||).  fp32 class: every linear and the
// down-sampling convolution as GEMMs on the tcgen05 3xTF32 kernel (ua2_umma.cu), attention on the fp32 kernel of ua2_dit.cu (head size
// 128).  (On a GPU the reference runs this attention through flash-attn in fp16; the oracle and this path keep fp32.)
//
//   th_im2col_cf_kernel    rows [x[b, :, 2 t] | x[b, :, 2 t + 1]] of the channels-first Whisper features (reads coalesced along t)
//   th_transpose_kernel    BEST-RQ features (B, C, T) -> columns [off, off + C) of the concatenated rows
//   th_epi_kernel<MODE>    bias into a column block of wider rows; bias + scatter to the rows that set_masking leaves for frames; GLU;
//                          LayerScale residual add
//   th_cls_kernel          the query token rows
//   th_qkv_kernel          one warp per (row, head): LayerNorm of q and k, rotary embedding, q (M, D), k / v (B, H, N, 128)
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"
#include "ua2_umma.cuh"

namespace ua2 {
namespace {

unsigned th_grid(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32); }

// col[(b, t)][j * C + c] = x[b, c, s * t + j], t < Tout; x channels-first (B, C, Tin)
__global__ void th_im2col_cf_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int C, int Tin, int Tout, int k, int s) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = (long long)B * k * C * Tout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % Tout);
    long long r = i / Tout;
    const int c = (int)(r % C);
    r /= C;
    const int j = (int)(r % k), b = (int)(r / k);
    col[((size_t)b * Tout + t) * ((size_t)k * C) + (size_t)j * C + c] = x[((size_t)b * C + c) * Tin + (size_t)s * t + j];
  }
}

// out[(b, t)][off + c] = x[b, c, t], t < T; rows of `ld` floats
__global__ void th_transpose_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C, int Tin, int T, int ld, int off) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = (long long)B * C * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const long long r = i / T;
    const int c = (int)(r % C), b = (int)(r / C);
    out[((size_t)b * T + t) * ld + off + c] = x[((size_t)b * C + c) * Tin + t];
  }
}

// torch Conv1d weight (Cout, Cin, k) -> GEMM weight (Cout, k * Cin), column j * Cin + c
__global__ void th_repack_conv_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int k) {
  const long long n = (long long)Cout * Cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int j = (int)((i / Cin) % k);
    const long long co = i / ((long long)Cin * k);
    out[i] = w[(co * Cin + c) * k + j];
  }
}

enum ThMode : int {
  TH_BIAS_LD = 0,    // out[m * ld + off + c] = v + bias[c]                                   down-sampled Whisper half of the concatenation
  TH_BIAS_ROWS = 1,  // out[(b Tn + t + t / interval) * N + c] = v + bias[c], m = b T + t     semantic_merge_proj into the set_masking layout
  TH_GLU = 2,        // out[m * F + c] = (v[m, c] + b[c]) * sigmoid(v[m, F + c] + b[F + c])   GLU, N = 2 F
  TH_SCALE_RES = 3   // out[m * N + c] += scale[c] * (v + bias[c])  (bias may be NULL)        LayerScale + residual
};

struct ThEpi {
  const float* src;  // (M, N) raw product
  const float* bias;
  const float* scale;
  float* out;
  int M, N, ld, off, T, Tn, interval;
};

// one row per blockIdx.x, 4 columns per thread
template <int MODE>
__global__ void __launch_bounds__(256) th_epi_kernel(const ThEpi e) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, c = (blockIdx.y * 256 + threadIdx.x) * 4;
  const int width = MODE == TH_GLU ? e.N / 2 : e.N;
  if (c >= width) return;
  float4 v = *reinterpret_cast<const float4*>(e.src + (size_t)m * e.N + c);
  if (e.bias != nullptr) {
    const float4 b = *reinterpret_cast<const float4*>(e.bias + c);
    v.x += b.x;
    v.y += b.y;
    v.z += b.z;
    v.w += b.w;
  }
  if (MODE == TH_BIAS_LD) {
    *reinterpret_cast<float4*>(e.out + (size_t)m * e.ld + e.off + c) = v;
  } else if (MODE == TH_BIAS_ROWS) {
    const int b = m / e.T, t = m - b * e.T;
    *reinterpret_cast<float4*>(e.out + ((size_t)b * e.Tn + t + t / e.interval) * e.N + c) = v;
  } else if (MODE == TH_GLU) {
    float4 g = *reinterpret_cast<const float4*>(e.src + (size_t)m * e.N + width + c);
    const float4 gb = *reinterpret_cast<const float4*>(e.bias + width + c);
    g.x = 1.f / (1.f + expf(-(g.x + gb.x)));
    g.y = 1.f / (1.f + expf(-(g.y + gb.y)));
    g.z = 1.f / (1.f + expf(-(g.z + gb.z)));
    g.w = 1.f / (1.f + expf(-(g.w + gb.w)));
    *reinterpret_cast<float4*>(e.out + (size_t)m * width + c) = make_float4(v.x * g.x, v.y * g.y, v.z * g.z, v.w * g.w);
  } else {
    const float4 s = *reinterpret_cast<const float4*>(e.scale + c);
    float4* hp = reinterpret_cast<float4*>(e.out + (size_t)m * e.N + c);
    float4 h = *hp;
    h.x = __fadd_rn(h.x, __fmul_rn(v.x, s.x));  // x + out * scale: no contraction, like the reference's two ops
    h.y = __fadd_rn(h.y, __fmul_rn(v.y, s.y));
    h.z = __fadd_rn(h.z, __fmul_rn(v.z, s.z));
    h.w = __fadd_rn(h.w, __fmul_rn(v.w, s.w));
    *hp = h;
  }
}

template <int MODE>
cudaError_t launch_th_epi(const LaunchCtx& lc, const ThEpi& e) {
  const int width = MODE == TH_GLU ? e.N / 2 : e.N;
  return launch(lc, th_epi_kernel<MODE>, dim3(e.M, (width + 1023) / 1024), dim3(256), 0, e);
}

// h[(b Tn + (g + 1) (interval + 1) - 1)] = cls  for g < T / interval
__global__ void th_cls_kernel(const float* __restrict__ cls, float* __restrict__ h, int B, int n_tok, int Tn, int interval, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = (long long)B * n_tok * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const long long r = i / D;
    const int g = (int)(r % n_tok), b = (int)(r / n_tok);
    h[((size_t)b * Tn + (size_t)(g + 1) * (interval + 1) - 1) * D + c] = cls[c];
  }
}

// one warp per (row, head), head size 128: lane l holds dims l, l + 32, l + 64, l + 96; the rotary pairs (i, i + 32), i < 32, are the
// lane's first two values.  src (M, 3 D) = [q | k | v]; q -> (M, D), k / v -> (B, H, N, 128).
__global__ void __launch_bounds__(256) th_qkv_kernel(const float* __restrict__ src, const float* __restrict__ qg, const float* __restrict__ qb,
                                                     const float* __restrict__ kg, const float* __restrict__ kb, const float* __restrict__ inv_freq,
                                                     float* __restrict__ q, float* __restrict__ k, float* __restrict__ v, int M, int N, int H) {
  constexpr int HD = 128;
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * 8 + warp;
  if (idx >= (long long)M * H) return;  // whole warps leave together
  const int m = (int)(idx / H), hh = (int)(idx - (long long)m * H);
  const int b = m / N, n = m - b * N, D = H * HD;
  const float ang = __fmul_rn((float)n, inv_freq[lane]);
  const float cs = cosf(ang), sn = sinf(ang);
  const size_t hd_off = (((size_t)b * H + hh) * N + n) * HD;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const float* row = src + (size_t)m * 3 * D + (size_t)which * D + (size_t)hh * HD;
    const float* g = which ? kg : qg;
    const float* be = which ? kb : qb;
    float x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = row[lane + 32 * e];
    const float mean = warp_sum((x[0] + x[1]) + (x[2] + x[3])) * (1.f / HD);
    float d[4], sq = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      d[e] = x[e] - mean;
      sq = fmaf(d[e], d[e], sq);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / HD) + 1e-5f);
    float y[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) y[e] = fmaf(d[e] * rstd, g[lane + 32 * e], be[lane + 32 * e]);
    // t * cos + rotate_half(t) * sin on dims [0, 64): (y0, y1) -> (y0 c - y1 s, y1 c + y0 s)
    const float r0 = __fadd_rn(__fmul_rn(y[0], cs), __fmul_rn(-y[1], sn));
    const float r1 = __fadd_rn(__fmul_rn(y[1], cs), __fmul_rn(y[0], sn));
    y[0] = r0;
    y[1] = r1;
    float* dst = which ? k + hd_off : q + (size_t)m * D + (size_t)hh * HD;
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[lane + 32 * e] = y[e];
  }
  const float* vr = src + (size_t)m * 3 * D + 2 * (size_t)D + (size_t)hh * HD;
#pragma unroll
  for (int e = 0; e < 4; ++e) v[hd_off + lane + 32 * e] = vr[lane + 32 * e];
}

cudaError_t launch_th_qkv(const LaunchCtx& lc, const float* src, const float* qg, const float* qb, const float* kg, const float* kb,
                          const float* inv_freq, float* q, float* k, float* v, int B, int N, int H) {
  const long long items = (long long)B * N * H;
  return launch(lc, th_qkv_kernel, dim3((unsigned)((items + 7) / 8)), dim3(256), 0, src, qg, qb, kg, kb, inv_freq, q, k, v, B * N, N, H);
}

}  // namespace
}  // namespace ua2

using namespace ua2;

namespace {

struct ThLin {
  const float* w = nullptr;
  const float* b = nullptr;
};
struct ThLayer {
  ThLin qkv, out, ff1, ff2;
  const float *qg = nullptr, *qb = nullptr, *kg = nullptr, *kb = nullptr, *s1 = nullptr, *s2 = nullptr, *inv_freq = nullptr;
};

}  // namespace

struct ua2_thinking {
  ua2_thinking_cfg cfg{};
  const float* cls = nullptr;
  ThLin ds, merge;
  float* ds_wr = nullptr;  // owned GEMM-form weight of the down-sampling convolution
  std::vector<ThLayer> layers;
  std::vector<void*> owned;
  bool ready = false;
  int max_B = 0, max_T = 0;
  float *col = nullptr, *cat = nullptr, *h = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *ff = nullptr, *stats = nullptr;
  size_t stats_floats = 0;
  TcWorkspace tc;
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

void th_free_ws(ua2_thinking* h) {
  for (float** p : {&h->col, &h->cat, &h->h, &h->q, &h->k, &h->v, &h->att, &h->ff, &h->stats, &h->tc.a, &h->tc.slots, &h->tc.c}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
}

int th_dmalloc(float** p, size_t floats) {
  UA2_CHECK_CUDA(cudaMalloc((void**)p, std::max<size_t>(floats, 4) * sizeof(float)));
  return UA2_OK;
}

int th_reserve(ua2_thinking* h, int B, int T) {
  if (B <= h->max_B && T <= h->max_T) return UA2_OK;
  B = std::max(B, h->max_B);
  T = std::max(T, h->max_T);
  const ua2_thinking_cfg& c = h->cfg;
  const size_t D = c.dim, F = (size_t)c.dim * c.ff_mult, Cw = c.whisper_dim, Cm = c.mu_dim;
  const size_t M = (size_t)B * T, Mn = (size_t)B * (T + T / c.interval);
  if (h->max_B) UA2_CHECK_CUDA(cudaDeviceSynchronize());
  th_free_ws(h);
  RUN(th_dmalloc(&h->col, M * 2 * Cw));
  RUN(th_dmalloc(&h->cat, M * (Cw + Cm)));
  RUN(th_dmalloc(&h->h, Mn * D));
  RUN(th_dmalloc(&h->q, Mn * D));
  RUN(th_dmalloc(&h->k, Mn * D));
  RUN(th_dmalloc(&h->v, Mn * D));
  RUN(th_dmalloc(&h->att, Mn * D));
  RUN(th_dmalloc(&h->ff, Mn * F));
  h->stats_floats = 2 * std::max(M, Mn) + 16;
  RUN(th_dmalloc(&h->stats, h->stats_floats));
  h->tc.a_floats = 2 * std::max({M * 2 * Cw, M * (Cw + Cm), Mn * F, Mn * D});
  h->tc.slots_floats = tc_slots_max_floats();
  h->tc.c_floats = std::max({M * Cw, M * D, Mn * 3 * D, Mn * 2 * F});
  RUN(th_dmalloc(&h->tc.a, h->tc.a_floats));
  RUN(th_dmalloc(&h->tc.slots, h->tc.slots_floats));
  RUN(th_dmalloc(&h->tc.c, h->tc.c_floats));
  h->max_B = B;
  h->max_T = T;
  return UA2_OK;
}

// x (M, K) @ W (N, K)^T -> raw product (tensor-core workspace for many rows); *src points at it
int th_linear(ua2_thinking* h, const LaunchCtx& lc, const float* x, const float* W, int M, int N, int K, const float** src) {
  GemvParams p;
  p.W = W;
  p.N = N;
  p.K = K;
  p.M = M;
  p.X = x;
  p.ldx = K;
  p.Y = h->tc.c;
  p.ldy = N;
  p.ws = h->stats;
  p.ws_floats = h->stats_floats;
  p.tc = &h->tc;
  const float* raw = nullptr;
  p.raw_out = &raw;
  CU(launch_gemv(lc, PRO_PLAIN, EPI_STORE, p));
  *src = raw ? raw : h->tc.c;
  return UA2_OK;
}

int th_forward(ua2_thinking* h, const LaunchCtx& lc, const float* whisper, const float* mu, int B, int Tw, int Tb, float* out) {
  const ua2_thinking_cfg& c = h->cfg;
  const int D = c.dim, H = D / c.dim_heads, F = D * c.ff_mult, Cw = c.whisper_dim, Cm = c.mu_dim, iv = c.interval;
  const int T = std::min(Tw / 2, Tb);
  UA2_REQUIRE(T >= iv && T % iv == 0, "min(Tw / 2, Tb) must be a positive multiple of the interval (the reference's set_masking reshapes by it)");
  const int n_tok = T / iv, Tn = T + n_tok, M = B * T, Mn = B * Tn;
  RUN(th_reserve(h, B, T));
  const float* src = nullptr;
  ThEpi e{};
  // ---- down_sampling_layer_whisper (k 2, s 2) as a GEMM -> columns [0, Cw) of the concatenated rows; BEST-RQ features -> [Cw, Cw + Cm)
  CU(launch(lc, th_im2col_cf_kernel, dim3(th_grid((long long)M * 2 * Cw)), dim3(256), 0, whisper, h->col, B, Cw, Tw, T, 2, 2));
  RUN(th_linear(h, lc, h->col, h->ds_wr, M, Cw, 2 * Cw, &src));
  e.src = src;
  e.bias = h->ds.b;
  e.out = h->cat;
  e.M = M;
  e.N = Cw;
  e.ld = Cw + Cm;
  e.off = 0;
  CU(launch_th_epi<TH_BIAS_LD>(lc, e));
  CU(launch(lc, th_transpose_kernel, dim3(th_grid((long long)M * Cm)), dim3(256), 0, mu, h->cat, B, Cm, Tb, T, Cw + Cm, Cw));
  // ---- semantic_merge_proj into the rows that set_masking leaves for frames; query tokens into the others
  RUN(th_linear(h, lc, h->cat, h->merge.w, M, D, Cw + Cm, &src));
  e = ThEpi{};
  e.src = src;
  e.bias = h->merge.b;
  e.out = h->h;
  e.M = M;
  e.N = D;
  e.T = T;
  e.Tn = Tn;
  e.interval = iv;
  CU(launch_th_epi<TH_BIAS_ROWS>(lc, e));
  CU(launch(lc, th_cls_kernel, dim3(th_grid((long long)B * n_tok * D)), dim3(256), 0, h->cls, h->h, B, n_tok, Tn, iv, D));
  // ---- blocks
  for (const ThLayer& L : h->layers) {
    RUN(th_linear(h, lc, h->h, L.qkv.w, Mn, 3 * D, D, &src));
    CU(launch_th_qkv(lc, src, L.qg, L.qb, L.kg, L.kb, L.inv_freq, h->q, h->k, h->v, B, Tn, H));
    CU(launch_dense_attn_f32(lc, h->q, h->k, h->v, h->att, B, Tn, H, c.dim_heads));
    RUN(th_linear(h, lc, h->att, L.out.w, Mn, D, D, &src));
    e = ThEpi{};
    e.src = src;
    e.scale = L.s1;
    e.out = h->h;
    e.M = Mn;
    e.N = D;
    CU(launch_th_epi<TH_SCALE_RES>(lc, e));
    RUN(th_linear(h, lc, h->h, L.ff1.w, Mn, 2 * F, D, &src));
    e = ThEpi{};
    e.src = src;
    e.bias = L.ff1.b;
    e.out = h->ff;
    e.M = Mn;
    e.N = 2 * F;
    CU(launch_th_epi<TH_GLU>(lc, e));
    RUN(th_linear(h, lc, h->ff, L.ff2.w, Mn, D, F, &src));
    e = ThEpi{};
    e.src = src;
    e.bias = L.ff2.b;
    e.scale = L.s2;
    e.out = h->h;
    e.M = Mn;
    e.N = D;
    CU(launch_th_epi<TH_SCALE_RES>(lc, e));
  }
  UA2_CHECK_CUDA(cudaMemcpyAsync(out, h->h, (size_t)Mn * D * sizeof(float), cudaMemcpyDeviceToDevice, lc.stream));
  return UA2_OK;
}

bool th_parse_index(const std::string& key, const std::string& pre, int& idx, std::string& rest) {
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

int ua2_thinking_create(const ua2_thinking_cfg* cfg, ua2_thinking** out) {
  UA2_REQUIRE(cfg && out, "null argument");
  const ua2_thinking_cfg& c = *cfg;
  UA2_REQUIRE(c.dim_heads == 128, "dim_heads must be 128 (AudioDiffusion1D.py:178)");
  UA2_REQUIRE(c.dim >= 128 && c.dim % 128 == 0 && c.dim <= 4096, "dim must be a multiple of dim_heads (<= 4096)");
  UA2_REQUIRE(c.depth >= 1 && c.interval >= 1 && c.ff_mult >= 1, "depth, interval and ff_mult must be positive");
  UA2_REQUIRE(c.whisper_dim >= 4 && c.whisper_dim % 4 == 0 && c.mu_dim >= 4 && c.mu_dim % 4 == 0, "whisper_dim and mu_dim must be multiples of 4");
  ua2_thinking* h = new ua2_thinking();
  h->cfg = c;
  h->layers.resize(c.depth);
  *out = h;
  return UA2_OK;
}

int ua2_thinking_destroy(ua2_thinking* h) {
  if (!h) return UA2_OK;
  cudaDeviceSynchronize();
  th_free_ws(h);
  for (void* p : h->owned) cudaFree(p);
  delete h;
  return UA2_OK;
}

int ua2_thinking_load_weight(ua2_thinking* h, const char* key_c, const float* dptr, const int64_t* shape, int ndim) {
  UA2_REQUIRE(h && key_c && dptr && shape && ndim >= 1, "null argument");
  const std::string key(key_c);
  const ua2_thinking_cfg& c = h->cfg;
  const int64_t D = c.dim, F = (int64_t)c.dim * c.ff_mult, Cw = c.whisper_dim, Cm = c.mu_dim, hd = c.dim_heads;
  const int64_t n_freq = std::max<int64_t>(hd / 2, 32) / 2;
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
  const Ent tops[] = {{"cls_token", &h->cls, {1, D}},
                      {"down_sampling_layer_whisper.weight", &h->ds.w, {Cw, Cw, 2}},
                      {"down_sampling_layer_whisper.bias", &h->ds.b, {Cw}},
                      {"semantic_merge_proj.weight", &h->merge.w, {D, Cw + Cm}},
                      {"semantic_merge_proj.bias", &h->merge.b, {D}}};
  for (const Ent& t : tops)
    if (key == t.name) {
      UA2_REQUIRE(is(t.shp), key + ": shape mismatch");
      *t.dst = dptr;
      return UA2_OK;
    }
  int idx = -1;
  std::string rest;
  UA2_REQUIRE(th_parse_index(key, "encoder_transformers.", idx, rest) && idx >= 0 && idx < c.depth, "unexpected key " + key);
  ThLayer& L = h->layers[idx];
  const Ent ents[] = {{"self_attn.to_qkv.weight", &L.qkv.w, {3 * D, D}},   {"self_attn.to_out.weight", &L.out.w, {D, D}},
                      {"self_attn.q_norm.weight", &L.qg, {hd}},            {"self_attn.q_norm.bias", &L.qb, {hd}},
                      {"self_attn.k_norm.weight", &L.kg, {hd}},            {"self_attn.k_norm.bias", &L.kb, {hd}},
                      {"self_attn_scale.scale", &L.s1, {D}},               {"ff.ff.0.proj.weight", &L.ff1.w, {2 * F, D}},
                      {"ff.ff.0.proj.bias", &L.ff1.b, {2 * F}},            {"ff.ff.2.weight", &L.ff2.w, {D, F}},
                      {"ff.ff.2.bias", &L.ff2.b, {D}},                     {"ff_scale.scale", &L.s2, {D}},
                      {"rope.inv_freq", &L.inv_freq, {n_freq}}};
  for (const Ent& t : ents)
    if (rest == t.name) {
      UA2_REQUIRE(is(t.shp), key + ": shape mismatch");
      *t.dst = dptr;
      return UA2_OK;
    }
  UA2_REQUIRE(false, "unexpected key " + key);
}

int ua2_thinking_finalize(ua2_thinking* h, void* stream) {
  UA2_REQUIRE(h, "null handle");
  const ua2_thinking_cfg& c = h->cfg;
  UA2_REQUIRE(h->cls && h->ds.w && h->ds.b && h->merge.w && h->merge.b,
              "missing top-level parameters (cls_token, down_sampling_layer_whisper.*, semantic_merge_proj.*)");
  for (int i = 0; i < c.depth; ++i) {
    const ThLayer& L = h->layers[i];
    UA2_REQUIRE(L.qkv.w && L.out.w && L.qg && L.qb && L.kg && L.kb && L.s1 && L.ff1.w && L.ff1.b && L.ff2.w && L.ff2.b && L.s2 && L.inv_freq,
                "missing parameters of encoder_transformers." + std::to_string(i));
  }
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const size_t n = (size_t)c.whisper_dim * c.whisper_dim * 2;
  if (!h->ready) {
    UA2_CHECK_CUDA(cudaMalloc((void**)&h->ds_wr, n * sizeof(float)));
    h->owned.push_back(h->ds_wr);
  }
  CU(launch(lc, th_repack_conv_kernel, dim3(th_grid((long long)n)), dim3(256), 0, h->ds.w, h->ds_wr, c.whisper_dim, c.whisper_dim, 2));
  h->ready = true;
  return UA2_OK;
}

long long ua2_thinking_rows(ua2_thinking* h, int Tw, int Tb) {
  if (!h) return -1;
  const int T = std::min(Tw / 2, Tb), iv = h->cfg.interval;
  if (T < iv || T % iv != 0) return 0;
  return T + T / iv;
}

int ua2_thinking_encode(ua2_thinking* h, const float* whisper, const float* mu, int B, int Tw, int Tb, float* out, void* stream) {
  UA2_REQUIRE(h && whisper && mu && out, "null argument");
  UA2_REQUIRE(h->ready, "ua2_thinking_finalize has not run");
  UA2_REQUIRE(B >= 1 && B <= 4096 && Tw >= 2 && Tb >= 1, "bad batch / frame counts");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.pdl = true;
  int launches = 0;
  lc.launch_counter = &launches;
  const int rc = th_forward(h, lc, whisper, mu, B, Tw, Tb, out);
  h->last_launches = launches;
  return rc;
}

int ua2_thinking_last_launch_count(ua2_thinking* h) { return h ? h->last_launches : 0; }

}  // extern "C"
