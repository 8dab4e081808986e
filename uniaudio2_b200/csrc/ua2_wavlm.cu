// WavLM encoder of ReasoningCodec_film's tokenize direction (the second SSL front-end):
//   tools/tokenizer/ReasoningCodec_film/models/AudioDiffusion1D.py
//     :226      self.wavlm_encoder = AutoModel.from_pretrained(wav_lm_path)            (transformers WavLMModel, 768-wide = base / base-plus)
//     :359-370  get_wavlm_feature: Resample(24000, 16000) -> + 160 zero samples -> wavlm_encoder(wav_16k, output_hidden_states=True)
//               .hidden_states -> stack -> [:, 6:10].mean(1) -> (B, 768, T)
// The model code is third-party (transformers==4.57.0, pyproject.toml:25; absent from the reference tree, present in this image as
// 5.5.0): transformers/models/wavlm/modeling_wavlm.py - WavLMFeatureEncoder (7 Conv1d, GroupNorm on the first, GELU),
// WavLMFeatureProjection (LayerNorm + Linear), WavLMPositionalConvEmbedding (weight-normed grouped Conv1d k 128 + GELU),
// WavLMEncoder (post-LayerNorm layers), WavLMAttention (gated relative position bias through F.multi_head_attention_forward).
// oracle/wavlm_oracle.py restates it and is pinned against the installed class.
//
// Served configuration: feat_extract_norm "group", do_stable_layer_norm false (base / base-plus), erf GELU, no attention mask,
// eval mode.  Two arithmetic modes, like the Whisper encoder (ua2_enc.cu):
//   fp32 class (default; what the 2e-4 parity tests run): linears and strided convolutions as GEMMs on the tcgen05 3xTF32 kernel
//       (ua2_umma.cu), attention on the fp32 SIMT kernel of ua2_dit.cu with the bias added to its scores
//   bf16 (option "bf16" = the reference's arithmetic: fetch_codes_batch runs under torch.autocast(bfloat16), reason_tokenizer.py:114-118):
//       GEMMs on tcgen05 kind::f16 with fp32 accumulation, both attention contractions on tcgen05 with the bias added in the softmax
//       warps (ua2_flash.cu, head size 64), operands handed from kernel to kernel as bf16; residual stream, LayerNorm / GroupNorm
//       statistics, the first convolution, the positional convolution and the gate stay fp32
// Activations are channels-last (B, T, C) from the first convolution on, so every later convolution is a GEMM over im2col rows and
// the transformer consumes the stem's output as it lies.
//
//   wl_conv0_stats_kernel / wl_gn_finalize_kernel / wl_conv0_apply_kernel
//                          conv0 (1 -> C0, k 10, s 5) is 10 FMA per output: it is evaluated twice instead of storing its raw output -
//                          pass 1 reduces per-(clip, channel) sum / sum of squares in double (deterministic: per-chunk partials summed in
//                          order), pass 2 recomputes, applies GroupNorm(C0 groups) + GELU and writes (B, T0, C0)
//   wl_im2col_kernel       rows [x[s t] | ... | x[s t + k - 1]] of a channels-last tensor (float4 along channels)
//   wl_posconv_kernel      grouped convolution (k 128, 16 groups of 48 channels, padding 64, last output dropped) + bias + GELU, fp32 FMA:
//                          CTA = 64 positions of one (clip, group), the (64 + k - 1) x 48 input tile in shared memory (row stride 49:
//                          conflict-free along positions), one tap's 48 x 48 weight slab staged per step and read as broadcasts
//   wl_gate_kernel         gate[b, h, t] = a (b' c_h - 1) + 2 with a, b' = sigmoid of the two 4-sums of gru_rel_pos_linear(x[b, t, head h])
//   wl_bias_table_kernel   tab[h, j - i + T - 1] = rel_attn_embed[bucket(j - i), h]; the (B H, T, T) bias itself is never formed: the
//                          attention kernel (ua2_dit.cu, dit_attn_kernel<.., BIAS>) adds gate * tab to its scores
//   wl_axpy_kernel         running mean of the selected hidden states
#include <algorithm>
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"
#include "ua2_umma.cuh"
#include "ua2_enc_dev.cuh"

namespace ua2 {
namespace {

constexpr int WL_TT = 256;     // conv0: output frames per CTA
constexpr int WL_MAXK0 = 16;   // conv0: kernel bound
constexpr int WL_MAXS0 = 8;    // conv0: stride bound
constexpr int PC_TT = 64;      // pos conv: positions per CTA
constexpr int PC_MAXK = 128;   // pos conv: kernel bound
constexpr int PC_MAXCG = 48;   // pos conv: channels per group bound

__device__ __forceinline__ float wl_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

unsigned wl_grid(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32); }

// stage the samples one CTA's frames read: xs[i] = x[b, s0 * t0 + i], zero past the clip
__device__ __forceinline__ void wl_conv0_stage(float* xs, const float* __restrict__ xb, int L, int t0, int nt, int k0, int s0) {
  const int n = (nt - 1) * s0 + k0;
  const long long base = (long long)t0 * s0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = (base + i < L) ? xb[base + i] : 0.f;
}

// pass 1: part[b, chunk, c] = {sum_t y, sum_t y^2} over the chunk's frames, y = conv0(x)[b, t, c]
__global__ void __launch_bounds__(256) wl_conv0_stats_kernel(const float* __restrict__ x, long long ld, const float* __restrict__ w0,
                                                             const float* __restrict__ b0, double* __restrict__ part, int L, int T0, int C0,
                                                             int k0, int s0) {
  __shared__ float xs[WL_TT * WL_MAXS0 + WL_MAXK0];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, chunk = blockIdx.x, t0 = chunk * WL_TT, nt = min(WL_TT, T0 - t0);
  wl_conv0_stage(xs, x + (size_t)b * ld, L, t0, nt, k0, s0);
  __syncthreads();
  for (int c = threadIdx.x; c < C0; c += blockDim.x) {
    float w[WL_MAXK0];
#pragma unroll
    for (int k = 0; k < WL_MAXK0; ++k) w[k] = k < k0 ? w0[(size_t)c * k0 + k] : 0.f;
    const float bias = b0 ? b0[c] : 0.f;
    double s = 0.0, q = 0.0;
    for (int t = 0; t < nt; ++t) {
      float v = bias;
#pragma unroll
      for (int k = 0; k < WL_MAXK0; ++k)
        if (k < k0) v = fmaf(w[k], xs[t * s0 + k], v);
      s += (double)v;
      q = fma((double)v, (double)v, q);
    }
    double* o = part + (((size_t)b * gridDim.x + chunk) * C0 + c) * 2;
    o[0] = s;
    o[1] = q;
  }
}

// stat[b, c] = {mean, 1 / sqrt(biased variance + eps)} - nn.GroupNorm(num_groups = C0, num_channels = C0)
__global__ void wl_gn_finalize_kernel(const double* __restrict__ part, float* __restrict__ stat, int B, int n_chunks, int C0, int T0, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C0) return;
  const int b = i / C0, c = i - b * C0;
  double s = 0.0, q = 0.0;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const double* o = part + (((size_t)b * n_chunks + ch) * C0 + c) * 2;
    s += o[0];
    q += o[1];
  }
  const double mean = s / (double)T0;
  double var = q / (double)T0 - mean * mean;
  if (var < 0.0) var = 0.0;
  stat[2 * (size_t)i] = (float)mean;
  stat[2 * (size_t)i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// pass 2: out[b, t, c] = gelu((y - mean) * rstd * gamma[c] + beta[c])
__global__ void __launch_bounds__(256) wl_conv0_apply_kernel(const float* __restrict__ x, long long ld, const float* __restrict__ w0,
                                                             const float* __restrict__ b0, const float* __restrict__ stat,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ out, int L, int T0, int C0, int k0, int s0) {
  __shared__ float xs[WL_TT * WL_MAXS0 + WL_MAXK0];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, t0 = blockIdx.x * WL_TT, nt = min(WL_TT, T0 - t0);
  wl_conv0_stage(xs, x + (size_t)b * ld, L, t0, nt, k0, s0);
  __syncthreads();
  for (int c = threadIdx.x; c < C0; c += blockDim.x) {
    float w[WL_MAXK0];
#pragma unroll
    for (int k = 0; k < WL_MAXK0; ++k) w[k] = k < k0 ? w0[(size_t)c * k0 + k] : 0.f;
    const float bias = b0 ? b0[c] : 0.f;
    const float mean = stat[2 * ((size_t)b * C0 + c)], rstd = stat[2 * ((size_t)b * C0 + c) + 1];
    const float g = gamma[c], be = beta[c];
    float* o = out + ((size_t)b * T0 + t0) * C0 + c;
    for (int t = 0; t < nt; ++t) {
      float v = bias;
#pragma unroll
      for (int k = 0; k < WL_MAXK0; ++k)
        if (k < k0) v = fmaf(w[k], xs[t * s0 + k], v);
      o[(size_t)t * C0] = wl_gelu(fmaf((v - mean) * rstd, g, be));
    }
  }
}

// col[(b, t)][j * C + c] = in[b, s * t + j, c], in channels-last (B, Tin, C); C % 4 == 0
template <typename OUT>
__global__ void wl_im2col_kernel(const float* __restrict__ in, OUT* __restrict__ col, int B, int Tin, int Tout, int C, int k, int s) {
  pdl_launch_dependents();
  pdl_wait();
  const int C4 = C / 4;
  const long long n4 = (long long)B * Tout * k * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long r = i / C4;
    const int j = (int)(r % k);
    r /= k;
    const int t = (int)(r % Tout), b = (int)(r / Tout);
    const float4 v = *reinterpret_cast<const float4*>(in + ((size_t)b * Tin + (size_t)s * t + j) * C + 4 * c4);
    OUT* dst = col + ((size_t)b * Tout + t) * ((size_t)k * C) + (size_t)j * C + 4 * c4;
    if (sizeof(OUT) == 2) {
      *reinterpret_cast<uint2*>(dst) = pack4_bf16(v.x, v.y, v.z, v.w);
    } else {
      *reinterpret_cast<float4*>(dst) = v;
    }
  }
}

// torch Conv1d weight (Cout, Cin, k) -> GEMM weight (Cout, k * Cin), column j * Cin + c
__global__ void wl_repack_conv_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int k) {
  const long long n = (long long)Cout * Cin * k;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cin);
    const int j = (int)((i / Cin) % k);
    const long long co = i / ((long long)Cin * k);
    out[i] = w[(co * Cin + c) * k + j];
  }
}

// grouped Conv1d weight (D, cg, K) -> [g][j][ci][co_local]: one tap of one group is a contiguous cg x cg slab
__global__ void wl_repack_posconv_kernel(const float* __restrict__ w, float* __restrict__ out, int D, int cg, int K) {
  const long long n = (long long)D * cg * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % cg);
    long long r = i / cg;
    const int ci = (int)(r % cg);
    r /= cg;
    const int j = (int)(r % K), g = (int)(r / K);
    out[i] = w[(((size_t)g * cg + col) * cg + ci) * K + j];
  }
}

// p[b, t, g cg + co] = gelu(bias + sum_j sum_ci w[g cg + co, ci, j] * h[b, t + j - K / 2, g cg + ci]),  t in [0, T)
// thread = two positions (t0 + tl, t0 + tl + 32) x up to 6 output channels (cq + 8 i): per input channel 2 conflict-free loads of the
// input tile and up to 6 broadcast loads of the tap's weight slab feed 12 FMAs; the next tap's slab is fetched into registers while
// this tap is computed (ncu on the first version - one position per thread, slab fetched between the barriers: 6.2 ms at 6 x 1500
// frames, shared-memory-load bound).
__global__ void __launch_bounds__(256) wl_posconv_kernel(const float* __restrict__ h, const float* __restrict__ wr, const float* __restrict__ bias,
                                                         float* __restrict__ p, int T, int D, int cg, int K) {
  __shared__ float Xs[(PC_TT + PC_MAXK - 1) * (PC_MAXCG + 1)];
  __shared__ float Ws[PC_MAXCG * PC_MAXCG];
  constexpr int WPT = (PC_MAXCG * PC_MAXCG + 255) / 256;  // slab elements per thread
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, tl = tid & 31, cq = tid >> 5;
  const int t0 = blockIdx.x * PC_TT, g = blockIdx.y, b = blockIdx.z;
  const int pad = K / 2, xst = cg + 1, rows = PC_TT + K - 1, slab = cg * cg;
  for (int i = tid; i < rows * cg; i += 256) {
    const int r = i / cg, ci = i - r * cg;
    const int ts = t0 + r - pad;
    Xs[r * xst + ci] = (ts >= 0 && ts < T) ? h[((size_t)b * T + ts) * D + g * cg + ci] : 0.f;
  }
  float acc0[PC_MAXCG / 8], acc1[PC_MAXCG / 8], wreg[WPT];
#pragma unroll
  for (int i = 0; i < PC_MAXCG / 8; ++i) acc0[i] = acc1[i] = 0.f;
  const float* wg = wr + (size_t)g * K * slab;
#pragma unroll
  for (int q = 0; q < WPT; ++q) wreg[q] = (tid + 256 * q < slab) ? wg[tid + 256 * q] : 0.f;
  for (int j = 0; j < K; ++j) {
    __syncthreads();  // the input tile is complete (first step) / the previous tap's slab is consumed
#pragma unroll
    for (int q = 0; q < WPT; ++q)
      if (tid + 256 * q < slab) Ws[tid + 256 * q] = wreg[q];
    __syncthreads();
    if (j + 1 < K) {
      const float* wn = wg + (size_t)(j + 1) * slab;
#pragma unroll
      for (int q = 0; q < WPT; ++q) wreg[q] = (tid + 256 * q < slab) ? wn[tid + 256 * q] : 0.f;
    }
    const float* xr0 = Xs + (tl + j) * xst;
    const float* xr1 = xr0 + 32 * xst;
    for (int ci = 0; ci < cg; ++ci) {
      const float x0 = xr0[ci], x1 = xr1[ci];
      const float* wrow = Ws + ci * cg + cq;
#pragma unroll
      for (int i = 0; i < PC_MAXCG / 8; ++i)
        if (cq + 8 * i < cg) {
          const float w = wrow[8 * i];
          acc0[i] = fmaf(x0, w, acc0[i]);
          acc1[i] = fmaf(x1, w, acc1[i]);
        }
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int t = t0 + tl + 32 * half;
    if (t < T) {
#pragma unroll
      for (int i = 0; i < PC_MAXCG / 8; ++i) {
        const int co = cq + 8 * i;
        if (co < cg) p[((size_t)b * T + t) * D + g * cg + co] = wl_gelu((half ? acc1[i] : acc0[i]) + bias[g * cg + co]);
      }
    }
  }
}

// one warp per (row m, head): WavLMAttention.forward steps 1-3
__global__ void __launch_bounds__(256) wl_gate_kernel(const float* __restrict__ h, const float* __restrict__ Wg, const float* __restrict__ bg,
                                                      const float* __restrict__ cst, float* __restrict__ gate, int M, int T, int H, int hs) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * 8 + warp;
  if (idx >= (long long)M * H) return;  // whole warps leave together
  const int m = (int)(idx / H), hh = (int)(idx - (long long)m * H);
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = 0.f;
  const float* xr = h + (size_t)m * H * hs + (size_t)hh * hs;
  for (int d = lane; d < hs; d += 32) {
    const float xv = xr[d];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = fmaf(xv, Wg[o * hs + d], acc[o]);
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = warp_sum(acc[o]);
  if (lane == 0) {
    const float pa = ((acc[0] + bg[0]) + (acc[1] + bg[1])) + ((acc[2] + bg[2]) + (acc[3] + bg[3]));
    const float pb = ((acc[4] + bg[4]) + (acc[5] + bg[5])) + ((acc[6] + bg[6]) + (acc[7] + bg[7]));
    const float ga = 1.f / (1.f + expf(-pa)), gb = 1.f / (1.f + expf(-pb));
    const int b = m / T, t = m - b * T;
    gate[((size_t)b * H + hh) * T + t] = ga * (gb * cst[hh] - 1.f) + 2.f;
  }
}

// tab[h, r] = emb[bucket[r], h], r = j - i + T - 1 in [0, 2 T - 2]
__global__ void wl_bias_table_kernel(const float* __restrict__ emb, const int32_t* __restrict__ bucket, float* __restrict__ tab, int H, int n) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * n) return;
  const int hh = i / n, r = i - hh * n;
  tab[i] = emb[(size_t)bucket[r] * H + hh];
}

// out = (first ? 0 : out) + alpha * h
__global__ void wl_axpy_kernel(const float* __restrict__ h, float* __restrict__ out, float alpha, int first, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(h)[i];
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!first) o = reinterpret_cast<const float4*>(out)[i];
    o.x = fmaf(alpha, v.x, o.x);
    o.y = fmaf(alpha, v.y, o.y);
    o.z = fmaf(alpha, v.z, o.z);
    o.w = fmaf(alpha, v.w, o.w);
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------- launchers
cudaError_t launch_wl_conv0(const LaunchCtx& lc, const float* x, long long ld, const float* w0, const float* b0, const float* gamma,
                            const float* beta, double* part, float* stat, float* out, int B, int L, int T0, int C0, int k0, int s0, float eps) {
  const int n_chunks = (T0 + WL_TT - 1) / WL_TT;
  const dim3 grid((unsigned)n_chunks, (unsigned)B);
  cudaError_t e = launch(lc, wl_conv0_stats_kernel, grid, dim3(256), 0, x, ld, w0, b0, part, L, T0, C0, k0, s0);
  if (e != cudaSuccess) return e;
  e = launch(lc, wl_gn_finalize_kernel, dim3((unsigned)((B * C0 + 255) / 256)), dim3(256), 0, (const double*)part, stat, B, n_chunks, C0, T0, eps);
  if (e != cudaSuccess) return e;
  return launch(lc, wl_conv0_apply_kernel, grid, dim3(256), 0, x, ld, w0, b0, (const float*)stat, gamma, beta, out, L, T0, C0, k0, s0);
}

template <typename OUT>
cudaError_t launch_wl_im2col(const LaunchCtx& lc, const float* in, OUT* col, int B, int Tin, int Tout, int C, int k, int s) {
  return launch(lc, wl_im2col_kernel<OUT>, dim3(wl_grid((long long)B * Tout * k * (C / 4))), dim3(256), 0, in, col, B, Tin, Tout, C, k, s);
}

cudaError_t launch_wl_posconv(const LaunchCtx& lc, const float* h, const float* wr, const float* bias, float* p, int B, int T, int D, int cg,
                              int K) {
  if (cg < 1 || cg > PC_MAXCG || K < 1 || K > PC_MAXK || D % cg != 0) return cudaErrorInvalidValue;
  return launch(lc, wl_posconv_kernel, dim3((unsigned)((T + PC_TT - 1) / PC_TT), (unsigned)(D / cg), (unsigned)B), dim3(256), 0, h, wr, bias, p, T,
                D, cg, K);
}

cudaError_t launch_wl_gate(const LaunchCtx& lc, const float* h, const float* Wg, const float* bg, const float* cst, float* gate, int B, int T,
                           int H, int hs) {
  const long long items = (long long)B * T * H;
  return launch(lc, wl_gate_kernel, dim3((unsigned)((items + 7) / 8)), dim3(256), 0, h, Wg, bg, cst, gate, B * T, T, H, hs);
}

// WavLMAttention._relative_positions_bucket for relative position rp = j - i (host; num_buckets 320, max_distance 800 in the checkpoints)
int wl_rel_bucket(int rp, int num_buckets, int max_distance) {
  const int nb = num_buckets / 2;
  int bucket = rp > 0 ? nb : 0;
  const int a = rp < 0 ? -rp : rp;
  const int max_exact = nb / 2;
  if (a < max_exact) return bucket + a;
  float v = std::log((float)a / (float)max_exact);                         // torch.log(relative_positions.float() / max_exact)
  v = v / (float)std::log((double)max_distance / (double)max_exact);      // / math.log(max_distance / max_exact)
  v = v * (float)(nb - max_exact);
  long long large = (long long)((float)max_exact + v);                     // (max_exact + x).to(torch.long): truncation
  if (large > nb - 1) large = nb - 1;
  return bucket + (int)large;
}

}  // namespace
}  // namespace ua2

using namespace ua2;

namespace {

constexpr int WL_MAX_CONV = 8;

struct WlLayer {
  Lin q, k, v, o, ff1, ff2;
  const float *gru_w = nullptr, *gru_b = nullptr, *gru_c = nullptr;
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  float *wqkv = nullptr, *bqkv = nullptr;  // owned
};

}  // namespace

struct ua2_wavlm {
  ua2_wavlm_cfg cfg{};
  Lin conv[WL_MAX_CONV];
  const float *gn_g = nullptr, *gn_b = nullptr, *fp_ln_g = nullptr, *fp_ln_b = nullptr;
  Lin fp, pos;
  const float *enc_ln_g = nullptr, *enc_ln_b = nullptr, *rel_embed = nullptr;
  std::vector<WlLayer> layers;
  float* conv_wr[WL_MAX_CONV] = {};  // owned GEMM-form weights of conv layers 1..
  float *pos_wr = nullptr, *zeros = nullptr;
  std::vector<void*> owned;
  bool ready = false;
  // workspace, sized for (max_B clips, max_L samples)
  int max_B = 0, max_L = 0;
  float *x0 = nullptr, *x1 = nullptr, *col = nullptr, *gnstat = nullptr, *h = nullptr, *n = nullptr, *q = nullptr, *k = nullptr, *v = nullptr,
        *att = nullptr, *ff = nullptr, *p = nullptr, *gate = nullptr, *tab = nullptr, *stats = nullptr;
  double* part = nullptr;
  int32_t* bucket_dev = nullptr;
  std::vector<int32_t> bucket_host;
  int tab_T = 0;  // the T the device table was built for (0 = none; reset when weights change)
  size_t stats_floats = 0;
  TcWorkspace tc;
  int opt_bf16 = 0;
  __nv_bfloat16* a16 = nullptr;                   // bf16 operand rows of the next tensor-core linear
  std::map<const float*, __nv_bfloat16*> w16;     // bf16 copies of the weights, made at first use
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

// frames after each convolution of the feature encoder (no padding): T_i = (T_{i-1} - k_i) / s_i + 1; false when the clip is too short
bool wl_frames(const ua2_wavlm_cfg& c, long long L, long long* T) {
  long long t = L;
  for (int i = 0; i < c.num_feat_extract_layers; ++i) {
    if (t < c.conv_kernel[i]) return false;
    t = (t - c.conv_kernel[i]) / c.conv_stride[i] + 1;
    T[i] = t;
  }
  return true;
}

void wl_free_ws(ua2_wavlm* h) {
  for (float** p : {&h->x0, &h->x1, &h->col, &h->gnstat, &h->h, &h->n, &h->q, &h->k, &h->v, &h->att, &h->ff, &h->p, &h->gate, &h->tab, &h->stats,
                    &h->tc.a, &h->tc.slots, &h->tc.c}) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
  if (h->part) cudaFree(h->part);
  h->part = nullptr;
  if (h->bucket_dev) cudaFree(h->bucket_dev);
  h->bucket_dev = nullptr;
  if (h->a16) cudaFree(h->a16);
  h->a16 = nullptr;
  h->tab_T = 0;
}

int wl_dmalloc(float** p, size_t floats) {
  UA2_CHECK_CUDA(cudaMalloc((void**)p, std::max<size_t>(floats, 4) * sizeof(float)));
  return UA2_OK;
}

int wl_reserve(ua2_wavlm* h, int B, int L) {
  if (B <= h->max_B && L <= h->max_L) return UA2_OK;
  B = std::max(B, h->max_B);
  L = std::max(L, h->max_L);
  const ua2_wavlm_cfg& c = h->cfg;
  long long Ts[WL_MAX_CONV];
  UA2_REQUIRE(wl_frames(c, L, Ts), "clip shorter than the feature encoder's receptive field");
  const int nc = c.num_feat_extract_layers;
  const size_t D = c.hidden_size, F = c.intermediate_size, H = c.num_attention_heads, T = Ts[nc - 1], M = (size_t)B * T;
  if (h->max_B) UA2_CHECK_CUDA(cudaDeviceSynchronize());
  wl_free_ws(h);
  size_t even = 4, odd = 4, col = 4, rows = M, cmax = M * std::max(3 * D, F);
  for (int i = 0; i < nc; ++i) {
    const size_t sz = (size_t)((size_t)B * Ts[i] * c.conv_dim[i]);
    if (i % 2 == 0) even = std::max(even, sz); else odd = std::max(odd, sz);
    if (i >= 1) {
      col = std::max(col, (size_t)((size_t)B * Ts[i] * c.conv_kernel[i] * c.conv_dim[i - 1]));
      rows = std::max(rows, (size_t)((size_t)B * Ts[i]));
      cmax = std::max(cmax, sz);
    }
  }
  col = std::max(col, M * std::max({D, F, (size_t)c.conv_dim[nc - 1]}));  // widest operand rows of the transformer's linears
  UA2_REQUIRE(col < (size_t)1 << 31 && cmax < (size_t)1 << 31, "batch x clip length too large for one call: split the batch");
  const size_t n_chunks = (Ts[0] + WL_TT - 1) / WL_TT;
  RUN(wl_dmalloc(&h->x0, even));
  RUN(wl_dmalloc(&h->x1, odd));
  RUN(wl_dmalloc(&h->col, col));
  RUN(wl_dmalloc(&h->gnstat, 2 * (size_t)B * c.conv_dim[0]));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->part, (size_t)B * n_chunks * c.conv_dim[0] * 2 * sizeof(double)));
  RUN(wl_dmalloc(&h->h, M * D));
  RUN(wl_dmalloc(&h->n, M * std::max(D, (size_t)c.conv_dim[nc - 1])));
  RUN(wl_dmalloc(&h->q, M * D));
  RUN(wl_dmalloc(&h->k, M * D));
  RUN(wl_dmalloc(&h->v, M * D));
  RUN(wl_dmalloc(&h->att, M * D));
  RUN(wl_dmalloc(&h->ff, M * F));
  RUN(wl_dmalloc(&h->p, M * D));
  RUN(wl_dmalloc(&h->gate, M * H));
  RUN(wl_dmalloc(&h->tab, H * (2 * T)));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->bucket_dev, 2 * T * sizeof(int32_t)));
  h->stats_floats = 2 * rows + 16;
  RUN(wl_dmalloc(&h->stats, h->stats_floats));
  h->tc.a_floats = 2 * col;
  h->tc.slots_floats = tc_slots_max_floats();
  h->tc.c_floats = cmax;
  RUN(wl_dmalloc(&h->tc.a, h->tc.a_floats));
  RUN(wl_dmalloc(&h->tc.slots, h->tc.slots_floats));
  RUN(wl_dmalloc(&h->tc.c, h->tc.c_floats));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->a16, col * sizeof(__nv_bfloat16)));
  h->max_B = B;
  h->max_L = L;
  return UA2_OK;
}

// x (M, K) @ W (N, K)^T, raw product left in the tensor-core workspace; the epilogue description comes back filled in
int wl_linear(ua2_wavlm* h, const LaunchCtx& lc, const float* x, const float* W, const float* bias, int M, int N, int K, int T, EncEpi* e) {
  *e = EncEpi{};
  e->bias = bias;
  e->M = M;
  e->N = N;
  e->T = T;
  e->eps = h->cfg.layer_norm_eps;
#ifndef UA2_CPU_SHIM
  if (h->opt_bf16) {  // operand rows are already bf16 in h->a16; the stream-K side slots are left for the epilogue to sum
    auto it = h->w16.find(W);
    const bool fresh = it == h->w16.end();
    if (fresh) {
      __nv_bfloat16* wb = nullptr;
      UA2_CHECK_CUDA(cudaMalloc((void**)&wb, (size_t)N * K * sizeof(__nv_bfloat16)));
      it = h->w16.emplace(W, wb).first;
      const long long n4 = (long long)N * K / 4;
      CU(launch(lc, enc_to_bf16_kernel, dim3(wl_grid(n4)), dim3(256), 0, W, wb, n4));
    }
    const UmmaPlan pl = umma_plan(M, N, 1, K, true);
    UA2_REQUIRE(pl.grid >= 1 && pl.slot_floats <= h->tc.slots_floats && (size_t)M * N <= h->tc.c_floats && (K % 8) == 0 && (N % 4) == 0,
                "encoder linear outside the tensor-core path's shapes");
    LaunchCtx lg = lc;
    if (fresh) lg.pdl = false;  // the GEMM prefetches weights before griddepcontrol.wait: full stream order behind the conversion kernel
    CU(run_umma_bf16(lg, h->a16, it->second, h->tc.c, N, h->tc.slots, M, N, K, pl));
    e->src = h->tc.c;
    e->slots = h->tc.slots;
    e->pl = pl;
    e->has_split = umma_has_split_tiles(pl) ? 1 : 0;
    return UA2_OK;
  }
#endif
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
  e->src = raw ? raw : h->tc.c;
  return UA2_OK;
}

int wl_forward(ua2_wavlm* h, const LaunchCtx& lc, const float* wav, long long ld, int B, int L, int hs_lo, int hs_hi, float* out, float* all_hidden) {
  const ua2_wavlm_cfg& c = h->cfg;
  const int nc = c.num_feat_extract_layers, D = c.hidden_size, F = c.intermediate_size, H = c.num_attention_heads, hs = D / H;
  long long Ts[WL_MAX_CONV];
  UA2_REQUIRE(wl_frames(c, L, Ts), "clip shorter than the feature encoder's receptive field");
  const int T = (int)Ts[nc - 1], M = B * T;
  UA2_REQUIRE(M >= 32, "fewer than 32 frames in the batch");
  const bool bf = h->opt_bf16 != 0;
  EncEpi e;
  // ---- feature encoder: conv0 + GroupNorm + GELU, then GEMM convolutions + GELU, ping-pong between x0 / x1
  CU(launch_wl_conv0(lc, wav, ld, h->conv[0].w, h->conv[0].b, h->gn_g, h->gn_b, h->part, h->gnstat, h->x0, B, L, (int)Ts[0], c.conv_dim[0],
                     c.conv_kernel[0], c.conv_stride[0], 1e-5f));
  float* cur = h->x0;
  for (int i = 1; i < nc; ++i) {
    float* nxt = (i % 2) ? h->x1 : h->x0;
    const int Tin = (int)Ts[i - 1], Tout = (int)Ts[i], Cin = c.conv_dim[i - 1], Cout = c.conv_dim[i], k = c.conv_kernel[i], s = c.conv_stride[i];
    if (bf) {
      CU(launch_wl_im2col(lc, cur, h->a16, B, Tin, Tout, Cin, k, s));
    } else {
      CU(launch_wl_im2col(lc, cur, h->col, B, Tin, Tout, Cin, k, s));
    }
    RUN(wl_linear(h, lc, h->col, h->conv_wr[i], h->conv[i].b ? h->conv[i].b : h->zeros, B * Tout, Cout, k * Cin, Tout, &e));
    e.y32 = nxt;
    CU(launch_enc_epi<EE_GELU>(lc, e));
    cur = nxt;
  }
  // ---- feature projection: LayerNorm(C_last) -> Linear -> h (B, T, D)
  const int CL = c.conv_dim[nc - 1];
  {
    EncEpi l{};
    l.M = M;
    l.N = CL;
    l.res = cur;
    l.ln_g = h->fp_ln_g;
    l.ln_b = h->fp_ln_b;
    l.eps = c.layer_norm_eps;
    l.y32 = h->n;
    l.y16 = bf ? h->a16 : nullptr;
    CU(launch(lc, enc_res_ln_kernel<false>, dim3(M), dim3((unsigned)(((CL / 4) + 31) / 32 * 32)), 0, l));
  }
  RUN(wl_linear(h, lc, h->n, h->fp.w, h->fp.b, M, D, CL, T, &e));
  e.y32 = h->h;
  CU(launch_enc_epi<EE_BIAS>(lc, e));
  // ---- h = LayerNorm(h + gelu(pos_conv(h)))
  const unsigned ln_threads = (unsigned)(((D / 4) + 31) / 32 * 32);
  CU(launch_wl_posconv(lc, h->h, h->pos_wr, h->pos.b, h->p, B, T, D, D / c.num_conv_pos_embedding_groups, c.num_conv_pos_embeddings));
  {
    EncEpi l{};
    l.src = h->p;
    l.bias = h->zeros;
    l.M = M;
    l.N = D;
    l.T = T;
    l.res = h->h;
    l.ln_g = h->enc_ln_g;
    l.ln_b = h->enc_ln_b;
    l.eps = c.layer_norm_eps;
    l.y32 = h->h;  // the normalised row replaces the residual stream (post-LayerNorm encoder)
    l.y16_also = bf ? h->a16 : nullptr;
    CU(launch(lc, enc_res_ln_kernel<true>, dim3(M), dim3(ln_threads), 0, l));
  }
  const long long n4 = (long long)M * D / 4;
  const float alpha = 1.f / (float)(hs_hi - hs_lo);
  auto emit = [&](int idx) -> int {  // hidden_states[idx] is in h->h
    if (all_hidden) UA2_CHECK_CUDA(cudaMemcpyAsync(all_hidden + (size_t)idx * M * D, h->h, (size_t)M * D * sizeof(float), cudaMemcpyDeviceToDevice, lc.stream));
    if (idx >= hs_lo && idx < hs_hi) CU(launch(lc, wl_axpy_kernel, dim3(wl_grid(n4)), dim3(256), 0, (const float*)h->h, out, alpha, idx == hs_lo ? 1 : 0, n4));
    return UA2_OK;
  };
  RUN(emit(0));
  const int n_run = hs_hi - 1;  // layers whose output is needed
  if (n_run > 0 && h->tab_T != T) {
    const int n = 2 * T - 1;
    h->bucket_host.resize(n);
    for (int r = 0; r < n; ++r) h->bucket_host[r] = wl_rel_bucket(r - (T - 1), c.num_buckets, c.max_bucket_distance);
    UA2_CHECK_CUDA(cudaMemcpyAsync(h->bucket_dev, h->bucket_host.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice, lc.stream));
    UA2_CHECK_CUDA(cudaStreamSynchronize(lc.stream));  // the host table may be rebuilt by the next call
    LaunchCtx l0 = lc;
    l0.pdl = false;
    CU(launch(l0, wl_bias_table_kernel, dim3((unsigned)((H * n + 255) / 256)), dim3(256), 0, h->rel_embed, (const int32_t*)h->bucket_dev, h->tab, H, n));
    h->tab_T = T;
  }
  for (int li = 0; li < n_run; ++li) {
    const WlLayer& Ly = h->layers[li];
    CU(launch_wl_gate(lc, h->h, Ly.gru_w, Ly.gru_b, Ly.gru_c, h->gate, B, T, H, hs));
    RUN(wl_linear(h, lc, h->h, Ly.wqkv, Ly.bqkv, M, 3 * D, D, T, &e));
    e.H = H;
    e.hs = hs;
#ifndef UA2_CPU_SHIM
    if (bf) {
      e.q16 = reinterpret_cast<__nv_bfloat16*>(h->q);
      e.k16 = reinterpret_cast<__nv_bfloat16*>(h->k);
      e.v16 = reinterpret_cast<__nv_bfloat16*>(h->v);
      CU(launch_enc_epi<EE_QKV16>(lc, e));
      CU(launch_flash_bf16_bias(lc, e.q16, e.k16, e.v16, nullptr, h->a16, B, T, H, hs, h->gate, h->tab));
    } else
#endif
    {
      e.q = h->q;
      e.k = h->k;
      e.v = h->v;
      CU(launch_enc_epi<EE_QKV>(lc, e));
      CU(launch_dense_attn_bias_f32(lc, h->q, h->k, h->v, h->att, B, T, H, hs, h->gate, h->tab));
    }
    // h = layer_norm(h + out_proj(att))
    RUN(wl_linear(h, lc, h->att, Ly.o.w, Ly.o.b, M, D, D, T, &e));
    e.res = h->h;
    e.ln_g = Ly.ln1_g;
    e.ln_b = Ly.ln1_b;
    e.y32 = h->h;
    e.y16_also = bf ? h->a16 : nullptr;
    CU(launch(lc, enc_res_ln_kernel<true>, dim3(M), dim3(ln_threads), 0, e));
    // h = final_layer_norm(h + output_dense(gelu(intermediate_dense(h))))
    RUN(wl_linear(h, lc, h->h, Ly.ff1.w, Ly.ff1.b, M, F, D, T, &e));
    e.y32 = h->ff;
    e.y16 = bf ? h->a16 : nullptr;
    CU(launch_enc_epi<EE_GELU>(lc, e));
    RUN(wl_linear(h, lc, h->ff, Ly.ff2.w, Ly.ff2.b, M, D, F, T, &e));
    e.res = h->h;
    e.ln_g = Ly.ln2_g;
    e.ln_b = Ly.ln2_b;
    e.y32 = h->h;
    e.y16_also = bf ? h->a16 : nullptr;
    CU(launch(lc, enc_res_ln_kernel<true>, dim3(M), dim3(ln_threads), 0, e));
    RUN(emit(li + 1));
  }
  return UA2_OK;
}

bool wl_parse_index(const std::string& key, const std::string& pre, int& idx, std::string& rest) {
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

int ua2_wavlm_create(const ua2_wavlm_cfg* cfg, ua2_wavlm** out) {
  UA2_REQUIRE(cfg && out, "null argument");
  const ua2_wavlm_cfg& c = *cfg;
  UA2_REQUIRE(c.hidden_size >= 32 && c.num_attention_heads >= 1 && c.num_hidden_layers >= 1 && c.intermediate_size >= 8, "bad dimensions");
  UA2_REQUIRE(c.hidden_size % c.num_attention_heads == 0, "hidden_size must be divisible by num_attention_heads");
  const int hs = c.hidden_size / c.num_attention_heads;
  UA2_REQUIRE(hs == 32 || hs == 64 || hs == 128, "head size must be 32 / 64 / 128");
  UA2_REQUIRE(c.hidden_size % 8 == 0 && c.hidden_size <= 4096 && c.intermediate_size % 8 == 0, "hidden_size (<= 4096) and intermediate_size must be multiples of 8");
  UA2_REQUIRE(c.num_feat_extract_layers >= 2 && c.num_feat_extract_layers <= WL_MAX_CONV, "2 .. 8 feature-encoder layers");
  UA2_REQUIRE(c.conv_kernel[0] >= 1 && c.conv_kernel[0] <= WL_MAXK0 && c.conv_stride[0] >= 1 && c.conv_stride[0] <= WL_MAXS0,
              "first convolution: kernel <= 16, stride <= 8");
  for (int i = 0; i < c.num_feat_extract_layers; ++i)
    UA2_REQUIRE(c.conv_dim[i] >= 8 && c.conv_dim[i] % 8 == 0 && c.conv_dim[i] <= 4096 && c.conv_kernel[i] >= 1 && c.conv_stride[i] >= 1,
                "conv_dim must be multiples of 8 (<= 4096), kernels and strides positive");
  UA2_REQUIRE(c.num_conv_pos_embedding_groups >= 1 && c.hidden_size % c.num_conv_pos_embedding_groups == 0 &&
                  c.hidden_size / c.num_conv_pos_embedding_groups <= PC_MAXCG && c.num_conv_pos_embeddings >= 2 &&
                  c.num_conv_pos_embeddings <= PC_MAXK && c.num_conv_pos_embeddings % 2 == 0,
              "positional convolution: even kernel <= 128, at most 48 channels per group");
  UA2_REQUIRE(c.num_buckets >= 4 && c.num_buckets % 4 == 0 && c.max_bucket_distance > c.num_buckets / 4, "bad relative-position bucket geometry");
  UA2_REQUIRE(c.layer_norm_eps > 0.f, "layer_norm_eps must be positive");
  ua2_wavlm* h = new ua2_wavlm();
  h->cfg = c;
  h->layers.resize(c.num_hidden_layers);
  *out = h;
  return UA2_OK;
}

int ua2_wavlm_destroy(ua2_wavlm* h) {
  if (!h) return UA2_OK;
  cudaDeviceSynchronize();
  wl_free_ws(h);
  for (void* p : h->owned) cudaFree(p);
  for (auto& kv : h->w16) cudaFree(kv.second);
  delete h;
  return UA2_OK;
}

int ua2_wavlm_load_weight(ua2_wavlm* h, const char* key_c, const float* dptr, const int64_t* shape, int ndim) {
  UA2_REQUIRE(h && key_c && dptr && shape && ndim >= 1, "null argument");
  const std::string key(key_c);
  const ua2_wavlm_cfg& c = h->cfg;
  const int64_t D = c.hidden_size, F = c.intermediate_size, H = c.num_attention_heads, hs = D / H, CL = c.conv_dim[c.num_feat_extract_layers - 1];
  const int64_t cg = D / c.num_conv_pos_embedding_groups, PK = c.num_conv_pos_embeddings;
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
  h->tab_T = 0;  // rel_attn_embed may be what changes
  const Ent tops[] = {{"feature_extractor.conv_layers.0.layer_norm.weight", &h->gn_g, {c.conv_dim[0]}},
                      {"feature_extractor.conv_layers.0.layer_norm.bias", &h->gn_b, {c.conv_dim[0]}},
                      {"feature_projection.layer_norm.weight", &h->fp_ln_g, {CL}},
                      {"feature_projection.layer_norm.bias", &h->fp_ln_b, {CL}},
                      {"feature_projection.projection.weight", &h->fp.w, {D, CL}},
                      {"feature_projection.projection.bias", &h->fp.b, {D}},
                      {"encoder.pos_conv_embed.conv.weight", &h->pos.w, {D, cg, PK}},
                      {"encoder.pos_conv_embed.conv.bias", &h->pos.b, {D}},
                      {"encoder.layer_norm.weight", &h->enc_ln_g, {D}},
                      {"encoder.layer_norm.bias", &h->enc_ln_b, {D}},
                      {"encoder.layers.0.attention.rel_attn_embed.weight", &h->rel_embed, {c.num_buckets, H}}};
  for (const Ent& t : tops)
    if (key == t.name) {
      UA2_REQUIRE(is(t.shp), key + ": shape mismatch");
      *t.dst = dptr;
      return UA2_OK;
    }
  int idx = -1;
  std::string rest;
  if (wl_parse_index(key, "feature_extractor.conv_layers.", idx, rest)) {
    UA2_REQUIRE(idx >= 0 && idx < c.num_feat_extract_layers, "unexpected key " + key);
    const int64_t Cin = idx == 0 ? 1 : c.conv_dim[idx - 1];
    if (rest == "conv.weight") {
      UA2_REQUIRE(is({c.conv_dim[idx], Cin, c.conv_kernel[idx]}), key + ": shape mismatch");
      h->conv[idx].w = dptr;
      return UA2_OK;
    }
    if (rest == "conv.bias") {
      UA2_REQUIRE(c.conv_bias, key + ": the configuration has conv_bias = false");
      UA2_REQUIRE(is({c.conv_dim[idx]}), key + ": shape mismatch");
      h->conv[idx].b = dptr;
      return UA2_OK;
    }
    UA2_REQUIRE(false, "unexpected key " + key + " (feat_extract_norm 'group' has a norm on layer 0 only)");
  }
  UA2_REQUIRE(wl_parse_index(key, "encoder.layers.", idx, rest) && idx >= 0 && idx < c.num_hidden_layers, "unexpected key " + key);
  WlLayer& L = h->layers[idx];
  const Ent ents[] = {{"attention.q_proj.weight", &L.q.w, {D, D}},
                      {"attention.q_proj.bias", &L.q.b, {D}},
                      {"attention.k_proj.weight", &L.k.w, {D, D}},
                      {"attention.k_proj.bias", &L.k.b, {D}},
                      {"attention.v_proj.weight", &L.v.w, {D, D}},
                      {"attention.v_proj.bias", &L.v.b, {D}},
                      {"attention.out_proj.weight", &L.o.w, {D, D}},
                      {"attention.out_proj.bias", &L.o.b, {D}},
                      {"attention.gru_rel_pos_linear.weight", &L.gru_w, {8, hs}},
                      {"attention.gru_rel_pos_linear.bias", &L.gru_b, {8}},
                      {"attention.gru_rel_pos_const", &L.gru_c, {1, H, 1, 1}},
                      {"layer_norm.weight", &L.ln1_g, {D}},
                      {"layer_norm.bias", &L.ln1_b, {D}},
                      {"feed_forward.intermediate_dense.weight", &L.ff1.w, {F, D}},
                      {"feed_forward.intermediate_dense.bias", &L.ff1.b, {F}},
                      {"feed_forward.output_dense.weight", &L.ff2.w, {D, F}},
                      {"feed_forward.output_dense.bias", &L.ff2.b, {D}},
                      {"final_layer_norm.weight", &L.ln2_g, {D}},
                      {"final_layer_norm.bias", &L.ln2_b, {D}}};
  for (const Ent& t : ents)
    if (rest == t.name) {
      UA2_REQUIRE(is(t.shp), key + ": shape mismatch");
      *t.dst = dptr;
      return UA2_OK;
    }
  UA2_REQUIRE(false, "unexpected key " + key);
}

int ua2_wavlm_finalize(ua2_wavlm* h, void* stream) {
  UA2_REQUIRE(h, "null handle");
  const ua2_wavlm_cfg& c = h->cfg;
  const size_t D = c.hidden_size, F = c.intermediate_size;
  const int nc = c.num_feat_extract_layers;
  UA2_REQUIRE(h->gn_g && h->gn_b && h->fp_ln_g && h->fp_ln_b && h->fp.w && h->fp.b && h->pos.w && h->pos.b && h->enc_ln_g && h->enc_ln_b && h->rel_embed,
              "missing top-level parameters (conv_layers.0.layer_norm.*, feature_projection.*, encoder.pos_conv_embed.conv.*, encoder.layer_norm.*, "
              "encoder.layers.0.attention.rel_attn_embed.weight)");
  for (int i = 0; i < nc; ++i)
    UA2_REQUIRE(h->conv[i].w && (!c.conv_bias || h->conv[i].b), "missing parameters of feature_extractor.conv_layers." + std::to_string(i));
  for (int i = 0; i < c.num_hidden_layers; ++i) {
    const WlLayer& L = h->layers[i];
    UA2_REQUIRE(L.q.w && L.q.b && L.k.w && L.k.b && L.v.w && L.v.b && L.o.w && L.o.b && L.gru_w && L.gru_b && L.gru_c && L.ln1_g && L.ln1_b && L.ff1.w &&
                    L.ff1.b && L.ff2.w && L.ff2.b && L.ln2_g && L.ln2_b,
                "missing parameters of encoder.layers." + std::to_string(i));
  }
  cudaStream_t st = (cudaStream_t)stream;
  LaunchCtx lc;
  lc.stream = st;
  auto own = [&](float** p, size_t floats) -> int {
    UA2_CHECK_CUDA(cudaMalloc((void**)p, floats * sizeof(float)));
    h->owned.push_back(*p);
    return UA2_OK;
  };
  size_t zmax = std::max(D, F);
  for (int i = 0; i < nc; ++i) zmax = std::max(zmax, (size_t)c.conv_dim[i]);
  if (!h->ready) {
    for (int i = 1; i < nc; ++i) RUN(own(&h->conv_wr[i], (size_t)c.conv_dim[i] * c.conv_dim[i - 1] * c.conv_kernel[i]));
    RUN(own(&h->pos_wr, D * (D / c.num_conv_pos_embedding_groups) * c.num_conv_pos_embeddings));
    RUN(own(&h->zeros, zmax));
    for (WlLayer& L : h->layers) {
      RUN(own(&L.wqkv, 3 * D * D));
      RUN(own(&L.bqkv, 3 * D));
    }
  }
  UA2_CHECK_CUDA(cudaMemsetAsync(h->zeros, 0, zmax * sizeof(float), st));
  for (int i = 1; i < nc; ++i) {
    const long long n = (long long)c.conv_dim[i] * c.conv_dim[i - 1] * c.conv_kernel[i];
    CU(launch(lc, wl_repack_conv_kernel, dim3(wl_grid(n)), dim3(256), 0, h->conv[i].w, h->conv_wr[i], c.conv_dim[i], c.conv_dim[i - 1], c.conv_kernel[i]));
  }
  {
    const int cg = (int)(D / c.num_conv_pos_embedding_groups);
    const long long n = (long long)D * cg * c.num_conv_pos_embeddings;
    CU(launch(lc, wl_repack_posconv_kernel, dim3(wl_grid(n)), dim3(256), 0, h->pos.w, h->pos_wr, (int)D, cg, c.num_conv_pos_embeddings));
  }
  for (WlLayer& L : h->layers) {
    const size_t wb = D * D * sizeof(float), bb = D * sizeof(float);
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.wqkv, L.q.w, wb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.wqkv + D * D, L.k.w, wb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.wqkv + 2 * D * D, L.v.w, wb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.bqkv, L.q.b, bb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.bqkv + D, L.k.b, bb, cudaMemcpyDeviceToDevice, st));
    UA2_CHECK_CUDA(cudaMemcpyAsync(L.bqkv + 2 * D, L.v.b, bb, cudaMemcpyDeviceToDevice, st));
  }
  for (auto& kv : h->w16) cudaFree(kv.second);  // weights may have changed: bf16 copies are rebuilt at next use
  h->w16.clear();
  h->tab_T = 0;
  h->ready = true;
  return UA2_OK;
}

long long ua2_wavlm_frames(ua2_wavlm* h, long long L) {
  if (!h) return -1;
  long long Ts[WL_MAX_CONV];
  if (!wl_frames(h->cfg, L, Ts)) return 0;
  return Ts[h->cfg.num_feat_extract_layers - 1];
}

int ua2_wavlm_forward(ua2_wavlm* h, const float* wav16, long long ld, int B, int L, int hs_lo, int hs_hi, float* out, float* all_hidden, void* stream) {
  UA2_REQUIRE(h && wav16 && out, "null argument");
  UA2_REQUIRE(h->ready, "ua2_wavlm_finalize has not run");
  UA2_REQUIRE(B >= 1 && B <= 65535 && L >= 1 && ld >= L, "bad batch / clip length / row stride");
  UA2_REQUIRE(hs_lo >= 0 && hs_hi > hs_lo && hs_hi <= h->cfg.num_hidden_layers + 1, "hidden-state range outside [0, num_hidden_layers]");
  UA2_REQUIRE(!h->opt_bf16 || h->cfg.hidden_size / h->cfg.num_attention_heads == 64, "bf16 mode serves head size 64");
  RUN(wl_reserve(h, B, L));
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.pdl = true;
  int launches = 0;
  lc.launch_counter = &launches;
  const int rc = wl_forward(h, lc, wav16, ld, B, L, hs_lo, hs_hi, out, all_hidden);
  h->last_launches = launches;
  return rc;
}

int ua2_wavlm_set_option(ua2_wavlm* h, const char* name, int value) {
  UA2_REQUIRE(h && name, "null argument");
  const std::string n(name);
  if (n == "bf16") {
#ifdef UA2_CPU_SHIM
    UA2_REQUIRE(!value, "bf16 mode needs the tensor cores");
#endif
    h->opt_bf16 = value ? 1 : 0;
    return UA2_OK;
  }
  UA2_REQUIRE(false, "unknown option " + n);
}

int ua2_wavlm_last_launch_count(ua2_wavlm* h) { return h ? h->last_launches : 0; }

int ua2_wavlm_rel_bucket_table(int T, int num_buckets, int max_distance, int32_t* out_host) {
  UA2_REQUIRE(out_host && T >= 1, "null argument");
  UA2_REQUIRE(num_buckets >= 4 && num_buckets % 4 == 0 && max_distance > num_buckets / 4, "bad bucket geometry");
  for (int r = 0; r < 2 * T - 1; ++r) out_host[r] = wl_rel_bucket(r - (T - 1), num_buckets, max_distance);
  return UA2_OK;
}

int ua2_wavlm_ops_f32(int op, const float* a, const float* b, const float* c, const float* d, float* y, int i0, int i1, int i2, int i3, int i4, void* stream) {
  // stand-alone launches of the encoder's own kernels for operator-level parity tests (tests/test_zz_wavlm_gpu.py):
  //   op 0  positional convolution: a = h (B = i0, T = i1, D = i2), b = weight (D, cg = i3, K = i4) torch layout, c = bias (D) -> y (B, T, D)
  //   op 1  gate: a = h (B = i0, T = i1, H = i2, hs = i3), b = gru weight (8, hs), c = gru bias (8), d = const (H) -> y (B, H, T)
  //   op 2  biased attention: a = q (B T, H hs), b = k, c = v (B, H, T, hs), d = [gate (B H T) | tab (H (2T - 1))] -> y (B T, H hs); B = i0, T = i1, H = i2, hs = i3
  UA2_REQUIRE(a && b && c && y, "null argument");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  if (op == 0) {
    const int B = i0, T = i1, D = i2, cg = i3, K = i4;
    UA2_REQUIRE(B >= 1 && T >= 1 && cg >= 1 && cg <= PC_MAXCG && D % cg == 0 && K >= 2 && K <= PC_MAXK && K % 2 == 0, "bad positional-convolution geometry");
    float* wr = nullptr;
    UA2_CHECK_CUDA(cudaMalloc((void**)&wr, (size_t)D * cg * K * sizeof(float)));
    cudaError_t e = launch(lc, wl_repack_posconv_kernel, dim3(wl_grid((long long)D * cg * K)), dim3(256), 0, b, wr, D, cg, K);
    if (e == cudaSuccess) e = launch_wl_posconv(lc, a, wr, c, y, B, T, D, cg, K);
    if (e == cudaSuccess) e = cudaStreamSynchronize(lc.stream);
    cudaFree(wr);
    UA2_CHECK_CUDA(e);
    return UA2_OK;
  }
  if (op == 1) {
    UA2_REQUIRE(d != nullptr && i0 >= 1 && i1 >= 1 && i2 >= 1 && i3 >= 1, "bad gate geometry");
    UA2_CHECK_CUDA(launch_wl_gate(lc, a, b, c, d, y, i0, i1, i2, i3));
    return UA2_OK;
  }
  if (op == 2) {
    UA2_REQUIRE(d != nullptr && i0 >= 1 && i1 >= 1 && i2 >= 1 && (i3 == 32 || i3 == 64 || i3 == 128), "bad attention geometry");
    UA2_CHECK_CUDA(launch_dense_attn_bias_f32(lc, a, b, c, y, i0, i1, i2, i3, d, d + (size_t)i0 * i2 * i1));
    return UA2_OK;
  }
#ifndef UA2_CPU_SHIM
  if (op == 3) {  // op 2 on the tensor cores: a / b / c = q / k / v as (B, H, T, 64) BF16 (device pointers passed as float*), d as in op 2
    UA2_REQUIRE(d != nullptr && i0 >= 1 && i1 >= 1 && i2 >= 1 && i3 == 64, "bad attention geometry (head size 64)");
    UA2_CHECK_CUDA(launch_flash_bf16_bias(lc, a, b, c, y, nullptr, i0, i1, i2, i3, d, d + (size_t)i0 * i2 * i1));
    return UA2_OK;
  }
#endif
  UA2_REQUIRE(false, "unknown op");
}

}  // extern "C"
