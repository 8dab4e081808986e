// Flow-matching decoder of ReasoningCodec_film: the DiT estimator (Transformer1DModel, adaLN-single) and the Euler solver
// with classifier-free guidance (BASECFM.solve_euler) - SURVEY.md section 8(f) rank 1, the "tokens -> latent" cost of
// `--stage all` (32 layers x 24 heads x 64, ~1.9 TFLOP per Euler step for a 20 s window).
//
// Replaces (paths relative to tools/tokenizer/ReasoningCodec_film/models/):
//   transformer_1d_flow.py  Transformer1DModel.forward :284-386, ProjectLayer :19-34, PixArtAlphaCombinedFlowEmbeddings
//                           :37-84, AdaLayerNormSingleFlow :87-117
//   attention.py            BasicTransformerBlock.forward :284-418 (ada_norm_single branch), FeedForward :623-681
//   AudioDiffusion1D.py     BASECFM.solve_euler :89-129
//   diffusers (un-vendored, >= 0.25): Attention / AttnProcessor2_0, GELU(tanh), TimestepEmbedding, SinusoidalPositionalEmbedding
//
// Roofline: TENSOR bound - every linear has M = B*T >= 1000 rows at production size and runs on the tcgen05 3xTF32 path
// of ua2_tcgemm.cu (fp32-class accuracy; the reference autocasts these linears to bf16); the kernels in this file are the
// HBM-bound glue around the GEMMs, one pass over the activations each:
//   dit_im2col3_kernel   k = 3 'same' convolution of ProjectLayer as a GEMM over [x[t-1] | x[t] | x[t+1]]
//   dit_epilogue_kernel  bias (+ scale | + positional table | + GELU-tanh | + SiLU | gate * . + residual | q/k/v head split)
//   dit_ln_mod_kernel    LayerNorm without affine, then * (1 + scale) + shift with the adaLN-single table + timestep rows
//   dit_attn_kernel      unmasked self-attention, CTA = 32 query rows x one head; K/V tiles of 32 keys staged in shared
//                        memory and shared by all rows; online softmax per row (one warp = 4 rows), fp32 FMA
//   dit_euler_*          in-context blend, CFG batch assembly, guidance mix and Euler update of the solver
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include <cuda_bf16.h>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"
#include "ua2_umma.cuh"

namespace ua2 {
namespace {

// fp32 -> bf16 (round to nearest even), 4 elements per thread; n must be a multiple of 4 (K % 8 == 0 on this path)
__global__ void dit_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<__nv_bfloat162*>(y)[2 * i] = a;
    reinterpret_cast<__nv_bfloat162*>(y)[2 * i + 1] = b;
  }
}

// ------------------------------------------------------------------------------------------------ conv k3 as GEMM input
// out[m, k*C + c] = x[b, t + k - 1, c] (0 outside the sequence), m = b*T + t
__global__ void dit_im2col3_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int T, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = (long long)B * T * 3 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int k = (int)((i / C) % 3);
    const long long m = i / (3LL * C);
    const int t = (int)(m % T);
    const int ts = t + k - 1;
    out[i] = (ts >= 0 && ts < T) ? x[(m + (k - 1)) * C + c] : 0.f;
  }
}

// Conv1d weight (Cout, Cin, 3) -> GEMM weight (Cout, 3*Cin) with column k*Cin + ci
__global__ void dit_repack_conv3_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin) {
  const long long n = (long long)Cout * Cin * 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int k = (int)((i / Cin) % 3);
    const long long co = i / (3LL * Cin);
    out[i] = w[(co * Cin + ci) * 3 + k];
  }
}

// ------------------------------------------------------------------------------------------------ fused epilogues
enum : int {
  DE_BIAS = 0,        // y = y + bias
  DE_BIAS_SCALE = 1,  // y = (y + bias) * s                      ProjectLayer: ffn_1 then * kernel_size ** -0.5
  DE_BIAS_PE = 2,     // y = (y + bias) + pe[t]                  proj_in.ffn_2 then SinusoidalPositionalEmbedding
  DE_BIAS_GELU = 3,   // y = gelu_tanh(y + bias)                 FeedForward 'gelu-approximate'
  DE_BIAS_SILU = 4,   // y = silu(y + bias)                      TimestepEmbedding.act
  DE_BIAS_KEEP_SILU = 5,  // y = y + bias, y2 = silu(y)          embedded_timestep and the input of adaln_single.linear
  DE_GATE_RES = 6,    // res = gate_b * (y + bias) + res         attention.py:350-353, :409-412
  DE_QKV_SPLIT = 7,   // q (M, D) = y[:, :D] + b; k, v -> (B, H, T, hs)
  DE_QKV_SPLIT_BF16 = 8  // q, k, v -> (B, H, T, hs) bf16: operands of the tensor-core attention (ua2_flash.cu)
};

struct DitEpi {
  const float* src;   // (M, N) raw GEMM output (the tensor-core path's product buffer, or y itself)
  float* y;           // (M, N) destination of the element-wise modes
  const float* bias;  // (N)
  int M, N, T;
  float s;
  const float* pe;    // (positions, N)
  float* y2;          // KEEP_SILU: silu copy; GATE_RES: the residual stream (M, N), updated in place
  const float* table; // GATE_RES: scale_shift_table (6, N)
  const float* t6;    // GATE_RES: timestep modulation (B, 6N)
  int gate_idx;
  float *q, *k, *v;   // QKV_SPLIT destinations
  __nv_bfloat16 *q16, *k16, *v16;  // QKV_SPLIT_BF16 destinations
  int H, hs;
  // fused bf16 block path (dit_epi4_kernel): the GEMM's stream-K side slots are summed here (no separate fix-up pass), and the
  // result goes out as bf16 when the only consumer is the next tensor-core linear
#ifndef UA2_CPU_SHIM
  const float* slots;
  UmmaPlan pl;
  int has_split;
  __nv_bfloat16* y16;
#endif
};

__device__ __forceinline__ float gelu_tanh(float x) {  // F.gelu(approximate='tanh')
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;
  const float inner = kBeta * (x + kKappa * x * x * x);
  return 0.5f * x * (1.f + tanhf(inner));
}
__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }

template <int MODE>
__global__ void dit_epilogue_kernel(const DitEpi e) {
  pdl_launch_dependents();
  pdl_wait();
  const long long n = (long long)e.M * e.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % e.N);
    const long long m = i / e.N;
    const float v = __fadd_rn(e.src[i], e.bias[c]);
    if (MODE == DE_BIAS) {
      e.y[i] = v;
    } else if (MODE == DE_BIAS_SCALE) {
      e.y[i] = __fmul_rn(v, e.s);
    } else if (MODE == DE_BIAS_PE) {
      e.y[i] = __fadd_rn(v, e.pe[(size_t)(m % e.T) * e.N + c]);
    } else if (MODE == DE_BIAS_GELU) {
      e.y[i] = gelu_tanh(v);
    } else if (MODE == DE_BIAS_SILU) {
      e.y[i] = silu(v);
    } else if (MODE == DE_BIAS_KEEP_SILU) {
      e.y[i] = v;
      e.y2[i] = silu(v);
    } else if (MODE == DE_GATE_RES) {
      const int b = (int)(m / e.T);
      const float gate = __fadd_rn(e.table[(size_t)e.gate_idx * e.N + c], e.t6[((size_t)b * 6 + e.gate_idx) * e.N + c]);
      e.y2[i] = __fadd_rn(__fmul_rn(gate, v), e.y2[i]);
    } else {  // DE_QKV_SPLIT: N = 3*D, columns ordered [q | k | v], each (h d)
      const int D = e.N / 3;
      const int part = c / D, cc = c - part * D;
      if (part == 0) {
        e.q[m * D + cc] = v;
      } else {
        const int hh = cc / e.hs, d = cc - hh * e.hs;
        const int b = (int)(m / e.T), t = (int)(m % e.T);
        (part == 1 ? e.k : e.v)[(((size_t)b * e.H + hh) * e.T + t) * e.hs + d] = v;
      }
    }
  }
}

template <int MODE>
cudaError_t launch_epi(const LaunchCtx& lc, const DitEpi& e) {
  const long long n = (long long)e.M * e.N;
  const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32);
  return launch(lc, dit_epilogue_kernel<MODE>, dim3(grid), dim3(256), 0, e);
}

#ifndef UA2_CPU_SHIM  // (tensor-core path only: not part of the CPU shim's build)
// The block's epilogues on the fused bf16 path, 4 columns per thread, one row per blockIdx.x: raw product (+ the continuation CTAs'
// side slots) + bias, then  DE_QKV_SPLIT_BF16: q / k / v (B, H, T, hs) bf16 | DE_BIAS_GELU: gelu_tanh -> y16 (M, N) bf16 |
// DE_GATE_RES: y2 = gate_b * v + y2 (fp32 residual stream).
template <int MODE>
__global__ void __launch_bounds__(256) dit_epi4_kernel(const DitEpi e) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const int c = (blockIdx.y * 256 + threadIdx.x) * 4;
  if (c >= e.N) return;
  float4 v = *reinterpret_cast<const float4*>(e.src + (size_t)m * e.N + c);
  if (e.has_split) {
    const float4 sd = umma_side_sum4(e.pl, e.slots, m, c, e.N);
    v.x += sd.x;
    v.y += sd.y;
    v.z += sd.z;
    v.w += sd.w;
  }
  const float4 bs = *reinterpret_cast<const float4*>(e.bias + c);
  v.x = __fadd_rn(v.x, bs.x);
  v.y = __fadd_rn(v.y, bs.y);
  v.z = __fadd_rn(v.z, bs.z);
  v.w = __fadd_rn(v.w, bs.w);
  if (MODE == DE_BIAS_GELU) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(gelu_tanh(v.x), gelu_tanh(v.y)), b2 = __floats2bfloat162_rn(gelu_tanh(v.z), gelu_tanh(v.w));
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b2);
    *reinterpret_cast<uint2*>(e.y16 + (size_t)m * e.N + c) = o;
  } else if (MODE == DE_GATE_RES) {
    const int b = m / e.T;
    const float4 tb = *reinterpret_cast<const float4*>(e.table + (size_t)e.gate_idx * e.N + c);
    const float4 t6 = *reinterpret_cast<const float4*>(e.t6 + ((size_t)b * 6 + e.gate_idx) * e.N + c);
    float4* rp = reinterpret_cast<float4*>(e.y2 + (size_t)m * e.N + c);
    float4 r = *rp;
    r.x = __fadd_rn(__fmul_rn(__fadd_rn(tb.x, t6.x), v.x), r.x);
    r.y = __fadd_rn(__fmul_rn(__fadd_rn(tb.y, t6.y), v.y), r.y);
    r.z = __fadd_rn(__fmul_rn(__fadd_rn(tb.z, t6.z), v.z), r.z);
    r.w = __fadd_rn(__fmul_rn(__fadd_rn(tb.w, t6.w), v.w), r.w);
    *rp = r;
  } else {  // DE_QKV_SPLIT_BF16: the 4 columns lie inside one head (hs % 4 == 0)
    const int D = e.N / 3;
    const int part = c / D, cc = c - part * D;
    const int hh = cc / e.hs, d = cc - hh * e.hs;
    const int b = m / e.T, t = m - b * e.T;
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b2 = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b2);
    *reinterpret_cast<uint2*>((part == 0 ? e.q16 : part == 1 ? e.k16 : e.v16) + (((size_t)b * e.H + hh) * e.T + t) * e.hs + d) = o;
  }
}

// DE_GATE_RES of one full row per CTA (thread = 4 columns, blockDim = N / 4 rounded up to a warp) followed by the NEXT LayerNorm +
// modulation of that row (attention.py:399-401 after :350-353, or the next block's :312-317 after :409-412): the residual stream is
// updated in fp32, and the normalised, modulated row goes out as bf16, the next linear's operand.  Saves the dit_ln_mod launch and
// its re-read of the row.  ln_table: the (6, N) table whose rows shift_idx / scale_idx modulate the LayerNorm output.
__global__ void __launch_bounds__(1024) dit_gate_res_ln_kernel(const DitEpi e, const float* __restrict__ ln_table, int shift_idx, int scale_idx,
                                                               float eps) {
  __shared__ float red[32];
  __shared__ float stat[2];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int c = tid * 4;
  const bool on = c < e.N;
  const int b = m / e.T;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) {
    float4 v = *reinterpret_cast<const float4*>(e.src + (size_t)m * e.N + c);
    if (e.has_split) {
      const float4 sd = umma_side_sum4(e.pl, e.slots, m, c, e.N);
      v.x += sd.x;
      v.y += sd.y;
      v.z += sd.z;
      v.w += sd.w;
    }
    const float4 bs = *reinterpret_cast<const float4*>(e.bias + c);
    const float4 tb = *reinterpret_cast<const float4*>(e.table + (size_t)e.gate_idx * e.N + c);
    const float4 t6 = *reinterpret_cast<const float4*>(e.t6 + ((size_t)b * 6 + e.gate_idx) * e.N + c);
    float4* rp = reinterpret_cast<float4*>(e.y2 + (size_t)m * e.N + c);
    r = *rp;
    r.x = __fadd_rn(__fmul_rn(__fadd_rn(tb.x, t6.x), __fadd_rn(v.x, bs.x)), r.x);
    r.y = __fadd_rn(__fmul_rn(__fadd_rn(tb.y, t6.y), __fadd_rn(v.y, bs.y)), r.y);
    r.z = __fadd_rn(__fmul_rn(__fadd_rn(tb.z, t6.z), __fadd_rn(v.z, bs.z)), r.z);
    r.w = __fadd_rn(__fmul_rn(__fadd_rn(tb.w, t6.w), __fadd_rn(v.w, bs.w)), r.w);
    *rp = r;
  }
  float s = warp_sum(on ? (r.x + r.y) + (r.z + r.w) : 0.f);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    float t = warp_sum(lane < nw ? red[lane] : 0.f);
    if (lane == 0) stat[0] = t / (float)e.N;
  }
  __syncthreads();
  const float mean = stat[0];
  const float dx = r.x - mean, dy = r.y - mean, dz = r.z - mean, dw = r.w - mean;
  float q = warp_sum(on ? (dx * dx + dy * dy) + (dz * dz + dw * dw) : 0.f);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  if (warp == 0) {
    float t = warp_sum(lane < nw ? red[lane] : 0.f);
    if (lane == 0) stat[1] = rsqrtf(t / (float)e.N + eps);
  }
  __syncthreads();
  if (!on) return;
  const float rstd = stat[1];
  const float4 sc_t = *reinterpret_cast<const float4*>(ln_table + (size_t)scale_idx * e.N + c);
  const float4 sc_b = *reinterpret_cast<const float4*>(e.t6 + ((size_t)b * 6 + scale_idx) * e.N + c);
  const float4 sh_t = *reinterpret_cast<const float4*>(ln_table + (size_t)shift_idx * e.N + c);
  const float4 sh_b = *reinterpret_cast<const float4*>(e.t6 + ((size_t)b * 6 + shift_idx) * e.N + c);
  const float o0 = __fadd_rn(__fmul_rn(dx * rstd, __fadd_rn(1.f, __fadd_rn(sc_t.x, sc_b.x))), __fadd_rn(sh_t.x, sh_b.x));
  const float o1 = __fadd_rn(__fmul_rn(dy * rstd, __fadd_rn(1.f, __fadd_rn(sc_t.y, sc_b.y))), __fadd_rn(sh_t.y, sh_b.y));
  const float o2 = __fadd_rn(__fmul_rn(dz * rstd, __fadd_rn(1.f, __fadd_rn(sc_t.z, sc_b.z))), __fadd_rn(sh_t.z, sh_b.z));
  const float o3 = __fadd_rn(__fmul_rn(dw * rstd, __fadd_rn(1.f, __fadd_rn(sc_t.w, sc_b.w))), __fadd_rn(sh_t.w, sh_b.w));
  const __nv_bfloat162 a = __floats2bfloat162_rn(o0, o1), b2 = __floats2bfloat162_rn(o2, o3);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&a);
  o.y = *reinterpret_cast<const uint32_t*>(&b2);
  *reinterpret_cast<uint2*>(e.y16 + (size_t)m * e.N + c) = o;
}

template <int MODE>
cudaError_t launch_epi4(const LaunchCtx& lc, const DitEpi& e) {
  return launch(lc, dit_epi4_kernel<MODE>, dim3(e.M, (e.N + 1023) / 1024), dim3(256), 0, e);
}
#endif

// ------------------------------------------------------------------------------------------------ LayerNorm + modulation
// out[m] = LN(x[m]; eps, no affine) * (1 + scale_b) + shift_b,  scale_b = table[scale_idx] + t[b, scale_idx * t_stride ...]
// (attention.py:312-317 / :399-401 with t = the 6*D rows of adaln_single; transformer_1d_flow.py:378-381 with t = the
// embedded timestep, t_stride = 0: both rows of the (2, D) table get the same D-vector added)
__global__ void __launch_bounds__(256) dit_ln_mod_kernel(const float* __restrict__ x, float* __restrict__ out, __nv_bfloat16* __restrict__ out16,
                                                         const float* __restrict__ table, const float* __restrict__ t, int t_row,
                                                         int t_stride, int shift_idx, int scale_idx, float eps, int T, int D) {
  __shared__ float red[8];
  __shared__ float stat[2];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, b = m / T;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* xr = x + (size_t)m * D;
  float s = 0.f;
  for (int c = tid; c < D; c += 256) s += xr[c];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    stat[0] = tot / (float)D;
  }
  __syncthreads();
  const float mean = stat[0];
  float q = 0.f;
  for (int c = tid; c < D; c += 256) {
    const float d = xr[c] - mean;
    q += d * d;
  }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    stat[1] = rsqrtf(tot / (float)D + eps);
  }
  __syncthreads();
  const float rstd = stat[1];
  const float* tb = t + (size_t)b * t_row;
  for (int c = tid; c < D; c += 256) {
    const float scale = __fadd_rn(table[(size_t)scale_idx * D + c], tb[(size_t)scale_idx * t_stride + c]);
    const float shift = __fadd_rn(table[(size_t)shift_idx * D + c], tb[(size_t)shift_idx * t_stride + c]);
    const float nrm = (xr[c] - mean) * rstd;
    const float o = __fadd_rn(__fmul_rn(nrm, __fadd_rn(1.f, scale)), shift);
    if (out16 != nullptr) {  // fused bf16 path: the only consumer is the next tensor-core linear
      out16[(size_t)m * D + c] = __float2bfloat16_rn(o);
    } else {
      out[(size_t)m * D + c] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
constexpr int DA_KEYS = 32;  // keys per shared-memory tile (one per lane in the score phase)

// CTA = 8 warps x RPW query rows of one (batch, head); K / V tiles of 32 keys are staged in shared memory once and shared by
// all rows.  Score phase: lane = key, K row read with 128-bit loads (row stride HS + 4 keeps them conflict-free), the RPW
// query rows of the warp are broadcast reads - (1 + RPW) loads per 4 * RPW FMAs.  Online softmax per row; the probabilities
// go through a per-warp shared tile so that the P @ V phase (lane = output dim) reads them as broadcast float4.
// BIAS (WavLM's gated relative position bias, transformers modeling_wavlm.py WavLMAttention.forward): the score of (query i, key j)
// gets gate[b, h, i] * tab[h, j - i + T - 1] added after the 1 / sqrt(hs) scaling - the (B H, T, T) additive attn_mask of
// F.multi_head_attention_forward, never materialised.
template <int HS, int RPW, bool BIAS>
__global__ void __launch_bounds__(256) dit_attn_kernel(const float* __restrict__ q, const float* __restrict__ kc,
                                                       const float* __restrict__ vc, float* __restrict__ out, int T, int H,
                                                       const float* __restrict__ gate, const float* __restrict__ tab) {
  constexpr int DPL = HS / 32;  // output dims per lane
  constexpr int ROWS = 8 * RPW;
  constexpr int KST = HS + 4;
  __shared__ __align__(16) float Qs[ROWS][HS];
  __shared__ __align__(16) float Ks[DA_KEYS][KST];
  __shared__ __align__(16) float Vs[DA_KEYS][HS];
  __shared__ __align__(16) float Ps[8][RPW][DA_KEYS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * ROWS, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HS;
  pdl_launch_dependents();
  pdl_wait();
  for (int i = tid; i < ROWS * HS / 4; i += 256) {
    const int r = i / (HS / 4), d4 = i - r * (HS / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t0 + r < T) v = *reinterpret_cast<const float4*>(q + ((size_t)b * T + t0 + r) * D + h * HS + d4 * 4);
    *reinterpret_cast<float4*>(&Qs[r][d4 * 4]) = v;
  }
  const float* Kb = kc + ((size_t)b * H + h) * (size_t)T * HS;
  const float* Vb = vc + ((size_t)b * H + h) * (size_t)T * HS;
  const float scale = rsqrtf((float)HS);
  float mx[RPW], l[RPW], acc[RPW][DPL];
  float gt[RPW];                 // BIAS: the row's gate
  const float* tb[RPW];          // BIAS: tab[h] shifted so that tb[r][j] is the entry of key j for this row
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    gt[r] = 0.f;
    tb[r] = nullptr;
    if (BIAS) {
      const int t = t0 + warp * RPW + r;
      if (t < T) {
        gt[r] = gate[((size_t)b * H + h) * T + t];
        tb[r] = tab + (size_t)h * (2 * T - 1) + (T - 1 - t);
      }
    }
    mx[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[r][d] = 0.f;
  }
  for (int j0 = 0; j0 < T; j0 += DA_KEYS) {
    __syncthreads();  // previous tile fully consumed (also orders the Qs fill before the first use)
    for (int i = tid; i < DA_KEYS * HS / 4; i += 256) {
      const int j = i / (HS / 4), d4 = i - j * (HS / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j0 + j < T) {
        kv = *reinterpret_cast<const float4*>(Kb + (size_t)(j0 + j) * HS + d4 * 4);
        vv = *reinterpret_cast<const float4*>(Vb + (size_t)(j0 + j) * HS + d4 * 4);
      }
      *reinterpret_cast<float4*>(&Ks[j][d4 * 4]) = kv;
      *reinterpret_cast<float4*>(&Vs[j][d4 * 4]) = vv;
    }
    __syncthreads();
    // ---- scores of this lane's key against the warp's rows
    float s[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) s[r] = 0.f;
#pragma unroll 4
    for (int d4 = 0; d4 < HS / 4; ++d4) {
      const float4 k4 = *reinterpret_cast<const float4*>(&Ks[lane][d4 * 4]);
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const float4 q4 = *reinterpret_cast<const float4*>(&Qs[warp * RPW + r][d4 * 4]);
        s[r] = fmaf(q4.x, k4.x, s[r]);
        s[r] = fmaf(q4.y, k4.y, s[r]);
        s[r] = fmaf(q4.z, k4.z, s[r]);
        s[r] = fmaf(q4.w, k4.w, s[r]);
      }
    }
    const bool valid = j0 + lane < T;
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      float sv = valid ? s[r] * scale : -INFINITY;
      if (BIAS) {
        if (valid && tb[r] != nullptr) sv = fmaf(gt[r], tb[r][j0 + lane], sv);
      }
      const float mn = fmaxf(mx[r], warp_max(sv));  // every tile holds at least one valid key, so mn is finite
      const float corr = expf(mx[r] - mn);
      const float p = valid ? expf(sv - mn) : 0.f;
      l[r] = l[r] * corr + warp_sum(p);
#pragma unroll
      for (int d = 0; d < DPL; ++d) acc[r][d] *= corr;
      mx[r] = mn;
      Ps[warp][r][lane] = p;
    }
    __syncwarp();
    // ---- P @ V: lane owns output dims lane + 32 * d
#pragma unroll 2
    for (int j4 = 0; j4 < DA_KEYS / 4; ++j4) {
      float4 p4[RPW];
#pragma unroll
      for (int r = 0; r < RPW; ++r) p4[r] = *reinterpret_cast<const float4*>(&Ps[warp][r][j4 * 4]);
#pragma unroll
      for (int d = 0; d < DPL; ++d) {
        const float v0 = Vs[j4 * 4 + 0][lane + 32 * d], v1 = Vs[j4 * 4 + 1][lane + 32 * d];
        const float v2 = Vs[j4 * 4 + 2][lane + 32 * d], v3 = Vs[j4 * 4 + 3][lane + 32 * d];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          acc[r][d] = fmaf(p4[r].x, v0, acc[r][d]);
          acc[r][d] = fmaf(p4[r].y, v1, acc[r][d]);
          acc[r][d] = fmaf(p4[r].z, v2, acc[r][d]);
          acc[r][d] = fmaf(p4[r].w, v3, acc[r][d]);
        }
      }
    }
    __syncwarp();  // Ps is rewritten by the next tile
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int t = t0 + warp * RPW + r;
    if (t < T) {
#pragma unroll
      for (int d = 0; d < DPL; ++d) out[((size_t)b * T + t) * D + h * HS + lane + 32 * d] = acc[r][d] / l[r];
    }
  }
}

cudaError_t launch_dit_attn(const LaunchCtx& lc, const float* q, const float* kc, const float* vc, float* out, int B, int T, int H,
                            int hs) {
  const dim3 block(256);
  const float* none = nullptr;
  switch (hs) {  // 4 rows per warp (32 per CTA); 2 at head size 128 to stay inside 48 KB of static shared memory
    case 32: return launch(lc, dit_attn_kernel<32, 4, false>, dim3((T + 31) / 32, H, B), block, 0, q, kc, vc, out, T, H, none, none);
    case 64: return launch(lc, dit_attn_kernel<64, 4, false>, dim3((T + 31) / 32, H, B), block, 0, q, kc, vc, out, T, H, none, none);
    case 128: return launch(lc, dit_attn_kernel<128, 2, false>, dim3((T + 15) / 16, H, B), block, 0, q, kc, vc, out, T, H, none, none);
    default: return cudaErrorInvalidValue;
  }
}

// the same attention with gate[b, h, i] * tab[h, j - i + T - 1] added to the scaled scores; gate (B, H, T), tab (H, 2 T - 1)
cudaError_t launch_dit_attn_bias(const LaunchCtx& lc, const float* q, const float* kc, const float* vc, float* out, int B, int T, int H,
                                 int hs, const float* gate, const float* tab) {
  const dim3 block(256);
  if (gate == nullptr || tab == nullptr) return cudaErrorInvalidValue;
  switch (hs) {
    case 32: return launch(lc, dit_attn_kernel<32, 4, true>, dim3((T + 31) / 32, H, B), block, 0, q, kc, vc, out, T, H, gate, tab);
    case 64: return launch(lc, dit_attn_kernel<64, 4, true>, dim3((T + 31) / 32, H, B), block, 0, q, kc, vc, out, T, H, gate, tab);
    case 128: return launch(lc, dit_attn_kernel<128, 2, true>, dim3((T + 15) / 16, H, B), block, 0, q, kc, vc, out, T, H, gate, tab);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------ timestep embedding
// proj[b] = [cos(args) | sin(args)], args = t[b] * freqs * 1000 (transformer_1d_flow.py:58-71); t from device memory or by value
__global__ void dit_tproj_kernel(const float* __restrict__ t_dev, float t_val, const float* __restrict__ freqs,
                                 float* __restrict__ proj, int B, int half) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float t = t_dev ? t_dev[b] : t_val;
  const float a = __fmul_rn(__fmul_rn(t, freqs[k]), 1000.f);
  proj[(size_t)b * 2 * half + k] = cosf(a);
  proj[(size_t)b * 2 * half + half + k] = sinf(a);
}

// ------------------------------------------------------------------------------------------------ Euler solver glue
// x[:, :ic] = (1 - (1 - sigma_min) * t) * noise[:, :ic] + t * incontext[:, :ic]   (AudioDiffusion1D.py:106), then the CFG
// batch of the estimator: rows [x | incontext | 0] and [x | incontext | mu]        (:108-113)
__global__ void dit_euler_pack_kernel(float* __restrict__ x, const float* __restrict__ noise, const float* __restrict__ incontext,
                                      const float* __restrict__ mu, float* __restrict__ inp, int T, int lat, int cond, int ic,
                                      float t, float one_minus_sigma) {
  pdl_launch_dependents();
  pdl_wait();
  const int in_ch = 2 * lat + cond;
  const long long n = (long long)T * in_ch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % in_ch), tt = (int)(i / in_ch);
    float u, cnd;
    if (c < lat) {
      float xv = x[(size_t)tt * lat + c];
      if (tt < ic) {
        const float c1 = __fsub_rn(1.f, __fmul_rn(one_minus_sigma, t));
        xv = __fadd_rn(__fmul_rn(c1, noise[(size_t)tt * lat + c]), __fmul_rn(t, incontext[(size_t)tt * lat + c]));
        x[(size_t)tt * lat + c] = xv;
      }
      u = cnd = xv;
    } else if (c < 2 * lat) {
      u = cnd = incontext[(size_t)tt * lat + (c - lat)];
    } else {
      u = 0.f;
      cnd = mu[(size_t)tt * cond + (c - 2 * lat)];
    }
    inp[i] = u;
    inp[n + i] = cnd;
  }
}

// dphi = uncond + g * (cond - uncond); x = x + dt * dphi   (AudioDiffusion1D.py:116-123)
__global__ void dit_euler_update_kernel(float* __restrict__ x, const float* __restrict__ d, long long n, float g, float dt) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float u = d[i], c = d[n + i];
    const float dphi = __fadd_rn(u, __fmul_rn(g, __fsub_rn(c, u)));
    x[i] = __fadd_rn(x[i], __fmul_rn(dt, dphi));
  }
}

unsigned grid_for(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32); }

struct Lin {
  const float *w = nullptr, *b = nullptr;
};
struct DitBlock {
  const float* table = nullptr;  // (6, D)
  Lin q, k, v, o, ff1, ff2;
  float *wqkv = nullptr, *bqkv = nullptr;  // concatenated [to_q; to_k; to_v]
};

}  // namespace

// unmasked softmax(q k^T / sqrt(hs)) v in fp32 for other handles of the library (ua2_enc.cu): q (B * T, H * hs), k / v (B, H, T, hs)
cudaError_t launch_dense_attn_f32(const LaunchCtx& lc, const float* q, const float* kc, const float* vc, float* out, int B, int T, int H, int hs) {
  return launch_dit_attn(lc, q, kc, vc, out, B, T, H, hs);
}
// the same with the additive bias gate[b, h, i] * tab[h, j - i + T - 1] (ua2_wavlm.cu): gate (B, H, T), tab (H, 2 T - 1)
cudaError_t launch_dense_attn_bias_f32(const LaunchCtx& lc, const float* q, const float* kc, const float* vc, float* out, int B, int T, int H, int hs,
                                       const float* gate, const float* tab) {
  return launch_dit_attn_bias(lc, q, kc, vc, out, B, T, H, hs, gate, tab);
}
}  // namespace ua2

using namespace ua2;

struct ua2_dit {
  ua2_dit_cfg cfg{};
  std::vector<DitBlock> blocks;
  const float *table = nullptr, *pe = nullptr, *tfreqs = nullptr;
  Lin in1, in2, out1, out2, te1, te2, ada;
  float *in1_w = nullptr, *out1_w = nullptr;  // conv weights repacked (Cout, 3*Cin)
  bool ready = false;
  std::vector<void*> owned;
  // workspace for M rows / B batch rows
  size_t rows = 0, brows = 0;
  float *col = nullptr, *h = nullptr, *n = nullptr, *qkv = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr,
        *ff = nullptr, *stats = nullptr;
  float *tproj = nullptr, *temb = nullptr, *tsemb = nullptr, *t6 = nullptr, *ttmp = nullptr;
  // solver
  size_t srows = 0;
  float *noise = nullptr, *sinp = nullptr, *sout = nullptr;
  TcWorkspace tc;
  // option "bf16": linears of >= 32 rows on bf16 operands with fp32 accumulation (the reference's autocast arithmetic)
  int opt_bf16 = 0;
  int opt_flash = 1;  // with bf16: tensor-core attention (0 = keep the fp32 SIMT attention; measurement switch)
  __nv_bfloat16* a16 = nullptr;  // (M, kmax) activations converted per call
  std::map<const float*, __nv_bfloat16*> w16;  // weights converted once, keyed by the fp32 tensor
  int last_launches = 0;
};

namespace {

#define RUN(expr)                  \
  do {                             \
    int _rc = (expr);              \
    if (_rc != UA2_OK) return _rc; \
  } while (0)
#define CU(expr)                                                                   \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
      return UA2_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

void free_list(std::initializer_list<float**> ps) {
  for (float** p : ps) {
    if (*p) cudaFree(*p);
    *p = nullptr;
  }
}

int dmalloc(float** p, size_t floats) {
  UA2_CHECK_CUDA(cudaMalloc((void**)p, std::max<size_t>(floats, 4) * sizeof(float)));
  return UA2_OK;
}

int reserve(ua2_dit* h, size_t M, size_t B) {
  const ua2_dit_cfg& c = h->cfg;
  const size_t D = (size_t)c.num_attention_heads * c.attention_head_dim, I = c.in_channels, O = c.out_channels;
  if (M > h->rows) {
    if (h->rows) UA2_CHECK_CUDA(cudaDeviceSynchronize());
    free_list({&h->col, &h->h, &h->n, &h->qkv, &h->q, &h->k, &h->v, &h->att, &h->ff, &h->stats, &h->tc.a, &h->tc.slots, &h->tc.c});
    RUN(dmalloc(&h->col, M * 3 * std::max(I, D)));
    RUN(dmalloc(&h->h, M * D));
    RUN(dmalloc(&h->n, M * D));
    RUN(dmalloc(&h->qkv, M * 3 * D));
    RUN(dmalloc(&h->q, M * D));
    RUN(dmalloc(&h->k, M * D));
    RUN(dmalloc(&h->v, M * D));
    RUN(dmalloc(&h->att, M * D));
    RUN(dmalloc(&h->ff, M * 4 * D));
    RUN(dmalloc(&h->stats, 2 * M + 8));
    if (tc_gemm_available()) {  // scratch of the tcgen05 3xTF32 path: split activations / weights, raw product
      const size_t kmax = std::max({3 * I, 4 * D, 3 * D}), nmax = std::max({4 * D, 3 * D, O});
      h->tc.a_floats = M * 2 * kmax;
      h->tc.slots_floats = tc_slots_max_floats();
      h->tc.c_floats = M * nmax;
      RUN(dmalloc(&h->tc.a, h->tc.a_floats));
      RUN(dmalloc(&h->tc.slots, h->tc.slots_floats));
      RUN(dmalloc(&h->tc.c, h->tc.c_floats));
      if (h->a16) cudaFree(h->a16);
      h->a16 = nullptr;
      UA2_CHECK_CUDA(cudaMalloc((void**)&h->a16, M * kmax * sizeof(__nv_bfloat16)));
    }
    h->rows = M;
  }
  if (B > h->brows) {
    if (h->brows) UA2_CHECK_CUDA(cudaDeviceSynchronize());
    free_list({&h->tproj, &h->temb, &h->tsemb, &h->t6, &h->ttmp});
    RUN(dmalloc(&h->tproj, B * c.flow_t_size));
    RUN(dmalloc(&h->temb, B * D));
    RUN(dmalloc(&h->tsemb, B * D));
    RUN(dmalloc(&h->ttmp, B * D));
    RUN(dmalloc(&h->t6, B * 6 * D));
    h->brows = B;
  }
  return UA2_OK;
}

// the bf16 copy of a weight, made at first use (2 B / parameter)
// *fresh: the copy was made by a kernel just launched on this stream.  The GEMM's weight producer starts its TMA loads BEFORE
// griddepcontrol.wait (weights are normally static), so the launch that follows must not be a programmatic dependent launch.
int weight16(ua2_dit* h, const LaunchCtx& lc, const float* W, int N, int K, const __nv_bfloat16** out, bool* fresh) {
  auto it = h->w16.find(W);
  *fresh = it == h->w16.end();
  if (it == h->w16.end()) {
    __nv_bfloat16* wb = nullptr;
    UA2_CHECK_CUDA(cudaMalloc((void**)&wb, (size_t)N * K * sizeof(__nv_bfloat16)));
    it = h->w16.emplace(W, wb).first;
    const long long n4 = (long long)N * K / 4;
    CU(launch(lc, dit_to_bf16_kernel, dim3(grid_for(n4)), dim3(256), 0, W, wb, n4));
  }
  *out = it->second;
  return UA2_OK;
}

// fused bf16 path: a16 (M, K) bf16 @ W^T on the tcgen05 kind::f16 mainloop -> raw product in the handle's buffer; the epilogue
// description comes back with the product, the side slots and the plan filled in (dit_epi4_kernel sums the slots itself)
int linear16(ua2_dit* h, const LaunchCtx& lc, const float* W, const float* bias, int M, int N, int K, int T, DitEpi* e) {
  const __nv_bfloat16* w16 = nullptr;
  bool fresh = false;
  RUN(weight16(h, lc, W, N, K, &w16, &fresh));
  const UmmaPlan pl = umma_plan(M, N, 1, K, true);
  UA2_REQUIRE(pl.slot_floats <= h->tc.slots_floats && (size_t)M * N <= h->tc.c_floats && (K % 8) == 0 && (N % 4) == 0, "flow decoder linear outside the tensor-core path's shapes");
  LaunchCtx lg = lc;
  if (fresh) lg.pdl = false;  // full stream order behind the conversion kernel (see weight16)
  CU(run_umma_bf16(lg, h->a16, w16, h->tc.c, N, h->tc.slots, M, N, K, pl));
  e->src = h->tc.c;
  e->bias = bias;
  e->M = M;
  e->N = N;
  e->T = T;
  e->slots = h->tc.slots;
  e->pl = pl;
  e->has_split = umma_has_split_tiles(pl) ? 1 : 0;
  return UA2_OK;
}

// x (M, K, row stride K) @ W^T, raw (no bias): tensor cores for M >= tc_min_rows (the product then stays in the path's own
// buffer, *src points at it and `y` is not written), skinny fp32 kernels below (product in y, *src = y)
int linear_raw(ua2_dit* h, const LaunchCtx& lc, const float* x, const float* W, float* y, int M, int N, int K, const float** src) {
  if (h->opt_bf16 && h->a16 != nullptr && M >= 32 && (K % 8) == 0 && (N % 4) == 0 && (size_t)M * N <= h->tc.c_floats) {
    const __nv_bfloat16* w16 = nullptr;
    bool fresh = false;
    RUN(weight16(h, lc, W, N, K, &w16, &fresh));
    const long long n4 = (long long)M * K / 4;
    LaunchCtx lg = lc;
    if (fresh) lg.pdl = false;  // full stream order behind the weight conversion (see weight16); the GEMM then follows this kernel
    CU(launch(lg, dit_to_bf16_kernel, dim3(grid_for(n4)), dim3(256), 0, x, h->a16, n4));
    // hand-written tcgen05 kind::f16 mainloop (ua2_umma.cu, bf16 mode): fp32 accumulation in TMEM, product in the handle's buffer
    const UmmaPlan pl = umma_plan(M, N, 1, K, true);
    cudaError_t e = pl.slot_floats <= h->tc.slots_floats ? run_umma_bf16(lc, h->a16, w16, h->tc.c, N, h->tc.slots, M, N, K, pl) : cudaErrorNotSupported;
    if (e == cudaSuccess) e = run_umma_fixup(lc, h->tc.c, N, h->tc.slots, M, N, 1, pl);
    if (e == cudaSuccess) {
      *src = h->tc.c;
      return UA2_OK;
    }
    if (e != cudaErrorNotSupported) CU(e);
  }
  GemvParams p;
  p.W = W;
  p.N = N;
  p.K = K;
  p.M = M;
  p.X = x;
  p.ldx = K;
  p.Y = y;
  p.ldy = N;
  p.ws = h->stats;
  p.ws_floats = 2 * h->rows + 8;
  p.tc = h->tc.a ? &h->tc : nullptr;
  const float* raw = nullptr;
  p.raw_out = &raw;
  CU(launch_gemv(lc, PRO_PLAIN, EPI_STORE, p));
  *src = raw ? raw : y;
  return UA2_OK;
}

DitEpi epi(const float* src, float* y, const float* bias, int M, int N, int T) {
  DitEpi e{};
  e.src = src;
  e.y = y;
  e.bias = bias;
  e.M = M;
  e.N = N;
  e.T = T;
  return e;
}

// ProjectLayer.forward: (B, T, Cin) -> (B, T, Cout); `pe` != nullptr adds the positional table after ffn_2 (proj_in)
int project_layer(ua2_dit* h, const LaunchCtx& lc, const float* x, const float* w1_repacked, const Lin& l1, const Lin& l2,
                  float* tmp, float* y, int B, int T, int Cin, int Cout, const float* pe) {
  const int M = B * T;
  CU(launch(lc, dit_im2col3_kernel, dim3(grid_for((long long)M * 3 * Cin)), dim3(256), 0, x, h->col, B, T, Cin));
  const float* src = nullptr;
  RUN(linear_raw(h, lc, h->col, w1_repacked, tmp, M, Cout, 3 * Cin, &src));
  DitEpi e1 = epi(src, tmp, l1.b, M, Cout, T);
  e1.s = 0.57735026918962576451f;  // 3 ** -0.5 rounded to fp32, transformer_1d_flow.py:32
  CU(launch_epi<DE_BIAS_SCALE>(lc, e1));
  RUN(linear_raw(h, lc, tmp, l2.w, y, M, Cout, Cout, &src));
  DitEpi e2 = epi(src, y, l2.b, M, Cout, T);
  if (pe) {
    e2.pe = pe;
    CU(launch_epi<DE_BIAS_PE>(lc, e2));
  } else {
    CU(launch_epi<DE_BIAS>(lc, e2));
  }
  return UA2_OK;
}

// Transformer1DModel.forward on device buffers: x (B, T, in) -> out (B, T, out_channels); timestep from t_dev (B) or t_val
int dit_forward(ua2_dit* h, const LaunchCtx& lc, const float* x, const float* t_dev, float t_val, float* out, int B, int T) {
  const ua2_dit_cfg& c = h->cfg;
  const int H = c.num_attention_heads, hs = c.attention_head_dim, D = H * hs, M = B * T;
  // ---- adaln_single: timestep -> embedded_timestep (temb) and the 6*D modulation rows (t6)
  const int half = c.flow_t_size / 2;
  CU(launch(lc, dit_tproj_kernel, dim3((B * half + 255) / 256), dim3(256), 0, t_dev, t_val, h->tfreqs, h->tproj, B, half));
  const float* src = nullptr;
  RUN(linear_raw(h, lc, h->tproj, h->te1.w, h->ttmp, B, D, c.flow_t_size, &src));
  CU(launch_epi<DE_BIAS_SILU>(lc, epi(src, h->ttmp, h->te1.b, B, D, 1)));
  RUN(linear_raw(h, lc, h->ttmp, h->te2.w, h->temb, B, D, D, &src));
  {
    DitEpi e = epi(src, h->temb, h->te2.b, B, D, 1);
    e.y2 = h->tsemb;
    CU(launch_epi<DE_BIAS_KEEP_SILU>(lc, e));
  }
  RUN(linear_raw(h, lc, h->tsemb, h->ada.w, h->t6, B, 6 * D, D, &src));
  CU(launch_epi<DE_BIAS>(lc, epi(src, h->t6, h->ada.b, B, 6 * D, 1)));
  // ---- proj_in + positional embedding
  RUN(project_layer(h, lc, x, h->in1_w, h->in1, h->in2, h->n, h->h, B, T, c.in_channels, D, h->pe));
  // ---- blocks
  const bool fused16 = h->opt_bf16 && h->opt_flash && hs == 64 && tc_gemm_available() && h->a16 != nullptr && M >= 32 && (D % 8) == 0;
  const bool fuse_ln = fused16 && D <= 4096;  // residual update + the next LayerNorm in one kernel (one CTA per row, 4 columns per thread)
  for (size_t li = 0; li < h->blocks.size(); ++li) {
    const DitBlock& bl = h->blocks[li];
    if (fused16) {
      // bf16 mode (the reference's autocast arithmetic), 9 launches per block: every linear and both contractions of the attention on
      // tcgen05; activations travel between them as bf16, written by the producing kernel; the residual stream h stays fp32
      if (li == 0 || !fuse_ln)
        CU(launch(lc, dit_ln_mod_kernel, dim3(M), dim3(256), 0, (const float*)h->h, (float*)nullptr, h->a16, bl.table, (const float*)h->t6,
                  6 * D, D, 0, 1, c.norm_eps, T, D));
      DitEpi e{};
      RUN(linear16(h, lc, bl.wqkv, bl.bqkv, M, 3 * D, D, T, &e));
      e.q16 = reinterpret_cast<__nv_bfloat16*>(h->q);
      e.k16 = reinterpret_cast<__nv_bfloat16*>(h->k);
      e.v16 = reinterpret_cast<__nv_bfloat16*>(h->v);
      e.H = H;
      e.hs = hs;
      CU(launch_epi4<DE_QKV_SPLIT_BF16>(lc, e));
      CU(launch_flash_bf16(lc, e.q16, e.k16, e.v16, nullptr, h->a16, B, T, H, hs));
      for (int half = 0; half < 2; ++half) {
        if (half == 1) {
          if (!fuse_ln)
            CU(launch(lc, dit_ln_mod_kernel, dim3(M), dim3(256), 0, (const float*)h->h, (float*)nullptr, h->a16, bl.table,
                      (const float*)h->t6, 6 * D, D, 3, 4, c.norm_eps, T, D));
          DitEpi f{};
          RUN(linear16(h, lc, bl.ff1.w, bl.ff1.b, M, 4 * D, D, T, &f));
          f.y16 = h->a16;  // the product is in the handle's fp32 buffer: the operand buffer is free again
          CU(launch_epi4<DE_BIAS_GELU>(lc, f));
        }
        DitEpi g{};
        RUN(linear16(h, lc, half ? bl.ff2.w : bl.o.w, half ? bl.ff2.b : bl.o.b, M, D, half ? 4 * D : D, T, &g));
        g.y2 = h->h;
        g.table = bl.table;
        g.t6 = h->t6;
        g.gate_idx = half ? 5 : 2;
        const bool last = half == 1 && li + 1 == h->blocks.size();  // norm_out follows: its own table, timestep vector and eps
        if (fuse_ln && !last) {
          g.y16 = h->a16;
          const float* ln_table = half ? h->blocks[li + 1].table : bl.table;
          CU(launch(lc, dit_gate_res_ln_kernel, dim3(M), dim3(((D / 4) + 31) / 32 * 32), 0, g, ln_table, half ? 0 : 3, half ? 1 : 4,
                    c.norm_eps));
        } else {
          CU(launch_epi4<DE_GATE_RES>(lc, g));
        }
      }
      continue;
    }
    CU(launch(lc, dit_ln_mod_kernel, dim3(M), dim3(256), 0, (const float*)h->h, h->n, (__nv_bfloat16*)nullptr, bl.table, (const float*)h->t6, 6 * D, D, 0, 1,
              c.norm_eps, T, D));
    RUN(linear_raw(h, lc, h->n, bl.wqkv, h->qkv, M, 3 * D, D, &src));
    {
      DitEpi e = epi(src, h->qkv, bl.bqkv, M, 3 * D, T);
      e.H = H;
      e.hs = hs;
      e.q = h->q;
      e.k = h->k;
      e.v = h->v;
      CU(launch_epi<DE_QKV_SPLIT>(lc, e));
    }
    CU(launch_dit_attn(lc, h->q, h->k, h->v, h->att, B, T, H, hs));
    RUN(linear_raw(h, lc, h->att, bl.o.w, h->n, M, D, D, &src));
    {
      DitEpi e = epi(src, h->n, bl.o.b, M, D, T);
      e.y2 = h->h;
      e.table = bl.table;
      e.t6 = h->t6;
      e.gate_idx = 2;
      CU(launch_epi<DE_GATE_RES>(lc, e));
    }
    CU(launch(lc, dit_ln_mod_kernel, dim3(M), dim3(256), 0, (const float*)h->h, h->n, (__nv_bfloat16*)nullptr, bl.table, (const float*)h->t6, 6 * D, D, 3, 4,
              c.norm_eps, T, D));
    RUN(linear_raw(h, lc, h->n, bl.ff1.w, h->ff, M, 4 * D, D, &src));
    CU(launch_epi<DE_BIAS_GELU>(lc, epi(src, h->ff, bl.ff1.b, M, 4 * D, T)));
    RUN(linear_raw(h, lc, h->ff, bl.ff2.w, h->n, M, D, 4 * D, &src));
    {
      DitEpi e = epi(src, h->n, bl.ff2.b, M, D, T);
      e.y2 = h->h;
      e.table = bl.table;
      e.t6 = h->t6;
      e.gate_idx = 5;
      CU(launch_epi<DE_GATE_RES>(lc, e));
    }
  }
  // ---- norm_out + modulation with (scale_shift_table + embedded_timestep) (eps 1e-6 hard-wired, :234), proj_out
  CU(launch(lc, dit_ln_mod_kernel, dim3(M), dim3(256), 0, (const float*)h->h, h->n, (__nv_bfloat16*)nullptr, h->table, (const float*)h->temb, D, 0, 0, 1, 1e-6f,
            T, D));
  RUN(project_layer(h, lc, h->n, h->out1_w, h->out1, h->out2, h->att, out, B, T, D, c.out_channels, nullptr));
  return UA2_OK;
}

bool parse_block_key(const std::string& key, int& idx, std::string& rest) {
  const std::string pre = "transformer_blocks.";
  if (key.compare(0, pre.size(), pre) != 0) return false;
  size_t i = pre.size(), j = i;
  while (j < key.size() && key[j] >= '0' && key[j] <= '9') ++j;
  if (j == i || j >= key.size() || key[j] != '.') return false;
  idx = std::stoi(key.substr(i, j - i));
  rest = key.substr(j + 1);
  return true;
}

}  // namespace

extern "C" {

int ua2_dit_create(const ua2_dit_cfg* cfg, ua2_dit** out) {
  UA2_REQUIRE(cfg && out, "null argument");
  const ua2_dit_cfg& c = *cfg;
  UA2_REQUIRE(c.num_attention_heads >= 1 && c.num_layers >= 1 && c.in_channels >= 4 && c.out_channels >= 2, "bad dimensions");
  UA2_REQUIRE(c.attention_head_dim == 32 || c.attention_head_dim == 64 || c.attention_head_dim == 128,
              "attention_head_dim must be 32 / 64 / 128");
  UA2_REQUIRE(c.in_channels % 4 == 0 && c.out_channels % 4 == 0, "in_channels and out_channels must be multiples of 4");
  UA2_REQUIRE(c.flow_t_size >= 8 && c.flow_t_size % 8 == 0, "flow_t_size must be a multiple of 8");
  UA2_REQUIRE(c.num_positional_embeddings >= 1, "num_positional_embeddings must be >= 1");
  ua2_dit* h = new ua2_dit();
  h->cfg = c;
  h->blocks.resize(c.num_layers);
  *out = h;
  return UA2_OK;
}

int ua2_dit_destroy(ua2_dit* h) {
  if (!h) return UA2_OK;
  cudaDeviceSynchronize();
  free_list({&h->col, &h->h, &h->n, &h->qkv, &h->q, &h->k, &h->v, &h->att, &h->ff, &h->stats, &h->tc.a, &h->tc.slots, &h->tc.c,
             &h->tproj, &h->temb, &h->tsemb, &h->t6, &h->ttmp, &h->noise, &h->sinp, &h->sout});
  for (void* p : h->owned) cudaFree(p);
  for (auto& kv : h->w16) cudaFree(kv.second);
  if (h->a16) cudaFree(h->a16);
  delete h;
  return UA2_OK;
}

int ua2_dit_load_weight(ua2_dit* h, const char* key_c, const float* dptr, const int64_t* shape, int ndim) {
  UA2_REQUIRE(h && key_c && dptr && shape && ndim >= 1, "null argument");
  const std::string key(key_c);
  const ua2_dit_cfg& c = h->cfg;
  const int64_t D = (int64_t)c.num_attention_heads * c.attention_head_dim, I = c.in_channels, O = c.out_channels;
  auto is = [&](std::initializer_list<int64_t> want) {
    if ((int)want.size() != ndim) return false;
    int i = 0;
    for (int64_t w : want)
      if (shape[i++] != w) return false;
    return true;
  };
#define WANT(cond) UA2_REQUIRE(cond, key + ": shape mismatch")
  struct Top {
    const char* name;
    const float** dst;
    std::initializer_list<int64_t> shp;
  };
  const Top tops[] = {
      {"scale_shift_table", &h->table, {2, D}},
      {"tfreqs", &h->tfreqs, {c.flow_t_size / 2}},
      {"proj_in.ffn_1.weight", &h->in1.w, {D, I, 3}},
      {"proj_in.ffn_1.bias", &h->in1.b, {D}},
      {"proj_in.ffn_2.weight", &h->in2.w, {D, D}},
      {"proj_in.ffn_2.bias", &h->in2.b, {D}},
      {"proj_out.ffn_1.weight", &h->out1.w, {O, D, 3}},
      {"proj_out.ffn_1.bias", &h->out1.b, {O}},
      {"proj_out.ffn_2.weight", &h->out2.w, {O, O}},
      {"proj_out.ffn_2.bias", &h->out2.b, {O}},
      {"adaln_single.emb.timestep_embedder.linear_1.weight", &h->te1.w, {D, c.flow_t_size}},
      {"adaln_single.emb.timestep_embedder.linear_1.bias", &h->te1.b, {D}},
      {"adaln_single.emb.timestep_embedder.linear_2.weight", &h->te2.w, {D, D}},
      {"adaln_single.emb.timestep_embedder.linear_2.bias", &h->te2.b, {D}},
      {"adaln_single.linear.weight", &h->ada.w, {6 * D, D}},
      {"adaln_single.linear.bias", &h->ada.b, {6 * D}},
  };
  for (const Top& t : tops)
    if (key == t.name) {
      WANT(is(t.shp));
      *t.dst = dptr;
      return UA2_OK;
    }
  if (key == "pos_embed.pe") {
    WANT(is({1, c.num_positional_embeddings, D}));
    h->pe = dptr;
    return UA2_OK;
  }
  int bi = -1;
  std::string rest;
  UA2_REQUIRE(parse_block_key(key, bi, rest) && bi >= 0 && bi < c.num_layers, "unexpected key " + key);
  DitBlock& b = h->blocks[bi];
  struct Blk {
    const char* name;
    const float** dst;
    std::initializer_list<int64_t> shp;
  };
  const Blk blks[] = {
      {"scale_shift_table", &b.table, {6, D}},
      {"attn1.to_q.weight", &b.q.w, {D, D}},
      {"attn1.to_q.bias", &b.q.b, {D}},
      {"attn1.to_k.weight", &b.k.w, {D, D}},
      {"attn1.to_k.bias", &b.k.b, {D}},
      {"attn1.to_v.weight", &b.v.w, {D, D}},
      {"attn1.to_v.bias", &b.v.b, {D}},
      {"attn1.to_out.0.weight", &b.o.w, {D, D}},
      {"attn1.to_out.0.bias", &b.o.b, {D}},
      {"ff.net.0.proj.weight", &b.ff1.w, {4 * D, D}},
      {"ff.net.0.proj.bias", &b.ff1.b, {4 * D}},
      {"ff.net.2.weight", &b.ff2.w, {D, 4 * D}},
      {"ff.net.2.bias", &b.ff2.b, {D}},
  };
  for (const Blk& t : blks)
    if (rest == t.name) {
      WANT(is(t.shp));
      *t.dst = dptr;
      return UA2_OK;
    }
#undef WANT
  UA2_REQUIRE(false, "unexpected key " + key);
}

int ua2_dit_finalize(ua2_dit* h, void* stream) {
  UA2_REQUIRE(h, "null handle");
  const ua2_dit_cfg& c = h->cfg;
  const size_t D = (size_t)c.num_attention_heads * c.attention_head_dim, I = c.in_channels, O = c.out_channels;
  UA2_REQUIRE(h->table && h->pe && h->tfreqs && h->in1.w && h->in1.b && h->in2.w && h->in2.b && h->out1.w && h->out1.b &&
                  h->out2.w && h->out2.b && h->te1.w && h->te1.b && h->te2.w && h->te2.b && h->ada.w && h->ada.b,
              "missing top-level parameters (scale_shift_table, pos_embed.pe, tfreqs, proj_in/out.*, adaln_single.*)");
  for (int i = 0; i < c.num_layers; ++i) {
    const DitBlock& b = h->blocks[i];
    UA2_REQUIRE(b.table && b.q.w && b.q.b && b.k.w && b.k.b && b.v.w && b.v.b && b.o.w && b.o.b && b.ff1.w && b.ff1.b && b.ff2.w &&
                    b.ff2.b,
                "missing parameters of transformer_blocks." + std::to_string(i));
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
    RUN(own(&h->in1_w, D * 3 * I));
    RUN(own(&h->out1_w, O * 3 * D));
    for (DitBlock& b : h->blocks) {
      RUN(own(&b.wqkv, 3 * D * D));
      RUN(own(&b.bqkv, 3 * D));
    }
  }
  UA2_CHECK_CUDA(launch(lc, dit_repack_conv3_kernel, dim3(grid_for((long long)D * 3 * I)), dim3(256), 0, h->in1.w, h->in1_w, (int)D, (int)I));
  UA2_CHECK_CUDA(launch(lc, dit_repack_conv3_kernel, dim3(grid_for((long long)O * 3 * D)), dim3(256), 0, h->out1.w, h->out1_w, (int)O, (int)D));
  for (DitBlock& b : h->blocks) {
    const Lin* src[3] = {&b.q, &b.k, &b.v};
    for (int j = 0; j < 3; ++j) {
      UA2_CHECK_CUDA(cudaMemcpyAsync(b.wqkv + (size_t)j * D * D, src[j]->w, D * D * 4, cudaMemcpyDeviceToDevice, st));
      UA2_CHECK_CUDA(cudaMemcpyAsync(b.bqkv + (size_t)j * D, src[j]->b, D * 4, cudaMemcpyDeviceToDevice, st));
    }
  }
  h->ready = true;
  return UA2_OK;
}

int ua2_dit_forward(ua2_dit* h, const float* hidden_states, const float* timestep, float* out, int B, int T, void* stream) {
  UA2_REQUIRE(h && h->ready, "handle not finalized");
  UA2_REQUIRE(hidden_states && timestep && out, "null argument");
  UA2_REQUIRE(B >= 1 && T >= 1 && B <= 65535, "need B, T >= 1");
  UA2_REQUIRE(T <= h->cfg.num_positional_embeddings, "sequence longer than the positional table");
  RUN(reserve(h, (size_t)B * T, (size_t)B));
  int launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.launch_counter = &launches;
  RUN(dit_forward(h, lc, hidden_states, timestep, 0.f, out, B, T));
  h->last_launches = launches;
  return UA2_OK;
}

int ua2_dit_solve_euler(ua2_dit* h, float* x, const float* incontext_x, int incontext_length, const float* t_span, int n_span,
                        const float* mu, int T, float guidance_scale, float sigma_min, void* stream) {
  UA2_REQUIRE(h && h->ready, "handle not finalized");
  UA2_REQUIRE(x && incontext_x && t_span && mu, "null argument");
  UA2_REQUIRE(n_span >= 2 && T >= 1, "need at least one step and one frame");
  UA2_REQUIRE(incontext_length >= 0 && incontext_length <= T, "incontext_length out of range");
  UA2_REQUIRE(T <= h->cfg.num_positional_embeddings, "sequence longer than the positional table");
  // the reference's other branch concatenates along time (AudioDiffusion1D.py:119) and cannot run this estimator
  UA2_REQUIRE(guidance_scale > 1.0f, "solve_euler is served with classifier-free guidance (guidance_scale > 1) only");
  const ua2_dit_cfg& c = h->cfg;
  const int lat = c.out_channels, cond = c.in_channels - 2 * lat;
  UA2_REQUIRE(cond > 0, "in_channels must exceed 2 * out_channels (latent | in-context latent | condition)");
  RUN(reserve(h, (size_t)2 * T, 2));
  if ((size_t)T > h->srows) {
    if (h->srows) UA2_CHECK_CUDA(cudaDeviceSynchronize());
    free_list({&h->noise, &h->sinp, &h->sout});
    RUN(dmalloc(&h->noise, (size_t)T * lat));
    RUN(dmalloc(&h->sinp, (size_t)2 * T * c.in_channels));
    RUN(dmalloc(&h->sout, (size_t)2 * T * lat));
    h->srows = T;
  }
  int launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.launch_counter = &launches;
  const long long n_lat = (long long)T * lat;
  UA2_CHECK_CUDA(cudaMemcpyAsync(h->noise, x, n_lat * sizeof(float), cudaMemcpyDeviceToDevice, lc.stream));  // noise = x.clone()
  // fp32 scalar arithmetic of the reference's 0-dim tensors (AudioDiffusion1D.py:98, :122-126), one rounding per operation
  float t = t_span[0], dt = t_span[1] - t_span[0];
  const float one_minus_sigma = (float)(1.0 - (double)sigma_min);
  for (int step = 1; step < n_span; ++step) {
    UA2_CHECK_CUDA(launch(lc, dit_euler_pack_kernel, dim3(grid_for((long long)T * c.in_channels)), dim3(256), 0, x,
                          (const float*)h->noise, incontext_x, mu, h->sinp, T, lat, cond, incontext_length, t, one_minus_sigma));
    RUN(dit_forward(h, lc, h->sinp, nullptr, t, h->sout, 2, T));
    UA2_CHECK_CUDA(launch(lc, dit_euler_update_kernel, dim3(grid_for(n_lat)), dim3(256), 0, x, (const float*)h->sout, n_lat,
                          guidance_scale, dt));
    t = t + dt;
    if (step < n_span - 1) dt = t_span[step + 1] - t;
  }
  h->last_launches = launches;
  return UA2_OK;
}

int ua2_dit_set_option(ua2_dit* h, const char* name, int value) {
  UA2_REQUIRE(h && name, "null argument");
  const std::string n(name);
  if (n == "bf16") {
    h->opt_bf16 = value ? 1 : 0;
    return UA2_OK;
  }
  if (n == "flash_attn") {
    h->opt_flash = value ? 1 : 0;
    return UA2_OK;
  }
  UA2_REQUIRE(false, "unknown option " + n);
}

int ua2_dit_last_launch_count(ua2_dit* h) { return h ? h->last_launches : 0; }

}  // extern "C"
