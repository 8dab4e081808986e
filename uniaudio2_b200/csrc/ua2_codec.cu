// Codec kernels: causal / strided 1-D convolutions of the SEANet stacks and the residual-VQ search.
//
// Replaces (paths relative to tools/tokenizer/MimiCodec/model/, byte-identical twins of llm_modules/*):
//   modules/conv.py:232-254   StreamingConv1d.forward (causal left pad k_eff - stride, right "extra" pad so the
//                             last window is full, constant or replicate mode) + the nn.ELU placed in front of every
//                             conv by modules/seanet.py (pre-activation) + SEANetResnetBlock's skip add (:92-94)
//   modules/conv.py:306-329   StreamingConvTranspose1d.forward (kernel 2*stride, trim the right k - stride samples)
//   modules/resample.py:68-119 ConvTrUpsample1d (depthwise transposed conv)
//   quantization/core_vq.py:179-185, :365-384 and quantization/vq.py:134-155  residual VQ encode / decode
//
// Layout: activations (B, C, T) fp32 like the reference; conv weights are repacked once by the host side to
// (Cin, K, Cout) so that a (ci, k) slice of output channels is contiguous.
// Rooflines: the convolutions sit at 10-130 FLOP/B, i.e. fp32-SIMT compute bound; the
// kernels below are register-tiled (8 co x 4 t per thread, 64 x 128 output tile per CTA) and read strided inputs
// through a polyphase shared-memory layout so that every LDS is conflict free.
#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {

constexpr int CO_T = 64;   // output channels per CTA
constexpr int T_T = 128;   // output time steps per CTA
constexpr int CI_T = 8;    // input channels per shared-memory stage

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }  // nn.ELU(alpha=1)

struct ConvParams {
  const float* x;     // (B, Cin, T_in)
  const float* w;     // (Cin, K, Cout) repacked
  const float* bias;  // (Cout) or null
  const float* res;   // (B, Cout, T_out) or null: out = res + conv
  float* y;           // (B, Cout, T_out)
  int B, Cin, Cout, T_in, T_out, K, stride, dilation, pad_left;
  int pre_elu;    // apply ELU to the input samples (activation placed before the conv in SEANet)
  int replicate;  // padding mode: 0 constant (zeros), 1 replicate (edge value)
};

// y[b, co, t] = bias[co] + sum_{ci, k} w[ci, k, co] * f(x[b, ci, t*stride + k*dilation - pad_left])
__global__ void __launch_bounds__(256) conv1d_kernel(const ConvParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  const int t0 = blockIdx.x * T_T, co0 = blockIdx.y * CO_T, b = blockIdx.z;
  const int s = p.stride;
  // input positions needed by this tile: [p0, p0 + span)
  const int span = (T_T - 1) * s + (p.K - 1) * p.dilation + 1;
  const int p0 = t0 * s - p.pad_left;
  // polyphase layout: position q (relative to p0) lives at xs[ci][(q % s) * plen + q / s]
  const int plen = (span + s - 1) / s + 1;
  const int xrow = s * plen;
  float* xs = smem;                      // [CI_T][xrow]
  float* ws = smem + CI_T * xrow;        // [CI_T][K][CO_T]
  float acc[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;

  pdl_launch_dependents();
  pdl_wait();
  for (int ci0 = 0; ci0 < p.Cin; ci0 += CI_T) {
    __syncthreads();
    // ---- stage inputs (coalesced along time, activation + padding applied once per sample)
    for (int i = tid; i < CI_T * span; i += 256) {
      const int cc = i / span, q = i - cc * span;
      const int ci = ci0 + cc;
      float v = 0.f;
      if (ci < p.Cin) {
        int pos = p0 + q;
        bool inside = pos >= 0 && pos < p.T_in;
        if (!inside && p.replicate) {
          pos = pos < 0 ? 0 : p.T_in - 1;
          inside = true;
        }
        if (inside) {
          v = p.x[((size_t)b * p.Cin + ci) * p.T_in + pos];
          if (p.pre_elu) v = elu1(v);
        }
      }
      xs[cc * xrow + (q % s) * plen + q / s] = v;
    }
    // ---- stage weights
    for (int i = tid; i < CI_T * p.K * CO_T; i += 256) {
      const int cc = i / (p.K * CO_T), r = i - cc * (p.K * CO_T);
      const int k = r / CO_T, co = r - k * CO_T;
      const int ci = ci0 + cc;
      float v = 0.f;
      if (ci < p.Cin && co0 + co < p.Cout) v = p.w[((size_t)ci * p.K + k) * p.Cout + co0 + co];
      ws[i] = v;
    }
    __syncthreads();
    const int ncc = min(CI_T, p.Cin - ci0);
    for (int cc = 0; cc < ncc; ++cc) {
      for (int k = 0; k < p.K; ++k) {
        const int off = k * p.dilation;
        const float* xr = xs + cc * xrow + (off % s) * plen + off / s;  // element for output t: xr[t_local]
        float xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = xr[lane + 32 * j];
        const float4 w0 = *reinterpret_cast<const float4*>(ws + (cc * p.K + k) * CO_T + wy * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + (cc * p.K + k) * CO_T + wy * 8 + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[c][j] = fmaf(wv[c], xv[j], acc[c][j]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co0 + wy * 8 + c;
    if (co >= p.Cout) continue;
    const float bv = p.bias ? p.bias[co] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + lane + 32 * j;
      if (t < p.T_out) {
        const size_t o = ((size_t)b * p.Cout + co) * p.T_out + t;
        float v = acc[c][j] + bv;
        if (p.res) v = p.res[o] + v;  // SEANetResnetBlock: shortcut(x) + block(x)
        p.y[o] = v;
      }
    }
  }
}

struct ConvTrParams {
  const float* x;     // (B, Cin, T_in)
  const float* w;     // (Cin, K, Cout) repacked from torch's (Cin, Cout, K)
  const float* bias;  // (Cout) or null
  float* y;           // (B, Cout, T_in * stride)   (right k - stride samples already trimmed)
  int B, Cin, Cout, T_in, K, stride;
  int pre_elu;
};

constexpr int J_T = 32;   // input positions per CTA
constexpr int MAX_S = 8;  // strides up to 8 (SEANet ratios)

// kernel = 2 * stride: output t = j*s + ph gets x[j] * w[.., ph] + x[j-1] * w[.., ph + s]
__global__ void __launch_bounds__(256) convtr1d_kernel(const ConvTrParams p) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
  const int j0 = blockIdx.x * J_T, co0 = blockIdx.y * CO_T, b = blockIdx.z;
  const int s = p.stride;
  float* xs = smem;                        // [CI_T][J_T + 1]   (index 0 = position j0 - 1)
  float* ws = smem + CI_T * (J_T + 1);     // [CI_T][K][CO_T]
  float acc[8][MAX_S];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int ph = 0; ph < MAX_S; ++ph) acc[c][ph] = 0.f;
  pdl_launch_dependents();
  pdl_wait();
  for (int ci0 = 0; ci0 < p.Cin; ci0 += CI_T) {
    __syncthreads();
    for (int i = tid; i < CI_T * (J_T + 1); i += 256) {
      const int cc = i / (J_T + 1), q = i - cc * (J_T + 1);
      const int ci = ci0 + cc, pos = j0 - 1 + q;
      float v = 0.f;
      if (ci < p.Cin && pos >= 0 && pos < p.T_in) {
        v = p.x[((size_t)b * p.Cin + ci) * p.T_in + pos];
        if (p.pre_elu) v = elu1(v);
      }
      xs[i] = v;
    }
    for (int i = tid; i < CI_T * p.K * CO_T; i += 256) {
      const int cc = i / (p.K * CO_T), r = i - cc * (p.K * CO_T);
      const int k = r / CO_T, co = r - k * CO_T;
      const int ci = ci0 + cc;
      float v = 0.f;
      if (ci < p.Cin && co0 + co < p.Cout) v = p.w[((size_t)ci * p.K + k) * p.Cout + co0 + co];
      ws[i] = v;
    }
    __syncthreads();
    const int ncc = min(CI_T, p.Cin - ci0);
    for (int cc = 0; cc < ncc; ++cc) {
      const float xc = xs[cc * (J_T + 1) + lane + 1], xp = xs[cc * (J_T + 1) + lane];
#pragma unroll
      for (int ph = 0; ph < MAX_S; ++ph) {
        if (ph < s) {
          const float* wa = ws + (cc * p.K + ph) * CO_T + wy * 8;
          const float* wb = ws + (cc * p.K + ph + s) * CO_T + wy * 8;
          const float4 a0 = *reinterpret_cast<const float4*>(wa), a1 = *reinterpret_cast<const float4*>(wa + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(wb), b1 = *reinterpret_cast<const float4*>(wb + 4);
          const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c][ph] = fmaf(av[c], xc, fmaf(bv[c], xp, acc[c][ph]));
        }
      }
    }
  }
  const int j = j0 + lane;
  if (j < p.T_in) {
    const int T_out = p.T_in * s;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int co = co0 + wy * 8 + c;
      if (co >= p.Cout) continue;
      const float bv = p.bias ? p.bias[co] : 0.f;
      float* dst = p.y + ((size_t)b * p.Cout + co) * T_out + (size_t)j * s;
#pragma unroll
      for (int ph = 0; ph < MAX_S; ++ph)
        if (ph < s) dst[ph] = acc[c][ph] + bv;
    }
  }
}

// depthwise transposed conv (groups = C, kernel 2*stride), modules/resample.py:68-119 with learnt=True, channel_wise=True
__global__ void convtr1d_depthwise_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                                          int B, int C, int T_in, int stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int T_out = T_in * stride;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * C * T_out;
  if (idx >= total) return;
  const int t = (int)(idx % T_out);
  const size_t bc = idx / T_out;
  const int c = (int)(bc % C);
  const int j = t / stride, ph = t - j * stride;
  const float* xr = x + bc * T_in;
  const float xc = xr[j], xp = j > 0 ? xr[j - 1] : 0.f;
  // torch computes the two taps as separate products summed in tap order k = ph, ph + stride
  y[idx] = __fadd_rn(__fmul_rn(w[c * 2 * stride + ph], xc), __fmul_rn(w[c * 2 * stride + ph + stride], xp));
}

// ------------------------------------------------------------------------------------------------ residual VQ
constexpr int FT = 16;   // frames per CTA
constexpr int CT = 64;   // codes per shared-memory tile

struct RvqEncParams {
  const float* x;       // (B, D, T) projected input
  const float* emb;     // (n_q, K, D) codebooks (embedding_sum / clamp(cluster_usage))
  const float* enorm;   // (n_q, K)   squared norms of the codebook rows
  int64_t* codes;       // (B, n_q_total, T): this launch fills rows [q_off, q_off + n_q)
  int B, D, T, K, n_q, n_q_total, q_off;
};

// One CTA keeps FT frames resident and walks all n_q quantizers: argmin_j ||r - e_j||^2 (= argmin ||e_j||^2 - 2 r.e_j),
// r <- r - e_code  (core_vq.py:365-376).  Distances are accumulated in fp32 FMA order d = 0..D-1.
__global__ void __launch_bounds__(256) rvq_encode_kernel(const RvqEncParams p) {
  extern __shared__ __align__(16) float smem[];
  const int D = p.D, DP = D + 4;
  float* r = smem;               // [FT][DP]
  float* cb = smem + FT * DP;    // [CT][DP]
  __shared__ float best_d[FT][4];
  __shared__ int best_i[FT][4];
  __shared__ int sel[FT];
  const int tid = threadIdx.x, lane = tid & 31;
  const int f0 = blockIdx.x * FT;
  const int N = p.B * p.T;
  pdl_launch_dependents();
  pdl_wait();
  // load the frame tile (transposing (B, D, T) -> [frame][d])
  for (int i = tid; i < FT * D; i += 256) {
    const int d = i / FT, f = i - d * FT;
    const int n = f0 + f;
    float v = 0.f;
    if (n < N) {
      const int b = n / p.T, t = n - b * p.T;
      v = p.x[((size_t)b * D + d) * p.T + t];
    }
    r[f * DP + d] = v;
  }
  // thread -> (code slot c = tid % 64, frame group fg = tid / 64 handling frames fg*4 .. fg*4+3)
  const int c = tid & 63, fg = tid >> 6;
  for (int q = 0; q < p.n_q; ++q) {
    const float* E = p.emb + (size_t)q * p.K * D;
    const float* En = p.enorm + (size_t)q * p.K;
    float bd[4];
    int bi[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bd[j] = INFINITY;
      bi[j] = 0;
    }
    for (int c0 = 0; c0 < p.K; c0 += CT) {
      __syncthreads();  // previous tile consumed (and residual update of the previous quantizer visible)
      for (int i = tid * 4; i < CT * D; i += 256 * 4) {
        const int cc = i / D, d = i - cc * D;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + cc < p.K) v = *reinterpret_cast<const float4*>(E + (size_t)(c0 + cc) * D + d);
        *reinterpret_cast<float4*>(cb + cc * DP + d) = v;
      }
      __syncthreads();
      float dot[4] = {0.f, 0.f, 0.f, 0.f};
      const float* cr = cb + c * DP;
      const float* rr = r + (fg * 4) * DP;
      for (int d = 0; d < D; d += 4) {
        const float4 e = *reinterpret_cast<const float4*>(cr + d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 x4 = *reinterpret_cast<const float4*>(rr + j * DP + d);
          dot[j] = fmaf(e.x, x4.x, dot[j]);
          dot[j] = fmaf(e.y, x4.y, dot[j]);
          dot[j] = fmaf(e.z, x4.z, dot[j]);
          dot[j] = fmaf(e.w, x4.w, dot[j]);
        }
      }
      if (c0 + c < p.K) {
        const float en = En[c0 + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float dist = fmaf(-2.f, dot[j], en);
          if (dist < bd[j]) {  // strict <: the lowest index wins ties inside a thread (codes visited in order)
            bd[j] = dist;
            bi[j] = c0 + c;
          }
        }
      }
    }
    // reduce over the 64 code slots of each frame (2 warps per frame group): lowest distance, then lowest index
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float d = bd[j];
      int ix = bi[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
        if (od < d || (od == d && oi < ix)) {
          d = od;
          ix = oi;
        }
      }
      if (lane == 0) {
        best_d[fg * 4 + j][(tid >> 5) & 1] = d;
        best_i[fg * 4 + j][(tid >> 5) & 1] = ix;
      }
    }
    __syncthreads();
    if (tid < FT) {
      float d = best_d[tid][0];
      int ix = best_i[tid][0];
      const float od = best_d[tid][1];
      const int oi = best_i[tid][1];
      if (od < d || (od == d && oi < ix)) {
        d = od;
        ix = oi;
      }
      sel[tid] = ix;
      const int n = f0 + tid;
      if (n < N) {
        const int b = n / p.T, t = n - b * p.T;
        p.codes[((size_t)b * p.n_q_total + p.q_off + q) * p.T + t] = ix;
      }
    }
    __syncthreads();
    // residual update r <- r - E[code]
    for (int i = tid; i < FT * D; i += 256) {
      const int f = i / D, d = i - f * D;
      r[f * DP + d] -= E[(size_t)sel[f] * D + d];
    }
  }
}

// out[b, d, t] = sum_q emb[q][codes[b, q_off + q, t]][d]   (core_vq.py:378-384 before the output projection)
__global__ void __launch_bounds__(256) rvq_decode_kernel(const int64_t* __restrict__ codes, const float* __restrict__ emb,
                                                         float* __restrict__ out, int B, int D, int T, int K, int n_q,
                                                         int n_q_total, int q_off) {
  extern __shared__ float tile[];  // [32 frames][D + 1]
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y, t0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  const int DP = D + 1;
  for (int i = tid; i < 32 * D; i += 256) {
    const int f = i / D, d = i - f * D;
    float acc = 0.f;
    if (t0 + f < T) {
      for (int q = 0; q < n_q; ++q) {
        const int64_t cidx = codes[((size_t)b * n_q_total + q_off + q) * T + t0 + f];
        acc += emb[((size_t)q * K + cidx) * D + d];  // same order as the reference: quantized = quantized + layer.decode
      }
    }
    tile[f * DP + d] = acc;
  }
  __syncthreads();
  for (int i = tid; i < 32 * D; i += 256) {
    const int d = i / 32, f = i - d * 32;
    if (t0 + f < T) out[((size_t)b * D + d) * T + t0 + f] = tile[f * DP + d];
  }
}

// torch ConvTranspose1d weight (Cin, Cout, 2s) -> per-phase GEMM operand (s, Cout, Cin, 2): [ph][n][ci][which] = w[ci][n][ph + which*s]
__global__ void repack_convtr_phase_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cin, int Cout, int s) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)s * Cout * Cin * 2;
  if (i >= total) return;
  const int which = (int)(i & 1);
  const int ci = (int)((i >> 1) % Cin);
  const int n = (int)((i / ((size_t)2 * Cin)) % Cout);
  const int ph = (int)(i / ((size_t)2 * Cin * Cout));
  dst[i] = src[((size_t)ci * Cout + n) * (2 * s) + ph + which * s];
}


// One warp per frame: code = argmin_j (||e_j||^2 - 2 S[m, j]) with S = r E^T from the tiled GEMM, lowest index on ties
// (torch.argmin, core_vq.py:184); then r[m, :] -= E[code, :] (core_vq.py:372) and codes[b, q, t] = code.
__global__ void __launch_bounds__(256) rvq_argmin_update_kernel(const float* __restrict__ S, const float* __restrict__ emb,
                                                                const float* __restrict__ enorm, float* __restrict__ r,
                                                                int64_t* __restrict__ codes, int M, int K, int D, int T, int n_q_total,
                                                                int q_index) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  const float* s = S + (size_t)m * K;
  float bd = INFINITY;
  int bi = 0;
  for (int j = lane; j < K; j += 32) {
    const float dist = fmaf(-2.f, s[j], enorm[j]);
    if (dist < bd) {
      bd = dist;
      bi = j;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(0xffffffffu, bd, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (od < bd || (od == bd && oi < bi)) {
      bd = od;
      bi = oi;
    }
  }
  const float* e = emb + (size_t)bi * D;
  float* rr = r + (size_t)m * D;
  for (int d = lane * 4; d < D; d += 128) {
    float4 v = *reinterpret_cast<float4*>(rr + d);
    const float4 ev = *reinterpret_cast<const float4*>(e + d);
    v.x -= ev.x;
    v.y -= ev.y;
    v.z -= ev.z;
    v.w -= ev.w;
    *reinterpret_cast<float4*>(rr + d) = v;
  }
  if (lane == 0) {
    const int b = m / T, t = m - b * T;
    codes[((size_t)b * n_q_total + q_index) * T + t] = bi;
  }
}


// time_film (AudioDiffusion1D.py:428-438): gamma = 1 + g * tanh(params[..., :C]); beta = params[..., C:]; rows of batches whose
// zero-condition flag is set get gamma = 1, beta = 0 (the reference draws that flag with torch.rand(B,1,1) < 0.2 even at
// inference; the draw is an INPUT here so that results are reproducible); out = gamma * features + beta.
__global__ void film_kernel(const float* __restrict__ params, const float* __restrict__ feat, const uint8_t* __restrict__ zero_mask,
                            float* __restrict__ out, long long M, int C, int T, float g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const long long m = i / C;
  const int c = (int)(i - m * C);
  const int b = (int)(m / T);
  float gamma = 1.0f + g * tanhf(params[m * 2 * C + c]);
  float beta = params[m * 2 * C + C + c];
  const float mk = (zero_mask && zero_mask[b]) ? 1.f : 0.f;
  gamma = gamma * (1.f - mk) + 1.0f * mk;
  beta = beta * (1.f - mk) + 0.0f * mk;
  out[i] = gamma * feat[i] + beta;
}

// F.interpolate(x (B, C, T_in), scale_factor=s, mode='nearest') as used on the reasoning feature (AudioDiffusion1D.py:523):
// T_out = floor(T_in * s), src = min(floor(dst * (1 / s)), T_in - 1)  (torch's legacy nearest with a given scale factor)
__global__ void interp_nearest_kernel(const float* __restrict__ x, float* __restrict__ y, long long BC, int T_in, int T_out, float inv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC * T_out) return;
  const long long r = i / T_out;
  const int t = (int)(i - r * T_out);
  int src = (int)floorf((float)t * inv);
  if (src > T_in - 1) src = T_in - 1;
  y[i] = x[r * T_in + src];
}

// y[m, n] += bias[n]  (nn.Linear bias of the cond_fusion / film heads, AudioDiffusion1D.py:278-284)
__global__ void add_bias_kernel(float* __restrict__ y, const float* __restrict__ bias, long long MN, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < MN) y[i] += bias[i % N];
}

// op 0: round(param * x) / param  (scalar quantiser of SQ-codec, scalar24k.py round_func9; torch.round = half to even)
// op 1: tanh(x)
__global__ void elementwise_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int op, float param) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  y[i] = op == 0 ? rintf(param * v) / param : tanhf(v);
}

}  // namespace
}  // namespace ua2

using namespace ua2;

extern "C" {

int ua2_conv1d_causal_f32(const float* x, const float* w_ckc, const float* bias, const float* residual, float* y, int B,
                          int Cin, int Cout, int T_in, int K, int stride, int dilation, int pre_elu, int replicate_pad,
                          void* stream) {
  UA2_REQUIRE(x && w_ckc && y, "null argument");
  UA2_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && T_in >= 1 && K >= 1 && stride >= 1 && dilation >= 1, "bad shape");
  const int k_eff = (K - 1) * dilation + 1;
  const int padding_total = k_eff - stride;
  UA2_REQUIRE(padding_total >= 0, "kernel smaller than stride is not a causal SEANet conv");
  // modules/conv.py:50-58: right 'extra' padding so that the last window is full  ->  T_out = ceil(T_in / stride)
  const int T_out = (T_in + stride - 1) / stride;
  ConvParams p{x, w_ckc, bias, residual, y, B, Cin, Cout, T_in, T_out, K, stride, dilation, padding_total, pre_elu, replicate_pad};
  const int span = (T_T - 1) * stride + (K - 1) * dilation + 1;
  const int plen = (span + stride - 1) / stride + 1;
  const size_t smem = ((size_t)CI_T * stride * plen + (size_t)CI_T * K * CO_T) * sizeof(float);
  UA2_REQUIRE(smem <= 200 * 1024, "conv tile does not fit shared memory");
  static DeviceOnce once;
  if (once.need()) {
    UA2_CHECK_CUDA(cudaFuncSetAttribute(conv1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const dim3 grid((T_out + T_T - 1) / T_T, (Cout + CO_T - 1) / CO_T, B);
  UA2_CHECK_CUDA(launch(lc, conv1d_kernel, grid, dim3(256), smem, p));
  return UA2_OK;
}

// implicit-GEMM variant: same semantics, weights in the reference's own (Cout, Cin, K) layout, 128-position x 128-channel
// register-tiled fp32 GEMM tiles (ua2_sgemm.cu).  This is the path the codec handle uses.
int ua2_conv1d_causal_gemm_f32(const float* x, const float* w_torch, const float* bias, const float* residual, float* y, int B,
                               int Cin, int Cout, int T_in, int K, int stride, int dilation, int pre_elu, int replicate_pad,
                               void* stream) {
  UA2_REQUIRE(x && w_torch && y, "null argument");
  UA2_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && T_in >= 1 && K >= 1 && stride >= 1 && dilation >= 1, "bad shape");
  const int k_eff = (K - 1) * dilation + 1;
  const int padding_total = k_eff - stride;
  UA2_REQUIRE(padding_total >= 0, "kernel smaller than stride is not a causal SEANet conv");
  const int T_out = (T_in + stride - 1) / stride;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_conv1d_gemm(lc, x, w_torch, bias, residual, y, B, Cin, Cout, T_in, T_out, K, stride, dilation, padding_total,
                                    pre_elu, replicate_pad));
  return UA2_OK;
}

// SEANetResnetBlock as one kernel (ua2_resblock.cu); served shape: C = 64, hidden 32, kernel 3 / 1, dilation 1
int ua2_resblock_f32(const float* x, const float* w1, const float* b1, const float* w2, const float* b2, float* y, int B, int C, int H,
                     int T, void* stream) {
  UA2_REQUIRE(x && w1 && b1 && w2 && b2 && y, "null argument");
  UA2_REQUIRE(B >= 1 && T >= 1, "bad shape");
  UA2_REQUIRE(x != y, "y must not alias x");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const cudaError_t e = launch_resblock_fused(lc, x, w1, b1, w2, b2, y, B, C, H, T);
  UA2_REQUIRE(e != cudaErrorNotSupported, "the fused residual block is served for C = 64, hidden = 32 only");
  UA2_CHECK_CUDA(e);
  return UA2_OK;
}

int ua2_convtr1d_causal_f32(const float* x, const float* w_ckc, const float* bias, float* y, int B, int Cin, int Cout,
                            int T_in, int stride, int pre_elu, void* stream) {
  UA2_REQUIRE(x && w_ckc && y, "null argument");
  UA2_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && T_in >= 1 && stride >= 1 && stride <= MAX_S, "stride must be in 1..8");
  const int K = 2 * stride;
  ConvTrParams p{x, w_ckc, bias, y, B, Cin, Cout, T_in, K, stride, pre_elu};
  const size_t smem = ((size_t)CI_T * (J_T + 1) + (size_t)CI_T * K * CO_T) * sizeof(float);
  static DeviceOnce once;
  if (once.need()) {
    UA2_CHECK_CUDA(cudaFuncSetAttribute(convtr1d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const dim3 grid((T_in + J_T - 1) / J_T, (Cout + CO_T - 1) / CO_T, B);
  UA2_CHECK_CUDA(launch(lc, convtr1d_kernel, grid, dim3(256), smem, p));
  return UA2_OK;
}

int ua2_convtr1d_repack_phase_f32(const float* w_torch, float* w_phase, int Cin, int Cout, int stride, void* stream) {
  UA2_REQUIRE(w_torch && w_phase && Cin >= 1 && Cout >= 1 && stride >= 1, "bad argument");
  const size_t total = (size_t)stride * Cout * Cin * 2;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch(lc, repack_convtr_phase_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, w_torch, w_phase, Cin, Cout,
                        stride));
  return UA2_OK;
}

int ua2_convtr1d_causal_gemm_f32(const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin, int Cout,
                                 int T_in, int stride, int pre_elu, void* stream) {
  UA2_REQUIRE(x && w_phase && y, "null argument");
  UA2_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && T_in >= 1 && stride >= 1 && stride <= 64, "bad shape");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_convtr1d_gemm(lc, x, w_phase, bias, y, B, Cin, Cout, T_in, stride, pre_elu, 0, T_in * stride));
  return UA2_OK;
}

// ---- general (non-causal capable) forms used by the ScalarModel wave decoder of ReasoningCodec_film (models/scalar24k.py)
int ua2_conv1d_f32(const float* x, const float* w_torch, const float* bias, const float* prelu_slope, const float* residual, float* y,
                   int B, int Cin, int Cout, int T_in, int K, int stride, int dilation, int pad_left, int pad_right, void* stream) {
  UA2_REQUIRE(x && w_torch && y, "null argument");
  UA2_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && T_in >= 1 && K >= 1 && stride >= 1 && dilation >= 1 && pad_left >= 0 && pad_right >= 0,
              "bad shape");
  const int k_eff = (K - 1) * dilation + 1;
  const int span = T_in + pad_left + pad_right - k_eff;
  UA2_REQUIRE(span >= 0, "input shorter than the kernel");
  const int T_out = span / stride + 1;  // nn.Conv1d output length
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_conv1d_gemm(lc, x, w_torch, bias, residual, y, B, Cin, Cout, T_in, T_out, K, stride, dilation, pad_left, 0, 0,
                                    prelu_slope));
  return UA2_OK;
}

int ua2_convtr1d_f32(const float* x, const float* w_phase, const float* bias, const float* prelu_slope, float* y, int B, int Cin,
                     int Cout, int T_in, int stride, int crop_left, int T_out, void* stream) {
  UA2_REQUIRE(x && w_phase && y, "null argument");
  UA2_REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && T_in >= 1 && stride >= 1 && crop_left >= 0 && T_out >= 1, "bad shape");
  UA2_REQUIRE(crop_left + T_out <= (T_in + 1) * stride, "crop window exceeds the full transposed-conv output ((T_in + 1) * stride)");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch_convtr1d_gemm(lc, x, w_phase, bias, y, B, Cin, Cout, T_in, stride, 0, crop_left, T_out, prelu_slope));
  return UA2_OK;
}

int ua2_elementwise_f32(const float* x, float* y, long long n, int op, float param, void* stream) {
  UA2_REQUIRE(x && y && n >= 1 && op >= 0 && op <= 1, "bad argument");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch(lc, elementwise_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, x, y, n, op, param));
  return UA2_OK;
}

int ua2_film_f32(const float* params, const float* features, const uint8_t* zero_mask, float* out, int B, int T, int C, float gamma_scale,
                 void* stream) {
  UA2_REQUIRE(params && features && out && B >= 1 && T >= 1 && C >= 1, "bad argument");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const long long M = (long long)B * T;
  UA2_CHECK_CUDA(launch(lc, film_kernel, dim3((unsigned)((M * C + 255) / 256)), dim3(256), 0, params, features, zero_mask, out, M, C, T,
                        gamma_scale));
  return UA2_OK;
}

int ua2_interp_nearest_f32(const float* x, float* y, int B, int C, int T_in, int T_out, float scale_factor, void* stream) {
  UA2_REQUIRE(x && y && B >= 1 && C >= 1 && T_in >= 1 && T_out >= 1 && scale_factor > 0.f, "bad argument");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const long long BC = (long long)B * C;
  UA2_CHECK_CUDA(launch(lc, interp_nearest_kernel, dim3((unsigned)((BC * T_out + 255) / 256)), dim3(256), 0, x, y, BC, T_in, T_out,
                        1.0f / scale_factor));
  return UA2_OK;
}

int ua2_linear_bias_f32(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, void* stream) {
  UA2_REQUIRE(x && W && y, "null argument");
  int rc = ua2_linear_f32(x, W, nullptr, 0.f, nullptr, y, M, N, K, stream);
  if (rc != UA2_OK || bias == nullptr) return rc;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const long long MN = (long long)M * N;
  UA2_CHECK_CUDA(launch(lc, add_bias_kernel, dim3((unsigned)((MN + 255) / 256)), dim3(256), 0, y, bias, MN, N));
  return UA2_OK;
}

int ua2_convtr1d_depthwise_f32(const float* x, const float* w, float* y, int B, int C, int T_in, int stride, void* stream) {
  UA2_REQUIRE(x && w && y && B >= 1 && C >= 1 && T_in >= 1 && stride >= 1, "bad argument");
  const size_t total = (size_t)B * C * T_in * stride;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const dim3 grid((unsigned)((total + 255) / 256));
  UA2_CHECK_CUDA(launch(lc, convtr1d_depthwise_kernel, grid, dim3(256), 0, x, w, y, B, C, T_in, stride));
  return UA2_OK;
}

int ua2_rvq_encode_f32(const float* x, const float* emb, const float* emb_sqnorm, int64_t* codes, int B, int D, int T, int K,
                       int n_q, int n_q_total, int q_off, void* stream) {
  UA2_REQUIRE(x && emb && emb_sqnorm && codes, "null argument");
  UA2_REQUIRE(B >= 1 && T >= 1 && K >= 1 && n_q >= 1 && D >= 4 && (D % 4) == 0, "bad shape (D must be a multiple of 4)");
  UA2_REQUIRE(q_off >= 0 && q_off + n_q <= n_q_total, "quantizer range outside the codes tensor");
  RvqEncParams p{x, emb, emb_sqnorm, codes, B, D, T, K, n_q, n_q_total, q_off};
  const size_t smem = (size_t)(FT + CT) * (D + 4) * sizeof(float);
  UA2_REQUIRE(smem <= 200 * 1024, "codebook dimension too large for the shared-memory tile");
  static DeviceOnce once;
  if (once.need()) {
    UA2_CHECK_CUDA(cudaFuncSetAttribute(rvq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const int N = B * T;
  UA2_CHECK_CUDA(launch(lc, rvq_encode_kernel, dim3((N + FT - 1) / FT), dim3(256), smem, p));
  return UA2_OK;
}

// Many-frame RVQ encode: per quantizer one 128 x 128-tiled fp32 GEMM (scores S = r E^T) + one argmin/residual-update
// kernel.  r_md: residual in frame-major layout (M = B*T, D), updated in place; S: (M, K) scratch.
int ua2_rvq_encode_gemm_f32(float* r_md, const float* emb, const float* emb_sqnorm, float* S, int64_t* codes, int B, int D, int T,
                            int K, int n_q, int n_q_total, int q_off, void* stream) {
  UA2_REQUIRE(r_md && emb && emb_sqnorm && S && codes, "null argument");
  UA2_REQUIRE(B >= 1 && T >= 1 && n_q >= 1 && (D % 8) == 0 && (K % 2) == 0, "bad shape (D % 8, even K)");
  UA2_REQUIRE(q_off >= 0 && q_off + n_q <= n_q_total, "quantizer range outside the codes tensor");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const int M = B * T;
  for (int q = 0; q < n_q; ++q) {
    GemvParams p;
    p.W = emb + (size_t)q * K * D;
    p.N = K;
    p.K = D;
    p.M = M;
    p.X = r_md;
    p.ldx = D;
    p.Y = S;
    p.ldy = K;
    UA2_CHECK_CUDA(launch_sgemm_linear(lc, PRO_PLAIN, EPI_STORE, p, nullptr));
    UA2_CHECK_CUDA(launch(lc, rvq_argmin_update_kernel, dim3((M * 32 + 255) / 256), dim3(256), 0, (const float*)S,
                          emb + (size_t)q * K * D, emb_sqnorm + (size_t)q * K, r_md, codes, M, K, D, T, n_q_total, q_off + q));
  }
  return UA2_OK;
}

int ua2_rvq_decode_f32(const int64_t* codes, const float* emb, float* out, int B, int D, int T, int K, int n_q, int n_q_total,
                       int q_off, void* stream) {
  UA2_REQUIRE(codes && emb && out, "null argument");
  UA2_REQUIRE(B >= 1 && T >= 1 && K >= 1 && n_q >= 1 && D >= 1, "bad shape");
  UA2_REQUIRE(q_off >= 0 && q_off + n_q <= n_q_total, "quantizer range outside the codes tensor");
  const size_t smem = (size_t)32 * (D + 1) * sizeof(float);
  static DeviceOnce once;
  if (once.need()) {
    UA2_CHECK_CUDA(cudaFuncSetAttribute(rvq_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  UA2_CHECK_CUDA(launch(lc, rvq_decode_kernel, dim3((T + 31) / 32, B), dim3(256), smem, codes, emb, out, B, D, T, K, n_q,
                        n_q_total, q_off));
  return UA2_OK;
}

}  // extern "C"
