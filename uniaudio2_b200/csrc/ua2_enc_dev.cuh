// Device side shared by the encoder handles (csrc/ua2_enc.cu: Whisper encoder; csrc/ua2_wavlm.cu: WavLM encoder): the epilogue kernels
// that consume a raw GEMM product (bias, GELU, positional table, q / k / v head split, residual add) and the one-row-per-CTA
// residual + LayerNorm kernel.  Included inside each translation unit (anonymous namespace: every unit gets its own instances).
#pragma once
#include <algorithm>

#ifndef UA2_CPU_SHIM
#include <cuda_bf16.h>
#endif

#include "ua2_kernels.cuh"
#include "ua2_umma.cuh"

namespace ua2 {
namespace {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }  // F.gelu default
__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 o;
  o.x = *reinterpret_cast<const uint32_t*>(&lo);
  o.y = *reinterpret_cast<const uint32_t*>(&hi);
  return o;
}

#ifndef UA2_CPU_SHIM
// fp32 -> bf16 copy of a weight matrix (bf16 modes: made at first use, 2 B per parameter); n4 = elements / 4
__global__ void enc_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = pack4_bf16(v.x, v.y, v.z, v.w);
  }
}
#endif

// ---------------------------------------------------------------------------------------------------------------- epilogues
enum EncMode : int {
  EE_GELU = 0,      // y = gelu(v + bias)                        conv1; fc1
  EE_GELU_POS = 1,  // h = gelu(v + bias) + pos[m % P]           conv2 + embed_positions (:806-811)
  EE_QKV = 2,       // q (M, D) fp32 (+ bias), k / v (B, H, T, hs) fp32     fp32-class attention operands
  EE_QKV16 = 3,     // q / k / v (B, H, T, hs) bf16                         tensor-core attention operands
  EE_RES = 4,       // h += v + bias                             residual adds (:416, :424)
  EE_BIAS = 5       // y = v + bias                              WavLM feature projection
};

struct EncEpi {
  const float* src;    // (M, N) raw product
  const float* slots;  // stream-K side slots of the GEMM (bf16 mode: summed here) or nullptr
  UmmaPlan pl;
  int has_split;
  const float* bias;   // (N)
  int M, N, T;         // T = rows per batch element
  float* y32;
  __nv_bfloat16* y16;  // when set, the result goes out as bf16 only
  const float* pos;    // (T, N)
  float *q, *k, *v;
  __nv_bfloat16 *q16, *k16, *v16;
  int H, hs;
  float* res;          // residual stream (M, N)
  // residual + LayerNorm kernel
  const float *ln_g, *ln_b;
  float eps;
  __nv_bfloat16* y16_also;  // residual + LayerNorm kernel: with y32, a bf16 copy of the normalised row (post-LayerNorm encoders: the row is both
                            // the residual stream and the next tensor-core linear's operand)
};

__device__ __forceinline__ float4 enc_load4(const EncEpi& e, int m, int c) {
  float4 v = *reinterpret_cast<const float4*>(e.src + (size_t)m * e.N + c);
#ifndef UA2_CPU_SHIM  // (the stream-K side slots belong to the tcgen05 GEMM, which the CPU shim of tests/ replaces by a plain GEMM)
  if (e.has_split) {
    const float4 sd = umma_side_sum4(e.pl, e.slots, m, c, e.N);
    v.x += sd.x;
    v.y += sd.y;
    v.z += sd.z;
    v.w += sd.w;
  }
#endif
  const float4 bs = *reinterpret_cast<const float4*>(e.bias + c);
  v.x += bs.x;
  v.y += bs.y;
  v.z += bs.z;
  v.w += bs.w;
  return v;
}

// one row per blockIdx.x, 4 columns per thread
template <int MODE>
__global__ void __launch_bounds__(256) enc_epi_kernel(const EncEpi e) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, c = (blockIdx.y * 256 + threadIdx.x) * 4;
  if (c >= e.N) return;
  float4 v = enc_load4(e, m, c);
  if (MODE == EE_GELU || MODE == EE_GELU_POS) {
    v = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
    if (MODE == EE_GELU_POS) {
      const float4 pe = *reinterpret_cast<const float4*>(e.pos + (size_t)(m % e.T) * e.N + c);
      v.x += pe.x;
      v.y += pe.y;
      v.z += pe.z;
      v.w += pe.w;
    }
    if (e.y16 != nullptr) {
      *reinterpret_cast<uint2*>(e.y16 + (size_t)m * e.N + c) = pack4_bf16(v.x, v.y, v.z, v.w);
    } else {
      *reinterpret_cast<float4*>(e.y32 + (size_t)m * e.N + c) = v;
    }
  } else if (MODE == EE_RES) {
    float4* rp = reinterpret_cast<float4*>(e.res + (size_t)m * e.N + c);
    float4 r = *rp;
    r.x += v.x;
    r.y += v.y;
    r.z += v.z;
    r.w += v.w;
    *rp = r;
  } else if (MODE == EE_BIAS) {
    *reinterpret_cast<float4*>(e.y32 + (size_t)m * e.N + c) = v;
  } else {  // EE_QKV / EE_QKV16: N = 3 D, columns [q | k | v], each (h d); the 4 columns lie inside one head
    const int D = e.N / 3;
    const int part = c / D, cc = c - part * D;
    const int hh = cc / e.hs, d = cc - hh * e.hs;
    const int b = m / e.T, t = m - b * e.T;
    const size_t hd = (((size_t)b * e.H + hh) * e.T + t) * e.hs + d;
    if (MODE == EE_QKV16) {
      *reinterpret_cast<uint2*>((part == 0 ? e.q16 : part == 1 ? e.k16 : e.v16) + hd) = pack4_bf16(v.x, v.y, v.z, v.w);
    } else if (part == 0) {
      *reinterpret_cast<float4*>(e.q + (size_t)m * D + cc) = v;
    } else {
      *reinterpret_cast<float4*>((part == 1 ? e.k : e.v) + hd) = v;
    }
  }
}

template <int MODE>
cudaError_t launch_enc_epi(const LaunchCtx& lc, const EncEpi& e) {
  return launch(lc, enc_epi_kernel<MODE>, dim3(e.M, (e.N + 1023) / 1024), dim3(256), 0, e);
}

// LayerNorm (affine) of one row per CTA, thread = 4 columns; HAS_RES: first h += product + bias (the residual add in front of it).
// Output: the normalised row as bf16 (next linear's operand) or fp32.
template <bool HAS_RES>
__global__ void __launch_bounds__(1024) enc_res_ln_kernel(const EncEpi e) {
  __shared__ float red[32];
  __shared__ float stat[2];
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const int c = tid * 4;
  const bool on = c < e.N;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (on) {
    float4* rp = reinterpret_cast<float4*>(e.res + (size_t)m * e.N + c);
    r = *rp;
    if (HAS_RES) {
      const float4 v = enc_load4(e, m, c);
      r.x += v.x;
      r.y += v.y;
      r.z += v.z;
      r.w += v.w;
      *rp = r;
    }
  }
  float s = warp_sum(on ? (r.x + r.y) + (r.z + r.w) : 0.f);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_sum(lane < nw ? red[lane] : 0.f);
    if (lane == 0) stat[0] = t / (float)e.N;
  }
  __syncthreads();
  const float mean = stat[0];
  const float dx = r.x - mean, dy = r.y - mean, dz = r.z - mean, dw = r.w - mean;
  const float q = warp_sum(on ? (dx * dx + dy * dy) + (dz * dz + dw * dw) : 0.f);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  if (warp == 0) {
    const float t = warp_sum(lane < nw ? red[lane] : 0.f);
    if (lane == 0) stat[1] = rsqrtf(t / (float)e.N + e.eps);
  }
  __syncthreads();
  if (!on) return;
  const float rstd = stat[1];
  const float4 g = *reinterpret_cast<const float4*>(e.ln_g + c), bb = *reinterpret_cast<const float4*>(e.ln_b + c);
  const float o0 = dx * rstd * g.x + bb.x, o1 = dy * rstd * g.y + bb.y, o2 = dz * rstd * g.z + bb.z, o3 = dw * rstd * g.w + bb.w;
  if (e.y16 != nullptr) {
    *reinterpret_cast<uint2*>(e.y16 + (size_t)m * e.N + c) = pack4_bf16(o0, o1, o2, o3);
  } else {
    *reinterpret_cast<float4*>(e.y32 + (size_t)m * e.N + c) = make_float4(o0, o1, o2, o3);
    if (e.y16_also != nullptr) *reinterpret_cast<uint2*>(e.y16_also + (size_t)m * e.N + c) = pack4_bf16(o0, o1, o2, o3);
  }
}

unsigned grid_for(long long n) { return (unsigned)std::min<long long>((n + 255) / 256, 148LL * 32); }

struct Lin {
  const float* w = nullptr;
  const float* b = nullptr;
};
}  // namespace
}  // namespace ua2
