// Device helpers shared by the v1/v2 skinny linears (ua2_gemv.cu) and the tiled fp32 GEMM (ua2_sgemm.cu):
// output-unit -> weight-row mapping and the fused epilogues.
#pragma once
#include "ua2_kernels.cuh"

namespace ua2 {
namespace v1dev {

template <int EPI>
__device__ __forceinline__ void unit_rows(const GemvParams& p, int u, const float*& rowA, const float*& rowB, int& nA,
                                          int& nB) {
  if (EPI == EPI_SWIGLU) {
    nA = nB = u;
    rowA = p.W + (size_t)u * p.K;
    rowB = p.W2 + (size_t)u * p.K;
  } else if (EPI == EPI_QKV) {
    const int half = p.hs >> 1;
    const int hh = u / half, i = u - hh * half;
    nA = hh * p.hs + i;
    nB = nA + half;
    rowA = p.W + (size_t)nA * p.K;
    rowB = p.W + (size_t)nB * p.K;
  } else {
    nA = 2 * u;
    nB = nA + 1;
    rowA = p.W + (size_t)nA * p.K;
    rowB = p.W + (size_t)nB * p.K;
  }
}


// ---- fused epilogue: lane m (< mcount) holds the two row sums (a, b) of output rows (nA, nB) for activation row m0+m
template <int EPI>
__device__ __forceinline__ void epilogue(const GemvParams& p, int lane, int mcount, int m0, float a, float b, int nA,
                                         int nB) {
  if (lane < mcount) {
      const int m = m0 + lane;
      if (EPI == EPI_STORE) {
        *reinterpret_cast<float2*>(p.Y + (size_t)m * p.ldy + nA) = make_float2(a, b);
      } else if (EPI == EPI_RESADD) {
        const float2 r = *reinterpret_cast<const float2*>(p.R + (size_t)m * p.ldr + nA);
        *reinterpret_cast<float2*>(p.Y + (size_t)m * p.ldy + nA) = make_float2(a + r.x, b + r.y);
      } else if (EPI == EPI_SWIGLU) {
        const float s = a / (1.0f + expf(-a));  // F.silu, lit_model.py:594
        p.Y[(size_t)m * p.ldy + nA] = s * b;
      } else if (EPI == EPI_GELU) {  // F.gelu (exact erf form), transformer.py:553
        const float ga = 0.5f * a * (1.0f + erff(a * 0.70710678118654752440f));
        const float gb = 0.5f * b * (1.0f + erff(b * 0.70710678118654752440f));
        *reinterpret_cast<float2*>(p.Y + (size_t)m * p.ldy + nA) = make_float2(ga, gb);
      } else if (EPI == EPI_SCALE_RESADD) {  // x_orig + layer_scale(update), transformer.py:569, :578
        const float2 r = *reinterpret_cast<const float2*>(p.R + (size_t)m * p.ldr + nA);
        const float2 sc = *reinterpret_cast<const float2*>(p.scale + nA);
        *reinterpret_cast<float2*>(p.Y + (size_t)m * p.ldy + nA) = make_float2(r.x + sc.x * a, r.y + sc.y * b);
      } else if (EPI == EPI_QKV_IL) {
        // in_proj rows are ordered (p h d) (transformer.py:391-393); interleaved-pair RoPE on q and k with the angle
        // computed on the fly in fp32 (rope.py:40-58); K/V go to (B, H, T, D) buffers indexed by (bidx, pos)
        const int hs = p.hs, HD = p.n_head * hs;
        const int part = nA / HD, rem = nA - part * HD;
        const int hh = rem / hs, d = rem - hh * hs;  // d even
        const int ps = p.pos[m];
        float oa = a, ob = b;
        if (part < 2) {
          const float freq = expf((float)(d >> 1) * (-logf(p.rope_max_period) * 2.0f / (float)hs));
          const float ang = freq * (float)ps;
          const float c = cosf(ang), sn = sinf(ang);
          oa = __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, sn));
          ob = __fadd_rn(__fmul_rn(a, sn), __fmul_rn(b, c));
        }
        if (part == 0) {
          *reinterpret_cast<float2*>(p.q_out + (size_t)m * HD + hh * hs + d) = make_float2(oa, ob);
        } else {
          float* dst = (part == 1 ? p.k_cache : p.v_cache) + (((size_t)p.bidx[m] * p.n_head + hh) * p.S_max + ps) * hs + d;
          *reinterpret_cast<float2*>(dst) = make_float2(oa, ob);
        }
      } else {  // EPI_QKV: split, half-split RoPE (lit_model.py:795-806), KV-cache append (:854-855)
        const int hs = p.hs, half = hs >> 1;
        const int hh = nA / hs, i = nA - hh * hs;
        const int ps = p.pos[m];
        if (hh < p.n_head + p.n_groups) {
          const float c0 = p.cos[(size_t)ps * hs + i], s0 = p.sin[(size_t)ps * hs + i];
          const float c1 = p.cos[(size_t)ps * hs + i + half], s1 = p.sin[(size_t)ps * hs + i + half];
          const float ra = __fadd_rn(__fmul_rn(a, c0), __fmul_rn(-b, s0));
          const float rb = __fadd_rn(__fmul_rn(b, c1), __fmul_rn(a, s1));
          if (hh < p.n_head) {
            float* q = p.q_out + (size_t)m * (p.n_head * hs) + hh * hs + i;
            q[0] = ra;
            q[half] = rb;
          } else {
            const int g = hh - p.n_head;
            float* kc = p.k_cache + (((size_t)p.bidx[m] * p.n_groups + g) * p.S_max + ps) * hs + i;
            kc[0] = ra;
            kc[half] = rb;
          }
        } else {
          const int g = hh - p.n_head - p.n_groups;
          float* vc = p.v_cache + (((size_t)p.bidx[m] * p.n_groups + g) * p.S_max + ps) * hs + i;
          vc[0] = a;
          vc[half] = b;
        }
      }
    }
}


}  // namespace v1dev
}  // namespace ua2
