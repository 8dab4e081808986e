// Device-side building blocks of the v3 skinny linear (shared by ua2_gemv3.cu and the persistent chain kernel
// ua2_chain.cu): bulk-copy / mbarrier PTX wrappers, unit -> weight-row mapping, fused prologue and epilogues.
#pragma once
#include "ua2_kernels.cuh"

namespace ua2 {
namespace v3dev {


constexpr int NWARPS = 9;   // reduction slots (8 or 9 warps per CTA)
constexpr int MAXW_RED = NWARPS;
constexpr int MAX_STAGES = 6;
constexpr int ROUND_UNITS = 32;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(s32(bar)),
      "r"(parity)
      : "memory");
}

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

// activation tile -> shared memory in ONE L2 round trip; RMSNorm: xs = x*g, per-warp partial sum of squares -> red
template <int MT, int PRO>
__device__ __forceinline__ void stage_activations3(const GemvParams& p, float* xs, float (*red)[MAXW_RED], int Kp, int m0,
                                                   int mcount, int n_splits) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K;
  if (PRO == PRO_ATTN) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      for (int k = tid * 4; k < Kp; k += blockDim.x * 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < mcount && k < K) {
          const int hh = k / p.hs, d = k - hh * p.hs;
          const size_t base = ((size_t)(m0 + m) * p.n_head + hh) * p.max_splits;
          if (n_splits == 1) {  // common case: one split, y = o / l
            const float l = p.ml_part[base * 2 + 1];
            const float4 o = *reinterpret_cast<const float4*>(p.o_part + base * p.hs + d);
            const float inv = 1.f / l;
            v = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
          } else {
            float mx = -INFINITY;
            for (int s = 0; s < n_splits; ++s) mx = fmaxf(mx, p.ml_part[(base + s) * 2]);
            float den = 0.f;
            for (int s = 0; s < n_splits; ++s) {
              const float w = __expf(p.ml_part[(base + s) * 2] - mx);  // empty split: exp(-inf) = 0
              if (w > 0.f) {
                den += w * p.ml_part[(base + s) * 2 + 1];
                const float4 o = *reinterpret_cast<const float4*>(p.o_part + (base + s) * p.hs + d);
                v.x += w * o.x;
                v.y += w * o.y;
                v.z += w * o.z;
                v.w += w * o.w;
              }
            }
            const float inv = 1.f / den;
            v.x *= inv;
            v.y *= inv;
            v.z *= inv;
            v.w *= inv;
          }
        }
        *reinterpret_cast<float4*>(xs + m * Kp + k) = v;
      }
    }
  } else {
    float ss[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) ss[m] = 0.f;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float* src = nullptr;
      if (m < mcount) {
        if (PRO == PRO_GATHER) {
          const long long row = (long long)p.gidx[(size_t)(m0 + m) * p.gidx_stride] + p.gidx_offset;
          src = p.emb + (size_t)row * K;
        } else {
          src = p.X + (size_t)(m0 + m) * p.ldx;
        }
      }
      for (int k = tid * 4; k < Kp; k += blockDim.x * 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src != nullptr && k < K) {
          v = *reinterpret_cast<const float4*>(src + k);
          if (PRO == PRO_RMSNORM) {
            ss[m] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            const float4 g = *reinterpret_cast<const float4*>(p.norm_w + k);
            v.x *= g.x;
            v.y *= g.y;
            v.z *= g.z;
            v.w *= g.w;
          }
        }
        *reinterpret_cast<float4*>(xs + m * Kp + k) = v;
      }
    }
    if (PRO == PRO_RMSNORM) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float s = warp_sum(ss[m]);
        if (lane == 0) red[m][warp] = s;
      }
    }
  }
  __syncthreads();
}

// epilogue of one (activation row m, unit) pair given the two complete row sums
template <int EPI>
__device__ __forceinline__ void epilogue_one(const GemvParams& p, int m, float a, float b, int nA) {
  if (EPI == EPI_STORE) {
    *reinterpret_cast<float2*>(p.Y + (size_t)m * p.ldy + nA) = make_float2(a, b);
  } else if (EPI == EPI_RESADD) {
    const float2 r = *reinterpret_cast<const float2*>(p.R + (size_t)m * p.ldr + nA);
    *reinterpret_cast<float2*>(p.Y + (size_t)m * p.ldy + nA) = make_float2(a + r.x, b + r.y);
  } else if (EPI == EPI_SWIGLU) {
    const float s = a / (1.0f + expf(-a));  // F.silu, lit_model.py:594
    p.Y[(size_t)m * p.ldy + nA] = s * b;
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

struct V3Cfg_unused_marker {};
struct V3Cfg {
  int nsl;       // K slices per unit; one warp per (group, slice)
  int ngrp;      // unit groups per CTA; warps = ngrp * nsl (8 or 9)
  int SL;        // floats per slice (multiple of 128)
  int KCW;       // floats per bulk copy (<= 1024): sized so a copy is >= 4 KB whenever K allows (see microbench)
  int stages;    // ring depth per warp (slots of KCW floats)
  int n_splits;  // attention splits to merge (PRO_ATTN)
};

constexpr int MAXW = 9;

// ROWS = weight rows per unit: 2 for the paired epilogues (QKV rotation pair / SwiGLU fc_1+fc_2), 1 for store / residual
template <int EPI>
struct RowsOf {
  static constexpr int value = (EPI == EPI_QKV || EPI == EPI_SWIGLU) ? 2 : 1;
};

template <int EPI>
__device__ __forceinline__ const float* unit_row(const GemvParams& p, int u, int r, int& nA) {
  if (EPI == EPI_SWIGLU) {
    nA = u;
    return (r == 0 ? p.W : p.W2) + (size_t)u * p.K;
  } else if (EPI == EPI_QKV) {
    const int half = p.hs >> 1;
    const int hh = u / half, i = u - hh * half;
    nA = hh * p.hs + i;
    return p.W + (size_t)(nA + r * half) * p.K;
  } else {
    nA = u;
    return p.W + (size_t)u * p.K;
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_v3(const GemvParams& p, int m, float a, float b, int nA) {
  if (EPI == EPI_STORE) {
    p.Y[(size_t)m * p.ldy + nA] = a;
  } else if (EPI == EPI_RESADD) {
    p.Y[(size_t)m * p.ldy + nA] = a + p.R[(size_t)m * p.ldr + nA];
  } else {
    epilogue_one<EPI>(p, m, a, b, nA);
  }
}


}  // namespace v3dev
}  // namespace ua2
