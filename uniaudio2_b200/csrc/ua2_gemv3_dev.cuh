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

#ifndef UA2_CPU_SHIM
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
#else  // tests/cpu_shim: the emulated barrier / copy of ua2_common.cuh
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { ::ua2::smem_bar_init(bar, (uint32_t)count); }
__device__ __forceinline__ void fence_mbar_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { ::ua2::smem_bar_arrive_expect_tx(bar, bytes); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) { ::ua2::bulk_copy_g2s(dst, src, bytes, bar); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { ::ua2::smem_bar_wait(bar, parity); }
#endif

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
  if (PRO == PRO_ATTN_DIRECT) {
    // softmax(q k^T / sqrt(hs)) v over the <= 8 cached keys of a local-decoder step (lit_model.py:505,513-532 with
    // T = 1; keys j <= pos are the causal mask row).  One warp per head, two dims per lane (hs = 64): every CTA recomputes
    // the full (M, n_head*hs) attention output from L2 (q 8 KB + K/V <= 32 KB) instead of waiting for a separate kernel.
    const int nw = blockDim.x >> 5;
    const int hs = p.hs;
    const int qpg = p.n_head / p.n_groups;
    const float scale = rsqrtf((float)hs);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      for (int k = (m < mcount ? K : 0) + tid; k < Kp; k += blockDim.x) xs[m * Kp + k] = 0.f;  // padding / unused rows
      if (m < mcount) {
        const int nk = min(p.pos[m0 + m] + 1, ATTN_DIRECT_MAX_KEYS);
        const int b = p.bidx[m0 + m];
        for (int hh = warp; hh < p.n_head; hh += nw) {
          const int g = hh / qpg;
          const float2 qv = *reinterpret_cast<const float2*>(p.X + (size_t)(m0 + m) * p.ldx + hh * hs + lane * 2);
          const size_t cb = ((size_t)b * p.n_groups + g) * p.S_max * hs + lane * 2;
          float2 kr[ATTN_DIRECT_MAX_KEYS], vr[ATTN_DIRECT_MAX_KEYS];
#pragma unroll
          for (int j = 0; j < ATTN_DIRECT_MAX_KEYS; ++j) {
            kr[j] = make_float2(0.f, 0.f);
            vr[j] = make_float2(0.f, 0.f);
            if (j < nk) {
              kr[j] = *reinterpret_cast<const float2*>(p.k_cache + cb + (size_t)j * hs);
              vr[j] = *reinterpret_cast<const float2*>(p.v_cache + cb + (size_t)j * hs);
            }
          }
          float sc[ATTN_DIRECT_MAX_KEYS];
          float mxs = -INFINITY;
#pragma unroll
          for (int j = 0; j < ATTN_DIRECT_MAX_KEYS; ++j) {
            sc[j] = warp_sum(qv.x * kr[j].x + qv.y * kr[j].y) * scale;
            if (j < nk) mxs = fmaxf(mxs, sc[j]);
          }
          float l = 0.f;
          float2 acc = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < ATTN_DIRECT_MAX_KEYS; ++j) {
            if (j < nk) {
              const float pj = expf(sc[j] - mxs);
              l += pj;
              acc.x = fmaf(pj, vr[j].x, acc.x);
              acc.y = fmaf(pj, vr[j].y, acc.y);
            }
          }
          const float inv = 1.f / l;
          *reinterpret_cast<float2*>(xs + m * Kp + hh * hs + lane * 2) = make_float2(acc.x * inv, acc.y * inv);
        }
      }
    }
  } else if (PRO == PRO_ATTN) {
    // Every global load of a thread is issued before the first dependent use: written as a plain loop the compiler
    // emits one LDG -> STS round trip per iteration (SASS), i.e. 4-8 serialised L2 latencies (1.5-3 us) per kernel.
    const int step = blockDim.x * 4;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const bool row_ok = m < mcount;
      if (n_splits <= 4) {
        constexpr int UNR = 2, NS = 4;
        for (int kb = tid * 4; kb < Kp; kb += step * UNR) {
          float ms[UNR][NS], ls[UNR][NS];
          float4 os[UNR][NS];
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int k = kb + u * step;
            const bool ok = row_ok && k < K;
            const int hh = ok ? k / p.hs : 0, d = ok ? k - hh * p.hs : 0;
            const size_t base = ((size_t)(m0 + m) * p.n_head + hh) * p.max_splits;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
              const bool oks = ok && s < n_splits;
              ms[u][s] = oks ? p.ml_part[(base + s) * 2] : -INFINITY;
              ls[u][s] = oks ? p.ml_part[(base + s) * 2 + 1] : 0.f;
              os[u][s] = oks ? *reinterpret_cast<const float4*>(p.o_part + (base + s) * p.hs + d) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const int k = kb + u * step;
            if (k < Kp) {
              float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row_ok && k < K) {
                if (n_splits == 1) {  // common case: one split, y = o / l
                  const float inv = 1.f / ls[u][0];
                  v = make_float4(os[u][0].x * inv, os[u][0].y * inv, os[u][0].z * inv, os[u][0].w * inv);
                } else {
                  float mx = -INFINITY;
#pragma unroll
                  for (int s = 0; s < NS; ++s) mx = fmaxf(mx, ms[u][s]);
                  float den = 0.f;
#pragma unroll
                  for (int s = 0; s < NS; ++s) {
                    const float w = __expf(ms[u][s] - mx);  // empty / unused split: exp(-inf) = 0
                    if (w > 0.f) {
                      den += w * ls[u][s];
                      v.x += w * os[u][s].x;
                      v.y += w * os[u][s].y;
                      v.z += w * os[u][s].z;
                      v.w += w * os[u][s].w;
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
        }
      } else {
        for (int k = tid * 4; k < Kp; k += step) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && k < K) {
            const int hh = k / p.hs, d = k - hh * p.hs;
            const size_t base = ((size_t)(m0 + m) * p.n_head + hh) * p.max_splits;
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
          *reinterpret_cast<float4*>(xs + m * Kp + k) = v;
        }
      }
    }
  } else {
    float ss[MT];
#pragma unroll
    for (int m = 0; m < MT; ++m) ss[m] = 0.f;
    const int step = blockDim.x * 4;
    constexpr int UNR = (PRO == PRO_RMSNORM) ? 4 : 8;
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
      // all loads of a chunk of UNR iterations first (see PRO_ATTN above)
      for (int kb = tid * 4; kb < Kp; kb += step * UNR) {
        float4 v[UNR], g[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int k = kb + u * step;
          const bool ok = src != nullptr && k < K;
          v[u] = ok ? *reinterpret_cast<const float4*>(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (PRO == PRO_RMSNORM) g[u] = ok ? *reinterpret_cast<const float4*>(p.norm_w + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int k = kb + u * step;
          if (k < Kp) {
            float4 w = v[u];
            if (PRO == PRO_RMSNORM) {
              ss[m] += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
              w.x *= g[u].x;
              w.y *= g[u].y;
              w.z *= g[u].z;
              w.w *= g[u].w;
            }
            *reinterpret_cast<float4*>(xs + m * Kp + k) = w;
          }
        }
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
__device__ __forceinline__ void epilogue_one(const GemvParams& p, int m, float a, float b, int nA, int ps_known = -1, int b_known = -1) {
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
    const int ps = ps_known >= 0 ? ps_known : p.pos[m];  // v3 loads the row positions in its prologue (one L2 trip less at the tail)
    const int bb = b_known >= 0 ? b_known : p.bidx[m];
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
        float* kc = p.k_cache + (((size_t)bb * p.n_groups + g) * p.S_max + ps) * hs + i;
        kc[0] = ra;
        kc[half] = rb;
      }
    } else {
      const int g = hh - p.n_head - p.n_groups;
      float* vc = p.v_cache + (((size_t)bb * p.n_groups + g) * p.S_max + ps) * hs + i;
      vc[0] = a;
      vc[half] = b;
    }
  }
}

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
__device__ __forceinline__ void epilogue_v3(const GemvParams& p, int m, float a, float b, int nA, int ps_known = -1, int b_known = -1) {
  if (EPI == EPI_STORE) {
    p.Y[(size_t)m * p.ldy + nA] = a;
  } else if (EPI == EPI_RESADD) {
    p.Y[(size_t)m * p.ldy + nA] = a + p.R[(size_t)m * p.ldr + nA];
  } else {
    epilogue_one<EPI>(p, m, a, b, nA, ps_known, b_known);
  }
}


}  // namespace v3dev
}  // namespace ua2
