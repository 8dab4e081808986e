// Skinny fp32 linear family (M <= 8 rows per tile) for the AR-decode hot path.
//
// Replaces the F.linear call sites of llm_models/lit_model.py:424 (qkv), :511 (attn proj), :592-595 (LLaMAMLP) and
// llm_models/model_new.py:617 (lm_head), :631 (projection), :632 (audio_head mm) - all bias-free, fp32.
//
// Roofline: HBM.  At M <= 8 the arithmetic intensity is <= 4 FLOP/B, so the kernel is a weight streamer:
//   * each warp owns two weight rows ("unit") and streams them with 128-bit ld.global.nc.L1::no_allocate loads,
//     16 loads (8 KB) in flight per warp before the first FMA; the first batch is issued BEFORE the activation
//     prologue (and before griddepcontrol.wait), so the weight stream does not stall on the producer kernel;
//   * the M x K activation tile is produced once per CTA in shared memory by a fused prologue
//     (plain copy | RMSNorm | embedding gather | split-softmax attention combine) and read conflict-free;
//   * a fused epilogue consumes the two row sums (store | residual add | RoPE + KV-cache append | SwiGLU).
// Algorithmic bytes per launch = 4*N*K (weights) (+ 4*M*(K+N) activations, negligible).
#include "ua2_gemv_dev.cuh"
#include "ua2_kernels.cuh"

namespace ua2 {

namespace {

using namespace v1dev;

// ---- fused prologue: produce the MT x Kp activation tile in shared memory (all threads of the CTA)
template <int MT, int PRO>
__device__ __forceinline__ void stage_activations(const GemvParams& p, float* xs, float (*red)[8], int Kp, int m0,
                                                  int mcount) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const int K = p.K;
  if (PRO == PRO_PLAIN || PRO == PRO_RMSNORM || PRO == PRO_GATHER || PRO == PRO_LAYERNORM) {
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
        if (src != nullptr && k < K) v = *reinterpret_cast<const float4*>(src + k);
        if (PRO == PRO_RMSNORM) ss[m] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        if (PRO == PRO_LAYERNORM) ss[m] += (v.x + v.y) + (v.z + v.w);
        *reinterpret_cast<float4*>(xs + m * Kp + k) = v;
      }
    }
    if (PRO == PRO_RMSNORM) {
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float s = warp_sum(ss[m]);
        if (lane == 0) red[m][warp] = s;
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float tot = 0.f;
        for (int w = 0; w < nwarps; ++w) tot += red[m][w];
        const float rs = rsqrtf(tot / (float)K + p.eps);  // lit_model.py:887-888
        for (int k = tid * 4; k < K; k += blockDim.x * 4) {
          float4 v = *reinterpret_cast<float4*>(xs + m * Kp + k);
          const float4 g = *reinterpret_cast<const float4*>(p.norm_w + k);
          v.x = (v.x * rs) * g.x;
          v.y = (v.y * rs) * g.y;
          v.z = (v.z * rs) * g.z;
          v.w = (v.w * rs) * g.w;
          *reinterpret_cast<float4*>(xs + m * Kp + k) = v;
        }
      }
    }
    if (PRO == PRO_LAYERNORM) {  // nn.LayerNorm(eps=1e-5) with affine weight + bias (transformer.py create_norm_fn)
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float s1 = warp_sum(ss[m]);
        if (lane == 0) red[m][warp] = s1;
      }
      __syncthreads();
      float mean[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float tot = 0.f;
        for (int w = 0; w < nwarps; ++w) tot += red[m][w];
        mean[m] = tot / (float)K;
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float sq = 0.f;
        for (int k = tid * 4; k < K; k += blockDim.x * 4) {
          const float4 v = *reinterpret_cast<float4*>(xs + m * Kp + k);
          const float a = v.x - mean[m], b = v.y - mean[m], c = v.z - mean[m], d = v.w - mean[m];
          sq += a * a + b * b + c * c + d * d;
        }
        sq = warp_sum(sq);
        if (lane == 0) red[m][warp] = sq;
      }
      __syncthreads();
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        float tot = 0.f;
        for (int w = 0; w < nwarps; ++w) tot += red[m][w];
        const float rs = rsqrtf(tot / (float)K + p.eps);
        for (int k = tid * 4; k < K; k += blockDim.x * 4) {
          float4 v = *reinterpret_cast<float4*>(xs + m * Kp + k);
          const float4 g = *reinterpret_cast<const float4*>(p.norm_w + k);
          const float4 bb = *reinterpret_cast<const float4*>(p.norm_b + k);
          v.x = (v.x - mean[m]) * rs * g.x + bb.x;
          v.y = (v.y - mean[m]) * rs * g.y + bb.y;
          v.z = (v.z - mean[m]) * rs * g.z + bb.z;
          v.w = (v.w - mean[m]) * rs * g.w + bb.w;
          *reinterpret_cast<float4*>(xs + m * Kp + k) = v;
        }
      }
    }
  } else {  // PRO_ATTN: merge the split-softmax partials (flash-decoding combine) into y (M, n_head*hs)
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      // all launched splits are merged; empty / out-of-window ones carry (m = -inf, l = 0) and get weight 0
      const int n_s = (m < mcount) ? (p.n_splits > 0 ? p.n_splits : (p.pos[m0 + m] + ATTN_CHUNK) / ATTN_CHUNK) : 0;
      for (int k = tid * 4; k < Kp; k += blockDim.x * 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n_s > 0 && k < K) {
          const int hh = k / p.hs, d = k - hh * p.hs;
          const size_t base = ((size_t)(m0 + m) * p.n_head + hh) * p.max_splits;
          float mx = -INFINITY;
          for (int s = 0; s < n_s; ++s) mx = fmaxf(mx, p.ml_part[(base + s) * 2]);
          float den = 0.f;
          for (int s = 0; s < n_s; ++s) {
            const float w = __expf(p.ml_part[(base + s) * 2] - mx);
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
  __syncthreads();

}

// =====================================================================================================
// v2: same math, weights fetched by the bulk-copy engine (cp.async.bulk, SASS UBLKCP) instead of registers.
//
// Every warp owns a private ring of STAGES shared-memory slots (2 rows x KC floats each) with one mbarrier per
// slot: lane 0 arms the barrier with expect_tx and issues two 1-D bulk copies (row A chunk, row B chunk); all
// lanes wait on the barrier's phase, consume the chunk with conflict-free 128-bit LDS, __syncwarp, and lane 0
// immediately refills the slot with the chunk STAGES ahead.  The prefetch distance (16 KB per warp) is therefore
// independent of the register budget, continuous across unit boundaries, and is issued before the activation
// prologue and before griddepcontrol.wait - so under programmatic dependent launch the weight stream of kernel
// N+1 is already in flight while kernel N drains.
// =====================================================================================================
constexpr int KC = 1024;    // floats per row chunk: 4 KB bulk copies (1-2 KB copies cap at 2.8-5.8 TB/s, see microbench)
constexpr int STAGES = 2;   // ring depth per warp: 2 x 2 x 4 KB = 16 KB in flight per warp
constexpr int V2_WARPS = 4;

#ifndef UA2_CPU_SHIM
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
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
      "}\n" ::"r"(smem_u32(bar)),
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

template <int MT, int PRO, int EPI>
__global__ void __launch_bounds__(V2_WARPS * 32) gemv2_kernel(const GemvParams p) {
  extern __shared__ __align__(128) float smem2[];
  __shared__ float red[8][8];
  __shared__ __align__(8) uint64_t bars[V2_WARPS][STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K;
  const int nIt = (K + 127) >> 7;
  const int Kp = nIt << 7;
  const int nCh = (K + KC - 1) / KC;
  const int m0 = blockIdx.x * MT;
  const int mcount = min(MT, p.M - m0);
  const int n_units = (EPI == EPI_SWIGLU) ? p.N : (p.N >> 1);
  const int unit0 = blockIdx.y * V2_WARPS + warp;
  const int unit_stride = gridDim.y * V2_WARPS;
  float* xs = smem2;                                                   // MT x Kp
  float* ring = smem2 + (size_t)MT * Kp + (size_t)warp * STAGES * 2 * KC;  // this warp's slots
  const int my_units = unit0 < n_units ? (n_units - unit0 + unit_stride - 1) / unit_stride : 0;
  const int total = my_units * nCh;  // chunks this warp will consume

  // producer side (lane 0): issue chunk t of this warp's flattened (unit, chunk) sequence
  auto issue = [&](int t) {
    const int uo = t / nCh, c = t - uo * nCh;
    const float *rowA, *rowB;
    int nA, nB;
    unit_rows<EPI>(p, unit0 + uo * unit_stride, rowA, rowB, nA, nB);
    const int k0 = c * KC;
    const uint32_t bytes = (uint32_t)min(KC, K - k0) * 4u;
    const int st = t % STAGES;
    uint64_t* bar = &bars[warp][st];
    mbar_expect_tx(bar, 2u * bytes);
    bulk_g2s(ring + (size_t)st * 2 * KC, rowA + k0, bytes, bar);
    bulk_g2s(ring + (size_t)st * 2 * KC + KC, rowB + k0, bytes, bar);
  };

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp][s], 1);
    fence_mbar_init();
    const int pre = min(total, STAGES);
    for (int t = 0; t < pre; ++t) issue(t);  // weights do not depend on the producer kernel: prefetch first
  }
  __syncwarp();
  pdl_launch_dependents();
  pdl_wait();

  stage_activations<MT, PRO>(p, xs, red, Kp, m0, mcount);  // ends with __syncthreads()

  float accA[MT], accB[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) accA[m] = accB[m] = 0.f;
  for (int t = 0; t < total; ++t) {
    const int uo = t / nCh, c = t - uo * nCh;
    const int st = t % STAGES;
    mbar_wait(&bars[warp][st], (uint32_t)((t / STAGES) & 1));
    const float* sa = ring + (size_t)st * 2 * KC;
    const float* sb = sa + KC;
    const int k0 = c * KC;
#pragma unroll
    for (int it = 0; it < KC / 128; ++it) {
      const int kk = (it * 32 + lane) * 4;
      if (k0 + kk < K) {
        const float4 wa = *reinterpret_cast<const float4*>(sa + kk);
        const float4 wb = *reinterpret_cast<const float4*>(sb + kk);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + m * Kp + k0 + kk);
          accA[m] = fmaf(wa.x, xv.x, accA[m]);
          accA[m] = fmaf(wa.y, xv.y, accA[m]);
          accA[m] = fmaf(wa.z, xv.z, accA[m]);
          accA[m] = fmaf(wa.w, xv.w, accA[m]);
          accB[m] = fmaf(wb.x, xv.x, accB[m]);
          accB[m] = fmaf(wb.y, xv.y, accB[m]);
          accB[m] = fmaf(wb.z, xv.z, accB[m]);
          accB[m] = fmaf(wb.w, xv.w, accB[m]);
        }
      }
    }
    __syncwarp();  // every lane has consumed the slot -> refill it with the chunk STAGES ahead
    if (lane == 0 && t + STAGES < total) issue(t + STAGES);
    if (c == nCh - 1) {  // unit finished: reduce + fused epilogue
      const float *rowA, *rowB;
      int nA, nB;
      unit_rows<EPI>(p, unit0 + uo * unit_stride, rowA, rowB, nA, nB);
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        const float sa2 = warp_sum(accA[m]);
        const float sb2 = warp_sum(accB[m]);
        if (lane == m) {
          a = sa2;
          b = sb2;
        }
        accA[m] = accB[m] = 0.f;
      }
      epilogue<EPI>(p, lane, mcount, m0, a, b, nA, nB);
    }
  }
}

int g_sgemm_min_rows = 128;

int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

template <int MT, int PRO, int EPI>
cudaError_t launch_one(const LaunchCtx& lc, const GemvParams& p) {
  const int nIt = (p.K + 127) >> 7;
  const size_t xbytes = (size_t)MT * nIt * 128 * sizeof(float);
  const int n_units = (EPI == EPI_SWIGLU) ? p.N : p.N / 2;
  const int m_tiles = (p.M + MT - 1) / MT;
  const size_t kMaxSmem = 220 * 1024;
  auto kern = gemv2_kernel<MT, PRO, EPI>;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
    if (e != cudaSuccess) return e;
    e = prefer_max_smem(kern);
    if (e != cudaSuccess) return e;
  }
  const size_t smem = xbytes + (size_t)V2_WARPS * STAGES * 2 * KC * sizeof(float);
  if (smem > kMaxSmem) return cudaErrorInvalidValue;
  // one wave of co-resident CTAs (shared memory bounds the CTAs per SM); beyond that warps loop over units
  int per_sm = (int)((227 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int gy = (n_units + V2_WARPS - 1) / V2_WARPS;
  const int cap = sm_count() * per_sm;
  if (gy > cap) gy = cap;
  return launch(lc, kern, dim3(m_tiles, gy), dim3(V2_WARPS * 32), smem, p);
}

template <int PRO, int EPI>
cudaError_t launch_mt(const LaunchCtx& lc, const GemvParams& p) {
  // tile of M rows per CTA; bounded by shared memory (MT*K*4 <= 200 KB)
  const size_t rowb = (size_t)((p.K + 127) / 128) * 128 * 4;
  int mt = p.M >= 8 ? 8 : (p.M >= 3 ? 4 : p.M);
  while (mt > 1 && mt * rowb > 128 * 1024) mt >>= 1;
  switch (mt) {
    case 1: return launch_one<1, PRO, EPI>(lc, p);
    case 2: return launch_one<2, PRO, EPI>(lc, p);
    case 4: return launch_one<4, PRO, EPI>(lc, p);
    default: return launch_one<8, PRO, EPI>(lc, p);
  }
}

}  // namespace

void set_sgemm_min_rows(int v) { g_sgemm_min_rows = v < 1 ? 1 : v; }
int get_sgemm_min_rows() { return g_sgemm_min_rows; }

namespace {
// Record / replay the frame's sequence of linears (GemvSeq) and attach the tail-prefetch specs of the following ones.
const GemvParams& with_prefetch(const LaunchCtx& lc, int epi, const GemvParams& p, GemvParams& tmp) {
  GemvSeq* sq = lc.seq;
  if (sq == nullptr) return p;
  if (!sq->recorded) {
    GemvSeqEntry e;
    e.W = p.W;
    e.W2 = p.W2;
    e.N = p.N;
    e.K = p.K;
    e.M = p.M;
    e.epi = epi;
    e.hs = p.hs;
    sq->ops.push_back(e);
    return p;
  }
  const int n = (int)sq->ops.size();
  const int i = sq->pos++;
  if (i >= n || sq->ops[i].W != p.W || sq->ops[i].M != p.M) return p;  // not the recorded frame: no prefetch
  const size_t budget = gemv3_prefetch_budget(sq->ops[i].idle_after);
  if (budget == 0) return p;
  GemvSeqEntry nxt[PF_MAX];
  for (int j = 0; j < PF_MAX; ++j) nxt[j] = sq->ops[(i + 1 + j) % n];  // wraps into the next frame (same weights)
  tmp = p;
  tmp.n_pf = gemv3_make_pf(nxt, PF_MAX < n ? PF_MAX : n, budget, tmp.pf);
  return tmp;
}
}  // namespace

cudaError_t launch_gemv(const LaunchCtx& lc, int pro, int epi, const GemvParams& p_in) {
  GemvParams p_tmp;
  const GemvParams& p = with_prefetch(lc, epi, p_in, p_tmp);
  if (p.M <= 0 || p.N <= 0 || p.K <= 0 || (p.K & 3) || (epi != EPI_SWIGLU && (p.N & 1))) return cudaErrorInvalidValue;
  // many rows: 128 x 128 register-tiled fp32 GEMM (each weight element reused 128x) instead of re-streaming W per 8 rows
  const bool tc_on = get_tc_gemm() && p.tc != nullptr && p.M >= get_tc_min_rows();
  if ((p.M >= g_sgemm_min_rows || tc_on) && p.ws != nullptr && pro != PRO_ATTN_DIRECT) {
    GemvParams q = p;
    int pro2 = pro;
    const size_t stats_floats = ((2 * (size_t)p.M + 3) / 4) * 4;  // keeps the combine buffer behind it 16-byte aligned (odd M)
    size_t need = stats_floats;
    bool ok = true;
    if (pro == PRO_ATTN) {
      need += (size_t)p.M * p.K;
      if (p.ws_floats >= need) {
        AttnParams a;
        a.o_part = const_cast<float*>(p.o_part);
        a.ml_part = const_cast<float*>(p.ml_part);
        a.pos = p.pos;
        a.M = p.M;
        a.n_head = p.n_head;
        a.hs = p.hs;
        a.max_splits = p.max_splits;
        a.n_splits_launch = p.n_splits;
        cudaError_t e = launch_attn_combine(lc, a, p.ws + stats_floats);
        if (e != cudaSuccess) return e;
        q.X = p.ws + stats_floats;
        q.ldx = p.K;
        pro2 = PRO_PLAIN;
      } else {
        ok = false;
      }
    } else if (p.ws_floats < need) {
      ok = false;
    }
    if (ok) {
      if (tc_on) {  // tcgen05 3xTF32 path (ua2_tcgemm.cu); unsupported combinations fall through
        cudaError_t et = launch_tc_linear(lc, pro2, epi, q);
        if (et != cudaErrorNotSupported) return et;
      }
      if (p.M >= g_sgemm_min_rows) {
        cudaError_t e = launch_sgemm_linear(lc, pro2, epi, q, p.ws);
        if (e != cudaErrorNotSupported) return e;
      }
      // otherwise: the skinny path below, from the original operands (a combined attention tile in scratch is simply unused)
    }
  }
  if (pro == PRO_ATTN_DIRECT) {
    if (epi != EPI_RESADD || p.hs != 64 || p.S_max > ATTN_DIRECT_MAX_KEYS) return cudaErrorInvalidValue;
    return launch_gemv3(lc, pro, epi, p, 1);
  }
  if (pro < PRO_LAYERNORM && epi < EPI_GELU) return launch_gemv3(lc, pro, epi, p, p.n_splits > 0 ? p.n_splits : 1);
  // the remaining (prologue, epilogue) pairs belong to the Moshi-family layers (llm_modules/transformer.py:430-588, the codec
  // transformer and ua2_stream.cu): LayerNorm / GELU / LayerScale / interleaved RoPE, on the per-warp bulk-copy-ring kernel
#define UA2_CASE(P, E) \
  if (pro == P && epi == E) return launch_mt<P, E>(lc, p);
  UA2_CASE(PRO_LAYERNORM, EPI_QKV_IL)
  UA2_CASE(PRO_LAYERNORM, EPI_GELU)
  UA2_CASE(PRO_ATTN, EPI_SCALE_RESADD)
  UA2_CASE(PRO_PLAIN, EPI_SCALE_RESADD)
  UA2_CASE(PRO_PLAIN, EPI_QKV_IL)
  UA2_CASE(PRO_LAYERNORM, EPI_STORE)
  UA2_CASE(PRO_LAYERNORM, EPI_SWIGLU)
  UA2_CASE(PRO_RMSNORM, EPI_GELU)
#undef UA2_CASE
  return cudaErrorInvalidValue;
}

}  // namespace ua2
