// Causal GQA attention of a few query rows against the preallocated KV cache (decode and chunked prefill).
//
// Replaces llm_models/lit_model.py:468-532: slice cache to input_pos_maxp1, repeat_interleave K/V over the
// q_per_kv query heads, F.scaled_dot_product_attention with the tril bool mask (key j visible iff j <= pos),
// scale 1/sqrt(head_size).  The kernel never materialises the repeat: one CTA serves all q_per_kv heads of a
// KV group, so each K/V row is read once.
//
// Split-softmax ("flash decoding"): grid (split, group, row); each CTA handles ATTN_CHUNK keys and writes
// un-normalised partial outputs + (max, sum); the consumer (the attention-output projection's fused prologue,
// PRO_ATTN in ua2_gemv.cu, or launch_attn_combine) merges the splits.
// Roofline: HBM/L2 latency - per layer it moves only 2*G*S*hs*4 bytes (8 KB per cached position).
#include "ua2_kernels.cuh"

namespace ua2 {
namespace {

constexpr int MAX_QPK = 4;

__device__ __forceinline__ uint32_t smem_u32a(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One CTA = (split of ATTN_CHUNK keys, KV group, query row).  The K and V chunks are contiguous in the cache
// ((b, g, pos, hs) layout), so each is fetched with ONE bulk copy (cp.async.bulk -> mbarrier) into shared memory:
// every byte of the chunk is in flight at once and the whole kernel is a single memory round trip.
//   phase 1 (QK^T): 8 lanes per key (conflict-free 128-bit LDS), q in registers, 3-step shuffle reduce
//   phase 2       : warp h = softmax statistics of query head h over the chunk
//   phase 3 (PV)  : warp w = keys [w*C/4, (w+1)*C/4), lanes over head dims (float4), cross-warp reduce in smem
template <int HS>
__global__ void __launch_bounds__(128) attn_split_kernel(const AttnParams p) {
  extern __shared__ __align__(128) float kv_s[];  // K chunk [C][HS] then V chunk [C][HS]
  __shared__ float sc[MAX_QPK][ATTN_CHUNK];
  __shared__ __align__(16) float redp[4][MAX_QPK][HS];
  __shared__ __align__(8) uint64_t bar;
  constexpr int C = ATTN_CHUNK;
  constexpr int NI = HS / 32;  // float4 per lane per key in phase 1
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, g = blockIdx.y, m = blockIdx.z;
  pdl_launch_dependents();
  pdl_wait();
  // The K/V chunk copy does not wait for pos[m]: it always fetches the whole chunk (clipped to the cache), so the position
  // load, the q load and the two bulk copies are ONE L2 round trip instead of two.  Rows past n_keys are never read.
  const int start = split * C;
  const int b = p.bidx_identity ? m : p.bidx[m];
  const int span = max(0, min(C, p.S_max - start));
  const float* Kc = p.k_cache + (((size_t)b * p.n_groups + g) * p.S_max + start) * HS;
  const float* Vc = p.v_cache + (((size_t)b * p.n_groups + g) * p.S_max + start) * HS;
  float* Ks = kv_s;
  float* Vs = kv_s + C * HS;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32a(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)span * HS * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32a(&bar)), "r"(2u * bytes) : "memory");
    if (bytes > 0) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32a(Ks)),
                   "l"(Kc), "r"(bytes), "r"(smem_u32a(&bar))
                   : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32a(Vs)),
                   "l"(Vc), "r"(bytes), "r"(smem_u32a(&bar))
                   : "memory");
    }
  }
  const int n_keys = p.pos[m] + 1;
  const int qpk = p.n_head / p.n_groups;
  const int lo = (p.window > 0) ? max(0, n_keys - p.window) : 0;  // first visible key (Moshi context window)
  const bool empty = start >= n_keys || start + C <= lo;
  if (empty && tid < qpk) {  // empty split: publish weight-0 statistics so consumers can merge blindly
    const size_t idx = (((size_t)m * p.n_head + g * qpk + tid) * p.max_splits + split) * 2;
    p.ml_part[idx] = -INFINITY;
    p.ml_part[idx + 1] = 0.f;
  }
  const int cnt = min(C, n_keys - start);
  const float scale = rsqrtf((float)HS);

  // q -> registers while the copies fly: lane (kk = lane>>3, part = lane&7) needs q[h][part*4 + 32*i .. +3]
  const int kk = lane >> 3, part = lane & 7;
  float4 qr[MAX_QPK][NI];
#pragma unroll
  for (int h = 0; h < MAX_QPK; ++h)
#pragma unroll
    for (int i = 0; i < NI; ++i)
      qr[h][i] = (h < qpk) ? *reinterpret_cast<const float4*>(p.q + (size_t)m * p.n_head * HS + (g * qpk + h) * HS +
                                                               part * 4 + 32 * i)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();  // barrier init visible to every waiter
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32a(&bar)),
      "r"(0u)
      : "memory");
  if (empty) return;  // (only after the copies have landed: the CTA's shared memory must outlive them)

  // ---- phase 1: scores.  16 keys per iteration over the CTA (4 warps x 4 keys)
#pragma unroll
  for (int it = 0; it < C / 16; ++it) {
    const int j = it * 16 + warp * 4 + kk;
    float s[MAX_QPK];
#pragma unroll
    for (int h = 0; h < MAX_QPK; ++h) s[h] = 0.f;
    if (j < cnt) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float4 kv = *reinterpret_cast<const float4*>(Ks + (size_t)j * HS + part * 4 + 32 * i);
#pragma unroll
        for (int h = 0; h < MAX_QPK; ++h) {
          s[h] = fmaf(kv.x, qr[h][i].x, s[h]);
          s[h] = fmaf(kv.y, qr[h][i].y, s[h]);
          s[h] = fmaf(kv.z, qr[h][i].z, s[h]);
          s[h] = fmaf(kv.w, qr[h][i].w, s[h]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < MAX_QPK; ++h) {
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 1);
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 2);
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 4);
      if (part == 0 && j < cnt && h < qpk) sc[h][j] = (start + j >= lo) ? s[h] * scale : -INFINITY;
    }
  }
  __syncthreads();

  // ---- phase 2: warp h does the local softmax of head h
  if (warp < qpk) {
    const int h = warp;
    float mx = -INFINITY;
    for (int t = lane; t < cnt; t += 32) mx = fmaxf(mx, sc[h][t]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = lane; t < cnt; t += 32) {
      const float e = expf(sc[h][t] - mx);
      sc[h][t] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) {
      const size_t idx = (((size_t)m * p.n_head + g * qpk + h) * p.max_splits + split) * 2;
      p.ml_part[idx] = mx;
      p.ml_part[idx + 1] = sum;
    }
  }
  __syncthreads();

  // ---- phase 3: P @ V.  warp w owns keys [w*C/4, (w+1)*C/4); LPK lanes span one key row (float4 each)
  constexpr int LPK = HS / 4;    // lanes per key row: 32 / 16 / 8
  constexpr int KPI = 32 / LPK;  // keys per warp iteration: 1 / 2 / 4
  const int ksub = lane / LPK, d4 = lane - ksub * LPK;
  float4 acc[MAX_QPK];
#pragma unroll
  for (int h = 0; h < MAX_QPK; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int t0 = warp * (C / 4);
#pragma unroll 4
  for (int tt = 0; tt < C / 4; tt += KPI) {
    const int t = t0 + tt + ksub;
    if (t < cnt) {
      const float4 v = *reinterpret_cast<const float4*>(Vs + (size_t)t * HS + d4 * 4);
#pragma unroll
      for (int h = 0; h < MAX_QPK; ++h) {
        const float w = (h < qpk) ? sc[h][t] : 0.f;
        acc[h].x = fmaf(w, v.x, acc[h].x);
        acc[h].y = fmaf(w, v.y, acc[h].y);
        acc[h].z = fmaf(w, v.z, acc[h].z);
        acc[h].w = fmaf(w, v.w, acc[h].w);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < MAX_QPK; ++h) {
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) {
      acc[h].x += __shfl_xor_sync(0xffffffffu, acc[h].x, o);
      acc[h].y += __shfl_xor_sync(0xffffffffu, acc[h].y, o);
      acc[h].z += __shfl_xor_sync(0xffffffffu, acc[h].z, o);
      acc[h].w += __shfl_xor_sync(0xffffffffu, acc[h].w, o);
    }
    if (ksub == 0) *reinterpret_cast<float4*>(&redp[warp][h][d4 * 4]) = acc[h];
  }
  __syncthreads();
  for (int i = tid; i < qpk * HS; i += 128) {
    const int h = i / HS, d = i - h * HS;
    const float o = (redp[0][h][d] + redp[1][h][d]) + (redp[2][h][d] + redp[3][h][d]);
    p.o_part[(((size_t)m * p.n_head + g * qpk + h) * p.max_splits + split) * HS + d] = o;
  }
}

// stand-alone merge of the split partials -> y (M, n_head*hs); the handle path fuses this into PRO_ATTN instead
__global__ void attn_combine_kernel(const AttnParams p, float* y) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const int n_s = p.n_splits_launch > 0 ? p.n_splits_launch : (p.pos[m] + ATTN_CHUNK) / ATTN_CHUNK;  // empties weigh 0
  const int D = p.n_head * p.hs;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const int hh = k / p.hs, d = k - hh * p.hs;
    const size_t base = ((size_t)m * p.n_head + hh) * p.max_splits;
    float mx = -INFINITY;
    for (int s = 0; s < n_s; ++s) mx = fmaxf(mx, p.ml_part[(base + s) * 2]);
    float den = 0.f, num = 0.f;
    for (int s = 0; s < n_s; ++s) {
      const float w = __expf(p.ml_part[(base + s) * 2] - mx);
      if (w > 0.f) {
        den += w * p.ml_part[(base + s) * 2 + 1];
        num += w * p.o_part[(base + s) * p.hs + d];
      }
    }
    y[(size_t)m * D + k] = num / den;
  }
}

template <int HS>
cudaError_t launch_attn_hs(const LaunchCtx& lc, const AttnParams& p) {
  const size_t smem = (size_t)2 * ATTN_CHUNK * HS * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_split_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = prefer_max_smem(attn_split_kernel<HS>);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const dim3 grid(p.n_splits_launch, p.n_groups, p.M), block(128);
  return launch(lc, attn_split_kernel<HS>, grid, block, smem, p);
}

}  // namespace

cudaError_t launch_attn(const LaunchCtx& lc, const AttnParams& p) {
  if (p.n_head % p.n_groups != 0 || p.n_head / p.n_groups > MAX_QPK) return cudaErrorInvalidValue;
  switch (p.hs) {
    case 128: return launch_attn_hs<128>(lc, p);
    case 64: return launch_attn_hs<64>(lc, p);
    case 32: return launch_attn_hs<32>(lc, p);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_attn_combine(const LaunchCtx& lc, const AttnParams& p, float* y) {
  static bool once = false;
  if (!once) {
    prefer_max_smem(attn_combine_kernel);
    once = true;
  }
  return launch(lc, attn_combine_kernel, dim3(p.M), dim3(256), 0, p, y);
}

}  // namespace ua2
