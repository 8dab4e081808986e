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

template <int HS>
__global__ void __launch_bounds__(128) attn_split_kernel(const AttnParams p) {
  __shared__ __align__(16) float qs[MAX_QPK][HS];
  __shared__ float sc[MAX_QPK][ATTN_CHUNK];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, g = blockIdx.y, m = blockIdx.z;
  pdl_launch_dependents();
  pdl_wait();
  const int n_keys = p.pos[m] + 1;
  const int start = split * ATTN_CHUNK;
  if (start >= n_keys) return;
  const int cnt = min(ATTN_CHUNK, n_keys - start);
  const int qpk = p.n_head / p.n_groups;
  const int b = p.bidx[m];
  const float scale = rsqrtf((float)HS);
  const float* Kc = p.k_cache + (((size_t)b * p.n_groups + g) * p.S_max + start) * HS;
  const float* Vc = p.v_cache + (((size_t)b * p.n_groups + g) * p.S_max + start) * HS;

  for (int i = tid; i < qpk * HS; i += 128) {
    const int h = i / HS, d = i - h * HS;
    qs[h][d] = p.q[(size_t)m * p.n_head * HS + (g * qpk + h) * HS + d];
  }
  __syncthreads();

  // ---- phase 1: one thread per key, scores for all q heads of the group
  if (tid < cnt) {
    const float4* kr = reinterpret_cast<const float4*>(Kc + (size_t)tid * HS);
    float s[MAX_QPK];
#pragma unroll
    for (int h = 0; h < MAX_QPK; ++h) s[h] = 0.f;
#pragma unroll 8
    for (int d4 = 0; d4 < HS / 4; ++d4) {
      const float4 kv = kr[d4];
#pragma unroll
      for (int h = 0; h < MAX_QPK; ++h) {
        if (h < qpk) {
          const float4 qv = *reinterpret_cast<const float4*>(&qs[h][d4 * 4]);
          s[h] = fmaf(kv.x, qv.x, s[h]);
          s[h] = fmaf(kv.y, qv.y, s[h]);
          s[h] = fmaf(kv.z, qv.z, s[h]);
          s[h] = fmaf(kv.w, qv.w, s[h]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < MAX_QPK; ++h)
      if (h < qpk) sc[h][tid] = s[h] * scale;
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

  // ---- phase 3: P @ V, thread per output dim (HS==128) or per (key-half, dim) (HS==64, HS==32)
  constexpr int PARTS = 128 / HS;
  const int part = tid / HS, d = tid - part * HS;
  float acc[MAX_QPK];
#pragma unroll
  for (int h = 0; h < MAX_QPK; ++h) acc[h] = 0.f;
#pragma unroll 4
  for (int t = part; t < cnt; t += PARTS) {
    const float v = Vc[(size_t)t * HS + d];
#pragma unroll
    for (int h = 0; h < MAX_QPK; ++h)
      if (h < qpk) acc[h] = fmaf(sc[h][t], v, acc[h]);
  }
  if (PARTS > 1) {
    // reduce the key-parts through shared memory (PARTS = 2 or 4)
    __shared__ float redp[4][MAX_QPK][HS];
#pragma unroll
    for (int h = 0; h < MAX_QPK; ++h) redp[part][h][d] = acc[h];
    __syncthreads();
    if (part == 0) {
#pragma unroll
      for (int h = 0; h < MAX_QPK; ++h) {
        float s = 0.f;
        for (int q = 0; q < PARTS; ++q) s += redp[q][h][d];
        acc[h] = s;
      }
    }
  }
  if (part == 0) {
#pragma unroll
    for (int h = 0; h < MAX_QPK; ++h)
      if (h < qpk) p.o_part[(((size_t)m * p.n_head + g * qpk + h) * p.max_splits + split) * HS + d] = acc[h];
  }
}

// stand-alone merge of the split partials -> y (M, n_head*hs); the handle path fuses this into PRO_ATTN instead
__global__ void attn_combine_kernel(const AttnParams p, float* y) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x;
  const int n_s = (p.pos[m] + ATTN_CHUNK) / ATTN_CHUNK;
  const int D = p.n_head * p.hs;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    const int hh = k / p.hs, d = k - hh * p.hs;
    const size_t base = ((size_t)m * p.n_head + hh) * p.max_splits;
    float mx = -INFINITY;
    for (int s = 0; s < n_s; ++s) mx = fmaxf(mx, p.ml_part[(base + s) * 2]);
    float den = 0.f, num = 0.f;
    for (int s = 0; s < n_s; ++s) {
      const float w = __expf(p.ml_part[(base + s) * 2] - mx);
      den += w * p.ml_part[(base + s) * 2 + 1];
      num += w * p.o_part[(base + s) * p.hs + d];
    }
    y[(size_t)m * D + k] = num / den;
  }
}

}  // namespace

cudaError_t launch_attn(const LaunchCtx& lc, const AttnParams& p) {
  if (p.n_head % p.n_groups != 0 || p.n_head / p.n_groups > MAX_QPK) return cudaErrorInvalidValue;
  const dim3 grid(p.n_splits_launch, p.n_groups, p.M), block(128);
  switch (p.hs) {
    case 128: return launch(lc, attn_split_kernel<128>, grid, block, 0, p);
    case 64: return launch(lc, attn_split_kernel<64>, grid, block, 0, p);
    case 32: return launch(lc, attn_split_kernel<32>, grid, block, 0, p);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_attn_combine(const LaunchCtx& lc, const AttnParams& p, float* y) {
  return launch(lc, attn_combine_kernel, dim3(p.M), dim3(256), 0, p, y);
}

}  // namespace ua2
