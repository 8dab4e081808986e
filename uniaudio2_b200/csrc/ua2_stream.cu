// Moshi-family streaming transformer (ring KV cache) and sample_token sampler.
//
// Replaces, for fp32 inference, the modules BASELINE.json's north_star names for the AR decode (SURVEY.md section 0,
// section 8 row a15; paths relative to the reference root):
//   llm_modules/transformer.py  StreamingTransformer.forward :671-692, StreamingTransformerLayer :545-588,
//                               StreamingMultiheadAttention.forward :375-419, RingKVCache.complete :242-278,
//                               multi_linear :155-179, _rms_norm :34-46, create_sin_embedding :126-152
//   llm_modules/gating.py       ActivationGating / gating_forward_kernel :12-51
//   llm_modules/rope.py         apply_rope :11-68 (interleaved pairs)
//   llm_utils/sampling.py       sample_token :84-105, sample_top_k :49-61, sample_top_p :64-81, multinomial :15-46
//
// The linears are the skinny / tiled fp32 family of ua2_gemv*.cu / ua2_sgemm.cu (weight streamers, HBM bound) with the
// norm fused into the prologue and residual / LayerScale / SwiGLU / GELU into the epilogue; per-step weights
// (multi_linear) are pointer offsets into the stacked weight, one launch per time step.  New here:
//   rope_ring_append_kernel  q/k rotation + append of k, v at slot pos % capacity          (activations only)
//   ring_attn_kernel         one CTA per (head, query row, split of <= ~256 slots): keys are read once with coalesced
//                            128-bit loads (a group of hs/4 lanes per key, two keys in flight per group), online softmax
//                            per group, merged through shared memory; rings longer than 256 slots are split across CTAs
//                            and merged by ring_attn_combine_kernel.  Slot positions are recovered from end_offset exactly
//                            like RingKVCache.complete, including its `delta <= 0` quirk.
//                            Bytes per launch = 2 * B * H * min(end, cap) * hs * 4 (HBM/L2 bound).
//   sample_token_kernel      one CTA per row: softmax statistics, then plain / top-k (4-pass radix select + rank
//                            counting, one noise draw per RANK like torch.topk + multinomial) / top-p (bitonic sort in
//                            shared memory, double-precision running sum like ATen's CPU cumsum).  Bytes = 4 * V per pass.
#include <map>
#include <string>
#include <vector>

#include "../../include/ua2_b200.h"
#include "ua2_kernels.cuh"
#include "ua2_philox.cuh"

namespace ua2 {
namespace {

// ------------------------------------------------------------------------------------------------ ring attention
struct RingAttnParams {
  const float* q;   // (M, H*hs)
  const float* kc;  // (B, H, cap, hs)
  const float* vc;
  const int32_t* pos;
  const int32_t* bidx;
  float* y;  // (M, H*hs)
  int M, H, cap;
  long long end;  // keys written so far, this call's included
  const long long* end_ptr;  // non-null: read `end` from device memory instead (captured graphs stay valid as it advances)
  int ring, causal, context;
  // split-softmax over the slots (long rings): CTA z handles slots [z * per_split, (z + 1) * per_split) and, when
  // n_splits > 1, writes un-normalised partials (max, sum, acc) that ring_attn_combine_kernel merges
  int n_splits, per_split;
  float* part_ml;   // (M, H, n_splits, 2)
  float* part_acc;  // (M, H, n_splits, hs)
};

// RingKVCache.complete, transformer.py:254-276 (ring) / KVCacheResult.from_kv :200-205 (linear)
__device__ __forceinline__ long long slot_position(int j, long long end, int cap, int ring) {
  if ((long long)j >= end) return -1;
  if (!ring) return j;
  const int end_index = (int)(end % cap);
  const int delta = j - end_index;
  return delta <= 0 ? end + delta : end + delta - cap;
}

constexpr int RA_WARPS = 8;
constexpr int RA_UNROLL = 2;         // keys per group per iteration (all their loads are issued before the first use)
constexpr int RA_SPLIT_KEYS = 256;   // slots per CTA from which the ring is split across CTAs
constexpr int RA_MAX_SPLITS = 16;

template <int HS>
__global__ void __launch_bounds__(RA_WARPS * 32) ring_attn_kernel(const RingAttnParams p) {
  constexpr int LPK = HS / 4;          // lanes per key row (one float4 each): 8 / 16 / 32
  constexpr int KPW = 32 / LPK;        // keys per warp step: 4 / 2 / 1
  constexpr int NG = RA_WARPS * KPW;   // independent online-softmax groups in the CTA
  __shared__ float s_m[NG], s_l[NG];
  __shared__ __align__(16) float s_acc[NG][HS];
  const int h = blockIdx.x, m = blockIdx.y, z = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = lane / LPK, d4 = lane - sub * LPK;
  const int grp = warp * KPW + sub;
  pdl_launch_dependents();
  pdl_wait();
  const int b = p.bidx[m];
  const long long pq = p.pos[m];
  const long long end = p.end_ptr ? *p.end_ptr : p.end;
  const float4 qv = *reinterpret_cast<const float4*>(p.q + ((size_t)m * p.H + h) * HS + d4 * 4);
  const float scale = rsqrtf((float)HS);
  const float* Kb = p.kc + ((size_t)b * p.H + h) * (size_t)p.cap * HS;
  const float* Vb = p.vc + ((size_t)b * p.H + h) * (size_t)p.cap * HS;
  const int j_lo = z * p.per_split, j_hi = min(p.cap, j_lo + p.per_split);
  float mx = -INFINITY, l = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = j_lo; j0 < j_hi; j0 += RA_UNROLL * NG) {
    bool vis[RA_UNROLL];
    float4 kv[RA_UNROLL], vv[RA_UNROLL];
#pragma unroll
    for (int u = 0; u < RA_UNROLL; ++u) {
      const int j = j0 + u * NG + grp;
      vis[u] = false;
      if (j < j_hi) {
        const long long pk = slot_position(j, end, p.cap, p.ring);
        const long long delta = pq - pk;
        vis[u] = pk >= 0 && (!p.causal || (delta >= 0 && (p.context <= 0 || delta < p.context)));
      }
      kv[u] = vv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vis[u]) {  // uniform over the LPK lanes of a group
        kv[u] = *reinterpret_cast<const float4*>(Kb + (size_t)j * HS + d4 * 4);
        vv[u] = *reinterpret_cast<const float4*>(Vb + (size_t)j * HS + d4 * 4);
      }
    }
#pragma unroll
    for (int u = 0; u < RA_UNROLL; ++u) {
      float s = 0.f;
      s = fmaf(kv[u].x, qv.x, s);
      s = fmaf(kv[u].y, qv.y, s);
      s = fmaf(kv[u].z, qv.z, s);
      s = fmaf(kv[u].w, qv.w, s);
#pragma unroll
      for (int o = LPK >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (vis[u]) {
        s *= scale;
        const float mn = fmaxf(mx, s);
        const float c = expf(mx - mn);  // first key: exp(-inf) = 0
        const float e = expf(s - mn);
        l = fmaf(l, c, e);
        acc.x = fmaf(acc.x, c, e * vv[u].x);
        acc.y = fmaf(acc.y, c, e * vv[u].y);
        acc.z = fmaf(acc.z, c, e * vv[u].z);
        acc.w = fmaf(acc.w, c, e * vv[u].w);
        mx = mn;
      }
    }
  }
  if (d4 == 0) {
    s_m[grp] = mx;
    s_l[grp] = l;
  }
  *reinterpret_cast<float4*>(&s_acc[grp][d4 * 4]) = acc;
  __syncthreads();
  for (int d = tid; d < HS; d += RA_WARPS * 32) {
    float gm = -INFINITY;
#pragma unroll
    for (int g = 0; g < NG; ++g) gm = fmaxf(gm, s_m[g]);
    float den = 0.f, num = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (s_m[g] > -INFINITY) {
        const float w = expf(s_m[g] - gm);
        den = fmaf(w, s_l[g], den);
        num = fmaf(w, s_acc[g][d], num);
      }
    }
    const size_t row = (size_t)m * p.H + h;
    if (p.n_splits <= 1) {
      // a row without any visible key (e.g. T = capacity: the first query's own key sits in the slot that reports a future
      // position) gives 0 like torch's SDPA, which routes fully masked rows through its safe softmax
      p.y[row * HS + d] = den > 0.f ? num / den : 0.f;
    } else {
      p.part_acc[(row * p.n_splits + z) * HS + d] = num;
      if (d == 0) {
        p.part_ml[(row * p.n_splits + z) * 2] = gm;
        p.part_ml[(row * p.n_splits + z) * 2 + 1] = den;
      }
    }
  }
}

// merge of the split partials: one thread per output element
__global__ void ring_attn_combine_kernel(const RingAttnParams p, int hs) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)p.M * p.H * hs) return;
  const size_t row = (size_t)(idx / hs);
  const int d = (int)(idx - (long long)row * hs);
  float gm = -INFINITY;
  for (int z = 0; z < p.n_splits; ++z) gm = fmaxf(gm, p.part_ml[(row * p.n_splits + z) * 2]);
  float den = 0.f, num = 0.f;
  for (int z = 0; z < p.n_splits; ++z) {
    const float mz = p.part_ml[(row * p.n_splits + z) * 2];
    if (mz > -INFINITY) {
      const float w = expf(mz - gm);
      den = fmaf(w, p.part_ml[(row * p.n_splits + z) * 2 + 1], den);
      num = fmaf(w, p.part_acc[(row * p.n_splits + z) * hs + d], num);
    }
  }
  p.y[idx] = den > 0.f ? num / den : 0.f;
}

// splits for a ring of `cap` slots (1 = the attention kernel writes y itself)
int ring_attn_splits(int cap) {
  int s = (cap + RA_SPLIT_KEYS - 1) / RA_SPLIT_KEYS;
  return s < 1 ? 1 : (s > RA_MAX_SPLITS ? RA_MAX_SPLITS : s);
}

// p.n_splits > 1 needs p.part_ml / p.part_acc; per_split is derived here
cudaError_t launch_ring_attn(const LaunchCtx& lc, RingAttnParams p, int hs) {
  if (p.n_splits < 1) p.n_splits = 1;
  if (p.n_splits > 1 && (p.part_ml == nullptr || p.part_acc == nullptr)) return cudaErrorInvalidValue;
  p.per_split = (p.cap + p.n_splits - 1) / p.n_splits;
  const dim3 grid(p.H, p.M, p.n_splits), block(RA_WARPS * 32);
  cudaError_t e;
  switch (hs) {
    case 128: e = launch(lc, ring_attn_kernel<128>, grid, block, 0, p); break;
    case 64: e = launch(lc, ring_attn_kernel<64>, grid, block, 0, p); break;
    case 32: e = launch(lc, ring_attn_kernel<32>, grid, block, 0, p); break;
    default: return cudaErrorInvalidValue;
  }
  if (e != cudaSuccess || p.n_splits == 1) return e;
  const long long n = (long long)p.M * p.H * hs;
  return launch(lc, ring_attn_combine_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, p, hs);
}

// ------------------------------------------------------------------------------------------------ RoPE + cache append
// one thread per (row, head, pair): rope.py:40-58 on q and k, then k / v to slot pos % cap (ring) or pos (linear)
__global__ void rope_ring_append_kernel(const float* __restrict__ qkv, int ld, const int32_t* __restrict__ pos,
                                        const int32_t* __restrict__ bidx, const float* __restrict__ freqs,
                                        float* __restrict__ q_out, float* __restrict__ kc, float* __restrict__ vc, int M, int H,
                                        int hs, int cap, int ring) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = hs >> 1;
  const long long total = (long long)M * H * half;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int i = (int)(idx % half);
  const int hh = (int)((idx / half) % H);
  const int m = (int)(idx / ((long long)half * H));
  const int C = H * hs;
  const float* row = qkv + (size_t)m * ld + hh * hs + 2 * i;
  const float2 q = *reinterpret_cast<const float2*>(row);
  const float2 k = *reinterpret_cast<const float2*>(row + C);
  const float2 v = *reinterpret_cast<const float2*>(row + 2 * C);
  const int ps = pos[m];
  float2 qo = q, ko = k;
  if (freqs != nullptr) {
    const float ang = __fmul_rn(freqs[i], (float)ps);
    const float rr = cosf(ang), ri = sinf(ang);
    qo.x = __fsub_rn(__fmul_rn(q.x, rr), __fmul_rn(q.y, ri));
    qo.y = __fadd_rn(__fmul_rn(q.x, ri), __fmul_rn(q.y, rr));
    ko.x = __fsub_rn(__fmul_rn(k.x, rr), __fmul_rn(k.y, ri));
    ko.y = __fadd_rn(__fmul_rn(k.x, ri), __fmul_rn(k.y, rr));
  }
  *reinterpret_cast<float2*>(q_out + (size_t)m * C + hh * hs + 2 * i) = qo;
  const int slot = ring ? ps % cap : ps;
  const size_t off = (((size_t)bidx[m] * H + hh) * cap + slot) * hs + 2 * i;
  *reinterpret_cast<float2*>(kc + off) = ko;
  *reinterpret_cast<float2*>(vc + off) = v;
}

// xt = x + positional_scale * [cos(pos / denom) | sin(pos / denom)]  (create_sin_embedding, transformer.py:126-152);
// denoms == nullptr: plain copy.  pos[m] = offset + m % T, bidx[m] = m / T and the ring's end_offset after this call are
// written on the way (per-call scalars live in device memory so that the captured layer graph stays valid).
__global__ void stx_begin_kernel(const float* __restrict__ x, float* __restrict__ xt, int32_t* __restrict__ pos,
                                 int32_t* __restrict__ bidx, const float* __restrict__ denoms, float pscale, int M, int T, int C,
                                 int offset, long long end_after, long long* __restrict__ d_end) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) *d_end = end_after;
  if (idx >= (long long)M * C) return;
  const int m = (int)(idx / C), c = (int)(idx - (long long)m * C);
  const int ps = offset + m % T;
  if (c == 0) {
    pos[m] = ps;
    bidx[m] = m / T;
  }
  float v = x[idx];
  if (denoms != nullptr) {
    const int half = C >> 1;
    const float phase = __fdiv_rn((float)ps, denoms[c < half ? c : c - half]);
    const float e = c < half ? cosf(phase) : sinf(phase);
    v = __fadd_rn(v, __fmul_rn(pscale, e));
  }
  xt[idx] = v;
}

// ------------------------------------------------------------------------------------------------ sample_token
constexpr int ST_THREADS = 1024;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_MAX_K = 1024;
constexpr int ST_MAX_SORT = 4096;

struct SampleTokenArgs {
  const float* logits;
  int V, use_sampling;
  float temp;
  int top_k;
  float top_p;
  int end_token;
  const float* noise;
  unsigned long long seed, offset;
  long long* out;
};

__device__ __forceinline__ uint32_t st_key(float x) {  // order-preserving float -> uint
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float st_unkey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}
__device__ __forceinline__ void st_better(float& bv, int& bi, float v, int i) {  // torch.argmax: first maximum
  if (v > bv || (v == bv && i < bi)) {
    bv = v;
    bi = i;
  }
}
__device__ float st_block_max(float v, float* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < ST_WARPS; ++w) r = fmaxf(r, sh[w]);
  return r;
}
__device__ float st_block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sh[0];
  for (int w = 1; w < ST_WARPS; ++w) r += sh[w];
  return r;
}
__device__ void st_block_argmax(float& v, int& i, float* shv, int* shi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    st_better(v, i, ov, oi);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    shv[threadIdx.x >> 5] = v;
    shi[threadIdx.x >> 5] = i;
  }
  __syncthreads();
  v = shv[0];
  i = shi[0];
  for (int w = 1; w < ST_WARPS; ++w) st_better(v, i, shv[w], shi[w]);
}

__global__ void __launch_bounds__(ST_THREADS) sample_token_kernel(const SampleTokenArgs a) {
  extern __shared__ __align__(16) unsigned char st_dyn[];  // top-p only: keys (Vp) + ids (Vp)
  __shared__ float sh_f[ST_WARPS];
  __shared__ int sh_i[ST_WARPS];
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int s_remaining, s_ncand, s_total_eq;
  __shared__ uint32_t cand_key[ST_MAX_K];
  __shared__ int cand_idx[ST_MAX_K];
  __shared__ int rank_idx[ST_MAX_K];
  static_assert(ST_MAX_K >= ST_THREADS, "eq_cnt aliases rank_idx");
  int* eq_cnt = rank_idx;  // per-thread tie counts: dead before the ranks are written
  const int tid = threadIdx.x;
  const int row = blockIdx.x, V = a.V;
  const float* lg = a.logits + (size_t)row * V;
  pdl_launch_dependents();
  pdl_wait();

  if (!a.use_sampling || !(a.temp > 0.f)) {  // sampling.py:103: torch.argmax(logits)
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += ST_THREADS) st_better(bv, bi, lg[i], i);
    st_block_argmax(bv, bi, sh_f, sh_i);
    if (tid == 0) a.out[row] = bi;
    return;
  }

  // probs = softmax(logits / temp)
  float mx = -INFINITY;
  for (int i = tid; i < V; i += ST_THREADS) mx = fmaxf(mx, __fdiv_rn(lg[i], a.temp));
  mx = st_block_max(mx, sh_f);
  float sum = 0.f;
  for (int i = tid; i < V; i += ST_THREADS) sum += expf(__fdiv_rn(lg[i], a.temp) - mx);
  sum = st_block_sum(sum, sh_f);
  auto prob = [&](int i) -> float {
    if (a.end_token >= 0 && i >= a.end_token) return -INFINITY;  // sampling.py:120
    return __fdiv_rn(expf(__fdiv_rn(lg[i], a.temp) - mx), sum);
  };
  const int n_noise = (a.top_k > 0 && !(a.top_p > 0.f)) ? a.top_k : V;
  auto noise = [&](int r) -> float {
    return a.noise ? a.noise[(size_t)row * n_noise + r]
                   : philox_exp1(a.seed, a.offset, (uint32_t)((size_t)row * n_noise + r));
  };

  if (a.top_p > 0.f) {  // sample_top_p, sampling.py:64-81
    int Vp = 1;
    while (Vp < V) Vp <<= 1;
    uint32_t* skey = reinterpret_cast<uint32_t*>(st_dyn);
    int* sidx = reinterpret_cast<int*>(st_dyn) + Vp;
    for (int i = tid; i < Vp; i += ST_THREADS) {
      skey[i] = i < V ? st_key(prob(i)) : 0u;
      sidx[i] = i < V ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= Vp; k <<= 1) {  // bitonic sort: descending probability, ascending id among equals
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < Vp; i += ST_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const uint32_t ka = skey[i], kb = skey[ixj];
            const int ia = sidx[i], ib = sidx[ixj];
            const bool a_first = ka > kb || (ka == kb && ia < ib);
            const bool up = (i & k) == 0;
            if (up ? !a_first : a_first) {
              skey[i] = kb;
              skey[ixj] = ka;
              sidx[i] = ib;
              sidx[ixj] = ia;
            }
          }
        }
        __syncthreads();
      }
    }
    float* sp = reinterpret_cast<float*>(skey);
    if (tid == 0) {  // torch.cumsum on CPU accumulates fp32 inputs in double and rounds every prefix to fp32
      double cum = 0.0;
      for (int r = 0; r < V; ++r) {
        const float pr = st_unkey(skey[r]);
        cum += (double)pr;
        const bool masked = __fsub_rn((float)cum, pr) > a.top_p;
        sp[r] = masked ? __fmul_rn(pr, 0.f) : pr;
      }
    }
    __syncthreads();
    float part = 0.f;
    for (int r = tid; r < V; r += ST_THREADS) part += sp[r];
    const float norm = st_block_sum(part, sh_f);
    float bv = -INFINITY;
    int br = 0x7fffffff;
    for (int r = tid; r < V; r += ST_THREADS) st_better(bv, br, __fdiv_rn(__fdiv_rn(sp[r], norm), noise(r)), r);
    st_block_argmax(bv, br, sh_f, sh_i);
    if (tid == 0) a.out[row] = sidx[br];
    return;
  }

  if (a.top_k <= 0) {  // multinomial over the whole row, sampling.py:15-46
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += ST_THREADS) st_better(bv, bi, __fdiv_rn(prob(i), noise(i)), i);
    st_block_argmax(bv, bi, sh_f, sh_i);
    if (tid == 0) a.out[row] = bi;
    return;
  }

  // ---- sample_top_k, sampling.py:49-61: exact k-th largest probability by a 4-pass MSB-first radix select
  const int k = a.top_k;
  if (tid == 0) {
    s_prefix = 0u;
    s_mask = 0u;
    s_remaining = k;
    s_ncand = 0;
  }
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    for (int i = tid; i < V; i += ST_THREADS) {
      const uint32_t key = st_key(prob(i));
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    }
    __syncthreads();
    if (tid == 0) {
      const int rem = s_remaining;
      int cum = 0, bsel = 255;
      for (; bsel > 0; --bsel) {
        if (cum + hist[bsel] >= rem) break;
        cum += hist[bsel];
      }
      s_remaining = rem - cum;
      s_total_eq = hist[bsel];  // after the last pass: how many entries carry exactly the threshold key
      s_prefix = prefix | ((uint32_t)bsel << shift);
      s_mask = mask | (0xFFu << shift);
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;      // key of the k-th largest probability
  const int need_eq = s_remaining;    // how many entries equal to it belong to the top k (lowest ids first)
  const int n_gt = k - need_eq;
  const bool ties_cut = s_total_eq > need_eq;  // only some of the entries equal to the threshold belong to the top k
  for (int i = tid; i < V; i += ST_THREADS) {
    const uint32_t key = st_key(prob(i));
    if (key > thr || (!ties_cut && key == thr)) {
      const int slot = atomicAdd(&s_ncand, 1);
      if (slot < ST_MAX_K) {
        cand_key[slot] = key;
        cand_idx[slot] = i;
      }
    }
  }
  if (ties_cut) {  // the lowest ids among the tied entries: contiguous chunk per thread + exclusive prefix of the counts
    const int chunk = (V + ST_THREADS - 1) / ST_THREADS;
    const int lo = min(V, tid * chunk), hi = min(V, lo + chunk);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += st_key(prob(i)) == thr;
    eq_cnt[tid] = c;
    __syncthreads();
    int r = 0;
    for (int t = 0; t < tid; ++t) r += eq_cnt[t];
    for (int i = lo; i < hi && r < need_eq; ++i) {
      if (st_key(prob(i)) == thr) {
        cand_key[n_gt + r] = thr;
        cand_idx[n_gt + r] = i;
        ++r;
      }
    }
  }
  __syncthreads();
  // rank of every candidate in torch.topk's descending order; the noise is indexed by rank (sampling.py:59-60)
  float bv = -INFINITY;
  int br = 0x7fffffff;
  for (int c = tid; c < k; c += ST_THREADS) {
    const uint32_t kc = cand_key[c];
    const int ic = cand_idx[c];
    int rank = 0;
    for (int c2 = 0; c2 < k; ++c2) {
      const uint32_t k2 = cand_key[c2];
      rank += (k2 > kc) || (k2 == kc && cand_idx[c2] < ic);
    }
    rank_idx[rank] = ic;
    st_better(bv, br, __fdiv_rn(st_unkey(kc), noise(rank)), rank);
  }
  st_block_argmax(bv, br, sh_f, sh_i);  // (its barriers also publish rank_idx)
  if (tid == 0) a.out[row] = rank_idx[br];
}

}  // namespace
}  // namespace ua2

using namespace ua2;

// ------------------------------------------------------------------------------------------------ handle
namespace {

struct StxLayer {
  const float *in_proj = nullptr, *out_proj = nullptr;
  const float *n1w = nullptr, *n1b = nullptr, *n2w = nullptr, *n2b = nullptr;
  const float *lin1 = nullptr, *lin2 = nullptr;
  std::vector<const float*> g_in, g_out;  // per step (one entry without weights_per_step)
  const float *ls1 = nullptr, *ls2 = nullptr;
};

int gating_hidden(int d_model, int dim_ff) {  // gating.py:40-43
  return dim_ff == 4 * d_model ? (21 * d_model) / 8 : (2 * dim_ff) / 3;
}

}  // namespace

struct ua2_stx {
  ua2_stx_cfg cfg{};
  std::vector<StxLayer> layers;
  std::vector<int> hidden;  // gating hidden size per step
  int n_steps = 1, f_max = 0;
  const float *rope_freqs = nullptr, *sin_denoms = nullptr;
  bool ready = false;
  // streaming state (_MHAState / _LayerState / _TransformerState)
  bool streaming = false;
  int sB = 0, cap = 0;
  long long offset = 0;
  std::vector<float*> kc, vc;
  std::vector<long long> end;
  // workspace for M rows
  size_t rows = 0;
  float *xt = nullptr, *qkv = nullptr, *q = nullptr, *att = nullptr, *hb = nullptr, *tk = nullptr, *tv = nullptr, *stats = nullptr;
  float *part_ml = nullptr, *part_acc = nullptr;  // split-softmax partials of the ring attention (RA_MAX_SPLITS per row)
  long long* d_end = nullptr;  // end_offset of the rings after the current call (read by the attention kernels)
  // streaming steps replay the layer stack as a CUDA graph with programmatic-dependent-launch edges, one graph per
  // (batch, T, weight step): first use of a shape runs eagerly (lazy kernel attributes), second use captures
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
  };
  std::map<unsigned long long, GraphEntry> graphs;
  int opt_graph = 1, opt_pdl = 1;
  int last_launches = 0;
  int32_t *pos = nullptr, *bidx = nullptr;
};

namespace {

#define RUN(expr)                  \
  do {                             \
    int _rc = (expr);              \
    if (_rc != UA2_OK) return _rc; \
  } while (0)

void clear_graphs(ua2_stx* h) {
  for (auto& kv : h->graphs)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->graphs.clear();
}

void free_ws(ua2_stx* h) {
  clear_graphs(h);  // captured graphs hold the workspace addresses
  for (void* p : {(void*)h->xt, (void*)h->qkv, (void*)h->q, (void*)h->att, (void*)h->hb, (void*)h->tk, (void*)h->tv,
                  (void*)h->stats, (void*)h->part_ml, (void*)h->part_acc, (void*)h->pos, (void*)h->bidx})
    if (p) cudaFree(p);
  h->xt = h->qkv = h->q = h->att = h->hb = h->tk = h->tv = h->stats = h->part_ml = h->part_acc = nullptr;
  h->pos = h->bidx = nullptr;
  h->rows = 0;
}

void free_rings(ua2_stx* h) {
  clear_graphs(h);
  for (float* p : h->kc)
    if (p) cudaFree(p);
  for (float* p : h->vc)
    if (p) cudaFree(p);
  h->kc.clear();
  h->vc.clear();
  h->end.clear();
}

int reserve_rows(ua2_stx* h, size_t M) {
  if (M <= h->rows) return UA2_OK;
  if (h->rows) UA2_CHECK_CUDA(cudaDeviceSynchronize());
  free_ws(h);
  const size_t C = h->cfg.d_model, F = h->f_max;
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->xt, M * C * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->qkv, M * 3 * C * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->q, M * C * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->att, M * C * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->hb, M * F * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->tk, M * C * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->tv, M * C * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->stats, (2 * M + 8) * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->part_ml, M * h->cfg.num_heads * RA_MAX_SPLITS * 2 * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->part_acc, M * C * RA_MAX_SPLITS * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->pos, M * 4));
  UA2_CHECK_CUDA(cudaMalloc((void**)&h->bidx, M * 4));
  h->rows = M;
  return UA2_OK;
}

// One linear of the layer over the (B, T) rows.  Without per-step weights: one launch of M = B*T rows.  With them
// (multi_linear, transformer.py:155-179): step t uses the slab `W + (offset + t) * slab` on the rows x[:, t] (row stride T*ld).
struct LinSpec {
  int pro = PRO_PLAIN, epi = EPI_STORE;
  const float *W = nullptr, *W2 = nullptr;
  size_t slab = 0;  // floats between consecutive per-step weights inside W (0: a single weight)
  int N = 0, K = 0;
  const float* X = nullptr;
  int ldx = 0;
  float* Y = nullptr;
  int ldy = 0;
  const float* R = nullptr;
  int ldr = 0;
  const float *scale = nullptr, *norm_w = nullptr, *norm_b = nullptr;
  float eps = 0.f;
};

int run_linear(ua2_stx* h, const LaunchCtx& lc, const LinSpec& s, int B, int T, bool per_step) {
  const int n_launch = per_step ? T : 1;
  for (int t = 0; t < n_launch; ++t) {
    GemvParams p;
    p.W = s.W + (per_step ? (size_t)(h->offset + t) * s.slab : 0);
    p.W2 = s.W2 ? s.W2 + (per_step ? (size_t)(h->offset + t) * s.slab : 0) : nullptr;
    p.N = s.N;
    p.K = s.K;
    p.M = per_step ? B : B * T;
    const int mul = per_step ? T : 1;
    p.X = s.X + (per_step ? (size_t)t * s.ldx : 0);
    p.ldx = s.ldx * mul;
    p.Y = s.Y + (per_step ? (size_t)t * s.ldy : 0);
    p.ldy = s.ldy * mul;
    if (s.R) {
      p.R = s.R + (per_step ? (size_t)t * s.ldr : 0);
      p.ldr = s.ldr * mul;
    }
    p.scale = s.scale;
    p.norm_w = s.norm_w;
    p.norm_b = s.norm_b;
    p.eps = s.eps;
    p.ws = h->stats;  // row statistics of the tiled many-row path
    p.ws_floats = 2 * h->rows + 8;
    UA2_CHECK_CUDA(launch_gemv(lc, s.pro, s.epi, p));
  }
  return UA2_OK;
}

bool parse_layer_key(const std::string& key, int& layer, std::string& rest) {
  const std::string pre = "layers.";
  if (key.compare(0, pre.size(), pre) != 0) return false;
  size_t i = pre.size(), j = i;
  while (j < key.size() && key[j] >= '0' && key[j] <= '9') ++j;
  if (j == i || j >= key.size() || key[j] != '.') return false;
  layer = std::stoi(key.substr(i, j - i));
  rest = key.substr(j + 1);
  return true;
}

// The layer stack over the M = B*T rows staged in h->xt (positions in h->pos / h->bidx, end_offset in h->d_end): every
// launch takes only device pointers and shape constants, so the sequence can be captured once per shape.
int run_layers(ua2_stx* h, const LaunchCtx& lc, int B, int T, long long offset) {
  const ua2_stx_cfg& c = h->cfg;
  const int C = c.d_model, H = c.num_heads, hs = C / H;
  const int M = B * T;
  const bool per_step = c.weights_per_step > 0;
  const bool rope = c.positional_embedding >= 2;
  const bool ln = c.norm <= 1;
  const float eps = (c.norm == 0 || c.norm == 2) ? 1e-5f : 1e-8f;  // create_norm_fn, transformer.py:111-121
  const int pro_norm = ln ? PRO_LAYERNORM : PRO_RMSNORM;
  const int epi_res = c.layer_scale ? EPI_SCALE_RESADD : EPI_RESADD;
  const long long saved_offset = h->offset;
  h->offset = offset;  // run_linear indexes the per-step slabs with it (0 when not streaming)
  int rc = UA2_OK;
  for (int li = 0; li < c.num_layers && rc == UA2_OK; ++li) {
    const StxLayer& l = h->layers[li];
    {  // norm1 -> in_proj
      LinSpec s;
      s.pro = pro_norm;
      s.epi = EPI_STORE;
      s.W = l.in_proj;
      s.slab = (size_t)3 * C * C;
      s.N = 3 * C;
      s.K = C;
      s.X = h->xt;
      s.ldx = C;
      s.Y = h->qkv;
      s.ldy = 3 * C;
      s.norm_w = l.n1w;
      s.norm_b = l.n1b;
      s.eps = eps;
      if ((rc = run_linear(h, lc, s, B, T, per_step)) != UA2_OK) break;
    }
    float* kc = h->streaming ? h->kc[li] : h->tk;
    float* vc = h->streaming ? h->vc[li] : h->tv;
    const int cap = h->streaming ? h->cap : T;
    {
      const long long total = (long long)M * H * (hs / 2);
      cudaError_t e = launch(lc, rope_ring_append_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                             (const float*)h->qkv, 3 * C, (const int32_t*)h->pos, (const int32_t*)h->bidx,
                             rope ? h->rope_freqs : (const float*)nullptr, h->q, kc, vc, M, H, hs, cap, h->streaming ? 1 : 0);
      if (e == cudaSuccess) {
        RingAttnParams a{};
        a.q = h->q;
        a.kc = kc;
        a.vc = vc;
        a.pos = h->pos;
        a.bidx = h->bidx;
        a.y = h->att;
        a.M = M;
        a.H = H;
        a.cap = cap;
        a.end_ptr = h->d_end;
        a.ring = h->streaming ? 1 : 0;
        a.causal = c.causal;
        a.context = c.context;
        a.n_splits = ring_attn_splits(cap);
        a.part_ml = h->part_ml;
        a.part_acc = h->part_acc;
        e = launch_ring_attn(lc, a, hs);
      }
      if (e != cudaSuccess) {
        set_error(std::string("attention launch: ") + cudaGetErrorString(e));
        rc = UA2_ERR_CUDA;
        break;
      }
    }
    {  // out_proj -> x + layer_scale_1 * update
      LinSpec s;
      s.pro = PRO_PLAIN;
      s.epi = epi_res;
      s.W = l.out_proj;
      s.slab = (size_t)C * C;
      s.N = C;
      s.K = C;
      s.X = h->att;
      s.ldx = C;
      s.Y = h->xt;
      s.ldy = C;
      s.R = h->xt;
      s.ldr = C;
      s.scale = l.ls1;
      if ((rc = run_linear(h, lc, s, B, T, per_step)) != UA2_OK) break;
    }
    if (!c.gating) {  // norm2 -> linear1 -> gelu -> linear2 -> x + layer_scale_2 * update
      LinSpec s;
      s.pro = pro_norm;
      s.epi = EPI_GELU;
      s.W = l.lin1;
      s.N = h->hidden[0];
      s.K = C;
      s.X = h->xt;
      s.ldx = C;
      s.Y = h->hb;
      s.ldy = h->f_max;
      s.norm_w = l.n2w;
      s.norm_b = l.n2b;
      s.eps = eps;
      if ((rc = run_linear(h, lc, s, B, T, false)) != UA2_OK) break;
      LinSpec o;
      o.pro = PRO_PLAIN;
      o.epi = epi_res;
      o.W = l.lin2;
      o.N = C;
      o.K = h->hidden[0];
      o.X = h->hb;
      o.ldx = h->f_max;
      o.Y = h->xt;
      o.ldy = C;
      o.R = h->xt;
      o.ldr = C;
      o.scale = l.ls2;
      if ((rc = run_linear(h, lc, o, B, T, false)) != UA2_OK) break;
    } else {  // ActivationGating: linear_in viewed (2, hidden): silu(first half) * second half, then linear_out
      const int n_launch = per_step ? T : 1;
      for (int t = 0; t < n_launch && rc == UA2_OK; ++t) {
        const int step = per_step ? (int)offset + t : 0;
        const int Hh = h->hidden[step];
        const int rows = per_step ? B : M, mul = per_step ? T : 1;
        GemvParams p;
        p.W = l.g_in[step];
        p.W2 = l.g_in[step] + (size_t)Hh * C;
        p.N = Hh;
        p.K = C;
        p.M = rows;
        p.X = h->xt + (size_t)t * C * (per_step ? 1 : 0);
        p.ldx = C * mul;
        p.Y = h->hb + (size_t)t * h->f_max * (per_step ? 1 : 0);
        p.ldy = h->f_max * mul;
        p.norm_w = l.n2w;
        p.norm_b = l.n2b;
        p.eps = eps;
        p.ws = h->stats;
        p.ws_floats = 2 * h->rows + 8;
        cudaError_t e = launch_gemv(lc, pro_norm, EPI_SWIGLU, p);
        if (e == cudaSuccess) {
          GemvParams o;
          o.W = l.g_out[step];
          o.N = C;
          o.K = Hh;
          o.M = rows;
          o.X = p.Y;
          o.ldx = p.ldy;
          o.Y = h->xt + (size_t)t * C * (per_step ? 1 : 0);
          o.ldy = C * mul;
          o.R = o.Y;
          o.ldr = o.ldy;
          o.scale = l.ls2;
          o.ws = h->stats;
          o.ws_floats = 2 * h->rows + 8;
          e = launch_gemv(lc, PRO_PLAIN, epi_res, o);
        }
        if (e != cudaSuccess) {
          set_error(std::string("gating launch: ") + cudaGetErrorString(e));
          rc = UA2_ERR_CUDA;
        }
      }
      if (rc != UA2_OK) break;
    }
  }
  h->offset = saved_offset;
  return rc;
}

}  // namespace

extern "C" {

int ua2_rope_ring_append_f32(const float* qkv, int ld_qkv, const int32_t* pos, const int32_t* bidx, const float* freqs,
                             float* q_out, float* k_cache, float* v_cache, int M, int H, int hs, int cap, int ring,
                             void* stream) {
  UA2_REQUIRE(qkv && pos && bidx && q_out && k_cache && v_cache, "null argument");
  UA2_REQUIRE(M >= 1 && H >= 1 && hs >= 2 && (hs % 2) == 0 && cap >= 1 && ld_qkv >= 3 * H * hs && (ld_qkv % 2) == 0,
              "need M, H, cap >= 1, even head size, ld_qkv >= 3*H*hs");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  const long long total = (long long)M * H * (hs / 2);
  UA2_CHECK_CUDA(launch(lc, rope_ring_append_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, qkv, ld_qkv, pos, bidx,
                        freqs, q_out, k_cache, v_cache, M, H, hs, cap, ring));
  return UA2_OK;
}

int ua2_ring_attn_f32(const float* q, const float* k_cache, const float* v_cache, const int32_t* pos, const int32_t* bidx,
                      float* y, int M, int H, int hs, int cap, int64_t end_offset, int ring, int causal, int context,
                      void* stream) {
  UA2_REQUIRE(q && k_cache && v_cache && pos && bidx && y, "null argument");
  UA2_REQUIRE(M >= 1 && M <= 65535 && H >= 1 && cap >= 1 && end_offset >= 0, "need 1 <= M <= 65535, H, cap >= 1");
  UA2_REQUIRE(hs == 32 || hs == 64 || hs == 128, "head size must be 32 / 64 / 128");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  RingAttnParams p{};
  p.q = q;
  p.kc = k_cache;
  p.vc = v_cache;
  p.pos = pos;
  p.bidx = bidx;
  p.y = y;
  p.M = M;
  p.H = H;
  p.cap = cap;
  p.end = end_offset;
  p.ring = ring;
  p.causal = causal;
  p.context = context;
  p.n_splits = 1;  // the operator form has no workspace: one CTA per (head, row) walks the whole ring
  UA2_CHECK_CUDA(launch_ring_attn(lc, p, hs));
  return UA2_OK;
}

int ua2_sample_token_f32(const float* logits, int R, int V, int use_sampling, float temp, int top_k, float top_p,
                         int end_token, const float* noise, uint64_t seed, uint64_t offset, int64_t* out, void* stream) {
  UA2_REQUIRE(logits && out, "null argument");
  UA2_REQUIRE(R >= 1 && V >= 1, "need R, V >= 1");
  const bool sampling = use_sampling && temp > 0.f;
  size_t smem = 0;
  if (sampling && top_p > 0.f) {
    UA2_REQUIRE(V <= ST_MAX_SORT, "top_p sampling is served up to 4096 entries per row");
    int Vp = 1;
    while (Vp < V) Vp <<= 1;
    smem = (size_t)Vp * 8;
  } else if (sampling && top_k > 0) {
    UA2_REQUIRE(top_k <= V, "selected index k out of range");  // torch.topk's error
    UA2_REQUIRE(top_k <= ST_MAX_K, "top_k sampling is served up to k = 1024");
  }
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  SampleTokenArgs a{logits, V, use_sampling, temp, top_k, top_p, end_token, noise, seed, offset, (long long*)out};
  UA2_CHECK_CUDA(launch(lc, sample_token_kernel, dim3(R), dim3(ST_THREADS), smem, a));
  return UA2_OK;
}

int ua2_stx_create(const ua2_stx_cfg* cfg, ua2_stx** out) {
  UA2_REQUIRE(cfg && out, "null argument");
  const ua2_stx_cfg& c = *cfg;
  UA2_REQUIRE(c.d_model >= 32 && c.num_heads >= 1 && c.d_model % c.num_heads == 0 && c.num_layers >= 1, "bad dimensions");
  const int hs = c.d_model / c.num_heads;
  UA2_REQUIRE(hs == 32 || hs == 64 || hs == 128, "head size d_model / num_heads must be 32 / 64 / 128");
  UA2_REQUIRE(c.d_model % 4 == 0, "d_model must be a multiple of 4");
  UA2_REQUIRE(c.positional_embedding >= 0 && c.positional_embedding <= 3, "positional_embedding: 0 none, 1 sin, 2 rope, 3 sin_rope");
  UA2_REQUIRE(c.norm >= 0 && c.norm <= 3, "Unknown norm type");  // create_norm_fn's ValueError
  UA2_REQUIRE(c.gating == 0 || c.gating == 1, "gating: only 'none' and 'silu' are served");
  UA2_REQUIRE(c.weights_per_step >= 0 && c.weights_per_step <= UA2_STX_MAX_STEPS, "weights_per_step out of range");
  UA2_REQUIRE(!(c.gating == 0 && c.weights_per_step), "weights_per_step without gating not supported for now.");  // :510
  UA2_REQUIRE(c.context >= 0, "context must be >= 0 (0 = None)");
  ua2_stx* h = new ua2_stx();
  h->cfg = c;
  h->n_steps = c.weights_per_step ? c.weights_per_step : 1;
  h->layers.resize(c.num_layers);
  for (int s = 0; s < h->n_steps; ++s) {
    const int ff = c.dim_feedforward[s] > 0 ? c.dim_feedforward[s] : c.dim_feedforward[0];
    if (ff <= 0) {
      delete h;
      UA2_REQUIRE(false, "dim_feedforward must be positive");
    }
    const int width = c.gating ? gating_hidden(c.d_model, ff) : ff;
    if (width < 4 || width % 4 != 0) {
      delete h;
      UA2_REQUIRE(false, "feed-forward width (hidden size of the gating, or dim_feedforward) must be a multiple of 4");
    }
    h->hidden.push_back(width);
    h->f_max = std::max(h->f_max, width);
  }
  for (auto& l : h->layers) {
    l.g_in.assign(h->n_steps, nullptr);
    l.g_out.assign(h->n_steps, nullptr);
  }
  *out = h;
  return UA2_OK;
}

int ua2_stx_destroy(ua2_stx* h) {
  if (!h) return UA2_OK;
  cudaDeviceSynchronize();
  free_ws(h);
  free_rings(h);
  if (h->d_end) cudaFree(h->d_end);
  delete h;
  return UA2_OK;
}

int ua2_stx_load_weight(ua2_stx* h, const char* key_c, const float* dptr, const int64_t* shape, int ndim) {
  UA2_REQUIRE(h && key_c && dptr && shape && ndim >= 1, "null argument");
  const std::string key(key_c);
  const ua2_stx_cfg& c = h->cfg;
  const int64_t D = c.d_model, mult = h->cfg.weights_per_step ? h->cfg.weights_per_step : 1;
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= shape[i];
  auto is2d = [&](int64_t r, int64_t cc) { return ndim == 2 && shape[0] == r && shape[1] == cc; };
  if (key == "rope_freqs") {
    UA2_REQUIRE(numel == D / c.num_heads / 2, "rope_freqs: expected head_size / 2 entries");
    h->rope_freqs = dptr;
    return UA2_OK;
  }
  if (key == "sin_denoms") {
    UA2_REQUIRE(numel == D / 2, "sin_denoms: expected d_model / 2 entries");
    h->sin_denoms = dptr;
    return UA2_OK;
  }
  int li = -1;
  std::string rest;
  UA2_REQUIRE(parse_layer_key(key, li, rest) && li >= 0 && li < c.num_layers, "unexpected key " + key);
  StxLayer& l = h->layers[li];
  const bool ln = c.norm <= 1;
  if (rest == "self_attn.in_proj_weight") {
    UA2_REQUIRE(is2d(mult * 3 * D, D), key + ": shape mismatch");
    l.in_proj = dptr;
  } else if (rest == "self_attn.out_proj.weight") {
    UA2_REQUIRE(is2d(mult * D, D), key + ": shape mismatch");
    l.out_proj = dptr;
  } else if (ln && (rest == "norm1.weight" || rest == "norm1.bias" || rest == "norm2.weight" || rest == "norm2.bias")) {
    UA2_REQUIRE(numel == D, key + ": shape mismatch");
    (rest == "norm1.weight" ? l.n1w : rest == "norm1.bias" ? l.n1b : rest == "norm2.weight" ? l.n2w : l.n2b) = dptr;
  } else if (!ln && (rest == "norm1.alpha" || rest == "norm2.alpha")) {
    UA2_REQUIRE(numel == D, key + ": shape mismatch");
    (rest == "norm1.alpha" ? l.n1w : l.n2w) = dptr;
  } else if (!c.gating && rest == "linear1.weight") {
    UA2_REQUIRE(is2d(h->hidden[0], D), key + ": shape mismatch");
    l.lin1 = dptr;
  } else if (!c.gating && rest == "linear2.weight") {
    UA2_REQUIRE(is2d(D, h->hidden[0]), key + ": shape mismatch");
    l.lin2 = dptr;
  } else if (c.layer_scale && (rest == "layer_scale_1.scale" || rest == "layer_scale_2.scale")) {
    UA2_REQUIRE(numel == D, key + ": shape mismatch");
    (rest == "layer_scale_1.scale" ? l.ls1 : l.ls2) = dptr;
  } else if (c.gating && rest.compare(0, 7, "gating.") == 0) {
    std::string g = rest.substr(7);
    int step = 0;
    if (c.weights_per_step) {  // gating.{s}.linear_in.weight
      size_t j = 0;
      while (j < g.size() && g[j] >= '0' && g[j] <= '9') ++j;
      UA2_REQUIRE(j > 0 && j < g.size() && g[j] == '.', "unexpected key " + key);
      step = std::stoi(g.substr(0, j));
      g = g.substr(j + 1);
      UA2_REQUIRE(step < h->n_steps, "unexpected key " + key);
    }
    if (g == "linear_in.weight") {
      UA2_REQUIRE(is2d(2 * h->hidden[step], D), key + ": shape mismatch");
      l.g_in[step] = dptr;
    } else if (g == "linear_out.weight") {
      UA2_REQUIRE(is2d(D, h->hidden[step]), key + ": shape mismatch");
      l.g_out[step] = dptr;
    } else {
      UA2_REQUIRE(false, "unexpected key " + key);
    }
  } else {
    UA2_REQUIRE(false, "unexpected key " + key);
  }
  return UA2_OK;
}

int ua2_stx_finalize(ua2_stx* h) {
  UA2_REQUIRE(h, "null handle");
  const ua2_stx_cfg& c = h->cfg;
  for (int i = 0; i < c.num_layers; ++i) {
    const StxLayer& l = h->layers[i];
    const std::string pre = "layers." + std::to_string(i) + ".";
    UA2_REQUIRE(l.in_proj && l.out_proj, "missing key " + pre + "self_attn.in_proj_weight / out_proj.weight");
    UA2_REQUIRE(l.n1w && l.n2w && (c.norm > 1 || (l.n1b && l.n2b)), "missing key " + pre + "norm1 / norm2 parameters");
    if (c.gating) {
      for (int s = 0; s < h->n_steps; ++s)
        UA2_REQUIRE(l.g_in[s] && l.g_out[s], "missing key " + pre + "gating linear_in / linear_out");
    } else {
      UA2_REQUIRE(l.lin1 && l.lin2, "missing key " + pre + "linear1.weight / linear2.weight");
    }
    if (c.layer_scale) UA2_REQUIRE(l.ls1 && l.ls2, "missing key " + pre + "layer_scale_1.scale / layer_scale_2.scale");
  }
  if (c.positional_embedding >= 2) UA2_REQUIRE(h->rope_freqs, "missing table rope_freqs");
  if (c.positional_embedding == 1 || c.positional_embedding == 3) UA2_REQUIRE(h->sin_denoms, "missing table sin_denoms");
  h->ready = true;
  return UA2_OK;
}

int ua2_stx_start_streaming(ua2_stx* h, int batch_size, void* stream) {
  UA2_REQUIRE(h && h->ready, "handle not finalized");
  UA2_REQUIRE(batch_size >= 1, "batch_size must be >= 1");
  const ua2_stx_cfg& c = h->cfg;
  int cap = c.context;
  if (cap == 0) {  // transformer.py:337-346
    UA2_REQUIRE(c.weights_per_step > 0, "Cannot create a streaming KVCache without a context to estimate capacity.");
    cap = c.weights_per_step;
  }
  cudaStream_t st = (cudaStream_t)stream;
  free_rings(h);
  const size_t bytes = (size_t)batch_size * c.d_model * cap * sizeof(float);
  for (int i = 0; i < c.num_layers; ++i) {
    float *k = nullptr, *v = nullptr;
    UA2_CHECK_CUDA(cudaMalloc((void**)&k, bytes));
    h->kc.push_back(k);
    UA2_CHECK_CUDA(cudaMalloc((void**)&v, bytes));
    h->vc.push_back(v);
    UA2_CHECK_CUDA(cudaMemsetAsync(k, 0, bytes, st));
    UA2_CHECK_CUDA(cudaMemsetAsync(v, 0, bytes, st));
    h->end.push_back(0);
  }
  h->streaming = true;
  h->sB = batch_size;
  h->cap = cap;
  h->offset = 0;
  return UA2_OK;
}

int ua2_stx_stop_streaming(ua2_stx* h) {
  UA2_REQUIRE(h, "null handle");
  if (h->streaming) UA2_CHECK_CUDA(cudaDeviceSynchronize());
  free_rings(h);
  h->streaming = false;
  h->sB = h->cap = 0;
  h->offset = 0;
  return UA2_OK;
}

int ua2_stx_reset_streaming(ua2_stx* h) {
  UA2_REQUIRE(h, "null handle");
  UA2_REQUIRE(h->streaming, "Trying to reset streaming, but the transformer wasn't streaming.");  // streaming.py:118-121
  h->offset = 0;
  for (auto& e : h->end) e = 0;  // RingKVCache.reset only rewinds end_offset; stale slots become invalid (position -1)
  return UA2_OK;
}

int ua2_stx_get_kv(ua2_stx* h, int layer, float** k, float** v, int64_t* end_offset, int* capacity) {
  UA2_REQUIRE(h && h->streaming, "not streaming");
  UA2_REQUIRE(layer >= 0 && layer < h->cfg.num_layers, "layer out of range");
  if (k) *k = h->kc[layer];
  if (v) *v = h->vc[layer];
  if (end_offset) *end_offset = h->end[layer];
  if (capacity) *capacity = h->cap;
  return UA2_OK;
}

int ua2_stx_forward(ua2_stx* h, const float* x, float* y, int B, int T, void* stream) {
  UA2_REQUIRE(h && h->ready, "handle not finalized");
  UA2_REQUIRE(x && y, "null argument");
  UA2_REQUIRE(B >= 1 && T >= 1 && (long long)B * T <= 65535, "need B, T >= 1 and B * T <= 65535 rows per call");
  const ua2_stx_cfg& c = h->cfg;
  const int C = c.d_model;
  const int M = B * T;
  const bool per_step = c.weights_per_step > 0;
  const long long offset = h->streaming ? h->offset : 0;
  if (h->streaming) {
    UA2_REQUIRE(c.causal, "Streaming only available for causal");  // transformer.py:381
    UA2_REQUIRE(B == h->sB, "batch size differs from the one streaming was started with");
    UA2_REQUIRE(T <= h->cap, "more time steps than the ring holds");
  }
  if (per_step) UA2_REQUIRE(offset + T <= c.weights_per_step, "time step beyond weights_per_step");
  RUN(reserve_rows(h, (size_t)M));
  if (!h->d_end) UA2_CHECK_CUDA(cudaMalloc((void**)&h->d_end, sizeof(long long)));
  int launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.launch_counter = &launches;
  const bool sin = c.positional_embedding == 1 || c.positional_embedding == 3;
  const long long end_after = h->streaming ? h->end[0] + T : T;
  const long long n_el = (long long)M * C;
  UA2_CHECK_CUDA(launch(lc, stx_begin_kernel, dim3((unsigned)((n_el + 255) / 256)), dim3(256), 0, x, h->xt, h->pos, h->bidx,
                        sin ? h->sin_denoms : (const float*)nullptr, c.positional_scale, M, T, C, (int)offset, end_after,
                        h->d_end));
  if (h->streaming && h->opt_graph) {
    lc.pdl = h->opt_pdl != 0;
    const unsigned long long key = ((unsigned long long)B << 40) | ((unsigned long long)T << 16) |
                                   ((unsigned long long)(per_step ? offset + 1 : 0) << 1) | (h->opt_pdl ? 1ull : 0ull);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {  // first use of this shape: eager (also sets the lazily-initialised kernel attributes)
      RUN(run_layers(h, lc, B, T, offset));
      h->graphs.emplace(key, ua2_stx::GraphEntry());
    } else {
      if (it->second.exec == nullptr) {  // second use: capture on a private stream
        cudaStream_t cs;
        UA2_CHECK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        LaunchCtx lcc = lc;
        lcc.stream = cs;
        int glaunches = 0;
        lcc.launch_counter = &glaunches;
        cudaGraph_t graph = nullptr;
        UA2_CHECK_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
        const int rc = run_layers(h, lcc, B, T, offset);
        const cudaError_t e = cudaStreamEndCapture(cs, &graph);
        if (rc != UA2_OK || e != cudaSuccess) {
          if (rc == UA2_OK) set_error(std::string("graph capture failed: ") + cudaGetErrorString(e));
          if (graph) cudaGraphDestroy(graph);
          cudaStreamDestroy(cs);
          return rc != UA2_OK ? rc : UA2_ERR_CUDA;
        }
        UA2_CHECK_CUDA(cudaGraphInstantiate(&it->second.exec, graph, 0));
        it->second.launches = glaunches;
        cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
      }
      UA2_CHECK_CUDA(cudaGraphLaunch(it->second.exec, lc.stream));
      launches += it->second.launches;
    }
  } else {
    RUN(run_layers(h, lc, B, T, offset));
  }
  UA2_CHECK_CUDA(cudaMemcpyAsync(y, h->xt, (size_t)M * C * sizeof(float), cudaMemcpyDeviceToDevice, lc.stream));
  h->last_launches = launches;
  if (h->streaming) {
    h->offset += T;
    for (auto& e : h->end) e += T;
  }
  return UA2_OK;
}

int ua2_stx_set_option(ua2_stx* h, const char* name, int value) {
  UA2_REQUIRE(h && name, "null argument");
  const std::string n(name);
  if (n == "graph")
    h->opt_graph = value;
  else if (n == "pdl")
    h->opt_pdl = value;
  else
    UA2_REQUIRE(false, "unknown option " + n);
  return UA2_OK;
}

int ua2_stx_last_launch_count(ua2_stx* h) { return h ? h->last_launches : 0; }

}  // extern "C"
