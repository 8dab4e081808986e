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
    smem_bar_init(&bar, 1);
    smem_bar_fence_init();
    const uint32_t bytes = (uint32_t)span * HS * 4u;
    smem_bar_arrive_expect_tx(&bar, 2u * bytes);
    if (bytes > 0) {
      bulk_copy_g2s(Ks, Kc, bytes, &bar);
      bulk_copy_g2s(Vs, Vc, bytes, &bar);
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
  smem_bar_wait(&bar, 0u);
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

// ---------------------------------------------------------------------------------------------------------------------------
// Option "attn_ring" (default 1): the same split-softmax work items for batched contexts of moderate length, where the one-shot
// kernel's CTAs are too short-lived to keep the copy engine busy (a CTA fetches, waits, computes and exits).  Here a persistent CTA walks a
// contiguous range of the (row, group, split) items and streams their K and V chunks through a 3-slot ring of bulk copies:
// while the scores of item j are computed from K_j, V_j and K_{j+1} are in flight; every freed slot is refilled at the next
// __syncthreads.  Slots hold ONE chunk (K or V, 32 KB at hs 128), so two CTAs fit an SM with 4 chunks in flight between them.
// Partial outputs and statistics land exactly where attn_split_kernel puts them (same item -> (m, head, split) map, same
// arithmetic order inside an item), so the consumers (PRO_ATTN, attn_combine_kernel) are unchanged.
constexpr int RING_SLOTS = 3;
#ifndef UA2_ATTN_RING_MAX_SPLITS
#define UA2_ATTN_RING_MAX_SPLITS 16
#endif
#ifndef UA2_ATTN_RING_MIN_ITEMS
#define UA2_ATTN_RING_MIN_ITEMS (4 * 148)  // work items from which launch_attn takes the ring kernel (tests/cpu_shim lowers it)
#endif

struct RingItem {
  int m, g, split, start, cnt, lo;
  bool empty;
};

__device__ __forceinline__ RingItem ring_item(const AttnParams& p, int it) {
  RingItem r;
  const int nsl = p.n_splits_launch;
  r.split = it % nsl;
  const int mg = it / nsl;
  r.g = mg % p.n_groups;
  r.m = mg / p.n_groups;
  const int n_keys = p.pos[r.m] + 1;
  r.start = r.split * ATTN_CHUNK;
  r.lo = (p.window > 0) ? max(0, n_keys - p.window) : 0;
  r.empty = r.start >= n_keys || r.start + ATTN_CHUNK <= r.lo;
  r.cnt = min(ATTN_CHUNK, n_keys - r.start);
  return r;
}

template <int HS>
__global__ void __launch_bounds__(128) attn_ring_kernel(const AttnParams p, int n_items) {
  extern __shared__ __align__(128) float ring_s[];  // RING_SLOTS x [C][HS]
  __shared__ float sc[MAX_QPK][ATTN_CHUNK];
  __shared__ __align__(16) float redp[4][MAX_QPK][HS];
  __shared__ __align__(8) uint64_t full[RING_SLOTS];
  constexpr int C = ATTN_CHUNK;
  constexpr int NI = HS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int it0 = (int)((long long)n_items * blockIdx.x / gridDim.x), it1 = (int)((long long)n_items * (blockIdx.x + 1) / gridDim.x);
  const int qpk = p.n_head / p.n_groups;
  const float scale = rsqrtf((float)HS);
  pdl_launch_dependents();
  pdl_wait();

  // producer state (thread 0 only): the next copy is the K (half 0) or V (half 1) chunk of item p_it, sequence number p_seq
  int p_it = it0, p_half = 0;
  unsigned p_seq = 0;
  auto issue = [&]() {
    while (p_half == 0 && p_it < it1 && ring_item(p, p_it).empty) ++p_it;
    if (p_it >= it1) return;
    const RingItem r = ring_item(p, p_it);
    const int b = p.bidx_identity ? r.m : p.bidx[r.m];
    const float* src = (p_half == 0 ? p.k_cache : p.v_cache) + (((size_t)b * p.n_groups + r.g) * p.S_max + r.start) * HS;
    const unsigned slot = p_seq % RING_SLOTS;
    const uint32_t bytes = (uint32_t)r.cnt * HS * 4u;  // only the visible rows of the chunk
    smem_bar_arrive_expect_tx(&full[slot], bytes);
    bulk_copy_g2s(ring_s + (size_t)slot * C * HS, src, bytes, &full[slot]);
    ++p_seq;
    if (p_half == 0) {
      p_half = 1;
    } else {
      p_half = 0;
      ++p_it;
    }
  };
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < RING_SLOTS; ++s) smem_bar_init(&full[s], 1);
    smem_bar_fence_init();
#pragma unroll
    for (int s = 0; s < RING_SLOTS; ++s) issue();
  }
  __syncthreads();  // barrier init visible to every waiter

  const int kk = lane >> 3, part = lane & 7;
  float4 qr[MAX_QPK][NI];
  int q_m = -1, q_g = -1;
  unsigned c_seq = 0;  // sequence number of the next chunk this CTA consumes
  for (int it = it0; it < it1; ++it) {
    const RingItem r = ring_item(p, it);
    if (r.empty) {  // publish weight-0 statistics so consumers can merge blindly
      if (tid < qpk) {
        const size_t idx = (((size_t)r.m * p.n_head + r.g * qpk + tid) * p.max_splits + r.split) * 2;
        p.ml_part[idx] = -INFINITY;
        p.ml_part[idx + 1] = 0.f;
      }
      continue;
    }
    if (r.m != q_m || r.g != q_g) {  // a contiguous item range changes (row, group) once per n_splits_launch items
#pragma unroll
      for (int h = 0; h < MAX_QPK; ++h)
#pragma unroll
        for (int i = 0; i < NI; ++i)
          qr[h][i] = (h < qpk) ? *reinterpret_cast<const float4*>(p.q + (size_t)r.m * p.n_head * HS + (r.g * qpk + h) * HS + part * 4 + 32 * i)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
      q_m = r.m;
      q_g = r.g;
    }
    const int cnt = r.cnt;
    const float* Ks = ring_s + (size_t)(c_seq % RING_SLOTS) * C * HS;
    smem_bar_wait(&full[c_seq % RING_SLOTS], (c_seq / RING_SLOTS) & 1u);

    // ---- phase 1: scores (as attn_split_kernel)
#pragma unroll
    for (int i16 = 0; i16 < C / 16; ++i16) {
      const int j = i16 * 16 + warp * 4 + kk;
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
        if (part == 0 && j < cnt && h < qpk) sc[h][j] = (r.start + j >= r.lo) ? s[h] * scale : -INFINITY;
      }
    }
    __syncthreads();       // the K slot is free
    if (tid == 0) issue();

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
        const size_t idx = (((size_t)r.m * p.n_head + r.g * qpk + h) * p.max_splits + r.split) * 2;
        p.ml_part[idx] = mx;
        p.ml_part[idx + 1] = sum;
      }
    }
    __syncthreads();
    const float* Vs = ring_s + (size_t)((c_seq + 1) % RING_SLOTS) * C * HS;
    smem_bar_wait(&full[(c_seq + 1) % RING_SLOTS], ((c_seq + 1) / RING_SLOTS) & 1u);

    // ---- phase 3: P @ V (as attn_split_kernel)
    constexpr int LPK = HS / 4;
    constexpr int KPI = 32 / LPK;
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
    __syncthreads();       // the V slot is free; redp complete
    if (tid == 0) issue();
    for (int i = tid; i < qpk * HS; i += 128) {
      const int h = i / HS, d = i - h * HS;
      const float o = (redp[0][h][d] + redp[1][h][d]) + (redp[2][h][d] + redp[3][h][d]);
      p.o_part[(((size_t)r.m * p.n_head + r.g * qpk + h) * p.max_splits + r.split) * HS + d] = o;
    }
    c_seq += 2;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Many query rows per sequence (forward_prefix, lit_model.py:468-532 with T > 1; the codec transformer, transformer.py:375-419):
// the kernels above give every query row its own CTAs, so a 540-row prompt re-fetches each K/V chunk 540 times from L2 (the codec
// transformer spent 18 % of an encode + decode there).  Here a CTA owns a tile of consecutive query rows of ONE head and walks the
// keys once: K/V tiles of 32 keys are staged in shared memory for all its rows, scores by 128-bit shared reads, online softmax in
// registers, causal / window mask per element.  Rows of a tile that belong to different sequences (bidx) are served one sequence
// after the other.  The result is written in the split-partial format with everything in split 0 (weight 1) and the other splits
// empty, so the consumers (attn_combine_kernel, PRO_ATTN) stay as they are.
constexpr int AR_KEYS = 32;

template <int HS, int RPW>
__global__ void __launch_bounds__(256) attn_rows_kernel(const AttnParams p) {
  constexpr int DPL = HS / 32;  // output dims per lane
  constexpr int ROWS = 8 * RPW;
  constexpr int KST = HS + 4;
  __shared__ __align__(16) float Qs[ROWS][HS];
  __shared__ __align__(16) float Ks[AR_KEYS][KST];
  __shared__ __align__(16) float Vs[AR_KEYS][HS];
  __shared__ __align__(16) float Ps[8][RPW][AR_KEYS];
  __shared__ int s_pos[ROWS], s_b[ROWS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * ROWS, h = blockIdx.y;
  const int qpk = p.n_head / p.n_groups, g = h / qpk;
  const int D = p.n_head * HS;
  pdl_launch_dependents();
  pdl_wait();
  if (tid < ROWS) {
    const int m = m0 + tid;
    s_pos[tid] = m < p.M ? p.pos[m] : -1;
    s_b[tid] = m < p.M ? (p.bidx_identity ? m : p.bidx[m]) : -1;
  }
  for (int i = tid; i < ROWS * HS / 4; i += 256) {
    const int r = i / (HS / 4), d4 = i - r * (HS / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < p.M) v = *reinterpret_cast<const float4*>(p.q + (size_t)(m0 + r) * D + h * HS + d4 * 4);
    *reinterpret_cast<float4*>(&Qs[r][d4 * 4]) = v;
  }
  __syncthreads();
  const float scale = rsqrtf((float)HS);
  float mx[RPW], l[RPW], acc[RPW][DPL];
  int rpos[RPW], rlo[RPW], rb[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    mx[r] = -INFINITY;
    l[r] = 0.f;
    rpos[r] = s_pos[warp * RPW + r];
    rb[r] = s_b[warp * RPW + r];
    rlo[r] = (p.window > 0) ? max(0, rpos[r] + 1 - p.window) : 0;
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[r][d] = 0.f;
  }
  int seg = 0;
  while (seg < ROWS && s_b[seg] >= 0) {
    const int b = s_b[seg];
    int seg_end = seg + 1, j_hi = s_pos[seg], j_lo = (p.window > 0) ? max(0, s_pos[seg] + 1 - p.window) : 0;
    while (seg_end < ROWS && s_b[seg_end] == b) {
      j_hi = max(j_hi, s_pos[seg_end]);
      j_lo = min(j_lo, (p.window > 0) ? max(0, s_pos[seg_end] + 1 - p.window) : 0);
      ++seg_end;
    }
    const float* Kb = p.k_cache + ((size_t)b * p.n_groups + g) * (size_t)p.S_max * HS;
    const float* Vb = p.v_cache + ((size_t)b * p.n_groups + g) * (size_t)p.S_max * HS;
    for (int j0 = (j_lo / AR_KEYS) * AR_KEYS; j0 <= j_hi; j0 += AR_KEYS) {
      __syncthreads();  // previous tile fully consumed
      for (int i = tid; i < AR_KEYS * HS / 4; i += 256) {
        const int j = i / (HS / 4), d4 = i - j * (HS / 4);
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (j0 + j <= j_hi) {
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
      const int j = j0 + lane;
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const bool valid = rb[r] == b && j <= rpos[r] && j >= rlo[r];
        const float sv = valid ? s[r] * scale : -INFINITY;
        const float mn = fmaxf(mx[r], warp_max(sv));
        float pr = 0.f;
        if (mn > -INFINITY) {  // (a row with no visible key in the tiles so far keeps its empty state)
          const float corr = expf(mx[r] - mn);
          pr = valid ? expf(sv - mn) : 0.f;
          l[r] = l[r] * corr + warp_sum(pr);
#pragma unroll
          for (int d = 0; d < DPL; ++d) acc[r][d] *= corr;
          mx[r] = mn;
        }
        Ps[warp][r][lane] = pr;
      }
      __syncwarp();
      // ---- P @ V: lane owns output dims lane + 32 * d
#pragma unroll 2
      for (int j4 = 0; j4 < AR_KEYS / 4; ++j4) {
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
    seg = seg_end;
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int m = m0 + warp * RPW + r;
    if (m < p.M) {
      const size_t base = ((size_t)m * p.n_head + h) * p.max_splits;
      const float inv = 1.f / l[r];
#pragma unroll
      for (int d = 0; d < DPL; ++d) p.o_part[base * HS + lane + 32 * d] = acc[r][d] * inv;
      for (int sidx = lane; sidx < p.n_splits_launch; sidx += 32) {  // split 0 carries everything with weight exp(0) * 1
        p.ml_part[(base + sidx) * 2] = sidx == 0 ? 0.f : -INFINITY;
        p.ml_part[(base + sidx) * 2 + 1] = sidx == 0 ? 1.f : 0.f;
      }
    }
  }
}

// stand-alone merge of the split partials -> y (M, n_head*hs); the handle path fuses this into PRO_ATTN instead
__global__ void attn_combine_kernel(const AttnParams p, float* y) {  // grid (M, n_head): one CTA per (row, head)
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x, hh = blockIdx.y;  // rows on x: a batched prefill can exceed the 65535 limit of grid.y
  const int n_s = p.n_splits_launch > 0 ? p.n_splits_launch : (p.pos[m] + ATTN_CHUNK) / ATTN_CHUNK;  // empties weigh 0
  const int D = p.n_head * p.hs;
  const size_t base = ((size_t)m * p.n_head + hh) * p.max_splits;
  for (int d = threadIdx.x; d < p.hs; d += blockDim.x) {
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
    y[(size_t)m * D + hh * p.hs + d] = num / den;
  }
}

template <int HS>
cudaError_t launch_attn_hs(const LaunchCtx& lc, const AttnParams& p) {
  const size_t smem = (size_t)2 * ATTN_CHUNK * HS * sizeof(float);
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(attn_split_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = prefer_max_smem(attn_split_kernel<HS>);
    if (e != cudaSuccess) return e;
  }
  const dim3 grid(p.n_splits_launch, p.n_groups, p.M), block(128);
  return launch(lc, attn_split_kernel<HS>, grid, block, smem, p);
}

template <int HS>
cudaError_t launch_attn_ring_hs(const LaunchCtx& lc, const AttnParams& p, int n_items) {
  const size_t smem = (size_t)RING_SLOTS * ATTN_CHUNK * HS * sizeof(float);
  static DeviceOnce attr_set;
  static int n_sm = 148;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(attn_ring_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = prefer_max_smem(attn_ring_kernel<HS>);
    if (e != cudaSuccess) return e;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int per_sm = HS == 128 ? 2 : 4;  // 105 KB / 57 KB / 31 KB of shared memory per CTA
  const int grid = std::min(n_items, n_sm * per_sm);
  return launch(lc, attn_ring_kernel<HS>, dim3(grid), dim3(128), smem, p, n_items);
}

template <int HS, int RPW>
cudaError_t launch_attn_rows_hs(const LaunchCtx& lc, const AttnParams& p) {
  static DeviceOnce once;
  if (once.need()) {
    prefer_max_smem(attn_rows_kernel<HS, RPW>);
  }
  return launch(lc, attn_rows_kernel<HS, RPW>, dim3((p.M + 8 * RPW - 1) / (8 * RPW), p.n_head), dim3(256), 0, p);
}

int g_attn_rows = 1;  // many-row launches (>= 128 rows that share sequences) on the row-tile kernel
int g_attn_ring = 1;  // default since round 2 (measured on a B200: profiles/r2_kernel_rooflines.md)

}  // namespace

void set_attn_rows(int v) { g_attn_rows = v ? 1 : 0; }
int get_attn_rows() { return g_attn_rows; }
void set_attn_ring(int v) { g_attn_ring = v ? 1 : 0; }
int get_attn_ring() { return g_attn_ring; }

cudaError_t launch_attn(const LaunchCtx& lc, const AttnParams& p) {
  if (p.n_head % p.n_groups != 0 || p.n_head / p.n_groups > MAX_QPK) return cudaErrorInvalidValue;
  // prefill passes / the codec transformer: rows share sequences, so a tile of rows can share its K / V tiles.  (M >= 128 also
  // guarantees that the consumer merges the partials with attn_combine_kernel, which never reads the data of an empty split.)
  if (g_attn_rows && p.M >= 128 && !p.bidx_identity && p.n_head <= 65535) {
    switch (p.hs) {
      case 128: return launch_attn_rows_hs<128, 2>(lc, p);
      case 64: return launch_attn_rows_hs<64, 4>(lc, p);
      case 32: return launch_attn_rows_hs<32, 4>(lc, p);
      default: return cudaErrorInvalidValue;
    }
  }
  const long long n_items = (long long)p.M * p.n_groups * p.n_splits_launch;
  // The ring pays off once every CTA walks several items AND the chunks are few per (row, group): batch 32 x 540 keys runs at 0.41
  // of the HBM peak on the ring against 0.24 one-shot; at 2048 keys (32 splits) the one-shot grid already keeps 0.77 of the peak
  // in flight and the ring's two co-resident CTAs per SM (0.67) lose.  A decode frame at batch 1 (8-32 items) stays one-shot.
  if (g_attn_ring && n_items >= UA2_ATTN_RING_MIN_ITEMS && n_items < (1LL << 30) && p.n_splits_launch <= UA2_ATTN_RING_MAX_SPLITS) {
    switch (p.hs) {
      case 128: return launch_attn_ring_hs<128>(lc, p, (int)n_items);
      case 64: return launch_attn_ring_hs<64>(lc, p, (int)n_items);
      case 32: return launch_attn_ring_hs<32>(lc, p, (int)n_items);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (p.hs) {
    case 128: return launch_attn_hs<128>(lc, p);
    case 64: return launch_attn_hs<64>(lc, p);
    case 32: return launch_attn_hs<32>(lc, p);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_attn_combine(const LaunchCtx& lc, const AttnParams& p, float* y) {
  static DeviceOnce once;
  if (once.need()) {
    prefer_max_smem(attn_combine_kernel);
  }
  return launch(lc, attn_combine_kernel, dim3(p.M, p.n_head), dim3(std::min(128, p.hs)), 0, p, y);
}

}  // namespace ua2
