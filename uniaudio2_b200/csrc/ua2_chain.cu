// Persistent "chain" kernel: a whole transformer stack (or one local-decoder step) of a B = 1 AR frame in ONE launch.
//
// Why: the frame is ~356 dependent 3-35 us kernels; measured on B200 (profiles/r1_launches.md, r1_ncu_full_gemv2.md) the
// linears stream at 79-85 % of HBM peak in isolation but the frame only reaches 60-65 %, because every kernel boundary
// costs ~5 us of drain + launch + prologue + first weight round trip during which HBM idles.  Here the ops of a stack
// are executed by one cooperative grid of 2 CTAs per SM that never leaves the SMs:
//   * ops are separated by a grid barrier (monotonic counter in global memory) instead of a kernel boundary;
//   * each warp owns a 2 x 4 KB bulk-copy ring (cp.async.bulk + mbarrier) that lives across ops: as soon as a warp has
//     consumed its last weight chunk of op i it issues the first chunks of the NEXT linear's weights - before the
//     partial-sum exchange, the epilogue and the grid barrier of op i - so HBM keeps streaming through the barrier;
//   * op descriptors (weights, shapes, slab/K-split geometry) are static per model and live in global memory; per-frame
//     values (positions, tokens, masks) are read through the same device arrays as in the multi-kernel path.
// Replaces, for B = 1 decode, the per-op launches of run_block()/run_global()/run_heads() in ua2_llm.cu; the math and the
// device functions (prologues, epilogues, attention) are the ones of ua2_gemv3.cu / ua2_attn.cu, so results are
// bit-identical to the multi-kernel v3 path except for the attention split size (32 instead of 64 keys).
#include <cooperative_groups.h>

#include "ua2_chain.cuh"
#include "ua2_gemv3_dev.cuh"

namespace ua2 {
namespace {

using namespace v3dev;

constexpr int CW = 8;            // warps per CTA
constexpr int CST = 2;           // ring slots per warp
constexpr int CSLOT = 1024;      // floats per slot (4 KB bulk copies)
constexpr int CK_MAX = 8192;     // largest K of any linear (activation tile floats, MT = 1)
constexpr int C_ATT = CHAIN_ATTN_CHUNK;

struct Geom {  // where this warp sits in a linear op
  int rows, n_units, u_lo, slab, ngrp, grp, sl, ks, klen, nCh, per_unit, total;
};

__device__ __forceinline__ int rows_of(int epi) { return (epi == EPI_QKV || epi == EPI_SWIGLU) ? 2 : 1; }

__device__ __forceinline__ Geom make_geom(const ChainOp& op, int cta, int G, int warp) {
  Geom g;
  const GemvParams& p = op.g;
  g.rows = rows_of(op.epi);
  g.n_units = g.rows == 2 ? (op.epi == EPI_SWIGLU ? p.N : (p.N >> 1)) : p.N;
  g.u_lo = (int)(((long long)cta * g.n_units) / G);
  g.slab = (int)(((long long)(cta + 1) * g.n_units) / G) - g.u_lo;
  g.ngrp = op.c.ngrp;
  g.grp = warp / op.c.nsl;
  g.sl = warp - g.grp * op.c.nsl;
  g.ks = g.sl * op.c.SL;
  g.klen = max(0, min(p.K, g.ks + op.c.SL) - g.ks);
  g.nCh = (g.klen + op.c.KCW - 1) / op.c.KCW;
  g.per_unit = g.nCh * g.rows;
  const int my_units = (g.grp < g.ngrp && g.slab > g.grp) ? (g.slab - g.grp + g.ngrp - 1) / g.ngrp : 0;
  g.total = my_units * g.per_unit;
  return g;
}

__device__ __forceinline__ const float* unit_row_rt(const ChainOp& op, int u, int r, int& nA) {
  const GemvParams& p = op.g;
  if (op.epi == EPI_SWIGLU) {
    nA = u;
    return (r == 0 ? p.W : p.W2) + (size_t)u * p.K;
  } else if (op.epi == EPI_QKV) {
    const int half = p.hs >> 1;
    const int hh = u / half, i = u - hh * half;
    nA = hh * p.hs + i;
    return p.W + (size_t)(nA + r * half) * p.K;
  }
  nA = u;
  return p.W + (size_t)u * p.K;
}

// lane 0: issue chunk t of this warp's (unit, row, k-chunk) sequence into ring slot (base + t) % CST
__device__ __forceinline__ void issue_chunk(const ChainOp& op, const Geom& g, int t, uint32_t base, float* ring, uint64_t* bars) {
  const int uo = t / g.per_unit, rem = t - uo * g.per_unit;
  const int r = rem / g.nCh, ch = rem - r * g.nCh;
  int nA;
  const float* row = unit_row_rt(op, g.u_lo + g.grp + uo * g.ngrp, r, nA);
  const int k0 = g.ks + ch * op.c.KCW;
  const uint32_t bytes = (uint32_t)min(op.c.KCW, g.ks + g.klen - k0) * 4u;
  const uint32_t slot = (base + (uint32_t)t) % CST;
  mbar_expect_tx(&bars[slot], bytes);
  bulk_g2s(ring + (size_t)slot * CSLOT, row + k0, bytes, &bars[slot]);
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (cooperative, co-resident) grid meet here; `target` = arrivals expected so far
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (ld_acquire(counter) < target) {
      if (++spins > (1u << 27)) __trap();  // never hang the GPU: a lost CTA aborts the launch instead
    }
    __threadfence();
  }
  __syncthreads();
}

// ---- one linear op on M = 1 row (templated on the fused prologue / epilogue)
// Ring bookkeeping of one warp: sequence numbers of the chunks issued / consumed so far (across ops).  Chunk `seq` lives in
// slot seq % CST and completes phase (seq / CST) & 1 of that slot's mbarrier.  `pump` keeps up to CST chunks in flight,
// running from the current linear straight into the next one (look-ahead of one linear).
struct Ring {
  uint32_t issued, consumed;
};

__device__ __forceinline__ void pump(Ring& rg, const ChainOp* cur, const Geom* gc, uint32_t base, const ChainOp* nxt, const Geom* gn,
                                     float* ring, uint64_t* bars, int lane) {
  const uint32_t end_cur = base + (cur ? (uint32_t)gc->total : 0u);
  while (rg.issued < rg.consumed + CST) {
    const uint32_t sq = rg.issued;
    if (cur && sq < end_cur) {
      if (lane == 0) issue_chunk(*cur, *gc, (int)(sq - base), base, ring, bars);
    } else if (nxt && sq - end_cur < (uint32_t)gn->total) {
      if (lane == 0) issue_chunk(*nxt, *gn, (int)(sq - end_cur), end_cur, ring, bars);
    } else {
      break;
    }
    ++rg.issued;
  }
}

template <int PRO, int EPI>
__device__ __forceinline__ void run_gemv(const ChainOp& op, const Geom& g, float* xs, float* part, float* ring, uint64_t* bars,
                                         float (*red)[MAXW_RED], Ring& rg, const ChainOp* nxt, const Geom* gn,
                                         unsigned long long* pri) {
  constexpr int ROWS = RowsOf<EPI>::value;
  constexpr int MT = 1;
  const GemvParams& p = op.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, Kp = ((K + 127) >> 7) << 7;
  // attention splits of this frame (B = 1: all rows share the position): the producer used C_ATT-key splits
  const int n_splits = (PRO == PRO_ATTN) ? (p.pos[0] + C_ATT) / C_ATT : 1;
  const uint32_t base = rg.consumed;  // every chunk of earlier ops has been consumed
  pump(rg, &op, &g, base, nxt, gn, ring, bars, lane);
  stage_activations3<MT, PRO>(p, xs, red, Kp, 0, 1, n_splits);  // ends with __syncthreads()
  if (pri) pri[1] = clock64();

  float acc[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
  const int part_stride = ROUND_UNITS * op.c.nsl * ROWS;
  const int n_rounds = (g.slab + ROUND_UNITS - 1) / ROUND_UNITS;
  int t = 0;
  for (int rd = 0; rd < n_rounds; ++rd) {
    const int r_lo = rd * ROUND_UNITS, r_hi = min(g.slab, r_lo + ROUND_UNITS);
    float* pb = part + (rd & 1) * part_stride;
    while (t < g.total) {
      const int uo = t / g.per_unit, rem = t - uo * g.per_unit;
      const int ul = g.grp + uo * g.ngrp;
      if (ul >= r_hi) break;
      const int r = rem / g.nCh, ch = rem - r * g.nCh;
      const uint32_t seq = base + (uint32_t)t;
      const uint32_t slot = seq % CST;
      mbar_wait(&bars[slot], (seq / CST) & 1u);
      const float* sw = ring + (size_t)slot * CSLOT;
      const int k0 = g.ks + ch * op.c.KCW;
      const int kend = min(op.c.KCW, g.ks + g.klen - k0);
      float pa = 0.f;
#pragma unroll 4
      for (int kk = lane * 4; kk < kend; kk += 128) {
        const float4 w = *reinterpret_cast<const float4*>(sw + kk);
        const float4 xv = *reinterpret_cast<const float4*>(xs + k0 + kk);
        pa = fmaf(w.x, xv.x, pa);
        pa = fmaf(w.y, xv.y, pa);
        pa = fmaf(w.z, xv.z, pa);
        pa = fmaf(w.w, xv.w, pa);
      }
      __syncwarp();
      ++rg.consumed;  // slot free: refill it with the chunk CST ahead (this linear, then the next one's look-ahead)
      pump(rg, &op, &g, base, nxt, gn, ring, bars, lane);
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr)
        if (rr == r) acc[rr] += pa;
      if (rem == g.per_unit - 1) {
        float* dst = pb + ((size_t)(ul - r_lo) * op.c.nsl + g.sl) * ROWS;
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr) {
          const float s2 = warp_sum(acc[rr]);
          if (lane == 0) dst[rr] = s2;
          acc[rr] = 0.f;
        }
      }
      ++t;
    }
    if (pri && rd == n_rounds - 1) pri[2] = clock64();
    __syncthreads();
    const int n_ru = r_hi - r_lo;
    for (int ul = tid; ul < n_ru; ul += CW * 32) {
      float a = 0.f, b = 0.f;
      const float* src = pb + (size_t)ul * op.c.nsl * ROWS;
      for (int s = 0; s < op.c.nsl; ++s) {
        if (s * op.c.SL < K) {
          a += src[s * ROWS];
          if (ROWS == 2) b += src[s * ROWS + 1];
        }
      }
      if (PRO == PRO_RMSNORM) {
        float tot = 0.f;
        for (int w = 0; w < CW; ++w) tot += red[0][w];
        const float rs = rsqrtf(tot / (float)K + p.eps);
        a *= rs;
        b *= rs;
      }
      int nA;
      unit_row_rt(op, g.u_lo + r_lo + ul, 0, nA);
      epilogue_v3<EPI>(p, 0, a, b, nA);
    }
  }
  (void)warp;
}

// ---- attention for one (split, group) item with 128 threads (warps 0..3); smem: kv (2 x C_ATT x HS floats)
template <int HS>
__device__ __forceinline__ void run_attn_item(const AttnParams& p, int split, int g, float* kv_s, float (*sc)[C_ATT],
                                              float (*redp)[4][128], uint64_t* bar, uint32_t& bar_phase) {
  constexpr int C = C_ATT, NI = HS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = 0;
  const int n_keys = p.pos[m] + 1;
  const int start = split * C;
  const int qpk = p.n_head / p.n_groups;
  const int cnt = min(C, n_keys - start);  // caller guarantees start < n_keys
  const int b = p.bidx[m];
  const float scale = rsqrtf((float)HS);
  const float* Kc = p.k_cache + (((size_t)b * p.n_groups + g) * p.S_max + start) * HS;
  const float* Vc = p.v_cache + (((size_t)b * p.n_groups + g) * p.S_max + start) * HS;
  float* Ks = kv_s;
  float* Vs = kv_s + C * HS;
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)cnt * HS * 4u;
    mbar_expect_tx(bar, 2u * bytes);
    bulk_g2s(Ks, Kc, bytes, bar);
    bulk_g2s(Vs, Vc, bytes, bar);
  }
  const int kk = lane >> 3, part = lane & 7;
  float4 qr[4][NI];
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int i = 0; i < NI; ++i)
      qr[h][i] = (h < qpk) ? *reinterpret_cast<const float4*>(p.q + (size_t)m * p.n_head * HS + (g * qpk + h) * HS + part * 4 + 32 * i)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  mbar_wait(bar, bar_phase & 1u);
  bar_phase++;
#pragma unroll
  for (int it = 0; it < C / 16; ++it) {
    const int j = it * 16 + warp * 4 + kk;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < cnt) {
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const float4 kvv = *reinterpret_cast<const float4*>(Ks + (size_t)j * HS + part * 4 + 32 * i);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          s[h] = fmaf(kvv.x, qr[h][i].x, s[h]);
          s[h] = fmaf(kvv.y, qr[h][i].y, s[h]);
          s[h] = fmaf(kvv.z, qr[h][i].z, s[h]);
          s[h] = fmaf(kvv.w, qr[h][i].w, s[h]);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 1);
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 2);
      s[h] += __shfl_xor_sync(0xffffffffu, s[h], 4);
      if (part == 0 && j < cnt && h < qpk) sc[h][j] = s[h] * scale;
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
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
  asm volatile("bar.sync 1, 128;" ::: "memory");
  constexpr int LPK = HS / 4, KPI = 32 / LPK;
  const int ksub = lane / LPK, d4 = lane - ksub * LPK;
  float4 acc[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int t0 = warp * (C / 4);
#pragma unroll 4
  for (int tt = 0; tt < C / 4; tt += KPI) {
    const int t = t0 + tt + ksub;
    if (t < cnt) {
      const float4 v = *reinterpret_cast<const float4*>(Vs + (size_t)t * HS + d4 * 4);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float w = (h < qpk) ? sc[h][t] : 0.f;
        acc[h].x = fmaf(w, v.x, acc[h].x);
        acc[h].y = fmaf(w, v.y, acc[h].y);
        acc[h].z = fmaf(w, v.z, acc[h].z);
        acc[h].w = fmaf(w, v.w, acc[h].w);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 4; ++h) {
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) {
      acc[h].x += __shfl_xor_sync(0xffffffffu, acc[h].x, o);
      acc[h].y += __shfl_xor_sync(0xffffffffu, acc[h].y, o);
      acc[h].z += __shfl_xor_sync(0xffffffffu, acc[h].z, o);
      acc[h].w += __shfl_xor_sync(0xffffffffu, acc[h].w, o);
    }
    if (ksub == 0) *reinterpret_cast<float4*>(&redp[warp][h][d4 * 4]) = acc[h];
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  for (int i = tid; i < qpk * HS; i += 128) {
    const int h = i / HS, d = i - h * HS;
    const float o = (redp[0][h][d] + redp[1][h][d]) + (redp[2][h][d] + redp[3][h][d]);
    p.o_part[(((size_t)m * p.n_head + g * qpk + h) * p.max_splits + split) * HS + d] = o;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // smem reusable by the next item
}

// prof (nullable): per op, for CTA 0 and CTA G/2: clock64 at {op start, prologue done, weights consumed, op done, barrier passed}
__global__ void __launch_bounds__(CW * 32, 2) chain_kernel(const ChainOp* __restrict__ ops, int n_ops, unsigned* sync_ctr,
                                                           unsigned long long* prof) {
  extern __shared__ __align__(128) float smem[];
  __shared__ float red[8][MAXW_RED];
  __shared__ __align__(8) uint64_t bars[CW][CST];
  __shared__ __align__(8) uint64_t abar;
  __shared__ float sc[4][C_ATT];
  __shared__ __align__(16) float redp[4][4][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  float* xs = smem;                               // CK_MAX floats (also the attention K/V staging area)
  float* part = smem + CK_MAX;                    // 2 x ROUND_UNITS x 8 x 2
  float* ring = part + 2 * ROUND_UNITS * 8 * 2 + (size_t)warp * CST * CSLOT;
  if (lane == 0) {
    for (int s = 0; s < CST; ++s) mbar_init(&bars[warp][s], 1);
    if (warp == 0) mbar_init(&abar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  unsigned long long* pr = nullptr;  // this thread's profile row base (thread 0 of the two sampled CTAs)
  if (prof != nullptr && tid == 0 && (cta == 0 || cta == G / 2)) pr = prof + (size_t)(cta == 0 ? 0 : 1) * CHAIN_PROF_OPS * 5;
  Ring rg{0u, 0u};
  uint32_t aphase = 0;  // completed phases of the attention staging barrier
  unsigned arrivals = 0;

  // first linear of the chain: start its weights right away
  {
    int j = 0;
    while (j < n_ops && ops[j].type != OP_GEMV) ++j;
    if (j < n_ops) {
      const Geom gj = make_geom(ops[j], cta, G, warp);
      pump(rg, nullptr, nullptr, 0u, &ops[j], &gj, ring, bars[warp], lane);
    }
  }

  for (int i = 0; i < n_ops; ++i) {
    const ChainOp& op = ops[i];
    unsigned long long* pri = (pr != nullptr && i < CHAIN_PROF_OPS) ? pr + (size_t)i * 5 : nullptr;
    if (pri) pri[0] = clock64();
    if (op.type == OP_GEMV) {
      const Geom g = make_geom(op, cta, G, warp);
      const ChainOp* nxt = op.next_gemv >= 0 ? &ops[op.next_gemv] : nullptr;
      Geom gn;
      if (nxt) gn = make_geom(*nxt, cta, G, warp);
#define UA2_RUN(P, E) \
  if (op.pro == P && op.epi == E) run_gemv<P, E>(op, g, xs, part, ring, bars[warp], red, rg, nxt, nxt ? &gn : nullptr, pri);
      UA2_RUN(PRO_RMSNORM, EPI_QKV)
      else UA2_RUN(PRO_ATTN, EPI_RESADD)
      else UA2_RUN(PRO_RMSNORM, EPI_SWIGLU)
      else UA2_RUN(PRO_PLAIN, EPI_RESADD)
      else UA2_RUN(PRO_PLAIN, EPI_STORE)
      else UA2_RUN(PRO_RMSNORM, EPI_STORE)
      else UA2_RUN(PRO_GATHER, EPI_STORE)
#undef UA2_RUN
    } else if (op.type == OP_ATTN) {
      const AttnParams& a = op.a;
      const int n_keys = a.pos[0] + 1;
      const int n_sp = (n_keys + C_ATT - 1) / C_ATT;
      const int items = n_sp * a.n_groups;
      for (int it = cta; it < items; it += G) {
        const int split = it / a.n_groups, grp = it - split * a.n_groups;
        if (tid < 128) {
          if (a.hs == 128) run_attn_item<128>(a, split, grp, xs, sc, redp, &abar, aphase);
          else if (a.hs == 64) run_attn_item<64>(a, split, grp, xs, sc, redp, &abar, aphase);
          else run_attn_item<32>(a, split, grp, xs, sc, redp, &abar, aphase);
        }
      }
    } else if (op.type == OP_EMBED) {
      // model_new.py:598-604 on one row, spread over the CTAs
      const int D = op.D, nq = op.nq;
      for (int k = (cta * CW * 32 + tid) * 4; k < D; k += G * CW * 32 * 4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < nq; ++c) {
          const float4 e = *reinterpret_cast<const float4*>(op.audio_emb + ((size_t)op.tokens[c] + (size_t)c * op.V) * D + k);
          const float w = op.mask[c] ? 1.f : 0.f;
          acc.x += e.x * w;
          acc.y += e.y * w;
          acc.z += e.z * w;
          acc.w += e.w * w;
        }
        *reinterpret_cast<float4*>(op.audio_in + k) = acc;
        *reinterpret_cast<float4*>(op.text_emb + k) = *reinterpret_cast<const float4*>(op.wte + (size_t)op.tokens[nq] * D + k);
      }
    } else if (op.type == OP_NORMMIX) {
      // ln_f + mask mix between stacks (model_new.py:607, :610, :613); one row -> CTA 0 does it
      if (cta == 0) {
        const int D = op.D, nq = op.nq;
        float ss = 0.f;
        for (int k = tid * 4; k < D; k += CW * 32 * 4) {
          const float4 v = *reinterpret_cast<const float4*>(op.nm_x + k);
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        ss = warp_sum(ss);
        if (lane == 0) red[0][warp] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < CW; ++w) tot += red[0][w];
        const float rs = rsqrtf(tot / (float)D + op.nm_eps);
        const float ma = op.mask[0] ? 1.f : 0.f, mt = op.mask[nq] ? 1.f : 0.f;
        for (int k = tid * 4; k < D; k += CW * 32 * 4) {
          const float4 v = *reinterpret_cast<const float4*>(op.nm_x + k);
          const float4 gw = *reinterpret_cast<const float4*>(op.nm_w + k);
          float4 n;
          n.x = (v.x * rs) * gw.x;
          n.y = (v.y * rs) * gw.y;
          n.z = (v.z * rs) * gw.z;
          n.w = (v.w * rs) * gw.w;
          float4 o = n;
          if (op.nm_mode == MIX_UND_TO_BACKBONE || op.nm_mode == MIX_FINAL) {
            const float4 ad = *reinterpret_cast<const float4*>(op.nm_add + k);
            o.x = __fadd_rn(__fmul_rn(n.x, ma), __fmul_rn(ad.x, mt));
            o.y = __fadd_rn(__fmul_rn(n.y, ma), __fmul_rn(ad.y, mt));
            o.z = __fadd_rn(__fmul_rn(n.z, ma), __fmul_rn(ad.z, mt));
            o.w = __fadd_rn(__fmul_rn(n.w, ma), __fmul_rn(ad.w, mt));
          } else if (op.nm_mode == MIX_BACKBONE_TO_GEN) {
            *reinterpret_cast<float4*>(op.nm_keep + k) = n;
            o = make_float4(n.x * ma, n.y * ma, n.z * ma, n.w * ma);
          }
          *reinterpret_cast<float4*>(op.nm_out + k) = o;
        }
      }
    }
    arrivals += (unsigned)G;
    if (pri) pri[3] = clock64();
    if (i + 1 < n_ops) grid_barrier(&sync_ctr[0], arrivals);
    if (pri) pri[4] = clock64();
  }
  // leave the counters at zero for the next launch: the last CTA to get here resets them
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned old = atomicAdd(&sync_ctr[1], 1u);
    if (old == (unsigned)G - 1u) {
      // every CTA has left its last grid_barrier wait (it only increments sync_ctr[1] afterwards)
      sync_ctr[0] = 0u;
      sync_ctr[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace

size_t chain_smem_bytes() { return (size_t)(CK_MAX + 2 * ROUND_UNITS * 8 * 2 + CW * CST * CSLOT) * sizeof(float); }

// geometry of a linear inside the chain: 8 warps, K slices of <= 1024 floats (>= 4 KB copies whenever K allows)
void chain_cfg_for(int K, V3Cfg_public* out) {
  int nsl = (K + 1023) / 1024;
  if (nsl > 8) nsl = 8;
  if (nsl < 1) nsl = 1;
  if (nsl == 3) nsl = 4;  // 8 warps: 2 groups x 4 slices (768-float slices for K = 3072)
  if (nsl > 4 && nsl < 8) nsl = 8;
  const int ngrp = 8 / nsl;
  const int per = (K + nsl - 1) / nsl;
  const int SL = ((per + 127) / 128) * 128;
  out->nsl = nsl;
  out->ngrp = ngrp;
  out->SL = SL;
  out->KCW = SL < CSLOT ? SL : CSLOT;
  out->stages = CST;
  out->n_splits = 0;
}

cudaError_t launch_chain(cudaStream_t stream, const ChainOp* d_ops, int n_ops, unsigned* d_sync, int n_ctas,
                         unsigned long long* d_prof) {
  static bool once = false;
  const size_t smem = chain_smem_bytes();
  if (!once) {
    cudaError_t e = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    prefer_max_smem(chain_kernel);
    once = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(CW * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // guarantees co-residency of all CTAs (needed by the grid barrier)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, chain_kernel, d_ops, n_ops, d_sync, d_prof);
}

int chain_max_ctas() {
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem_bytes());
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_kernel, CW * 32, chain_smem_bytes());
  if (per_sm > 2) per_sm = 2;
  return sms * per_sm;
}

}  // namespace ua2
