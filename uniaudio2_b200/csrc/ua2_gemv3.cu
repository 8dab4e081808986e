// Skinny fp32 linear, v3: persistent CTAs (2 per SM), slab partition of the weight rows, K split across the warps of a
// CTA, per-warp bulk-copy weight rings.  The default linear of every decode frame (M < 128 rows).
//
// Why (measured, profiles/r1_launches.md): at M = 1 every linear of the AR frame is a 9-37 us weight stream and the frame
// has ~330 of them, so what is lost is not steady-state bandwidth but the bubble at every kernel boundary (drain ->
// launch -> activation prologue -> first weight round trip).  v3 keeps that bubble small:
//   * <= ~110 KB of shared memory and 6-8 warps per CTA, two CTAs per SM: under programmatic dependent launch the CTAs of
//     the NEXT kernel take over SM slots as this kernel's CTAs retire, fill their weight rings and park on
//     griddepcontrol.wait;
//   * each warp owns a ring of 3 shared-memory slots of 4 KB fed by cp.async.bulk (SASS UBLKCP) completing on per-slot
//     mbarriers - the copy size / depth the HBM microbenchmark calls for (profiles/r1_microbench_hbm_streaming.txt); the
//     ring is filled BEFORE griddepcontrol.wait / the activation prologue and refilled the moment a slot is consumed;
//   * work is balanced to one weight-row unit: CTA c owns the contiguous slab of units [c*U/G, (c+1)*U/G), the grid G is
//     chosen in [SMs, 2 SMs] for the best remainder balance, and the warps split K (nsl slices of ~1024 floats per unit);
//     partial sums meet in shared memory once per round of <= 32 units;
//   * the prologue is a single L2 round trip: every global load of a thread is issued before its first use
//     (ua2_gemv3_dev.cuh), RMSNorm's rsqrt(mean(x^2)+eps) is a per-row scalar applied in the epilogue
//     ((sum_k W[n,k] x[k] g[k]) * rs), the attention combine reads all launched splits at once, and the epilogue's
//     residual / position operands are fetched together with the activations.
#include "ua2_gemv3_dev.cuh"

namespace ua2 {
namespace {

using namespace v3dev;

// Tail L2 prefetch of the following linears' weights: measured SLOWER on B200 (1421 -> 1366 / 1306 / 1246 audio tok/s for
// 16 / 32 / 48 MB per boundary, profiles/r1_l2_prefetch_experiment.md), so the device side is compiled out by default;
// build with -DUA2_GEMV3_TAIL_PREFETCH=1 to reproduce.
#ifndef UA2_GEMV3_TAIL_PREFETCH
#define UA2_GEMV3_TAIL_PREFETCH 0
#endif

// fire-and-forget L2 prefetch of `bytes` (multiple of 16) starting at a 16-byte aligned global address
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// One lane per warp: prefetch this CTA's share of the following linears' first rows (see PfSpec).  CTA `cta` of a grid
// of G covers the slabs c2 = cta, cta + G, ... of each next kernel; the rows of a slab are dealt round-robin to warps.
__device__ __forceinline__ void tail_prefetch(const GemvParams& p, int cta, int G, int warp, int nwarps) {
  for (int j = 0; j < p.n_pf; ++j) {
    const PfSpec& f = p.pf[j];
    const uint32_t rbytes = (uint32_t)f.K * 4u;
    const int rows = f.mode == 0 ? 1 : 2;
    for (int c2 = cta; c2 < f.G; c2 += G) {
      const int u0 = (int)(((long long)c2 * f.n_units) / f.G), u1 = (int)(((long long)(c2 + 1) * f.n_units) / f.G);
      const int n = min(f.n, u1 - u0);
      for (int i = warp; i < n * rows; i += nwarps) {
        const int u = u0 + i / rows, r = i - (i / rows) * rows;
        const float* row;
        if (f.mode == 1) {
          row = (r == 0 ? f.W : f.W2) + (size_t)u * f.K;
        } else if (f.mode == 2) {
          const int half = f.hs >> 1;
          const int hh = u / half, ii = u - hh * half;
          row = f.W + (size_t)(hh * f.hs + ii + r * half) * f.K;
        } else {
          row = f.W + (size_t)u * f.K;
        }
        l2_prefetch_bulk(row, rbytes);
      }
    }
  }
}

template <int MT, int PRO, int EPI>
__global__ void __launch_bounds__(MAXW * 32, 1) gemv3_kernel(const GemvParams p, const V3Cfg c) {
  constexpr int ROWS = RowsOf<EPI>::value;
  extern __shared__ __align__(128) float smem3[];
  __shared__ float red[8][NWARPS];
  __shared__ __align__(8) uint64_t bars[MAXW][MAX_STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nthreads = blockDim.x;
  const int K = p.K;
  const int Kp = ((K + 127) >> 7) << 7;
  const int m0 = blockIdx.x * MT;
  const int mcount = min(MT, p.M - m0);
  const int n_units = (ROWS == 2) ? ((EPI == EPI_SWIGLU) ? p.N : (p.N >> 1)) : p.N;
  // slab of this CTA
  const int G = gridDim.y, cta = blockIdx.y;
  const int u_lo = (int)(((long long)cta * n_units) / G), u_hi = (int)(((long long)(cta + 1) * n_units) / G);
  const int slab = u_hi - u_lo;
  const int ngrp = c.ngrp;
  const int grp = warp / c.nsl, sl = warp - grp * c.nsl;
  const int ks = sl * c.SL;
  const int klen = max(0, min(K, ks + c.SL) - ks);
  const int nCh = (klen + c.KCW - 1) / c.KCW;
  const int per_unit = nCh * ROWS;                                        // bulk copies per unit for this warp
  const int my_units = slab > grp ? (slab - grp + ngrp - 1) / ngrp : 0;  // units grp, grp+ngrp, ... of the slab
  const int total = my_units * per_unit;

  float* xs = smem3;                       // MT x Kp
  float* part = smem3 + (size_t)MT * Kp;  // [2 buffers][ROUND_UNITS][nsl][ROWS][MT]
  const int part_stride = ROUND_UNITS * c.nsl * ROWS * MT;
  float* ring = part + 2 * part_stride + (size_t)warp * c.stages * c.KCW;

  auto issue = [&](int t) {
    const int uo = t / per_unit, rem = t - uo * per_unit;
    const int r = rem / nCh, ch = rem - r * nCh;
    int nA;
    const float* row = unit_row<EPI>(p, u_lo + grp + uo * ngrp, r, nA);
    const int k0 = ks + ch * c.KCW;
    const uint32_t bytes = (uint32_t)min(c.KCW, ks + klen - k0) * 4u;
    const int st = t % c.stages;
    uint64_t* bar = &bars[warp][st];
    mbar_expect_tx(bar, bytes);
    bulk_g2s(ring + (size_t)st * c.KCW, row + k0, bytes, bar);
  };

  if (lane == 0) {
    for (int s = 0; s < c.stages; ++s) mbar_init(&bars[warp][s], 1);
    fence_mbar_init();
    const int pre = min(total, c.stages);
    for (int t = 0; t < pre; ++t) issue(t);  // weights never depend on the producer kernel
#if UA2_GEMV3_TAIL_PREFETCH
    if (p.n_pf > 0 && total <= c.stages) tail_prefetch(p, cta, G, warp, nthreads >> 5);  // nothing more to issue
#endif
  }
  __syncwarp();
  pdl_launch_dependents();
  pdl_wait();

  int ps_row[MT];  // EPI_QKV: cache slot of each row, fetched with the activations instead of at the tail
#pragma unroll
  for (int m = 0; m < MT; ++m) ps_row[m] = (EPI == EPI_QKV && m < mcount) ? p.pos[m0 + m] : -1;
  int b_row[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) b_row[m] = (EPI == EPI_QKV && m < mcount) ? (p.bidx_identity ? m0 + m : p.bidx[m0 + m]) : -1;
  // EPI_RESADD: the residual element this thread will add in round 0 is fetched now, not at the tail
  float r_pre = 0.f;
  if (EPI == EPI_RESADD) {
    const int n_ru0 = min(slab, ROUND_UNITS);
    if (tid < n_ru0 * mcount) {
      const int ul = tid / mcount, m = tid - ul * mcount;
      r_pre = p.R[(size_t)(m0 + m) * p.ldr + (u_lo + ul)];
    }
  }
  stage_activations3<MT, PRO>(p, xs, red, Kp, m0, mcount, c.n_splits);  // ends with __syncthreads()

  float acc[ROWS][MT];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int m = 0; m < MT; ++m) acc[r][m] = 0.f;

  const int n_rounds = (slab + ROUND_UNITS - 1) / ROUND_UNITS;
  int t = 0;
  for (int rd = 0; rd < n_rounds; ++rd) {
    const int r_lo = rd * ROUND_UNITS, r_hi = min(slab, r_lo + ROUND_UNITS);  // slab-local units of this round
    float* pb = part + (rd & 1) * part_stride;
    while (t < total) {
      const int uo = t / per_unit, rem = t - uo * per_unit;
      const int ul = grp + uo * ngrp;
      if (ul >= r_hi) break;
      const int r = rem / nCh, ch = rem - r * nCh;
      const int st = t % c.stages;
      mbar_wait(&bars[warp][st], (uint32_t)((t / c.stages) & 1));
      const float* sw = ring + (size_t)st * c.KCW;
      const int k0 = ks + ch * c.KCW;
      const int kend = min(c.KCW, ks + klen - k0);
      float part_acc[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) part_acc[m] = 0.f;
#pragma unroll 4
      for (int kk = lane * 4; kk < kend; kk += 128) {
        const float4 w = *reinterpret_cast<const float4*>(sw + kk);
#pragma unroll
        for (int m = 0; m < MT; ++m) {
          const float4 xv = *reinterpret_cast<const float4*>(xs + m * Kp + k0 + kk);
          part_acc[m] = fmaf(w.x, xv.x, part_acc[m]);
          part_acc[m] = fmaf(w.y, xv.y, part_acc[m]);
          part_acc[m] = fmaf(w.z, xv.z, part_acc[m]);
          part_acc[m] = fmaf(w.w, xv.w, part_acc[m]);
        }
      }
      __syncwarp();
#if UA2_GEMV3_TAIL_PREFETCH
      if (lane == 0) {
        if (t + c.stages < total) {
          issue(t + c.stages);
          // this warp's last own copy is on its way: queue the next kernels' first rows behind it
          if (p.n_pf > 0 && t + c.stages == total - 1) tail_prefetch(p, cta, G, warp, nthreads >> 5);
        }
      }
#else
      if (lane == 0 && t + c.stages < total) issue(t + c.stages);
#endif
#pragma unroll
      for (int rr = 0; rr < ROWS; ++rr)
        if (rr == r) {
#pragma unroll
          for (int m = 0; m < MT; ++m) acc[rr][m] += part_acc[m];
        }
      if (rem == per_unit - 1) {  // slice of this unit finished: publish the partial sums
        float* dst = pb + ((size_t)(ul - r_lo) * c.nsl + sl) * ROWS * MT;
#pragma unroll
        for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
          for (int m = 0; m < MT; ++m) {
            const float s2 = warp_sum(acc[rr][m]);
            if (lane == 0) dst[rr * MT + m] = s2;
            acc[rr][m] = 0.f;
          }
      }
      ++t;
    }
    __syncthreads();  // all partial sums of the round are in shared memory
    const int n_ru = r_hi - r_lo;
    for (int idx = tid; idx < n_ru * mcount; idx += nthreads) {
      const int ul = idx / mcount, m = idx - ul * mcount;
      float a = 0.f, b = 0.f;
      const float* src = pb + (size_t)ul * c.nsl * ROWS * MT;
      for (int s = 0; s < c.nsl; ++s) {
        if (s * c.SL < K) {  // slices past the end of K never wrote
          a += src[s * ROWS * MT + m];
          if (ROWS == 2) b += src[s * ROWS * MT + MT + m];
        }
      }
      if (PRO == PRO_RMSNORM) {
        float tot = 0.f;
        for (int w = 0; w < (nthreads >> 5); ++w) tot += red[m][w];
        const float rs = rsqrtf(tot / (float)K + p.eps);  // lit_model.py:887-888, applied after the dot product
        a *= rs;
        b *= rs;
      }
      int nA;
      unit_row<EPI>(p, u_lo + r_lo + ul, 0, nA);
      int ps = -1, bq = -1;
      if (EPI == EPI_QKV) {
#pragma unroll
        for (int mm = 0; mm < MT; ++mm)
          if (mm == m) {
            ps = ps_row[mm];
            bq = b_row[mm];
          }
      }
      if (EPI == EPI_RESADD && rd == 0 && idx == tid)
        p.Y[(size_t)(m0 + m) * p.ldy + nA] = a + r_pre;
      else
        epilogue_v3<EPI>(p, m0 + m, a, b, nA, ps, bq);
    }
    // no second barrier: the next round writes the other partial buffer, and a thread only reaches the barrier of
    // round rd+1 after finishing its epilogue share of round rd
  }
}

int g_v3_ctas_per_sm = 2;
int g_v3_max_stages = 3;
int g_v3_kcw = 1024;
int g_v3_balance_grid = 1;
int g_v3_budget_kb = 110;
int g_sms = 0;
int sm_count3() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

// Launch shape of one v3 linear (shared by the launcher and by the prefetch planner of the PRECEDING kernel, which must
// reproduce the slab partition of the kernel it prefetches for).
struct V3Plan {
  V3Cfg c;
  int gy = 0, n_units = 0, nwarps = 0, m_tiles = 0, rows = 1;
  size_t smem = 0;
  bool ok = false;
};

V3Plan plan_v3(int MT, int epi, int M, int N, int K, int n_splits) {
  V3Plan pl;
  const int ROWS = (epi == EPI_QKV || epi == EPI_SWIGLU) ? 2 : 1;
  pl.rows = ROWS;
  const size_t kMaxSmem = 224 * 1024;
  const int Kp = ((K + 127) / 128) * 128;
  V3Cfg& c = pl.c;
  // K slices of ~1024 floats: every bulk copy is >= 4 KB, the size at which 8-9 warps x 2 slots saturate HBM
  // (profiles/r1_microbench_hbm_streaming.txt: 4 KB copies, depth 2, 8 warps -> 7.16 TB/s; 1 KB copies -> 2.8 TB/s)
  c.nsl = (K + 1023) / 1024;
  if (c.nsl > 8) c.nsl = 8;
  if (c.nsl < 1) c.nsl = 1;
  if (g_v3_ctas_per_sm <= 1) {
    c.ngrp = 8 / c.nsl;
    if (c.nsl == 3) c.ngrp = 3;  // 9 warps
  } else {
    // two smaller CTAs per SM: finer slabs (better static balance), dynamic CTA placement under PDL
    c.ngrp = c.nsl >= 5 ? 1 : (c.nsl >= 3 ? 2 : (c.nsl == 2 ? 4 : 8));
  }
  if (c.ngrp < 1) c.ngrp = 1;
  const int nwarps = c.nsl * c.ngrp;
  pl.nwarps = nwarps;
  const int per = (K + c.nsl - 1) / c.nsl;
  c.SL = ((per + 127) / 128) * 128;
  c.KCW = c.SL < g_v3_kcw ? c.SL : g_v3_kcw;
  c.n_splits = n_splits;
  const size_t xbytes = (size_t)MT * Kp * 4;
  const size_t pbytes = (size_t)2 * ROUND_UNITS * c.nsl * ROWS * MT * 4;
  const size_t stage_bytes = (size_t)nwarps * c.KCW * 4;
  // M <= 2 (decode): stay under ~half an SM's shared memory so the dependent kernel's CTA is co-resident (PDL)
  const size_t budget = (MT <= 2 ? (size_t)g_v3_budget_kb * 1024 : kMaxSmem);
  int stages = (int)((budget > xbytes + pbytes ? budget - xbytes - pbytes : 0) / stage_bytes);
  if (stages > g_v3_max_stages) stages = g_v3_max_stages;
  if (stages < 2) stages = 2;
  c.stages = stages;
  pl.smem = xbytes + pbytes + stage_bytes * stages;
  if (pl.smem > kMaxSmem) return pl;
  const int n_units = (ROWS == 2) ? ((epi == EPI_SWIGLU) ? N : N / 2) : N;
  pl.n_units = n_units;
  pl.m_tiles = (M + MT - 1) / MT;
  // Grid size: every (CTA, group) streams ceil(units / (G * ngrp)) units, so pick the G in [SMs, SMs * ctas_per_sm] that
  // wastes the least on the remainder (e.g. 2560 QKV pairs, 2 groups: G = 256 -> exactly 5 per group instead of 4.3 -> 5
  // on 296 CTAs).  All CTAs are co-resident either way; the kernel is bound by per-warp stream latency, not by SM count.
  const int g_max = sm_count3() * (g_v3_ctas_per_sm <= 1 ? 1 : g_v3_ctas_per_sm);
  int gy = g_max;
  if (g_v3_balance_grid && n_units > g_max * c.ngrp && n_units < 32 * g_max * c.ngrp) {
    auto eff_of = [&](int G) {
      const int slots = G * c.ngrp;
      const int per = (n_units + slots - 1) / slots;
      const int slab_max = (n_units + G - 1) / G;  // CTA slabs are floor-partitioned: some hold ceil(units / G)
      const int per2 = (slab_max + c.ngrp - 1) / c.ngrp;
      return (double)n_units / ((double)slots * (per2 > per ? per2 : per));
    };
    double best = 0.0;
    for (int G = sm_count3(); G <= g_max; ++G) best = eff_of(G) > best ? eff_of(G) : best;
    for (int G = g_max; G >= sm_count3(); --G)
      if (eff_of(G) >= best - 0.03) {  // largest grid within 3 % of the best balance (more warps in flight)
        gy = G;
        break;
      }
  }
  if (gy > n_units) gy = n_units;
  pl.gy = gy;
  pl.ok = true;
  return pl;
}

int mt_of(int M, int K) {
  const size_t rowb = (size_t)((K + 127) / 128) * 128 * 4;
  int mt = M >= 8 ? 8 : (M >= 3 ? 4 : M);
  while (mt > 1 && mt * rowb > 128 * 1024) mt >>= 1;
  return mt;
}

template <int MT, int PRO, int EPI>
cudaError_t launch3_one(const LaunchCtx& lc, const GemvParams& p, int n_splits) {
  auto kern = gemv3_kernel<MT, PRO, EPI>;
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return e;
    e = prefer_max_smem(kern);
    if (e != cudaSuccess) return e;
  }
  const V3Plan pl = plan_v3(MT, EPI, p.M, p.N, p.K, n_splits);
  if (!pl.ok) return cudaErrorInvalidValue;
  return launch(lc, kern, dim3(pl.m_tiles, pl.gy), dim3(pl.nwarps * 32), pl.smem, p, pl.c);
}

template <int PRO, int EPI>
cudaError_t launch3_mt(const LaunchCtx& lc, const GemvParams& p, int n_splits) {
  switch (mt_of(p.M, p.K)) {
    case 1: return launch3_one<1, PRO, EPI>(lc, p, n_splits);
    case 2: return launch3_one<2, PRO, EPI>(lc, p, n_splits);
    case 4: return launch3_one<4, PRO, EPI>(lc, p, n_splits);
    default: return launch3_one<8, PRO, EPI>(lc, p, n_splits);
  }
}

int g_v3_pf_mb = 0, g_v3_pf_idle_mb = 0;
}  // namespace

size_t gemv3_prefetch_budget(int idle_after) {
  size_t b = (size_t)g_v3_pf_mb << 20;
  if (idle_after) b += (size_t)g_v3_pf_idle_mb << 20;
  return b;
}
void set_gemv3_prefetch_mb(int mb, int idle_mb) {
  if (mb >= 0) g_v3_pf_mb = mb > 96 ? 96 : mb;
  if (idle_mb >= 0) g_v3_pf_idle_mb = idle_mb > 96 ? 96 : idle_mb;
}
// Prefetch specs for the next linears of the frame: walk the sequence until the byte budget is spent (at most PF_MAX
// kernels ahead).  Only decode-shaped launches (one M tile) are planned; anything else ends the walk.
int gemv3_make_pf(const GemvSeqEntry* next, int n_next, size_t budget_bytes, PfSpec* out) {
  int n_out = 0;
  for (int i = 0; i < n_next && n_out < PF_MAX && budget_bytes > 0; ++i) {
    const GemvSeqEntry& e = next[i];
    if (e.epi > EPI_SWIGLU || e.M > 2) break;
    const V3Plan pl = plan_v3(mt_of(e.M, e.K), e.epi, e.M, e.N, e.K, 0);
    if (!pl.ok || pl.m_tiles != 1) break;
    PfSpec f;
    f.W = e.W;
    f.W2 = e.W2;
    f.K = e.K;
    f.n_units = pl.n_units;
    f.G = pl.gy;
    f.mode = e.epi == EPI_SWIGLU ? 1 : (e.epi == EPI_QKV ? 2 : 0);
    f.hs = e.hs;
    const size_t unit_bytes = (size_t)pl.rows * e.K * 4;
    const size_t total = unit_bytes * pl.n_units;
    const size_t want = total < budget_bytes ? total : budget_bytes;
    int n = (int)((want + unit_bytes * pl.gy - 1) / (unit_bytes * pl.gy));
    const int slab_max = (pl.n_units + pl.gy - 1) / pl.gy;
    if (n > slab_max) n = slab_max;
    if (n < 1) break;
    f.n = n;
    out[n_out++] = f;
    const size_t used = (size_t)n * unit_bytes * pl.gy;
    budget_bytes = used >= budget_bytes ? 0 : budget_bytes - used;
    if (want < total) break;  // partial prefetch of this matrix: its own stream takes over from here
  }
  return n_out;
}
void set_gemv3_ctas_per_sm(int v) { g_v3_ctas_per_sm = v < 1 ? 1 : (v > 3 ? 3 : v); }
void set_gemv3_balance_grid(int v) { g_v3_balance_grid = v ? 1 : 0; }
void set_gemv3_kcw(int v) { g_v3_kcw = (v >= 128 && v <= 1024 && v % 128 == 0) ? v : 1024; }
void set_gemv3_budget_kb(int v) { g_v3_budget_kb = (v >= 24 && v <= 220) ? v : 110; }
void set_gemv3_max_stages(int v) { g_v3_max_stages = v < 2 ? 2 : (v > MAX_STAGES ? MAX_STAGES : v); }

cudaError_t launch_gemv3(const LaunchCtx& lc, int pro, int epi, const GemvParams& p, int n_splits) {
#define UA2_CASE3(P, E) \
  if (pro == P && epi == E) return launch3_mt<P, E>(lc, p, n_splits);
  UA2_CASE3(PRO_PLAIN, EPI_STORE)
  UA2_CASE3(PRO_PLAIN, EPI_RESADD)
  UA2_CASE3(PRO_PLAIN, EPI_SWIGLU)
  UA2_CASE3(PRO_PLAIN, EPI_QKV)
  UA2_CASE3(PRO_RMSNORM, EPI_STORE)
  UA2_CASE3(PRO_RMSNORM, EPI_RESADD)
  UA2_CASE3(PRO_RMSNORM, EPI_SWIGLU)
  UA2_CASE3(PRO_RMSNORM, EPI_QKV)
  UA2_CASE3(PRO_GATHER, EPI_STORE)
  UA2_CASE3(PRO_ATTN, EPI_RESADD)
  UA2_CASE3(PRO_ATTN_DIRECT, EPI_RESADD)
#undef UA2_CASE3
  return cudaErrorInvalidValue;
}

}  // namespace ua2
