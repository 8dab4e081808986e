// Top-k / temperature / Gumbel-style sampler, one thread-block CLUSTER per sampled row.
//
// Replaces llm_models/model_new.py:146-156 sample_topk, :158-187 audio_sample_topk and :141-143
// _multinomial_sample_one_no_sync:
//     logits = logits / T ; logits[:forbid] = -inf
//     keep   = logits >= kth_largest(logits)            (ties at the threshold are kept)
//     p      = softmax(log_softmax(masked logits))
//     token  = argmax(p / q),  q ~ Exp(1)               (first index wins ties, torch.argmax)
// plus the CFG mix of model_new.py:618-622 / :634-637 (row 0 = cond, row 1 = uncond):  u + (c - u) * scale.
//
// Design: the V logits of a row are split across the CTAs of a cluster (8 CTAs for the 128256-entry text head,
// 1 CTA for a 12300-entry codebook head); each CTA keeps its slice in shared memory, so the row is read from
// L2 exactly once.  The exact k-th largest value is found by a 4-pass MSB-first radix select on order-preserving
// keys; per-CTA 256-bin histograms are merged through distributed shared memory (each CTA reads its peers'
// bins), as are the max / sum / argmax reductions.  Integer outputs are bit-exact functions of the fp32 logits.
// Roofline: latency (a few cluster barriers); bytes = 4*V (+4*V noise) per row.
#include <cooperative_groups.h>

#include "ua2_kernels.cuh"
#include "ua2_philox.cuh"

namespace cg = cooperative_groups;

namespace ua2 {
namespace {

constexpr int SAMPLE_THREADS = 512;
constexpr int SLICE_CAP = 16384;  // floats of a row slice cached in shared memory per CTA (64 KB)
constexpr int MAX_CLUSTER = 8;
constexpr int FAST_GROUPS = SAMPLE_THREADS / 4;  // group maxima per CTA = largest k served by the candidate path
constexpr int CAND_CAP = 1536;                   // candidate list capacity (floats + ints in static shared memory)

__device__ __forceinline__ uint32_t f2key(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SampleArgs {
  const float* logits;  // (rows_in, V)
  int V;
  const FrameScalars* fs;
  int is_audio;  // audio heads apply forbid_prefix; the text head does not (model_new.py:623 vs :639)
  int out_col, out_ld;
  long long noise_off;
  unsigned long long stream_id;
  int B;
};

template <bool CLUSTERED>
__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const SampleArgs a) {
  extern __shared__ __align__(16) float vals[];      // slice of scaled logits (<= SLICE_CAP)
  __shared__ unsigned int hist[4][256];              // one merged histogram per radix pass (never reused -> 1 barrier/pass)
  __shared__ unsigned int hist_w[SAMPLE_THREADS / 32][256];  // per-warp private histograms of the current pass
  __shared__ float red_f[4][SAMPLE_THREADS / 32];    // block-level scratch
  __shared__ int red_i[SAMPLE_THREADS / 32];
  __shared__ float cl_f[4];                          // per-CTA results exchanged across the cluster
  __shared__ int cl_i[2];
  __shared__ unsigned int sel_bin, sel_k;
  __shared__ float gmax[FAST_GROUPS];                // fast path: group maxima, candidate list
  __shared__ float cand_v[CAND_CAP];
  __shared__ int cand_i[CAND_CAP];
  __shared__ int cand_n;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = SAMPLE_THREADS / 32;
  int csize = 1, crank = 0;
  if (CLUSTERED) {
    cg::cluster_group cl = cg::this_cluster();
    csize = cl.num_blocks();
    crank = cl.block_rank();
  }
  const int row = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();

  const FrameScalars fs = *a.fs;
  const int V = a.V;
  const bool use_cfg = fs.cfg_scale > 1.0f && fs.B > 1;
  const int forbid = a.is_audio ? fs.forbid_prefix : 0;
  const int per = (V + csize - 1) / csize;
  const int lo = min(V, crank * per), hi = min(V, lo + per);
  const int n = hi - lo;


  // value of (scaled, masked) logit i of this row; cached in smem when it fits
  auto raw = [&](int i) -> float {
    float x;
    if (use_cfg) {
      const float c = a.logits[i], u = a.logits[(size_t)V + i];
      x = u + (c - u) * fs.cfg_scale;  // model_new.py:619
    } else {
      x = a.logits[(size_t)row * V + i];
    }
    x = x / fs.temperature;
    if (i < forbid) x = -INFINITY;
    return x;
  };
  // slice -> shared memory (scaled, masked), returns the thread's running maximum.  All loads of a thread are issued before
  // their first use: a plain `for` over raw() serialises ~24-32 L2 round trips per thread (17 of the sampler's 21 us).
  auto load_slice = [&]() -> float {
    float lm = -INFINITY;
    constexpr int UNR = 8;
    const float* lrow = a.logits + (use_cfg ? (size_t)0 : (size_t)row * V) + lo;
    for (int i0 = tid; i0 < n; i0 += SAMPLE_THREADS * UNR) {
      float xc[UNR], xu[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int i = i0 + u * SAMPLE_THREADS;
        xc[u] = i < n ? lrow[i] : 0.f;
        xu[u] = (use_cfg && i < n) ? lrow[(size_t)V + i] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int i = i0 + u * SAMPLE_THREADS;
        if (i < n) {
          float x = use_cfg ? xu[u] + (xc[u] - xu[u]) * fs.cfg_scale : xc[u];  // model_new.py:619
          x = x / fs.temperature;
          if (lo + i < forbid) x = -INFINITY;
          if (i < SLICE_CAP) vals[i] = x;
          lm = fmaxf(lm, x);
        }
      }
    }
    return lm;
  };
  // cluster-wide all-reduce helpers (every CTA ends up with the same value)
  auto cluster_barrier = [&]() {
    if (CLUSTERED)
      cg::this_cluster().sync();
    else
      __syncthreads();
  };
  const float* noise = fs.noise ? fs.noise + a.noise_off + (size_t)row * V : nullptr;

  // ------------------------------------------------------------------------------------------------------------------
  // Fast path (1 < k <= FAST_GROUPS, slice cached in shared memory): candidate filtering instead of a radix select.
  //   * every group of 4 threads keeps the maximum of its elements: FAST_GROUPS disjoint group maxima per CTA.  The k-th
  //     largest group maximum `lb` is a lower bound of the row's k-th largest value (the k largest group maxima are k
  //     distinct elements >= lb), and so is the largest `lb` over the CTAs of the cluster;
  //   * the elements >= lb (typically ~1.2 k of them; all of the row's top k) are compacted into a candidate list;
  //   * the exact k-th largest (ties kept, like `logits >= kth`) is found by rank counting inside the list, and one warp
  //     finishes softmax(log_softmax) and argmax(p / q) over the <= CAND_CAP candidates - no more passes over V.
  // Falls through to the generic radix path when the list overflows (heavy ties / -inf plateaus).
  if (fs.topk > 1 && fs.topk <= FAST_GROUPS && n <= SLICE_CAP && per <= SLICE_CAP) {
    const int k = fs.topk;
    const float lm = load_slice();
    float g = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, 1));
    g = fmaxf(g, __shfl_xor_sync(0xffffffffu, g, 2));
    if ((lane & 3) == 0) gmax[tid >> 2] = g;
    const float wm = warp_max(g);
    if (lane == 0) red_f[0][warp] = wm;
    if (tid == 0) cand_n = 0;
    __syncthreads();
    if (tid < FAST_GROUPS) {
      const float mine = gmax[tid];
      int gt = 0;
#pragma unroll 8
      for (int j = 0; j < FAST_GROUPS; ++j) {
        const float o = gmax[j];
        gt += (o > mine || (o == mine && j < tid)) ? 1 : 0;
      }
      if (gt == k - 1) cl_f[1] = mine;  // exactly one thread: ranks are a permutation
    }
    if (tid == 0) {
      float m2 = -INFINITY;
      for (int w = 0; w < NW; ++w) m2 = fmaxf(m2, red_f[0][w]);
      cl_f[0] = m2;
    }
    cluster_barrier();
    float mx, lb;
    if (CLUSTERED) {
      cg::cluster_group cl = cg::this_cluster();
      mx = -INFINITY;
      lb = -INFINITY;
      for (int r = 0; r < csize; ++r) {
        mx = fmaxf(mx, *cl.map_shared_rank(&cl_f[0], r));
        lb = fmaxf(lb, *cl.map_shared_rank(&cl_f[1], r));
      }
    } else {
      mx = cl_f[0];
      lb = cl_f[1];
    }
    // compact the candidates (warp-aggregated append)
    for (int i0 = 0; i0 < n; i0 += SAMPLE_THREADS) {
      const int i = i0 + tid;
      const float x = i < n ? vals[i] : -INFINITY;
      const bool c = i < n && x >= lb;
      const unsigned m = __ballot_sync(0xffffffffu, c);
      if (m) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&cand_n, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        const int pos = base + __popc(m & ((1u << lane) - 1u));
        if (c && pos < CAND_CAP) {
          cand_v[pos] = x;
          cand_i[pos] = lo + i;
        }
      }
    }
    cluster_barrier();
    int C = 0;
    if (CLUSTERED) {
      cg::cluster_group cl = cg::this_cluster();
      int mine_off = 0;
      for (int r = 0; r < csize; ++r) {
        const int cr = *cl.map_shared_rank(&cand_n, r);
        if (r == 0) mine_off = cr;  // CTA 0 appends the peers' lists behind its own
        C += cr;
      }
      bool ok = C <= CAND_CAP;
      for (int r = 0; r < csize; ++r) ok = ok && (*cl.map_shared_rank(&cand_n, r) <= CAND_CAP);
      if (ok && crank == 0) {
        int off = mine_off;
        for (int r = 1; r < csize; ++r) {
          const int cr = *cl.map_shared_rank(&cand_n, r);
          const float* rv = cl.map_shared_rank(&cand_v[0], r);
          const int* ri = cl.map_shared_rank(&cand_i[0], r);
          for (int t = tid; t < cr; t += SAMPLE_THREADS) {
            cand_v[off + t] = rv[t];
            cand_i[off + t] = ri[t];
          }
          off += cr;
        }
      }
      cluster_barrier();  // peers' lists have been read: they may retire
      if (ok && crank != 0) return;
      if (!ok) C = CAND_CAP + 1;
    } else {
      C = cand_n;
    }
    if (C <= CAND_CAP) {
      // exact k-th largest inside the list
      for (int t = tid; t < C; t += SAMPLE_THREADS) {
        const float v = cand_v[t];
        int gt = 0, ge = 0;
        for (int j = 0; j < C; ++j) {
          const float o = cand_v[j];
          gt += o > v ? 1 : 0;
          ge += o >= v ? 1 : 0;
        }
        if (gt < k && k <= ge) cl_f[2] = v;  // every thread that qualifies holds the same value
      }
      __syncthreads();
      if (warp == 0) {
        const float thr = cl_f[2];
        float z = 0.f;
        for (int t = lane; t < C; t += 32) {
          const float x = cand_v[t];
          if (x >= thr) z += expf(x - mx);
        }
        z = warp_sum(z);
        const float lse = logf(z);
        const float max_ls = 0.f - lse;
        float z2 = 0.f;
        for (int t = lane; t < C; t += 32) {
          const float x = cand_v[t];
          if (x >= thr) z2 += expf(((x - mx) - lse) - max_ls);
        }
        z2 = warp_sum(z2);
        float best = -1.f;
        int best_i = 0x7fffffff;
        for (int t = lane; t < C; t += 32) {
          const float x = cand_v[t];
          if (x >= thr) {
            const int gi = cand_i[t];
            const float p = expf(((x - mx) - lse) - max_ls) / z2;
            const float q = noise ? noise[gi] : philox_exp1(fs.seed, fs.offset * 65536ull + a.stream_id * 1024ull + row, gi);
            const float sc = p / q;
            if (sc > best || (sc == best && gi < best_i)) {
              best = sc;
              best_i = gi;
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (ob > best || (ob == best && oi < best_i)) {
            best = ob;
            best_i = oi;
          }
        }
        if (lane == 0) {
          if (use_cfg) {
            for (int b = 0; b < fs.B; ++b) fs.out[(size_t)b * a.out_ld + a.out_col] = best_i;  // .repeat(2,1), model_new.py:621
          } else {
            fs.out[(size_t)row * a.out_ld + a.out_col] = best_i;
          }
        }
      }
      return;
    }
    __syncthreads();  // overflow: generic path below (re-reads the row)
  }

  const float lm_generic = load_slice();
  __syncthreads();
  auto val = [&](int i) -> float { return (i < SLICE_CAP) ? vals[i] : raw(lo + i); };

  // ---- 1. max
  float mx = warp_max(lm_generic);
  if (lane == 0) red_f[0][warp] = mx;
  __syncthreads();
  if (tid == 0) {
    float m2 = -INFINITY;
    for (int w = 0; w < NW; ++w) m2 = fmaxf(m2, red_f[0][w]);
    cl_f[0] = m2;
  }
  cluster_barrier();
  if (CLUSTERED) {
    cg::cluster_group cl = cg::this_cluster();
    float m2 = -INFINITY;
    for (int r = 0; r < csize; ++r) m2 = fmaxf(m2, *cl.map_shared_rank(&cl_f[0], r));
    mx = m2;
  } else {
    mx = cl_f[0];
  }

  // ---- 2. exact k-th largest via MSB-first radix select
  uint32_t thr_key;
  if (fs.topk <= 1) {
    thr_key = f2key(mx);
  } else {
    uint32_t prefix = 0, pmask = 0;
    unsigned int k_rem = (unsigned int)fs.topk;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      // per-warp private histograms: logits share their top bits, so a CTA-wide histogram would serialise every
      // atomic on a handful of bins (measured 6.5 us per pass); private rows confine the contention to one warp
      for (int i = tid; i < NW * 256; i += SAMPLE_THREADS) (&hist_w[0][0])[i] = 0u;
      __syncthreads();
      for (int i = tid; i < n; i += SAMPLE_THREADS) {
        const uint32_t key = f2key(val(i));
        if ((key & pmask) == prefix) atomicAdd(&hist_w[warp][(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid < 256) {
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) tot += hist_w[w][tid];
        hist[pass][tid] = tot;
      }
      cluster_barrier();
      // every CTA redundantly merges the histograms and picks the same bin
      if (warp == 0) {
        unsigned int cnt[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int bin = 255 - (lane * 8 + j);  // descending bins; lane 0 holds the largest values
          unsigned int c = 0;
          if (CLUSTERED) {
            cg::cluster_group cl = cg::this_cluster();
            for (int r = 0; r < csize; ++r) c += *cl.map_shared_rank(&hist[pass][bin], r);
          } else {
            c = hist[pass][bin];
          }
          cnt[j] = c;
        }
        unsigned int tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) tot += cnt[j];
        // exclusive prefix over lanes (descending value order)
        unsigned int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        unsigned int above = incl - tot;
        if (above < k_rem && k_rem <= incl) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (above < k_rem && k_rem <= above + cnt[j]) {
              sel_bin = 255u - (unsigned int)(lane * 8 + j);
              sel_k = k_rem - above;
            }
            above += cnt[j];
          }
        }
      }
      __syncthreads();
      prefix |= sel_bin << shift;
      pmask |= 255u << shift;
      k_rem = sel_k;
      __syncthreads();
    }
    thr_key = prefix;
  }

  // ---- 3. Z = sum_{kept} exp(x - mx)
  float z = 0.f;
  for (int i = tid; i < n; i += SAMPLE_THREADS) {
    const float x = val(i);
    if (f2key(x) >= thr_key) z += expf(x - mx);
  }
  z = warp_sum(z);
  if (lane == 0) red_f[1][warp] = z;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += red_f[1][w];
    cl_f[1] = s;
  }
  cluster_barrier();
  if (CLUSTERED) {
    cg::cluster_group cl = cg::this_cluster();
    float s = 0.f;
    for (int r = 0; r < csize; ++r) s += *cl.map_shared_rank(&cl_f[1], r);
    z = s;
  } else {
    z = cl_f[1];
  }
  const float lse = logf(z);
  // log_softmax value of element x is (x - mx) - lse; its max over the row is (0 - lse)
  const float max_ls = 0.f - lse;

  // ---- 4. Z2 = sum exp(ls - max_ls)   (the second softmax of model_new.py:153)
  float z2 = 0.f;
  for (int i = tid; i < n; i += SAMPLE_THREADS) {
    const float x = val(i);
    if (f2key(x) >= thr_key) z2 += expf(((x - mx) - lse) - max_ls);
  }
  z2 = warp_sum(z2);
  if (lane == 0) red_f[2][warp] = z2;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < NW; ++w) s += red_f[2][w];
    cl_f[2] = s;
  }
  cluster_barrier();
  if (CLUSTERED) {
    cg::cluster_group cl = cg::this_cluster();
    float s = 0.f;
    for (int r = 0; r < csize; ++r) s += *cl.map_shared_rank(&cl_f[2], r);
    z2 = s;
  } else {
    z2 = cl_f[2];
  }

  // ---- 5. argmax p / q  (first index wins ties)
  float best = -1.f;
  int best_i = 0x7fffffff;
  for (int i = tid; i < n; i += SAMPLE_THREADS) {
    const float x = val(i);
    if (f2key(x) >= thr_key) {
      const float p = expf(((x - mx) - lse) - max_ls) / z2;
      const float q = noise ? noise[lo + i] : philox_exp1(fs.seed, fs.offset * 65536ull + a.stream_id * 1024ull + row, lo + i);
      const float sc = p / q;
      if (sc > best || (sc == best && lo + i < best_i)) {
        best = sc;
        best_i = lo + i;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  if (lane == 0) {
    red_f[3][warp] = best;
    red_i[warp] = best_i;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 0; w < NW; ++w) {
      if (red_f[3][w] > best || (red_f[3][w] == best && red_i[w] < best_i)) {
        best = red_f[3][w];
        best_i = red_i[w];
      }
    }
    cl_f[3] = best;
    cl_i[0] = best_i;
  }
  cluster_barrier();
  if (crank == 0 && tid == 0) {
    if (CLUSTERED) {
      cg::cluster_group cl = cg::this_cluster();
      for (int r = 1; r < csize; ++r) {
        const float ob = *cl.map_shared_rank(&cl_f[3], r);
        const int oi = *cl.map_shared_rank(&cl_i[0], r);
        if (ob > best || (ob == best && oi < best_i)) {
          best = ob;
          best_i = oi;
        }
      }
    }
    if (use_cfg) {
      for (int b = 0; b < fs.B; ++b) fs.out[(size_t)b * a.out_ld + a.out_col] = best_i;  // .repeat(2,1), model_new.py:621
    } else {
      fs.out[(size_t)row * a.out_ld + a.out_col] = best_i;
    }
  }
  // keep every CTA's shared memory alive until all remote reads are done
  cluster_barrier();
}

}  // namespace

cudaError_t launch_sampler(const LaunchCtx& lc, const float* logits, int V, const FrameScalars* d_fs, int is_audio,
                           int out_col, int out_ld, long long noise_off, unsigned long long stream_id, int B,
                           int rows) {
  SampleArgs a{logits, V, d_fs, is_audio, out_col, out_ld, noise_off, stream_id, B};
  int csize = 1;
  while (csize < MAX_CLUSTER && (V + csize - 1) / csize > SLICE_CAP) csize <<= 1;
  const int per = (V + csize - 1) / csize;
  const size_t smem = (size_t)(per < SLICE_CAP ? per : SLICE_CAP) * sizeof(float);
  static DeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLICE_CAP * 4);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SLICE_CAP * 4);
    if (e != cudaSuccess) return e;
    prefer_max_smem(sample_kernel<true>);
    prefer_max_smem(sample_kernel<false>);
  }
  if (lc.launch_counter) ++*lc.launch_counter;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(csize, rows);
  cfg.blockDim = dim3(SAMPLE_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (csize > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = csize;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (lc.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (csize > 1) return cudaLaunchKernelEx(&cfg, sample_kernel<true>, a);
  return cudaLaunchKernelEx(&cfg, sample_kernel<false>, a);
}

}  // namespace ua2
