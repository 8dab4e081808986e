// Causal 1-D convolutions of the SEANet stacks on tcgen05, straight from the (B, C, T) activation layout - no im2col buffer.
//
//   y[b, co, t] = bias[co] + sum_{tap, ci} W[co, ci, tap] * f(x)[b, ci, t * stride + tap - pad_left]  (+ res[b, co, t]),   f = ELU or identity
// (modules/conv.py:232-254 StreamingConv1d with left zero padding; modules/seanet.py:21-94 the residual block's two convolutions).
// The narrow layers that run at the 24 kHz / 6 kHz rates (64 - 128 channels) did 10 - 30 TFLOP/s on the fp32 register-tiled core
// (profiles/r2_kernel_rooflines.md: shared-memory-bandwidth bound) although they are HBM-bound by their arithmetic.  As an implicit
// GEMM with M = 128 output positions (TMEM lanes), N = output channels, K = taps x input channels:
//
//   warps 4-11   A producers, two groups of four warps taking alternate k-blocks.  thread = output position: 32 coalesced scalar
//                loads of one tap of 32 input channels (zero outside the sequence: the causal padding), ELU, cvt.rna.tf32 hi / lo
//                split, two tcgen05.st 32x32b.x32 into a TMEM slot - the "A operand from tensor memory" form, so the activations
//                never touch shared memory
//   warp 0       weight producer: k-block tiles (Cout rows x 32 k, hi and lo planes, K-major SWIZZLE_128B) by TMA from a
//                re-packed copy [2][Cout][tap * Cin + ci]; the tiles stay resident when the whole filter fits the ring
//   warp 1       MMA issuer: per k-step of 8:  D += A_hi B_hi;  D += A_lo B_hi;  D += A_hi B_lo   (tcgen05.mma kind::tf32, 3xTF32)
//   warps 12-15  epilogue: tcgen05.ld of D (lane = position, column = output channel) -> + bias (+ residual) -> y, coalesced
//                over positions for every channel.  Two accumulators (Cout <= 128) let tile i drain while tile i + 1 accumulates.
//
// Algorithmic bytes per launch: 4 * B * (Cin * T_in + Cout * T_out) (+ residual); the weights (<= 1 MB) live in L2 / shared memory.
#include <algorithm>

#include "ua2_kernels.cuh"
#include "ua2_tcgen05.cuh"

namespace ua2 {
namespace {

using namespace tc;

constexpr int CU_BM = 128;       // output positions per tile
constexpr int CU_A_SLOTS = 6;    // capacity of the TMEM A ring's barrier arrays (64 columns per slot); a kernel uses A_SLOTS of them:
                                 // NO <= 64: 6 slots, accumulators at columns 384 (+ 64); NO >= 128: 4 slots, accumulators at 256 (+ 128).
                                 // (ncu, 64 -> 32 k 3 at 24 kHz: with 4 slots the splitter warps spent half their time waiting for a free
                                 // slot while the MMA warp spun on `slot full` - two slots per splitter group do not absorb the jitter)
constexpr int CU_MAX_STAGES = 24;
constexpr int CU_THREADS = 512;

struct ConvUmmaParams {
  const float* x;
  const float* bias;
  const float* res;
  float* y;
  const float* prelu;  // scalar nn.PReLU slope applied after the bias and before the residual add (ScalarModel: activation(conv(x)), scalar24k.py:139-150) or nullptr
  int B, Cin, Cout, T_in, T_out, Ktaps, stride, pad_left, pre_elu;
  int tap_step;     // dilation d of a convolution (tap k reads t * stride + k * d - pad_left), -1 transposed convolution (tap k reads j - k)
  int T_pos;        // positions (GEMM rows) per batch element: T_out, or the input grid Tj of a transposed convolution
  int tr_stride;    // 0 = convolution; s = transposed convolution: GEMM column n0 + c = phase * Cout + channel -> y[b, ch, j * s + phase - crop]
  int tr_crop;
  int n0;           // first GEMM column (weight row) of this launch; n_cols of them are valid
  int n_cols;
  int KB;           // k-blocks of 32: Ktaps * cb_per_tap
  int cb_per_tap;   // ceil(Cin / 32): the last channel block of a tap is zero-padded when Cin % 32 != 0 (weights and activations)
  int tiles_per_b;  // ceil(T_pos / 128)
  int n_tiles;
  int n_stages;     // weight ring stages
  int resident;     // KB <= n_stages: every k-block is loaded once and stays
  // stride-1 convolutions with T_in % 4 == 0: the 128 + (Ktaps - 1) * dilation samples x 32 channels that a (tile, channel block)
  // needs come by TMA (boxes of 32 samples x 32 channels, SWIZZLE_128B) into a shared-memory ring and serve all taps from there
  // (k-block order: channel block, then tap; the weight planes are repacked in that order); 0 = the register gather (strided /
  // transposed / odd lengths)
  int staged;
  int xw;           // 32-sample boxes per activation stage (4 KB each)
  int x_shift;      // 0..3: a stage starts x_shift samples LEFT of t0 - pad_left, so that every box starts on a 16-byte boundary of the row
  int n_xstages;
};

// ELU with the SFU exponential: exp(v) - 1 = ex2(v * log2 e) - 1 for v <= 0.  Absolute error ~1e-7 (the library expm1f of the fp32
// kernels is exact to an ulp); two orders below the tensor-core accumulation error of this path, codec indices stay bit-equal.
__device__ __forceinline__ float elu_fast(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
  return v > 0.f ? v : e - 1.f;
}

constexpr int CU_X_STAGES = 4;

template <int NO>
__global__ void __launch_bounds__(CU_THREADS, 1) conv_umma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                                                                 const ConvUmmaParams p) {
  constexpr int STAGE_BYTES = 2 * NO * 128;  // hi tile then lo tile, NO rows of 128 bytes each
  constexpr int NACC = NO <= 128 ? 2 : 1;
  constexpr int A_SLOTS = NO <= 64 ? 6 : 4;
  constexpr int CU_ACC_COL = A_SLOTS * 64;
  constexpr int ACC_STRIDE = NO <= 64 ? 64 : 128;
  extern __shared__ uint8_t cu_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(cu_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_ring = base;
  uint8_t* x_ring = w_ring + (size_t)p.n_stages * STAGE_BYTES;  // 1024-byte aligned (STAGE_BYTES is a multiple of 1024)
  const uint32_t x_bytes = (uint32_t)p.xw * 4096u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_ring + (size_t)p.n_xstages * x_bytes);
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + CU_MAX_STAGES;
  uint64_t* a_full = w_empty + CU_MAX_STAGES;
  uint64_t* a_empty = a_full + CU_A_SLOTS;
  uint64_t* acc_full = a_empty + CU_A_SLOTS;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* x_full = acc_empty + 2;
  uint64_t* x_empty = x_full + CU_X_STAGES;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(x_empty + CU_X_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < p.n_stages; ++i) {
      smem_bar_init(&w_full[i], 1);
      smem_bar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < CU_A_SLOTS; ++i) {
      smem_bar_init(&a_full[i], 4);
      smem_bar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      smem_bar_init(&acc_full[i], 1);
      smem_bar_init(&acc_empty[i], 4);
    }
    for (int i = 0; i < CU_X_STAGES; ++i) {
      smem_bar_init(&x_full[i], 1);
      smem_bar_init(&x_empty[i], 8);  // every splitter warp is done with the stage
    }
    if (p.staged) tma_prefetch_desc(&tmX);
    smem_bar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_smem;
  pdl_launch_dependents();
  const int my_tiles = cta < p.n_tiles ? (p.n_tiles - 1 - cta) / G + 1 : 0;

  if (warp == 0) {
    // ================= weight producer (the re-packed weights are written by the preceding kernel of this launch sequence)
    pdl_wait();
    const uint32_t total = p.resident ? (uint32_t)p.KB : (uint32_t)my_tiles * (uint32_t)p.KB;
    for (uint32_t it = 0; it < total; ++it) {
      const uint32_t sw = it % (uint32_t)p.n_stages, pw = (it / (uint32_t)p.n_stages) & 1;
      const int kb = (int)(it % (uint32_t)p.KB);
      if (!p.resident) smem_bar_wait(&w_empty[sw], pw ^ 1);
      if (elect_one()) {
        smem_bar_arrive_expect_tx(&w_full[sw], STAGE_BYTES);
        tma_load_3d(w_ring + (size_t)sw * STAGE_BYTES, &tmW, kb * 32, p.n0, 0, &w_full[sw], POLICY_EVICT_LAST);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NO >> 3) << 17) | ((uint32_t)(CU_BM >> 4) << 24);
    uint32_t it = 0;
    for (int j = 0; j < my_tiles; ++j) {
      const int ab = NACC == 2 ? (j & 1) : 0;
      const uint32_t use = NACC == 2 ? (uint32_t)(j >> 1) : (uint32_t)j;  // how often this accumulator has been used before
      smem_bar_wait(&acc_empty[ab], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem + CU_ACC_COL + ab * ACC_STRIDE;
      for (int kb = 0; kb < p.KB; ++kb, ++it) {
        const uint32_t sw = p.resident ? (uint32_t)kb : it % (uint32_t)p.n_stages;
        const uint32_t pw = p.resident ? 0u : (it / (uint32_t)p.n_stages) & 1;
        const uint32_t sa = it % A_SLOTS, pa = (it / A_SLOTS) & 1;
        smem_bar_wait(&w_full[sw], pw);
        smem_bar_wait(&a_full[sa], pa);
        tc_fence_after();
        const uint32_t ws = smem_addr_u32(w_ring + (size_t)sw * STAGE_BYTES);
        const uint64_t bh = smem_desc_sw128(ws), bl = smem_desc_sw128(ws + NO * 128);
        const uint32_t a_hi = tmem + sa * 64, a_lo = a_hi + 32;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            mma_tf32_ts(d_tmem, a_hi + ks * 8, bh + (uint64_t)(ks * 2), idesc, (kb == 0 && ks == 0) ? 0u : 1u);
            mma_tf32_ts(d_tmem, a_lo + ks * 8, bh + (uint64_t)(ks * 2), idesc, 1u);
            mma_tf32_ts(d_tmem, a_hi + ks * 8, bl + (uint64_t)(ks * 2), idesc, 1u);
          }
          tc_commit(&a_empty[sa]);
          if (!p.resident) tc_commit(&w_empty[sw]);
          if (kb + 1 == p.KB) tc_commit(&acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2 && p.staged) {
    // ================= activation producer (staged mode): 32 channels x (128 + halo) samples per (tile, channel block) as boxes of 32
    // samples; samples left of 0 / right of T_in arrive as zeros = the convolution's zero padding
    pdl_wait();  // x comes from the preceding kernel
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)p.cb_per_tap;
    int j = 0, cb = 0;
    for (uint32_t n = 0; n < total; ++n) {
      const uint32_t xs = n % (uint32_t)p.n_xstages, ph = (n / (uint32_t)p.n_xstages) & 1;
      smem_bar_wait(&x_empty[xs], ph ^ 1);
      const int tile = cta + j * G;
      const int b = tile / p.tiles_per_b, t0 = (tile - b * p.tiles_per_b) * CU_BM;
      if (elect_one()) {
        smem_bar_arrive_expect_tx(&x_full[xs], x_bytes);
        for (int k = 0; k < p.xw; ++k)
          tma_load_2d(x_ring + (size_t)xs * x_bytes + (size_t)k * 4096, &tmX, t0 - p.pad_left - p.x_shift + 32 * k, b * p.Cin + cb * 32,
                      &x_full[xs], POLICY_EVICT_FIRST);
      }
      __syncwarp();
      if (++cb == p.cb_per_tap) {
        cb = 0;
        ++j;
      }
    }
  } else if (warp >= 4 && warp < 12 && p.staged) {
    // ================= splitters (staged mode): thread = output position; tap k of a channel block reads column r + k * dilation of
    // the block's shared-memory stage (consecutive lanes, consecutive words: conflict-free), activates, splits hi / lo -> TMEM
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t n_it = (uint32_t)my_tiles * (uint32_t)p.KB;
    uint32_t it = (uint32_t)grp;
    int jt = (int)(it / (uint32_t)p.KB);
    int kb0 = (int)(it - (uint32_t)jt * (uint32_t)p.KB);
    int cb = kb0 / p.Ktaps, tap = kb0 - cb * p.Ktaps;
    for (; it < n_it; it += 2) {
      const uint32_t n = (uint32_t)jt * (uint32_t)p.cb_per_tap + (uint32_t)cb;
      const uint32_t xs = n % (uint32_t)p.n_xstages, xph = (n / (uint32_t)p.n_xstages) & 1;
      smem_bar_wait(&x_full[xs], xph);
      // sample c = r + tap * dilation of the stage: box c / 32, row = channel i (128 bytes), 16-byte chunk ((c % 32) / 4) ^ (i % 8)
      const int c = r + tap * p.tap_step + p.x_shift;
      const uint8_t* xb = x_ring + (size_t)xs * x_bytes + (size_t)(c >> 5) * 4096 + (c & 3) * 4;
      const int cq = (c & 31) >> 2;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = *reinterpret_cast<const float*>(xb + i * 128 + ((cq ^ (i & 7)) << 4));
      const uint32_t sa = it % A_SLOTS, pa = (it / A_SLOTS) & 1;
      smem_bar_wait(&a_empty[sa], pa ^ 1);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + sa * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float f = p.pre_elu ? elu_fast(v[half * 16 + i]) : v[half * 16 + i];
          const uint32_t h = tf32_rna_bits(f);
          hi[i] = h;
          lo[i] = tf32_rna_bits(f - __uint_as_float(h));
        }
        tmem_st16(taddr + half * 16, hi);
        tmem_st16(taddr + 32 + half * 16, lo);
      }
      tmem_wait_st();  // the stores carry the loaded values: the stage has been read when they are done
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        bar_arrive(&a_full[sa]);
        if (tap + 2 >= p.Ktaps) bar_arrive(&x_empty[xs]);  // this warp's last tap of the stage (Ktaps >= 2: both groups have one)
      }
      tap += 2;
      while (tap >= p.Ktaps) {
        tap -= p.Ktaps;
        if (++cb == p.cb_per_tap) {
          cb = 0;
          ++jt;
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================= A producers: f(x) of one tap x 32 channels for this thread's output position -> TMEM (hi, lo)
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t n_it = (uint32_t)my_tiles * (uint32_t)p.KB;
    pdl_wait();  // x comes from the preceding kernel
    // the loads of k-block it + 2 (this group's next one) are issued before the split of k-block it: two k-blocks of a warp in flight.
    // (tile, k-block) of the cursor advance incrementally - no integer division in the loop
    struct Cur {
      int kb, tap, cb;  // k-block, its tap and channel block
      int b, t0;        // batch element and first position of the tile
      int tile;
    };
    auto cur_init = [&](uint32_t it) -> Cur {
      Cur c;
      const int j = (int)(it / (uint32_t)p.KB);
      c.kb = (int)(it - (uint32_t)j * (uint32_t)p.KB);
      c.tap = c.kb / p.cb_per_tap;
      c.cb = c.kb - c.tap * p.cb_per_tap;
      c.tile = cta + j * G;
      c.b = c.tile / p.tiles_per_b;
      c.t0 = (c.tile - c.b * p.tiles_per_b) * CU_BM;
      return c;
    };
    auto cur_step2 = [&](Cur& c) {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        if (++c.cb == p.cb_per_tap) {
          c.cb = 0;
          ++c.tap;
        }
        if (++c.kb == p.KB) {
          c.kb = 0;
          c.tap = 0;
          c.tile += G;
          c.t0 += G * CU_BM;
          while (c.t0 >= p.tiles_per_b * CU_BM) {
            c.t0 -= p.tiles_per_b * CU_BM;
            ++c.b;
          }
        }
      }
    };
    auto gather = [&](const Cur& c, float (&v)[32]) {
      const int t_pos = c.t0 + r;
      const int t_in = t_pos * p.stride + c.tap * p.tap_step - p.pad_left;
      const bool ok = t_pos < p.T_pos && t_in >= 0 && t_in < p.T_in;
      const float* src = p.x + ((size_t)c.b * p.Cin + c.cb * 32) * p.T_in + (ok ? t_in : 0);
      const int nch = p.Cin - c.cb * 32;  // valid channels of this block (>= 32 except in a padded last block)
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = (ok && i < nch) ? __ldg(src + (size_t)i * p.T_in) : 0.f;
    };
    float v[32], vn[32];
    Cur cn = cur_init((uint32_t)grp);
    if ((uint32_t)grp < n_it) gather(cn, v);
    for (uint32_t it = (uint32_t)grp; it < n_it; it += 2) {
      const bool more = it + 2 < n_it;
      if (more) {
        cur_step2(cn);
        gather(cn, vn);
      }
      const uint32_t sa = it % A_SLOTS, pa = (it / A_SLOTS) & 1;
      smem_bar_wait(&a_empty[sa], pa ^ 1);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + sa * 64;
#pragma unroll
      for (int half = 0; half < 2; ++half) {  // 16 channels at a time: hi / lo of a half live in 32 registers next to v / vn
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float f = p.pre_elu ? elu_fast(v[half * 16 + i]) : v[half * 16 + i];
          const uint32_t h = tf32_rna_bits(f);
          hi[i] = h;
          lo[i] = tf32_rna_bits(f - __uint_as_float(h));
        }
        tmem_st16(taddr + half * 16, hi);
        tmem_st16(taddr + 32 + half * 16, lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(&a_full[sa]);
      if (more) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = vn[i];
      }
    }
  } else if (warp >= 12) {
    // ================= epilogue
    const int q = warp & 3;
    const int r = q * 32 + lane;
    pdl_wait();  // residual operand / output buffer ordering against the preceding kernels
    // Per tile: row base of this thread's position; per column only a multiply-add (conv) or a phase / channel split (transposed).
    struct Row {
      size_t base;   // conv: &y[b, n0, t_pos]; transposed: &y[b, 0, 0]
      int t_pos;
      bool ok;
    };
    auto row_of = [&](int j) -> Row {
      const int tile = cta + j * G;
      const int b = tile / p.tiles_per_b, t_pos = (tile - b * p.tiles_per_b) * CU_BM + r;
      Row rw;
      rw.t_pos = t_pos;
      rw.ok = t_pos < p.T_pos;
      rw.base = p.tr_stride ? (size_t)b * p.Cout * p.T_out : ((size_t)b * p.Cout + p.n0) * p.T_out + t_pos;
      return rw;
    };
    auto where = [&](const Row& rw, int col, size_t& idx, int& co) -> bool {
      if (!p.tr_stride) {
        co = p.n0 + col;
        idx = rw.base + (size_t)col * p.T_out;
        return rw.ok && col < p.n_cols;
      }
      // transposed convolution: column = phase * Cout + channel, output time j * s + phase - crop
      const int c = p.n0 + col, ph = c / p.Cout;
      co = c - ph * p.Cout;
      const int t = rw.t_pos * p.tr_stride + ph - p.tr_crop;
      idx = rw.base + (size_t)co * p.T_out + t;
      return rw.ok && col < p.n_cols && t >= 0 && t < p.T_out;
    };
    // bias + residual of 32 columns: read-only loads that do not depend on the accumulator, so they are issued one chunk AHEAD
    // (before the wait for the tile's MMAs / before the previous chunk's stores) and their latency hides behind the tensor cores.
    auto load_add = [&](const Row& rw, int c0, float (&add)[32]) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        size_t idx;
        int co;
        const bool ok = where(rw, c0 + i, idx, co);
        float a = (ok && p.bias) ? __ldg(p.bias + co) : 0.f;
        if (ok && p.res && !p.prelu) a += __ldg(p.res + idx);  // with a PReLU the residual is added after the activation (below)
        add[i] = a;
      }
    };
    // Transposed convolution with stride 4 and no left crop (the 24 kHz up-sampling layer, modules/seanet.py:330-340): the four phases
    // of a channel are the four consecutive output samples 4 j .. 4 j + 3 of this thread's position j - one 128-bit store per channel,
    // a warp writes 512 contiguous bytes (the generic path below stores one float per column with an integer division in front of it:
    // 4.46 ms for 128 -> 64 at batch 16 x 10 s, 88 us per tile, all of it in these four warps).
    const bool fast_tr = p.tr_stride == 4 && p.tr_crop == 0 && p.res == nullptr && p.n0 == 0 && p.n_cols == 4 * p.Cout && (p.Cout % 16) == 0 &&
                         (p.T_out % 4) == 0 && 4 * p.Cout <= NO;
    const float slope = p.prelu ? __ldg(p.prelu) : 1.f;
    float add_next[32];
    Row row_next{};
    if (my_tiles > 0) {
      row_next = row_of(0);
      if (!fast_tr && p.tr_stride) load_add(row_next, 0, add_next);
    }
    for (int j = 0; j < my_tiles; ++j) {
      const int ab = NACC == 2 ? (j & 1) : 0;
      const uint32_t use = NACC == 2 ? (uint32_t)(j >> 1) : (uint32_t)j;
      const Row rw = row_next;
      if (j + 1 < my_tiles) row_next = row_of(j + 1);
      smem_bar_wait(&acc_full[ab], use & 1);
      tc_fence_after();
      if (fast_tr) {
        const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + CU_ACC_COL + ab * ACC_STRIDE;
        const int t4 = rw.t_pos * 4;
        const bool ok = rw.ok && t4 + 3 < p.T_out;
#pragma unroll 1
        for (int ch0 = 0; ch0 < p.Cout; ch0 += 8) {
          uint32_t v0[8], v1[8], v2[8], v3[8];
          tmem_ld8(tbase + ch0, v0);
          tmem_ld8(tbase + p.Cout + ch0, v1);
          tmem_ld8(tbase + 2 * p.Cout + ch0, v2);
          tmem_ld8(tbase + 3 * p.Cout + ch0, v3);
          tmem_wait_ld();
          if (ok) {
            float* dst = p.y + rw.base + (size_t)ch0 * p.T_out + t4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float bv = p.bias ? __ldg(p.bias + ch0 + i) : 0.f;
              *reinterpret_cast<float4*>(dst + (size_t)i * p.T_out) = make_float4(__uint_as_float(v0[i]) + bv, __uint_as_float(v1[i]) + bv,
                                                                                 __uint_as_float(v2[i]) + bv, __uint_as_float(v3[i]) + bv);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&acc_empty[ab]);
        continue;
      }
      if (!p.tr_stride) {
        // convolution: column c of the tile is output channel n0 + c at this thread's position - one pointer per row, one multiply-add
        // of the channel stride per column.  (The generic path below re-derives (channel, time, validity) per element: ncu counted
        // 1 300 instructions per tile in each epilogue warp for 32 columns, 96 % of the warps' time, and the MMA warp waiting on
        // `accumulator free` half of its time: the kernel was epilogue-bound on the narrow layers.)
        const uint32_t tb = tmem + ((uint32_t)(q * 32) << 16) + CU_ACC_COL + ab * ACC_STRIDE;
        float* yp = p.y + rw.base;
        const float* rp = p.res ? p.res + rw.base : nullptr;
        const float* bp = p.bias ? p.bias + p.n0 : nullptr;
        const size_t ts = (size_t)p.T_out;
        const bool row_ok = rw.ok;
#pragma unroll 1
        for (int c0 = 0; c0 < p.n_cols; c0 += 32) {
          const int nc = min(32, p.n_cols - c0);  // warp-uniform
          float add[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = (bp && i < nc) ? __ldg(bp + c0 + i) : 0.f;
            if (rp && !p.prelu && row_ok && i < nc) a += __ldg(rp + (size_t)(c0 + i) * ts);
            add[i] = a;
          }
          uint32_t v[32];
          tmem_ld32(tb + c0, v);
          tmem_wait_ld();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < nc) {
                float o = __uint_as_float(v[i]) + add[i];
                if (p.prelu) {
                  o = o > 0.f ? o : o * slope;
                  if (rp) o += __ldg(rp + (size_t)(c0 + i) * ts);
                }
                yp[(size_t)(c0 + i) * ts] = o;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&acc_empty[ab]);
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < NO; c0 += 32) {
        if (c0 >= p.n_cols) break;
        float add[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) add[i] = add_next[i];
        if (c0 + 32 < p.n_cols && c0 + 32 < NO)
          load_add(rw, c0 + 32, add_next);
        else if (j + 1 < my_tiles)
          load_add(row_next, 0, add_next);
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + CU_ACC_COL + ab * ACC_STRIDE + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          size_t idx;
          int co;
          if (where(rw, c0 + i, idx, co)) {
            float o = __uint_as_float(v[i]) + add[i];
            if (p.prelu) {
              o = o > 0.f ? o : o * slope;
              if (p.res) o += __ldg(p.res + idx);
            }
            p.y[idx] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(&acc_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// torch Conv1d weight (Cout, Cin, Ktaps) -> [2][Cout][KT] hi / lo tf32 planes, CinP = Cin rounded up to 32 (zero columns); column
// order tap * CinP + ci (register-gather mode) or, cb_major, (ci / 32 * Ktaps + tap) * 32 + ci % 32 (staged mode: all taps of a
// 32-channel block are consecutive k-blocks)
__global__ void conv_umma_repack_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int CinP, int Ktaps, int cb_major) {
  pdl_launch_dependents();
  pdl_wait();
  const int KT = CinP * Ktaps;
  const long long n = (long long)Cout * KT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / KT), k = (int)(i - (long long)co * KT);
    int tap, ci;
    if (cb_major) {
      const int blk = k >> 5, cbk = blk / Ktaps;
      tap = blk - cbk * Ktaps;
      ci = cbk * 32 + (k & 31);
    } else {
      tap = k / CinP;
      ci = k - tap * CinP;
    }
    const float v = ci < Cin ? w[((size_t)co * Cin + ci) * Ktaps + tap] : 0.f;
    const uint32_t h = tf32_rna_bits(v);
    wp[i] = __uint_as_float(h);
    wp[n + i] = __uint_as_float(tf32_rna_bits(v - __uint_as_float(h)));
  }
}
// per-phase transposed-conv operand (s, Cout, Cin, 2) (repack_convtr_phase_kernel, ua2_codec.cu) -> [2][s * Cout][tap * Cin + ci] planes
__global__ void convtr_umma_repack_kernel(const float* __restrict__ w_phase, float* __restrict__ wp, int rows, int Cin) {
  pdl_launch_dependents();
  pdl_wait();
  const int KT = 2 * Cin;
  const long long n = (long long)rows * KT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / KT), k = (int)(i - (long long)row * KT);
    const int tap = k / Cin, ci = k - tap * Cin;
    const float v = w_phase[((size_t)row * Cin + ci) * 2 + tap];
    const uint32_t h = tf32_rna_bits(v);
    wp[i] = __uint_as_float(h);
    wp[n + i] = __uint_as_float(tf32_rna_bits(v - __uint_as_float(h)));
  }
}

struct ConvUmmaScratch {
  float* wp = nullptr;
  size_t floats = 0;
};
ConvUmmaScratch g_cu;
int g_conv_umma = 1;
int g_conv_umma_staged = 1;  // option "conv_umma_staged": measurement switch of the TMA-staged activation path

cudaError_t reserve_wp(const LaunchCtx& lc, size_t need) {
  if (need <= g_cu.floats) return cudaSuccess;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(lc.stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return cudaErrorNotSupported;  // the scratch would have to grow
  if (g_cu.wp) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    cudaFree(g_cu.wp);
    g_cu.wp = nullptr;
    g_cu.floats = 0;
  }
  const size_t want = std::max(need, (size_t)1 << 21);
  cudaError_t e = cudaMalloc((void**)&g_cu.wp, want * sizeof(float));
  if (e != cudaSuccess) return e;
  g_cu.floats = want;
  return cudaSuccess;
}

template <int NO>
cudaError_t launch_no(const LaunchCtx& lc, const CUtensorMap& tmW, const CUtensorMap& tmX, ConvUmmaParams p) {
  constexpr int STAGE_BYTES = 2 * NO * 128;
  size_t x_ring = 0;
  if (p.staged) {  // activation ring first (2-4 stages), the weights get the rest
    const size_t xb = (size_t)p.xw * 4096;
    p.n_xstages = (int)std::min<size_t>(CU_X_STAGES, std::max<size_t>(2, (72 * 1024) / xb));
    x_ring = p.n_xstages * xb;
    if (x_ring + 2 * STAGE_BYTES > 200 * 1024) return cudaErrorNotSupported;
  }
  int stages = std::min(CU_MAX_STAGES, (int)((200 * 1024 - x_ring) / STAGE_BYTES));
  if (p.KB <= stages) {
    stages = p.KB;
    p.resident = 1;
  } else {
    stages = std::min(stages, 4);
    p.resident = 0;
  }
  p.n_stages = stages;
  const size_t smem = 1024 + (size_t)stages * STAGE_BYTES + x_ring + (2 * CU_MAX_STAGES + 2 * CU_A_SLOTS + 4 + 2 * CU_X_STAGES) * 8 + 16;
  static DeviceOnce attr;
  if (attr.need()) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e != cudaSuccess) return e;
  }
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(p.n_tiles, sms);
  return launch(lc, conv_umma_kernel<NO>, dim3(grid), dim3(CU_THREADS), smem, tmW, tmX, p);
}

// columns [n0, n0 + n_cols) of a GEMM whose weight planes hold `rows_total` rows
cudaError_t launch_cols(const LaunchCtx& lc, ConvUmmaParams p, int KT, int rows_total) {
  const int NO = p.n_cols <= 32 ? 32 : p.n_cols <= 64 ? 64 : p.n_cols <= 128 ? 128 : 256;
  CUtensorMap tmW, tmX;
  if (!make_tmap(&tmW, g_cu.wp, KT, rows_total, 2, NO, false)) return cudaErrorNotSupported;
  tmX = tmW;  // unused unless staged
  if (p.staged && !make_tmap(&tmX, p.x, p.T_in, (long long)p.B * p.Cin, 1, 32, false)) return cudaErrorNotSupported;
  switch (NO) {
    case 32: return launch_no<32>(lc, tmW, tmX, p);
    case 64: return launch_no<64>(lc, tmW, tmX, p);
    case 128: return launch_no<128>(lc, tmW, tmX, p);
    default: return launch_no<256>(lc, tmW, tmX, p);
  }
}

}  // namespace

void set_conv_umma(int v) { g_conv_umma = v ? 1 : 0; }
void set_conv_umma_staged(int v) { g_conv_umma_staged = v ? 1 : 0; }
int get_conv_umma() { return g_conv_umma; }

// cudaErrorNotSupported when the layer is not served here (the caller continues on its other paths)
cudaError_t launch_conv1d_umma(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res, float* y, int B,
                               int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation, int pad_left, int pre_elu,
                               int replicate, const float* prelu) {
  if (!g_conv_umma || dilation < 1 || replicate || (Cin & 15) || Cin < 32 || Cout < 16 || (long long)B * T_out < 4 * CU_BM) return cudaErrorNotSupported;
  if (pre_elu && prelu) return cudaErrorNotSupported;
  // pointwise convolutions with fewer than 4 k-blocks per tile are all epilogue (measured at batch 16 x 10 s: 32 -> 64 at 24 kHz 3.0 ms
  // here against 1.6 ms on the fp32 register-tiled core, 64 -> 128 at 6 kHz 2.6 against 2.2): they stay on the SIMT core
  if (Ktaps == 1 && Cin < 128) return cudaErrorNotSupported;
  const int CinP = (Cin + 31) / 32 * 32;
  const long long KT = (long long)CinP * Ktaps;
  if (KT > (1 << 20)) return cudaErrorNotSupported;
  const int tiles_per_b = (T_out + CU_BM - 1) / CU_BM;
  const long long n_tiles = (long long)B * tiles_per_b;
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // few positions and a long reduction (the low-rate stages): one tile per SM leaves most SMs idle and every column chunk gathers the
  // activations again - the materialised-im2col GEMM (ua2_convtc.cu) spreads the same work over all SMs by stream-K
  if (Cout > 256 && n_tiles * 2 < sms && KT >= 256) return cudaErrorNotSupported;
  if (Ktaps == 1 && Cin >= 256 && Cout > 256) return cudaErrorNotSupported;  // a plain GEMM: transpose to rows + the linears' kernel (ua2_convtc.cu)
  // stride 1 and a row pitch TMA can address: activations staged in shared memory once per (tile, channel block) for all taps
  const int halo = (Ktaps - 1) * dilation;
  const int x_shift = (4 - (pad_left & 3)) & 3;  // TMA needs the first sample of a box on a 16-byte boundary: t0 - pad_left - x_shift is a multiple of 4
  const int xw = (CU_BM + halo + x_shift + 31) / 32;  // boxes of 32 samples
  const bool staged = g_conv_umma_staged && stride == 1 && Ktaps >= 2 && (T_in & 3) == 0 && xw <= 8 &&
                      (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  cudaError_t e = reserve_wp(lc, (size_t)2 * Cout * KT);
  if (e != cudaSuccess) return e;
  e = launch(lc, conv_umma_repack_kernel, dim3((unsigned)std::min<long long>((Cout * KT + 255) / 256, 148 * 8)), dim3(256), 0, w_torch, g_cu.wp, Cout,
             Cin, CinP, Ktaps, staged ? 1 : 0);
  if (e != cudaSuccess) return e;
  ConvUmmaParams p{};
  p.x = x;
  p.bias = bias;
  p.res = res;
  p.y = y;
  p.prelu = prelu;
  p.B = B;
  p.Cin = Cin;
  p.Cout = Cout;
  p.T_in = T_in;
  p.T_out = T_out;
  p.Ktaps = Ktaps;
  p.stride = stride;
  p.pad_left = pad_left;
  p.pre_elu = pre_elu;
  p.tap_step = dilation;
  p.staged = staged ? 1 : 0;
  p.xw = xw;
  p.x_shift = x_shift;
  p.T_pos = T_out;
  p.KB = (int)(KT / 32);
  p.cb_per_tap = CinP / 32;
  p.tiles_per_b = tiles_per_b;
  if (n_tiles * p.KB >= (1LL << 31)) return cudaErrorNotSupported;
  p.n_tiles = (int)n_tiles;
  for (int n0 = 0; n0 < Cout; n0 += 256) {  // 256 output channels per launch (a wider layer gathers its activations once per chunk)
    p.n0 = n0;
    p.n_cols = std::min(256, Cout - n0);
    if ((e = launch_cols(lc, p, (int)KT, Cout)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// Transposed convolution with kernel = 2 * stride (modules/conv.py:306-329) over the per-phase weights of launch_convtr1d_gemm:
//   y[b, n, j * s + ph - crop_left] = bias[n] + sum_ci ( w[ph][n][ci][0] f(x[b, ci, j]) + w[ph][n][ci][1] f(x[b, ci, j - 1]) )
// as ONE implicit GEMM over the input grid j with K = 2 Cin and the s * Cout (phase, channel) pairs as columns, 256 columns per launch.
cudaError_t launch_convtr1d_umma(const LaunchCtx& lc, const float* x, const float* w_phase, const float* bias, float* y, int B, int Cin, int Cout,
                                 int T_in, int stride, int pre_elu, int crop_left, int T_out) {
  const int Tj = (crop_left + T_out > T_in * stride) ? T_in + 1 : T_in;
  const int rows = stride * Cout, KT = 2 * Cin;
  if (!g_conv_umma || (Cin & 31) || Cout < 16 || (long long)B * Tj < 4 * CU_BM || rows > 4096) return cudaErrorNotSupported;
  cudaError_t e = reserve_wp(lc, (size_t)2 * rows * KT);
  if (e != cudaSuccess) return e;
  e = launch(lc, convtr_umma_repack_kernel, dim3((unsigned)std::min<long long>(((long long)rows * KT + 255) / 256, 148 * 8)), dim3(256), 0, w_phase,
             g_cu.wp, rows, Cin);
  if (e != cudaSuccess) return e;
  ConvUmmaParams p{};
  p.x = x;
  p.bias = bias;
  p.res = nullptr;
  p.y = y;
  p.B = B;
  p.Cin = Cin;
  p.Cout = Cout;
  p.T_in = T_in;
  p.T_out = T_out;
  p.Ktaps = 2;
  p.stride = 1;
  p.pad_left = 0;
  p.pre_elu = pre_elu;
  p.tap_step = -1;
  p.T_pos = Tj;
  p.tr_stride = stride;
  p.tr_crop = crop_left;
  p.KB = KT / 32;
  p.cb_per_tap = Cin / 32;
  p.tiles_per_b = (Tj + CU_BM - 1) / CU_BM;
  const long long n_tiles = (long long)B * p.tiles_per_b;
  if (n_tiles * p.KB >= (1LL << 31)) return cudaErrorNotSupported;
  p.n_tiles = (int)n_tiles;
  // whole phases per launch, at most 256 columns (Cout <= 256); a wider layer goes one phase at a time in chunks of 256 channels
  const int ph_per = Cout <= 256 ? std::max(1, 256 / Cout) : 0;
  if (ph_per == 0) return cudaErrorNotSupported;
  for (int ph0 = 0; ph0 < stride; ph0 += ph_per) {
    p.n0 = ph0 * Cout;
    p.n_cols = std::min(ph_per, stride - ph0) * Cout;
    if ((e = launch_cols(lc, p, KT, rows)) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace ua2
