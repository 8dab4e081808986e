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
constexpr int CU_A_SLOTS = 4;    // TMEM A ring (64 columns per slot)
constexpr int CU_ACC_COL = 256;  // accumulators: columns [256, 256 + NO) and, double-buffered, [384, 384 + NO)
constexpr int CU_MAX_STAGES = 24;
constexpr int CU_THREADS = 512;

struct ConvUmmaParams {
  const float* x;
  const float* bias;
  const float* res;
  float* y;
  int B, Cin, Cout, T_in, T_out, Ktaps, stride, pad_left, pre_elu;
  int KB;           // k-blocks of 32: Ktaps * Cin / 32
  int cb_per_tap;   // Cin / 32
  int tiles_per_b;  // ceil(T_out / 128)
  int n_tiles;
  int n_stages;     // weight ring stages
  int resident;     // KB <= n_stages: every k-block is loaded once and stays
};

__device__ __forceinline__ float elu_fast(float v) { return v > 0.f ? v : __expf(v) - 1.f; }

template <int NO>
__global__ void __launch_bounds__(CU_THREADS, 1) conv_umma_kernel(const __grid_constant__ CUtensorMap tmW, const ConvUmmaParams p) {
  constexpr int STAGE_BYTES = 2 * NO * 128;  // hi tile then lo tile, NO rows of 128 bytes each
  constexpr int NACC = NO <= 128 ? 2 : 1;
  extern __shared__ uint8_t cu_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(cu_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* w_ring = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring + (size_t)p.n_stages * STAGE_BYTES);
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + CU_MAX_STAGES;
  uint64_t* a_full = w_empty + CU_MAX_STAGES;
  uint64_t* a_empty = a_full + CU_A_SLOTS;
  uint64_t* acc_full = a_empty + CU_A_SLOTS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

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
        tma_load_3d(w_ring + (size_t)sw * STAGE_BYTES, &tmW, kb * 32, 0, 0, &w_full[sw], POLICY_EVICT_LAST);
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
      const uint32_t d_tmem = tmem + CU_ACC_COL + ab * 128;
      for (int kb = 0; kb < p.KB; ++kb, ++it) {
        const uint32_t sw = p.resident ? (uint32_t)kb : it % (uint32_t)p.n_stages;
        const uint32_t pw = p.resident ? 0u : (it / (uint32_t)p.n_stages) & 1;
        const uint32_t sa = it % CU_A_SLOTS, pa = (it / CU_A_SLOTS) & 1;
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
  } else if (warp >= 4 && warp < 12) {
    // ================= A producers: f(x) of one tap x 32 channels for this thread's output position -> TMEM (hi, lo)
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t n_it = (uint32_t)my_tiles * (uint32_t)p.KB;
    pdl_wait();  // x comes from the preceding kernel
    for (uint32_t it = (uint32_t)grp; it < n_it; it += 2) {
      const int j = (int)(it / (uint32_t)p.KB), kb = (int)(it - (uint32_t)j * (uint32_t)p.KB);
      const int tile = cta + j * G;
      const int b = tile / p.tiles_per_b, t_out = (tile - b * p.tiles_per_b) * CU_BM + r;
      const int tap = kb / p.cb_per_tap, c0 = (kb - tap * p.cb_per_tap) * 32;
      const int t_in = t_out * p.stride + tap - p.pad_left;
      const bool ok = t_out < p.T_out && t_in >= 0 && t_in < p.T_in;
      const float* src = p.x + ((size_t)b * p.Cin + c0) * p.T_in + (ok ? t_in : 0);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ok ? __ldg(src + (size_t)i * p.T_in) : 0.f;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float f = p.pre_elu ? elu_fast(v[i]) : v[i];
        const uint32_t h = tf32_rna_bits(f);
        hi[i] = h;
        lo[i] = tf32_rna_bits(f - __uint_as_float(h));
      }
      const uint32_t sa = it % CU_A_SLOTS, pa = (it / CU_A_SLOTS) & 1;
      smem_bar_wait(&a_empty[sa], pa ^ 1);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + sa * 64;
      tmem_st32(taddr, hi);
      tmem_st32(taddr + 32, lo);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(&a_full[sa]);
    }
  } else if (warp >= 12) {
    // ================= epilogue
    const int q = warp & 3;
    const int r = q * 32 + lane;
    pdl_wait();  // residual operand / output buffer ordering against the preceding kernels
    for (int j = 0; j < my_tiles; ++j) {
      const int ab = NACC == 2 ? (j & 1) : 0;
      const uint32_t use = NACC == 2 ? (uint32_t)(j >> 1) : (uint32_t)j;
      const int tile = cta + j * G;
      const int b = tile / p.tiles_per_b, t_out = (tile - b * p.tiles_per_b) * CU_BM + r;
      smem_bar_wait(&acc_full[ab], use & 1);
      tc_fence_after();
      const size_t row = ((size_t)b * p.Cout) * p.T_out + t_out;
#pragma unroll 1
      for (int c0 = 0; c0 < NO; c0 += 32) {
        if (c0 >= p.Cout) break;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + CU_ACC_COL + ab * 128 + c0, v);
        tmem_wait_ld();
        if (t_out < p.T_out) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int co = c0 + i;
            if (co < p.Cout) {
              float o = __uint_as_float(v[i]) + (p.bias ? p.bias[co] : 0.f);
              const size_t idx = row + (size_t)co * p.T_out;
              if (p.res) o += p.res[idx];
              p.y[idx] = o;
            }
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

// torch Conv1d weight (Cout, Cin, Ktaps) -> [2][Cout][tap * Cin + ci] hi / lo tf32 planes
__global__ void conv_umma_repack_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout, int Cin, int Ktaps) {
  pdl_launch_dependents();
  pdl_wait();
  const int KT = Cin * Ktaps;
  const long long n = (long long)Cout * KT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / KT), k = (int)(i - (long long)co * KT);
    const int tap = k / Cin, ci = k - tap * Cin;
    const float v = w[((size_t)co * Cin + ci) * Ktaps + tap];
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

template <int NO>
cudaError_t launch_no(const LaunchCtx& lc, const CUtensorMap& tmW, ConvUmmaParams p) {
  constexpr int STAGE_BYTES = 2 * NO * 128;
  int stages = std::min(CU_MAX_STAGES, (int)((200 * 1024) / STAGE_BYTES));
  if (p.KB <= stages) {
    stages = p.KB;
    p.resident = 1;
  } else {
    stages = std::min(stages, 4);
    p.resident = 0;
  }
  p.n_stages = stages;
  const size_t smem = 1024 + (size_t)stages * STAGE_BYTES + (2 * CU_MAX_STAGES + 2 * CU_A_SLOTS + 4) * 8 + 16;
  static size_t attr = 0;
  if (smem > attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_umma_kernel<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024));
    if (e != cudaSuccess) return e;
    attr = 220 * 1024;
  }
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(p.n_tiles, sms);
  return launch(lc, conv_umma_kernel<NO>, dim3(grid), dim3(CU_THREADS), smem, tmW, p);
}

}  // namespace

void set_conv_umma(int v) { g_conv_umma = v ? 1 : 0; }
int get_conv_umma() { return g_conv_umma; }

// cudaErrorNotSupported when the layer is not served here (the caller continues on its other paths)
cudaError_t launch_conv1d_umma(const LaunchCtx& lc, const float* x, const float* w_torch, const float* bias, const float* res, float* y, int B,
                               int Cin, int Cout, int T_in, int T_out, int Ktaps, int stride, int dilation, int pad_left, int pre_elu,
                               int replicate) {
  if (!g_conv_umma || dilation != 1 || replicate || (Cin & 31) || Cout < 16 || Cout > 256 || (long long)B * T_out < 4 * CU_BM) return cudaErrorNotSupported;
  const long long KT = (long long)Cin * Ktaps;
  if (KT > (1 << 20)) return cudaErrorNotSupported;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(lc.stream, &cs);
  const size_t need = (size_t)2 * Cout * KT;
  if (need > g_cu.floats) {
    if (cs != cudaStreamCaptureStatusNone) return cudaErrorNotSupported;  // the scratch would have to grow
    if (g_cu.wp) {
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) return e;
      cudaFree(g_cu.wp);
      g_cu.wp = nullptr;
      g_cu.floats = 0;
    }
    const size_t want = std::max(need, (size_t)1 << 20);
    cudaError_t e = cudaMalloc((void**)&g_cu.wp, want * sizeof(float));
    if (e != cudaSuccess) return e;
    g_cu.floats = want;
  }
  cudaError_t e = launch(lc, conv_umma_repack_kernel, dim3((unsigned)std::min<long long>((Cout * KT + 255) / 256, 148 * 8)), dim3(256), 0, w_torch,
                         g_cu.wp, Cout, Cin, Ktaps);
  if (e != cudaSuccess) return e;
  const int NO = Cout <= 32 ? 32 : Cout <= 64 ? 64 : Cout <= 128 ? 128 : 256;
  CUtensorMap tmW;
  if (!make_tmap(&tmW, g_cu.wp, (int)KT, Cout, 2, NO, false)) return cudaErrorNotSupported;
  ConvUmmaParams p{};
  p.x = x;
  p.bias = bias;
  p.res = res;
  p.y = y;
  p.B = B;
  p.Cin = Cin;
  p.Cout = Cout;
  p.T_in = T_in;
  p.T_out = T_out;
  p.Ktaps = Ktaps;
  p.stride = stride;
  p.pad_left = pad_left;
  p.pre_elu = pre_elu;
  p.KB = (int)(KT / 32);
  p.cb_per_tap = Cin / 32;
  p.tiles_per_b = (T_out + CU_BM - 1) / CU_BM;
  const long long n_tiles = (long long)B * p.tiles_per_b;
  if (n_tiles * p.KB >= (1LL << 31)) return cudaErrorNotSupported;
  p.n_tiles = (int)n_tiles;
  switch (NO) {
    case 32: return launch_no<32>(lc, tmW, p);
    case 64: return launch_no<64>(lc, tmW, p);
    case 128: return launch_no<128>(lc, tmW, p);
    default: return launch_no<256>(lc, tmW, p);
  }
}

}  // namespace ua2
