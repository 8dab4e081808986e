// Tensor-core flash attention for the flow-matching decoder's blocks (models/attention.py:308-415 -> diffusers Attention; unmasked
// self-attention over T = 500 latent frames, 24 heads x 64) in the reference's own precision for this block: bf16 operands, fp32
// accumulation and softmax statistics (reason_tokenizer.py:265 runs it under torch.autocast(bfloat16) -> SDPA in bf16).
//
// The fp32 SIMT kernel (dit_attn_kernel) is shared-memory-bandwidth bound at 22 % of the FMA pipe (profiles/r1_flow_decoder.md) and
// costs ~6 of the 14 ms of an estimator call.  Here both contractions run on tcgen05:
//
//   CTA = 128 query rows of one (batch, head).  Q (128 x 64 bf16) stays in shared memory; key blocks of 128 keys stream through a ring.
//   warp 0      TMA producer: Q once, then K_j / V_j tiles (128 x 64 bf16 = 128-byte rows, SWIZZLE_128B)
//   warp 1      MMA issuer:   S_j = Q K_j^T  (kind::f16, A and B from shared memory, K-major)      -> TMEM, double-buffered
//                             O  += P_j V_j  (A = P_j from TENSOR MEMORY, B = V_j from shared memory in MN-major form: rows are keys)
//   warps 4-7   softmax, thread = query row (TMEM lane): tcgen05.ld of its S row, running max / sum in registers (no shuffles),
//               p = exp2((s - m) * scale * log2 e) packed to bf16 pairs and stored back to tensor memory as the next MMA's A operand;
//               when the running max moves, the O accumulator (64 columns) is rescaled in place (tcgen05.ld / st)
//   epilogue    the same four warps: O / l -> out (B, T, H * 64) fp32
//
// TMEM: S0 [0,128) S1 [128,256) P [256,320) O [320,384) of 512 columns.
#include <cuda_bf16.h>

#include "ua2_kernels.cuh"
#include "ua2_tcgen05.cuh"
#include "ua2_umma.cuh"

namespace ua2 {
namespace {

using namespace tc;

constexpr int FA_BM = 128, FA_BN = 128, FA_HS = 64;
constexpr int FA_TILE_BYTES = FA_BN * FA_HS * 2;   // 16 KB
constexpr int FA_THREADS = 256;


// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// MN-major operand whose rows (the K index of the MMA) are 128 bytes = 64 bf16 of the MN index, SWIZZLE_128B: 8-row groups 1024 B apart
// (stride byte offset), one swizzle row wide (leading byte offset unused); a k-step of 16 rows advances the start by 2048 bytes
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// SBUF = 2: scores double-buffered (S_{j+1} accumulates while the softmax warps work on S_j), 512 TMEM columns, one CTA per SM.
// SBUF = 1: one score buffer, 256 columns and 80 KB of shared memory: two CTAs per SM overlap each other instead.
// BIAS (WavLM's gated relative position bias, csrc/ua2_wavlm.cu): logit(i, j) = q_i k_j / sqrt(hs) + gate[b, h, i] * tab[h, j - i + T - 1];
// the softmax warps work on z = s * scale * log2 e + gate * log2 e * tab in the exp2 domain, so the running maximum, the rescale
// factor and the probabilities all refer to z.  tab (H, 2 T - 1) is 12 KB per head: the two passes read it through L1.
template <int SBUF, bool BIAS>
__global__ void __launch_bounds__(FA_THREADS, SBUF == 1 ? 2 : 1)
flash_bf16_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                  float* __restrict__ out, __nv_bfloat16* __restrict__ out16, int T, int H, float scale_log2e, const float* __restrict__ gate,
                  const float* __restrict__ tab) {
  constexpr int FA_S_COL = 0, FA_P_COL = SBUF * 128, FA_O_COL = SBUF * 128 + 64;
  constexpr uint32_t FA_TMEM_COLS = SBUF == 2 ? 512u : 256u;
  constexpr int FA_STAGES = SBUF == 2 ? 3 : 2;  // K / V ring
  extern __shared__ uint8_t fa_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = base;                                   // 16 KB
  uint8_t* k_ring = q_s + FA_TILE_BYTES;                 // FA_STAGES x 16 KB
  uint8_t* v_ring = k_ring + FA_STAGES * FA_TILE_BYTES;  // FA_STAGES x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_ring + FA_STAGES * FA_TILE_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = q_full + 1;          // FA_STAGES
  uint64_t* kv_empty = kv_full + FA_STAGES;
  uint64_t* s_full = kv_empty + FA_STAGES; // 2 (S double buffer)
  uint64_t* s_empty = s_full + 2;          // 2
  uint64_t* p_full = s_empty + 2;          // 1
  uint64_t* o_done = p_full + 1;           // 1: the P V MMA of a key block has completed (P and O may be touched again)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(o_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row_base = (b * H + h) * T;   // first row of this (batch, head) in the (B * H * T, 64) operand matrices
  const int n_kb = (T + FA_BN - 1) / FA_BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    smem_bar_init(q_full, 1);
    for (int i = 0; i < FA_STAGES; ++i) {
      smem_bar_init(&kv_full[i], 1);
      smem_bar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      smem_bar_init(&s_full[i], 1);
      smem_bar_init(&s_empty[i], 4);
    }
    smem_bar_init(p_full, 4);
    smem_bar_init(o_done, 1);
    smem_bar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(tmem_base_smem)), "r"(FA_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_smem;
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= TMA producer
    pdl_wait();
    if (elect_one()) {
      smem_bar_arrive_expect_tx(q_full, FA_TILE_BYTES);
      tma_load_2d(q_s, &tmQ, 0, row_base + qt * FA_BM, q_full, POLICY_EVICT_FIRST);
    }
    __syncwarp();
    for (int j = 0; j < n_kb; ++j) {
      const int s = j % FA_STAGES;
      const uint32_t ph = (uint32_t)(j / FA_STAGES) & 1;
      smem_bar_wait(&kv_empty[s], ph ^ 1);
      if (elect_one()) {
        smem_bar_arrive_expect_tx(&kv_full[s], 2 * FA_TILE_BYTES);
        tma_load_2d(k_ring + s * FA_TILE_BYTES, &tmK, 0, row_base + j * FA_BN, &kv_full[s], POLICY_EVICT_LAST);
        tma_load_2d(v_ring + s * FA_TILE_BYTES, &tmV, 0, row_base + j * FA_BN, &kv_full[s], POLICY_EVICT_LAST);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================= MMA issuer
    // S = Q K^T: M 128, N 128, bf16 x bf16 -> fp32, both operands K-major
    constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FA_BN >> 3) << 17) | ((uint32_t)(FA_BM >> 4) << 24);
    // O += P V: M 128, N 64, A from tensor memory (K-major), B MN-major (bit 16)
    constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(FA_HS >> 3) << 17) | ((uint32_t)(FA_BM >> 4) << 24);
    smem_bar_wait(q_full, 0);
    const uint64_t qd = smem_desc_sw128(smem_addr_u32(q_s));
    auto issue_s = [&](int j) {  // S_j into buffer j & 1
      const int s = j % FA_STAGES;
      smem_bar_wait(&kv_full[s], (uint32_t)(j / FA_STAGES) & 1);
      smem_bar_wait(&s_empty[j % SBUF], ((uint32_t)(j / SBUF) & 1) ^ 1);
      tc_fence_after();
      const uint64_t kd = smem_desc_sw128(smem_addr_u32(k_ring + s * FA_TILE_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < FA_HS / 16; ++ks) mma_bf16_ss(tmem + FA_S_COL + (j % SBUF) * 128, qd + (uint64_t)(ks * 2), kd + (uint64_t)(ks * 2), idesc_s, ks ? 1u : 0u);
        tc_commit(&s_full[j % SBUF]);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int j = 0; j < n_kb; ++j) {
      if (SBUF == 2 && j + 1 < n_kb) issue_s(j + 1);  // the next block's scores accumulate while the softmax warps work on this one
      const int s = j % FA_STAGES;
      smem_bar_wait(p_full, (uint32_t)j & 1);
      tc_fence_after();
      const uint64_t vd = smem_desc_mn_sw128(smem_addr_u32(v_ring + s * FA_TILE_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < FA_BN / 16; ++ks)  // 16 keys per k-step: 8 TMEM columns of P, 16 rows (2048 B) of V
          mma_bf16_ts(tmem + FA_O_COL, tmem + FA_P_COL + ks * 8, vd + (uint64_t)(ks * 128), idesc_o, (j == 0 && ks == 0) ? 0u : 1u);
        tc_commit(&kv_empty[s]);
        tc_commit(o_done);
      }
      __syncwarp();
      if (SBUF == 1 && j + 1 < n_kb) issue_s(j + 1);  // the score buffer is free once P_j has been written
    }
  } else if (warp >= 4) {
    // ================= softmax + epilogue: thread = query row
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float gl = 0.f;              // BIAS: gate[b, h, row] * log2 e
    const float* tb = nullptr;   // BIAS: tab[h] shifted so that tb[key] is this row's entry; rows past T read row T - 1's (never stored)
    if (BIAS) {
      const int tq = min(qt * FA_BM + r, T - 1);
      gl = gate[((size_t)b * H + h) * T + tq] * 1.4426950408889634f;
      tb = tab + (size_t)h * (2 * T - 1) + (T - 1 - tq);
    }
    for (int j = 0; j < n_kb; ++j) {
      smem_bar_wait(&s_full[j % SBUF], (uint32_t)(j / SBUF) & 1);
      tc_fence_after();
      // ---- row maximum of this block (keys past T are masked)
      const int n_valid = min(FA_BN, T - j * FA_BN);
      float mx = m_run;
      uint32_t sv[32];
#pragma unroll 1
      for (int c0 = 0; c0 < FA_BN; c0 += 32) {
        tmem_ld32(tmem + lane_addr + FA_S_COL + (j % SBUF) * 128 + c0, sv);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c0 + i < n_valid) {
            if (BIAS) {
              mx = fmaxf(mx, fmaf(__uint_as_float(sv[i]), scale_log2e, gl * tb[j * FA_BN + c0 + i]));
            } else {
              mx = fmaxf(mx, __uint_as_float(sv[i]));
            }
          }
      }
      const float alpha = BIAS ? ex2(m_run - mx) : ex2((m_run - mx) * scale_log2e);  // 0 for the first block (m_run = -inf)
      // ---- the previous block's P V must have completed before P is overwritten and O rescaled
      if (j > 0) {
        smem_bar_wait(o_done, (uint32_t)(j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {  // warp-uniform: tcgen05.ld / st are .sync.aligned
#pragma unroll 1
          for (int c0 = 0; c0 < FA_HS; c0 += 32) {
            uint32_t ov[32];
            tmem_ld32(tmem + lane_addr + FA_O_COL + c0, ov);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st32(tmem + lane_addr + FA_O_COL + c0, ov);
          }
        }
      }
      // ---- p = exp2((s - mx) * scale * log2 e) -> bf16 pairs -> P (64 columns); row sum in fp32 of the ROUNDED values
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < FA_BN; c0 += 32) {
        tmem_ld32(tmem + lane_addr + FA_S_COL + (j % SBUF) * 128 + c0, sv);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = 0.f, p1 = 0.f;
          if (BIAS) {
            const int kj = j * FA_BN + c0 + 2 * i;
            if (c0 + 2 * i < n_valid) p0 = ex2(fmaf(__uint_as_float(sv[2 * i]), scale_log2e, gl * tb[kj]) - mx);
            if (c0 + 2 * i + 1 < n_valid) p1 = ex2(fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2e, gl * tb[kj + 1]) - mx);
          } else {
            if (c0 + 2 * i < n_valid) p0 = ex2((__uint_as_float(sv[2 * i]) - mx) * scale_log2e);
            if (c0 + 2 * i + 1 < n_valid) p1 = ex2((__uint_as_float(sv[2 * i + 1]) - mx) * scale_log2e);
          }
          pk[i] = pack_bf16(p0, p1);
          const __nv_bfloat162 rb = *reinterpret_cast<const __nv_bfloat162*>(&pk[i]);
          sum += __bfloat162float(rb.x) + __bfloat162float(rb.y);
        }
        tmem_st16(tmem + lane_addr + FA_P_COL + (c0 >> 1), pk);
      }
      tmem_wait_st();
      l_run = l_run * alpha + sum;
      m_run = mx;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        bar_arrive(&s_empty[j % SBUF]);
        bar_arrive(p_full);
      }
    }
    // ---- O / l -> out[b, t, h * 64 + d]
    smem_bar_wait(o_done, (uint32_t)(n_kb - 1) & 1);
    tc_fence_after();
    const int t = qt * FA_BM + r;
    const float inv = 1.f / l_run;
#pragma unroll 1
    for (int c0 = 0; c0 < FA_HS; c0 += 32) {
      uint32_t ov[32];
      tmem_ld32(tmem + lane_addr + FA_O_COL + c0, ov);
      tmem_wait_ld();
      if (t < T && out16 != nullptr) {
        __nv_bfloat16* dst = out16 + ((size_t)b * T + t) * (size_t)(H * FA_HS) + h * FA_HS + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dst + i) = make_uint4(pack_bf16(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv),
                                                          pack_bf16(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv),
                                                          pack_bf16(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv),
                                                          pack_bf16(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv));
      } else if (t < T) {
        float* dst = out + ((size_t)b * T + t) * (size_t)(H * FA_HS) + h * FA_HS + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv,
                                                            __uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(FA_TMEM_COLS) : "memory");
  }
}

int g_flash_sbuf = 0;  // 0 = by grid size; 1 / 2 force a variant (option "flash_sbuf", measurement switch)
int flash_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

void set_flash_sbuf(int v) { g_flash_sbuf = v == 1 || v == 2 ? v : 0; }

// q16 / k16 / v16: (B, H, T, 64) bf16; out: (B, T, H * 64) fp32, or out16 (same shape, bf16) when not NULL.
// gate (B, H, T) + tab (H, 2 T - 1), both or neither: the additive bias gate[b, h, i] * tab[h, j - i + T - 1] on the scaled scores.
// cudaErrorNotSupported unless head size 64.
template <int SBUF, bool BIAS>
cudaError_t launch_flash_variant(const LaunchCtx& lc, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, float* out, void* out16,
                                 int B, int T, int H, float scale_log2e, const float* gate, const float* tab) {
  const size_t smem = 1024 + (size_t)(1 + 2 * (SBUF == 1 ? 2 : 3)) * FA_TILE_BYTES + 32 * 8 + 16;
  static DeviceOnce once;
  if (once.need()) {
    cudaError_t e = cudaFuncSetAttribute(flash_bf16_kernel<SBUF, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const dim3 grid((T + FA_BM - 1) / FA_BM, H, B);
  return launch(lc, flash_bf16_kernel<SBUF, BIAS>, grid, dim3(FA_THREADS), smem, tmQ, tmK, tmV, out, static_cast<__nv_bfloat16*>(out16), T, H,
                scale_log2e, gate, tab);
}

cudaError_t launch_flash_bf16_bias(const LaunchCtx& lc, const void* q16, const void* k16, const void* v16, float* out, void* out16, int B, int T, int H,
                                   int hs, const float* gate, const float* tab) {
  if (hs != FA_HS || T < 1 || B < 1 || H < 1 || B > 65535 || H > 65535 || (gate == nullptr) != (tab == nullptr)) return cudaErrorNotSupported;
  const long long rows = (long long)B * H * T;
  CUtensorMap tmQ, tmK, tmV;
  if (!make_tmap(&tmQ, q16, FA_HS, rows, 1, FA_BM, false, true) || !make_tmap(&tmK, k16, FA_HS, rows, 1, FA_BN, false, true) ||
      !make_tmap(&tmV, v16, FA_HS, rows, 1, FA_BN, false, true))
    return cudaErrorNotSupported;
  const int ctas = ((T + FA_BM - 1) / FA_BM) * H * B;
  int sbuf = g_flash_sbuf;
  if (sbuf == 0) sbuf = (ctas > flash_sm_count() ? 1 : 2);  // more tiles than SMs: co-resident CTA pairs instead of a second wave
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)hs);
  if (gate != nullptr) {
    if (sbuf == 1) return launch_flash_variant<1, true>(lc, tmQ, tmK, tmV, out, out16, B, T, H, scale_log2e, gate, tab);
    return launch_flash_variant<2, true>(lc, tmQ, tmK, tmV, out, out16, B, T, H, scale_log2e, gate, tab);
  }
  if (sbuf == 1) return launch_flash_variant<1, false>(lc, tmQ, tmK, tmV, out, out16, B, T, H, scale_log2e, gate, tab);
  return launch_flash_variant<2, false>(lc, tmQ, tmK, tmV, out, out16, B, T, H, scale_log2e, gate, tab);
}

cudaError_t launch_flash_bf16(const LaunchCtx& lc, const void* q16, const void* k16, const void* v16, float* out, void* out16, int B, int T, int H,
                              int hs) {
  return launch_flash_bf16_bias(lc, q16, k16, v16, out, out16, B, T, H, hs, nullptr, nullptr);
}

}  // namespace ua2
