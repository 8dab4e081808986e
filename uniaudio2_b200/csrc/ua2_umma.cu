// Hand-written tcgen05 / TMEM / TMA mainloop for the many-row linears of the path (lit_model.py:424, 511, 592-595 F.linear;
// conv.py:232-254 as implicit GEMM; transformer_1d_flow.py linears) with fp32-class accuracy: "3xTF32, split on chip".
//
//   C (M x N) = X (M x K) * W (N x K)^T,  fp32 in HBM, every product evaluated as  X_hi W_hi + X_hi W_lo + X_lo W_hi  with
//   x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi)   (the dropped lo*lo term is ~2^-22 relative), fp32 accumulation in TMEM.
//
// Why not a library mainloop (round 1 used a CUTLASS collective over pre-split copies [hi|hi|lo]): the weights are the big
// operand (4 B / parameter, 12.4 GB per batch-32 decode frame) and a library mainloop cannot split them on chip - round 1 paid
// 12-28 B of HBM / L2 traffic per parameter.  Here the fp32 weight tile is loaded ONCE by TMA and split by eight warps between
// shared memory and tensor memory, which is exactly what the tcgen05 "A operand from TMEM" form exists for.
//
// Swap-AB: the MMA's M dimension (128 TMEM lanes) carries 128 weight rows, its N dimension the activation rows (NT = 32 .. 256),
// so a decode batch of 32 rows is one N = 32 instruction instead of a 75 % empty M = 128 tile.
//
//   warp 0       W producer (one thread): W tile 128 x 32 fp32 (TMA box {32, 128}, SWIZZLE_128B) into the W ring; weights do not
//                depend on the preceding kernel, so these loads start before griddepcontrol.wait (PDL overlap of the ring fill)
//   warp 2       X producer (one thread): X_hi / X_lo tiles NT x 32 (one 3-D TMA box {32, NT, 2}; pre-split by tc_split_a_kernel,
//                which also carries the RMSNorm / LayerNorm / gather prologue) into the X ring
//   warps 4-11   splitter, two groups of four warps taking alternate k-blocks: thread = weight row; 8 x LDS.128 of the swizzled
//                row, cvt.rna.tf32 split, two tcgen05.st 32x32b.x32 -> TMEM columns [slot * 64, +32) = hi, [+32, +64) = lo; the W
//                stage is released as soon as it is in registers
//   warp 1       MMA issuer (one thread): per k-step of 8:  D += A_hi B_hi;  D += A_lo B_hi;  D += A_hi B_lo   with
//                tcgen05.mma.cta_group::1.kind::tf32 [D tmem], [A tmem], B smem-descriptor (K-major, SWIZZLE_128B);
//                tcgen05.commit releases the X stage and the TMEM A slot, and at the end of a tile hands D to the epilogue
//   warps 12-15  epilogue: tcgen05.ld 32x32b.x32 of D (lane = weight row n, column = activation row m) -> C[m][n], coalesced over n
//
// Schedule: tiles are numbered in a rasterised order (groups of GM activation-row tiles, weight tiles inside a group, the GM row
// tiles innermost) so that the tiles in flight at any moment - G consecutive numbers - share GM activation tiles and G / GM weight
// tiles through L2.  The first floor(n_tiles / G) * G tiles are dealt round-robin as whole tiles; the remaining tiles (all of them
// in a decode-sized call) are stream-K: CTA c owns units [c Lr, (c + 1) Lr) of their flattened (tile, k-block) space, so every SM
// streams the same number of weight bytes whatever N / 128 is.  A CTA whose range starts inside a tile writes that partial tile
// to its side slot; the consumer (tc_epilogue_kernel / umma_fixup_kernel) adds the continuation slots in CTA order - deterministic.
//
// TMEM budget (512 columns): A ring 4 slots x 64 columns, accumulator NT columns at column 256.
#include "ua2_kernels.cuh"
#include "ua2_tcgen05.cuh"
#include "ua2_umma.cuh"

namespace ua2 {
namespace {

using namespace tc;

constexpr int UM_BM = 128;        // weight rows per tile (TMEM lanes)
constexpr int UM_BK = 32;         // fp32 per k-block = one 128-byte swizzle row
constexpr int UM_A_SLOTS = 4;     // TMEM A ring
constexpr int UM_ACC_COL = 256;   // accumulator base column
constexpr int UM_THREADS = 512;   // 16 warps, see the role table above

// BF16 = false: fp32 operands, 3xTF32 (k-block = 32 floats, X stage = hi + lo planes).
// BF16 = true : bf16 operands, one tcgen05.mma kind::f16 per k-step of 16 (k-block = 64 bf16 = the same 128-byte rows); the "splitter"
//               warps only move the weight tile from shared to tensor memory.  The flow decoder's option (the reference autocasts it to bf16).
template <int NT, bool BF16>
struct UmCfg {
  static constexpr int SW = NT <= 32 ? 8 : NT <= 64 ? 6 : NT <= 128 ? 4 : (BF16 ? 4 : 2);  // W ring stages (16 KB each)
  static constexpr int SX = NT <= 32 ? 8 : NT <= 64 ? 6 : NT <= 128 ? 4 : (BF16 ? 4 : 3);  // X ring stages
  static constexpr int W_BYTES = UM_BM * 128;
  static constexpr int XH_BYTES = NT * 128;
  static constexpr int X_BYTES = BF16 ? XH_BYTES : 2 * XH_BYTES;
  static constexpr int N_BARS = 2 * SW + 2 * SX + 2 * UM_A_SLOTS + 2;
  static constexpr size_t SMEM = 1024 + (size_t)SW * W_BYTES + (size_t)SX * X_BYTES + N_BARS * 8 + 16;
};

// The work of one CTA as a sequence of segments (tile, first k-block, end k-block): whole tiles of the data-parallel waves, then
// its range of the stream-K remainder.  Every warp role walks the same sequence.
struct UmSeg {
  const UmmaPlan& pl;
  int c, j, u, u_end;
  __device__ UmSeg(const UmmaPlan& p, int cta) : pl(p), c(cta), j(0) {
    u = cta * p.Lr < p.rem_units ? cta * p.Lr : p.rem_units;
    u_end = u + p.Lr < p.rem_units ? u + p.Lr : p.rem_units;
  }
  __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1) {
    if (j < pl.full_waves) {  // the last wave may be partial (dp_tiles is not a multiple of the grid then)
      tile = c + j * pl.grid;
      ++j;
      if (tile < pl.dp_tiles) {
        kb0 = 0;
        kb1 = pl.KB;
        return true;
      }
    }
    if (u >= u_end) return false;
    const int tr = u / pl.KB;
    kb0 = u - tr * pl.KB;
    const int e = (tr + 1) * pl.KB < u_end ? (tr + 1) * pl.KB : u_end;
    kb1 = kb0 + (e - u);
    tile = pl.dp_tiles + tr;
    u = e;
    return true;
  }
};
// rasterised tile number -> (activation-row tile, weight tile)
__device__ __forceinline__ void um_tile_coords(const UmmaPlan& pl, int tile, int& mt, int& nt) {
  const int per_group = pl.GM * pl.n_nt;
  const int g = tile / per_group, r = tile - g * per_group;
  const int gmg = pl.n_mt - g * pl.GM < pl.GM ? pl.n_mt - g * pl.GM : pl.GM;
  nt = r / gmg;
  mt = g * pl.GM + (r - nt * gmg);
}

// ------------------------------------------------------------------ the kernel
template <int NT, bool BF16>
__global__ void __launch_bounds__(UM_THREADS, 1)
umma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX,
                   float* __restrict__ C, int ldc, float* __restrict__ slots, int M, int N, const UmmaPlan pl) {
  using Cfg = UmCfg<NT, BF16>;
  constexpr int BKE = BF16 ? 64 : 32;  // elements per k-block (128 bytes)
  extern __shared__ uint8_t um_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(um_smem_raw) + 1023) & ~(uintptr_t)1023);  // swizzle atoms: 1024 B
  uint8_t* w_ring = base;
  uint8_t* x_ring = w_ring + (size_t)Cfg::SW * Cfg::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_ring + (size_t)Cfg::SX * Cfg::X_BYTES);
  uint64_t* w_full = bars;
  uint64_t* w_empty = w_full + Cfg::SW;
  uint64_t* x_full = w_empty + Cfg::SW;
  uint64_t* x_empty = x_full + Cfg::SX;
  uint64_t* a_full = x_empty + Cfg::SX;
  uint64_t* a_empty = a_full + UM_A_SLOTS;
  uint64_t* acc_full = a_empty + UM_A_SLOTS;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < Cfg::SW; ++i) {
      smem_bar_init(&w_full[i], 1);
      smem_bar_init(&w_empty[i], 4);
    }
    for (int i = 0; i < Cfg::SX; ++i) {
      smem_bar_init(&x_full[i], 1);
      smem_bar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < UM_A_SLOTS; ++i) {
      smem_bar_init(&a_full[i], 4);
      smem_bar_init(&a_empty[i], 1);
    }
    smem_bar_init(acc_full, 1);
    smem_bar_init(acc_empty, 4);
    smem_bar_fence_init();
  }
  if (warp == 1) {  // one warp owns the tensor-memory allocation (all 512 columns: one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr_u32(tmem_base_smem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_smem;
  pdl_launch_dependents();

  if (warp == 0) {
    // ================= W producer (no griddepcontrol.wait: weights are not written by the preceding kernels)
    {
      UmSeg seg(pl, cta);
      int tile, kb0, kb1;
      uint32_t it = 0;
      while (seg.next(tile, kb0, kb1)) {
        int mt, nt;
        um_tile_coords(pl, tile, mt, nt);
        const int mat = nt / pl.nt_per_mat, row0 = (nt - mat * pl.nt_per_mat) * UM_BM;
        const CUtensorMap* tm = mat ? &tmW2 : &tmW;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t sw = it % Cfg::SW, pw = (it / Cfg::SW) & 1;
          smem_bar_wait(&w_empty[sw], pw ^ 1);
          if (elect_one()) {
            smem_bar_arrive_expect_tx(&w_full[sw], Cfg::W_BYTES);
            tma_load_2d(w_ring + (size_t)sw * Cfg::W_BYTES, tm, kb * BKE, row0, &w_full[sw], POLICY_EVICT_FIRST);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 2) {
    // ================= X producer
    {
      pdl_wait();  // the split planes come from the preceding kernel
      UmSeg seg(pl, cta);
      int tile, kb0, kb1;
      uint32_t it = 0;
      while (seg.next(tile, kb0, kb1)) {
        int mt, nt;
        um_tile_coords(pl, tile, mt, nt);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t sx = it % Cfg::SX, px = (it / Cfg::SX) & 1;
          smem_bar_wait(&x_empty[sx], px ^ 1);
          if (elect_one()) {
            smem_bar_arrive_expect_tx(&x_full[sx], Cfg::X_BYTES);
            if constexpr (BF16)
              tma_load_2d(x_ring + (size_t)sx * Cfg::X_BYTES, &tmX, kb * BKE, mt * NT, &x_full[sx], POLICY_EVICT_LAST);
            else
              tma_load_3d(x_ring + (size_t)sx * Cfg::X_BYTES, &tmX, kb * BKE, mt * NT, 0, &x_full[sx], POLICY_EVICT_LAST);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (warp-uniform loop; one elected lane issues)
    {
      // instruction descriptor: D fp32 (bit 4), A/B format at bits 7 / 10 (tf32 = 2, bf16 = 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
      constexpr uint32_t fmt = BF16 ? 1u : 2u;
      constexpr uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(UM_BM >> 4) << 24);
      const uint32_t d_tmem = tmem + UM_ACC_COL;
      UmSeg seg(pl, cta);
      int tile, kb0, kb1;
      uint32_t it = 0, tile_j = 0;
      while (seg.next(tile, kb0, kb1)) {
        smem_bar_wait(acc_empty, (tile_j & 1) ^ 1);  // the epilogue has drained the previous tile
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t sx = it % Cfg::SX, px = (it / Cfg::SX) & 1, sa = it % UM_A_SLOTS, pa = (it / UM_A_SLOTS) & 1;
          smem_bar_wait(&x_full[sx], px);
          smem_bar_wait(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t xs = smem_addr_u32(x_ring + (size_t)sx * Cfg::X_BYTES);
          const uint64_t bh = smem_desc_sw128(xs), bl = smem_desc_sw128(xs + Cfg::XH_BYTES);
          const uint32_t a_hi = tmem + sa * 64, a_lo = a_hi + 32;
          if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // one k-step = 32 bytes of every row (8 tf32 / 16 bf16): +2 in the descriptor's 16-byte address units, +8 TMEM columns
            if constexpr (BF16) {
              mma_bf16_ts(d_tmem, a_hi + ks * 8, bh + (uint64_t)(ks * 2), idesc, (kb == kb0 && ks == 0) ? 0u : 1u);
            } else {
              mma_tf32_ts(d_tmem, a_hi + ks * 8, bh + (uint64_t)(ks * 2), idesc, (kb == kb0 && ks == 0) ? 0u : 1u);
              mma_tf32_ts(d_tmem, a_lo + ks * 8, bh + (uint64_t)(ks * 2), idesc, 1u);
              mma_tf32_ts(d_tmem, a_hi + ks * 8, bl + (uint64_t)(ks * 2), idesc, 1u);
            }
          }
          tc_commit(&x_empty[sx]);
          tc_commit(&a_empty[sa]);
          if (kb + 1 == kb1) tc_commit(acc_full);
          }
          __syncwarp();
        }
        ++tile_j;
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================= splitter: fp32 weight rows -> (hi, lo) tf32 planes in tensor memory; group g takes k-blocks it = g, g + 2, ...
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int r = q * 32 + lane;        // weight row inside the tile
    int n_dp = 0;  // whole tiles of this CTA: c, c + grid, ... below dp_tiles
    if (cta < pl.dp_tiles) n_dp = (pl.dp_tiles - 1 - cta) / pl.grid + 1;
    const uint32_t n_it = (uint32_t)(n_dp * pl.KB) + (uint32_t)UmSeg(pl, cta).u_end - (uint32_t)UmSeg(pl, cta).u;
    for (uint32_t it = (uint32_t)grp; it < n_it; it += 2) {
      const uint32_t sw = it % Cfg::SW, pw = (it / Cfg::SW) & 1, sa = it % UM_A_SLOTS, pa = (it / UM_A_SLOTS) & 1;
      smem_bar_wait(&w_full[sw], pw);
      const uint8_t* row = w_ring + (size_t)sw * Cfg::W_BYTES + r * 128;
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // logical 16-byte chunk j of row r sits at physical chunk j ^ (r & 7)
        const float4 v = *reinterpret_cast<const float4*>(row + ((j ^ (r & 7)) << 4));
        const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if constexpr (BF16) {
            hi[j * 4 + e] = __float_as_uint(f[e]);  // two bf16 per word, moved as they are
          } else {
            const uint32_t h = tf32_rna_bits(f[e]);
            hi[j * 4 + e] = h;
            lo[j * 4 + e] = tf32_rna_bits(f[e] - __uint_as_float(h));
          }
        }
      }
      // The stage may be refilled only after every lane's LDS has RETURNED: an arrive does not wait for outstanding shared-memory
      // loads (first hardware run of the free-running W producer: single weight rows of a tile came from the next k-block).  The
      // shuffle reduction consumes one word of each 128-bit load of every lane, so lane 0's arrive is issued after all of them.
      uint32_t dep = hi[0] ^ hi[4] ^ hi[8] ^ hi[12] ^ hi[16] ^ hi[20] ^ hi[24] ^ hi[28];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dep ^= __shfl_xor_sync(0xffffffffu, dep, o);
      asm volatile("and.b32 %0, %0, 0;" : "+r"(dep));  // opaque zero that still depends on the loads
      if (lane == 0) bar_arrive(&w_empty[sw] + dep);  // the tile is in registers: the W producer may refill the stage
      smem_bar_wait(&a_empty[sa], pa ^ 1);
      tc_fence_after();
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + sa * 64;
      tmem_st32(taddr, hi);
      if constexpr (!BF16) tmem_st32(taddr + 32, lo);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(&a_full[sa]);
    }
  } else if (warp >= 12) {
    // ================= epilogue: D (lane = weight row, column = activation row) -> C or the CTA's side slot
    const int q = warp & 3;
    UmSeg seg(pl, cta);
    int tile, kb0, kb1;
    uint32_t tile_j = 0;
    while (seg.next(tile, kb0, kb1)) {
      int mt, nt;
      um_tile_coords(pl, tile, mt, nt);
      const int mat = nt / pl.nt_per_mat, row0 = (nt - mat * pl.nt_per_mat) * UM_BM;
      const int n_local = q * 32 + lane;
      const bool n_ok = row0 + n_local < N;
      smem_bar_wait(acc_full, tile_j & 1);
      tc_fence_after();
      float* dst;
      int ld;
      int m_valid;
      if (kb0 == 0) {  // this CTA owns the tile's first k-block: its sum goes to C
        dst = C + (size_t)mt * NT * ldc + (size_t)mat * N + row0 + n_local;
        ld = ldc;
        m_valid = M - mt * NT < NT ? M - mt * NT : NT;
      } else {         // continuation of a tile another CTA started: side slot [NT][128]
        dst = slots + (size_t)cta * NT * UM_BM + n_local;
        ld = UM_BM;
        m_valid = NT;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + UM_ACC_COL + c0, v);
        tmem_wait_ld();
        if (n_ok || kb0 != 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < m_valid) dst[(size_t)(c0 + j) * ld] = __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(acc_empty);
      ++tile_j;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// C[m][col] += the side slots of the CTAs that continued the tile (for consumers that read C in place)
__global__ void __launch_bounds__(256) umma_fixup_kernel(float* __restrict__ C, int ldc, const float* __restrict__ slots, int M, int N, int n_mat,
                                                          const UmmaPlan pl) {
  pdl_launch_dependents();
  pdl_wait();
  const int col = blockIdx.y * blockDim.x + threadIdx.x, m = blockIdx.x;  // rows on grid.x: no 65535 limit
  if (col >= N * n_mat || m >= M) return;
  const float add = umma_side_sum(pl, slots, m, col, N);
  if (add != 0.f) C[(size_t)m * ldc + col] += add;
}

int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

template <int NT, bool BF16>
cudaError_t launch_nt(const LaunchCtx& lc, const CUtensorMap& tmW, const CUtensorMap& tmW2, const CUtensorMap& tmX, float* C, int ldc, float* slots,
                      int M, int N, const UmmaPlan& pl) {
  using Cfg = UmCfg<NT, BF16>;
  cudaError_t e = cudaFuncSetAttribute(umma_kernel<NT, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
  if (e != cudaSuccess) return e;
  return launch(lc, umma_kernel<NT, BF16>, dim3(pl.grid), dim3(UM_THREADS), Cfg::SMEM, tmW, tmW2, tmX, C, ldc, slots, M, N, pl);
}

}  // namespace

int umma_pick_nt(int M) {
  if (M <= 32) return 32;
  if (M <= 64) return 64;
  if (M <= 128) return 128;
  // many rows: 256-row tiles unless 128-row tiles waste less padding
  const long long p256 = ((M + 255) / 256) * 256LL, p128 = ((M + 127) / 128) * 128LL;
  return p128 < p256 ? 128 : 256;
}

UmmaPlan umma_plan(int M, int N, int n_mat, int K, bool bf16) {
  UmmaPlan pl;
  pl.NT = umma_pick_nt(M);
  const int bke = bf16 ? 64 : UM_BK;
  pl.KB = (K + bke - 1) / bke;
  pl.nt_per_mat = (N + UM_BM - 1) / UM_BM;
  pl.n_nt = pl.nt_per_mat * n_mat;
  pl.n_mt = (M + pl.NT - 1) / pl.NT;
  pl.GM = std::max(1, std::min(pl.n_mt, 1024 / pl.NT));
  const long long n_tiles = (long long)pl.n_mt * pl.n_nt;
  if (n_tiles * pl.KB >= (1LL << 30)) {
    pl.grid = 0;  // not served (index arithmetic is 32-bit)
    return pl;
  }
  pl.n_tiles = (int)n_tiles;
  const int sms = sm_count();
  pl.grid = sms;
  pl.full_waves = pl.n_tiles / sms;
  pl.dp_tiles = pl.full_waves * sms;
  // a remainder that fills >= 85 % of a wave runs as one more (partial) wave of whole tiles: splitting every one of its tiles
  // between two CTAs would cost two accumulator drains per CTA and a side-slot pass for nothing
  if ((long long)(pl.n_tiles - pl.dp_tiles) * 100 >= 85LL * sms) {
    pl.full_waves += 1;
    pl.dp_tiles = pl.n_tiles;
  }
  pl.rem_units = (pl.n_tiles - pl.dp_tiles) * pl.KB;
  if (pl.full_waves == 0) {  // decode-sized: pure stream-K, at least 4 k-blocks per CTA (a shorter range is all prologue)
    int g = pl.rem_units / 4;
    pl.grid = std::max(1, std::min(sms, g));
  }
  pl.Lr = pl.rem_units ? (pl.rem_units + pl.grid - 1) / pl.grid : 0;
  pl.slot_floats = (size_t)pl.grid * pl.NT * UM_BM;
  return pl;
}

// X2: [2][M][K] split planes (hi, lo) of the activation rows.  C: (M x n_mat * N) row-major, ldc floats per row.
cudaError_t run_umma_tf32x3(const LaunchCtx& lc, const float* X2, const float* W, const float* W2, float* C, int ldc, float* slots, int M, int N,
                            int K, const UmmaPlan& pl) {
  if ((K & 3) || M < 1 || N < 1 || pl.grid < 1 || (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(X2) & 15) ||
      (W2 != nullptr && (reinterpret_cast<uintptr_t>(W2) & 15)))
    return cudaErrorNotSupported;
  CUtensorMap tmW, tmW2, tmX;
  if (!make_tmap(&tmW, W, K, N, 1, UM_BM, true)) return cudaErrorNotSupported;
  if (!make_tmap(&tmW2, W2 ? W2 : W, K, N, 1, UM_BM, true)) return cudaErrorNotSupported;
  if (!make_tmap(&tmX, X2, K, M, 2, pl.NT, false)) return cudaErrorNotSupported;
  switch (pl.NT) {
    case 32: return launch_nt<32, false>(lc, tmW, tmW2, tmX, C, ldc, slots, M, N, pl);
    case 64: return launch_nt<64, false>(lc, tmW, tmW2, tmX, C, ldc, slots, M, N, pl);
    case 128: return launch_nt<128, false>(lc, tmW, tmW2, tmX, C, ldc, slots, M, N, pl);
    case 256: return launch_nt<256, false>(lc, tmW, tmW2, tmX, C, ldc, slots, M, N, pl);
  }
  return cudaErrorNotSupported;
}

// bf16 operands (X (M x K), W (N x K), both row-major bf16), fp32 accumulation and output; plan from umma_plan(..., bf16 = true); K % 8 == 0
cudaError_t run_umma_bf16(const LaunchCtx& lc, const void* X16, const void* W16, float* C, int ldc, float* slots, int M, int N, int K,
                          const UmmaPlan& pl) {
  if ((K & 7) || M < 1 || N < 1 || pl.grid < 1 || (reinterpret_cast<uintptr_t>(W16) & 15) || (reinterpret_cast<uintptr_t>(X16) & 15))
    return cudaErrorNotSupported;
  CUtensorMap tmW, tmX;
  if (!make_tmap(&tmW, W16, K, N, 1, UM_BM, true, true)) return cudaErrorNotSupported;
  if (!make_tmap(&tmX, X16, K, M, 1, pl.NT, false, true)) return cudaErrorNotSupported;
  switch (pl.NT) {
    case 32: return launch_nt<32, true>(lc, tmW, tmW, tmX, C, ldc, slots, M, N, pl);
    case 64: return launch_nt<64, true>(lc, tmW, tmW, tmX, C, ldc, slots, M, N, pl);
    case 128: return launch_nt<128, true>(lc, tmW, tmW, tmX, C, ldc, slots, M, N, pl);
    case 256: return launch_nt<256, true>(lc, tmW, tmW, tmX, C, ldc, slots, M, N, pl);
  }
  return cudaErrorNotSupported;
}

cudaError_t run_umma_fixup(const LaunchCtx& lc, float* C, int ldc, const float* slots, int M, int N, int n_mat, const UmmaPlan& pl) {
  if (!umma_has_split_tiles(pl)) return cudaSuccess;  // every CTA owns whole tiles: no side slots in use
  const dim3 grid(M, (N * n_mat + 255) / 256);
  return launch(lc, umma_fixup_kernel, grid, dim3(256), 0, C, ldc, slots, M, N, n_mat, pl);
}

}  // namespace ua2
