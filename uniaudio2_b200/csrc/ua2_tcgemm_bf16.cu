// CUTLASS 4.x sm_100a collective (TMA producer warp, single-thread tcgen05.mma kind::f16 issue on bf16 operands, TMEM fp32
// accumulators, tcgen05.ld + TMA-store epilogue) for C (M x N, row-major, fp32) = A (M x K, row-major, bf16) * B (N x K,
// row-major, bf16)^T.  The optional reduced-precision mainloop of the flow-matching decoder (ua2_dit.cu, option "bf16"): the
// reference runs those linears under torch.autocast(bfloat16) (reason_tokenizer.py:265), so bf16 operands with fp32
// accumulation are its own arithmetic; the default path stays 3xTF32 so that it can be checked against the fp32 oracle at 1e-4.
#ifdef UA2_HAVE_CUTLASS
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

namespace ua2 {
namespace {
using namespace cute;
using ElementAB = cutlass::bfloat16_t;
using LayoutA = cutlass::layout::RowMajor;
using LayoutB = cutlass::layout::ColumnMajor;  // (N, K) row-major = (K, N) column-major
using LayoutC = cutlass::layout::RowMajor;
constexpr int kAlignAB = 8;  // 16 bytes
constexpr int kAlignC = 4;
using MmaTile = Shape<_128, _128, _64>;
using Cluster = Shape<_1, _1, _1>;
using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTile, Cluster, cutlass::epilogue::collective::EpilogueTileAuto, float,
    float, float, LayoutC, kAlignC, float, LayoutC, kAlignC, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementAB, LayoutA, kAlignAB, ElementAB, LayoutB, kAlignAB, float, MmaTile,
    Cluster, cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
}  // namespace

cudaError_t run_bf16_gemm_128x128(cudaStream_t st, const __nv_bfloat16* A, const __nv_bfloat16* B, float* C, int M, int N, int K) {
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  if ((K % kAlignAB) != 0 || (N % kAlignC) != 0) return cudaErrorNotSupported;
  const StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K, 1));
  const StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K, 1));
  const StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, 1));
  const StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, 1));
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                {M, N, K, 1},
                                {reinterpret_cast<const ElementAB*>(A), sa, reinterpret_cast<const ElementAB*>(B), sb},
                                {{1.f, 0.f}, C, sc, C, sd}};
  static int sms = 0, dev = -1;
  if (dev < 0) {
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  args.hw_info.device_id = dev;
  args.hw_info.sm_count = sms;
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return cudaErrorNotSupported;
  if (Gemm::get_workspace_size(args) != 0) return cudaErrorNotSupported;
  if (gemm.initialize(args, nullptr, st) != cutlass::Status::kSuccess) return cudaErrorInvalidValue;
  if (gemm.run(st) != cutlass::Status::kSuccess) return cudaErrorLaunchFailure;
  return cudaSuccess;
}
}  // namespace ua2
#endif
