// Shared device/host helpers for the sm_100a kernels of the UniAudio2 hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace ua2 {

// ---- error plumbing (thread-local message surfaced through ua2_last_error) ----
void set_error(const std::string& msg);
#define UA2_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::ua2::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                       std::to_string(__LINE__));                                                 \
      return UA2_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)
#define UA2_REQUIRE(cond, msg)                    \
  do {                                            \
    if (!(cond)) {                                \
      ::ua2::set_error(std::string("invalid argument: ") + (msg)); \
      return UA2_ERR_INVALID;                     \
    }                                             \
  } while (0)

// ---- device helpers ----
// Streaming 128-bit load for weights: read-only path, do not allocate in L1 (weights are touched once per
// launch; activations staged in shared memory keep L1/smem for themselves).
#ifndef UA2_CPU_SHIM
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
#else  // tests/cpu_shim: the kernels of this file set compile with g++ and run one OS thread per CUDA thread
__device__ __forceinline__ float4 ldg_stream(const float* p) { return *reinterpret_cast<const float4*>(p); }
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Programmatic dependent launch (PDL): wait for the producer grid's memory to be visible / let the
// dependent grid start its prologue early.  No-ops when the launch carries no PDL attribute.
#ifndef UA2_CPU_SHIM
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
#endif

// mbarrier + bulk-copy (cp.async.bulk global -> shared) wrappers: one spelling of the PTX for the kernels that stream through
// shared-memory rings.  `bar` lives in shared memory.  smem_bar_wait(bar, parity) returns once the phase of that parity has completed.
#ifndef UA2_CPU_SHIM
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void smem_bar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(arrivals));
}
__device__ __forceinline__ void smem_bar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void smem_bar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {  // bytes % 16 == 0
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_addr_u32(bar))
               : "memory");
}
__device__ __forceinline__ void smem_bar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_addr_u32(bar)),
      "r"(parity)
      : "memory");
}
#else  // tests/cpu_shim: the copy completes inside the issuing call; the barrier word is {phase:1 | arrivals:15 | count:16 | tx:32}
__device__ __forceinline__ void smem_bar_init(uint64_t* bar, uint32_t arrivals) { shim::mbar_init(bar, arrivals); }
__device__ __forceinline__ void smem_bar_fence_init() {}
__device__ __forceinline__ void smem_bar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { shim::mbar_update(bar, 1, (int64_t)bytes); }
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  shim::bulk_copy(dst_smem, src, bytes);
  shim::mbar_update(bar, 0, -(int64_t)bytes);
}
__device__ __forceinline__ void smem_bar_wait(uint64_t* bar, uint32_t parity) { shim::mbar_wait(bar, parity); }
#endif

// Function attributes (dynamic shared-memory opt-in, carveout preference) are PER DEVICE: a once-flag must be too, or a process that
// runs handles on a second GPU never issues the opt-in there and its first large-shared-memory launch fails with invalid-argument.
struct DeviceOnce {
  bool done[64] = {};
  bool need() {  // true the first time it is asked on the current device
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

// Per-launch context shared by the host-side launchers.
struct GemvSeq;
struct LaunchCtx {
  cudaStream_t stream = nullptr;
  bool pdl = false;       // attach the programmatic-stream-serialization attribute
  int* launch_counter = nullptr;
  struct GemvSeq* seq = nullptr;  // frame-level launch sequence (tail L2 prefetch of the next linears' weights)
};

// Every kernel of the path asks for the maximum shared-memory carveout, so consecutive (and, under PDL,
// co-resident) kernels never force an L1/shared reconfiguration of the SM between launches.
#ifdef UA2_CPU_SHIM
template <typename K>
inline cudaError_t prefer_max_smem(K) {
  return cudaSuccess;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch(const LaunchCtx& lc, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
  if (lc.launch_counter) ++*lc.launch_counter;
  return shim::run_grid_smem(kernel, grid, block, smem, args...);
}
#else
template <typename K>
inline cudaError_t prefer_max_smem(K kern) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch(const LaunchCtx& lc, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = lc.stream;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (lc.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (lc.launch_counter) ++*lc.launch_counter;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace ua2
