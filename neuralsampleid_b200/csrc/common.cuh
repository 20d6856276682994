// Shared helpers for libgrafp_sm100a (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/grafp.h"

namespace grafp {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int fail(const char* fmt, ...);
int check_launch(const char* what);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (GRAFP_PDL=1) -------------------------------------------------------------------
// A kernel launched with the programmatic-stream-serialization attribute may start while its predecessor in the stream
// is still running.  Every such kernel here (a) triggers its own dependents at its first instruction and (b) executes
// griddepcontrol.wait -- which returns once the predecessor grid has completed and its memory is visible -- after the
// prologue that touches no global data (barrier init, TMEM allocation, descriptor prefetch, staging of constant
// per-channel parameters) and before ANY access to tensors another kernel writes.  What overlaps is the launch latency
// and that prologue; the data path is ordered exactly as without the attribute.  Without the attribute both
// instructions are no-ops.
int pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                             Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 0) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#define GRAFP_REQUIRE(cond, ...) \
  do { if (!(cond)) return ::grafp::fail(__VA_ARGS__); } while (0)

__device__ __forceinline__ float apply_act(float v, int act, float p) {
  switch (act) {
    case GRAFP_ACT_RELU:  return v < 0.0f ? 0.0f : v;      // NaN propagates (torch.relu), unlike fmaxf
    case GRAFP_ACT_LEAKY: return v > 0.0f ? v : v * p;
    case GRAFP_ACT_GELU:  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case GRAFP_ACT_ELU:   return v > 0.0f ? v : expm1f(v);
    case GRAFP_ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    default:              return v;
  }
}

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

// ---- mbarrier / bulk-copy (TMA) primitives ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware until the
      : "memory");                                         // phase completes instead of re-polling (the polls
                                                           // were 40 % of all issued instructions in the kNN)
  return ok != 0;
}
// Bounded wait: a protocol bug traps (visible as a launch failure) instead of hanging the box.  The bound is
// wall-clock (4 s on %globaltimer, polled only on the slow path): with the suspend-time hint one try_wait may
// sleep for milliseconds, so a spin count no longer bounds anything.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // slow path: a lean poll loop (the polls of idle warps compete with the working warps for issue slots: the
  // profile of the first large-graph kNN showed 19 instructions per poll, a quarter of everything issued), the
  // watchdog clock is read once per 4096 polls
  const uint64_t t0 = global_timer_ns();
  while (true) {
#pragma unroll 1
    for (int i = 0; i < 4096; ++i)
      if (mbar_try_wait(bar, parity)) return;
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("grafp: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// Latency-critical wait: plain polling without the suspend-time hint (a hardware-suspended warp wakes up noticeably
// later than a polling one; used where a hand-off sits on the critical path of a short dependent chain).
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint64_t t0 = global_timer_ns();
  while (true) {
#pragma unroll 1
    for (int i = 0; i < 65536; ++i) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
      if (ok) return;
    }
    if (global_timer_ns() - t0 > 4000000000ull) {
      printf("grafp: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// 1-D bulk async copy global -> shared, completion on an mbarrier (TMA engine, UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

int sm_count();

}  // namespace grafp
