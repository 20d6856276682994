// Micro-benchmark: tcgen05.mma kind::f16 issue/execute rate as a function of the operand swizzle (64 B rows vs 128 B
// rows, K-major) and of N.  One CTA per SM, one issuing thread, operands resident in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../neuralsampleid_b200/csrc -I../../include umma_rate.cu -o umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace grafp;

template <int MODE>
__global__ void __launch_bounds__(64, 1) rate_kernel(int N, int iters, int nstage, int acc_period, unsigned long long* out) {
  constexpr int mode = MODE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int rowb = mode == 0 ? 64 : 128;
  const int slices = rowb / 32;
  const uint32_t a_bytes = 128 * rowb, b_bytes = N * rowb;
  for (int i = threadIdx.x; i < nstage * (int)(a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const long long t0 = clock64();
    int s = 0, ap = 0; uint32_t acc = t;
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = smem_u32(smem + (size_t)s * (a_bytes + b_bytes)), b = a + a_bytes;
      const uint64_t da = mode == 0 ? umma_desc_sw64(a) : umma_desc_sw128(a);
      const uint64_t db = mode == 0 ? umma_desc_sw64(b) : umma_desc_sw128(b);
#pragma unroll
      for (int k = 0; k < (mode == 0 ? 2 : 4); ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, 1u);
      if (++s == nstage) s = 0;
      if (++ap == acc_period) { ap = 0; acc = acc == t ? t + (uint32_t)N : t; }
    }
    umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(t, 512); }
}

// Same work, issued the CUTLASS way: the whole (converged) warp runs the loop, one elected lane issues.
template <int MODE>
__global__ void __launch_bounds__(64, 1) rate_kernel_elect(int N, int iters, int nstage, int acc_period, unsigned long long* out) {
  constexpr int mode = MODE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int rowb = mode == 0 ? 64 : 128;
  const uint32_t a_bytes = 128 * rowb, b_bytes = N * rowb;
  for (int i = threadIdx.x; i < nstage * (int)(a_bytes + b_bytes) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) { mbar_init(&done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const long long t0 = clock64();
    const uint32_t base = smem_u32(smem);
    const uint32_t hi = mode == 0 ? ((512u >> 4) | (1u << 14) | (4u << 29)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
    int s = 0, ap = 0; uint32_t acc = t;
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = base + (uint32_t)s * (a_bytes + b_bytes), b = a + a_bytes;
      const uint32_t alo = ((a >> 4) & 0x3FFFu) | (1u << 16), blo = ((b >> 4) & 0x3FFFu) | (1u << 16);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < (mode == 0 ? 2 : 4); ++k) {
          asm volatile(
              "{\n\t.reg .b64 da, db;\n\t"
              "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, 1;\n\t}"
              ::"r"(acc), "r"(alo + 2 * k), "r"(blo + 2 * k), "r"(hi), "r"(idesc) : "memory");
        }
      }
      __syncwarp();
      if (++s == nstage) s = 0;
      if (++ap == acc_period) { ap = 0; acc = acc == t ? t + (uint32_t)N : t; }
    }
    if (elect_one()) umma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(t, 512); }
}

int main() {
  unsigned long long* d; cudaMalloc(&d, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel_elect<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel_elect<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int style = 0; style < 2; ++style)
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {64, 128, 256})
      for (int period : {1 << 30, 1}) {
        const int grid = sms;
        const int iters = 4000, nstage = 3;
        unsigned long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          if (style == 0) {
            if (mode == 0) rate_kernel<0><<<grid, 64, 200 * 1024>>>(N, iters, nstage, period, d);
            else rate_kernel<1><<<grid, 64, 200 * 1024>>>(N, iters, nstage, period, d);
          } else {
            if (mode == 0) rate_kernel_elect<0><<<grid, 64, 200 * 1024>>>(N, iters, nstage, period, d);
            else rate_kernel_elect<1><<<grid, 64, 200 * 1024>>>(N, iters, nstage, period, d);
          }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        const int slices = mode == 0 ? 2 : 4;
        const double per = (double)h / ((double)iters * slices);
        printf("%s %s N=%3d acc switch every %d k-blocks: %.1f cycles per 128xNx16 MMA (%.0f MAC/clk)\n", style ? "elect " : "thread0", mode == 0 ? "SW64 " : "SW128", N,
               period > 1000 ? 0 : period, per, 128.0 * N * 16 / per);
      }
  return 0;
}
