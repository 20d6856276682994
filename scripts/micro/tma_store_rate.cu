// Micro-benchmark: TMA tile-STORE throughput per SM (shared -> global, 128 rows x 128 B boxes) with 1, 2 or 4 staging tiles in
// flight per issuing thread and 1 or 2 issuing threads, every SM writing its own region of a 2 GB matrix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../neuralsampleid_b200/csrc -I../../include tma_store_rate.cu -o tma_store_rate -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace grafp;

template <int DEPTH>
__global__ void __launch_bounds__(128, 1) store_kernel(const __grid_constant__ CUtensorMap tm, int iters, int nthr, int rows_per_cta,
                                                       unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 8 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
  fence_proxy_async_smem();
  __syncthreads();
  const int me = (threadIdx.x & 31) == 0 ? (int)(threadIdx.x >> 5) : 999;
  if (me < nthr) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % DEPTH;
      bulk_wait_group_read<DEPTH - 1>();
      const int row = (int)blockIdx.x * rows_per_cta + ((it * nthr + me) * 128) % rows_per_cta;
      tma_store_2d(&tm, smem + (size_t)(me * DEPTH + s) * 16384, ((it >> 3) & 7) * 32, row);
      bulk_commit_group();
    }
    bulk_wait_group_all();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && me == 0) out[0] = (unsigned long long)(t1 - t0);
  }
}

int main() {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int rows_per_cta = 16384;
  const int64_t rows = (int64_t)sms * rows_per_cta, cols = 256;        // fp32: 1 KB rows, 2.4 GB
  void* g; cudaMalloc(&g, rows * cols * 4);
  unsigned long long* d; cudaMalloc(&d, 8);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(store_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  cudaFuncSetAttribute(store_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  cudaFuncSetAttribute(store_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  for (int nthr : {1, 2})
    for (int depth : {1, 2, 4}) {
      const int iters = 4000;
      unsigned long long h = 0;
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (depth == 1) store_kernel<1><<<sms, 128, 140 * 1024>>>(tm, iters, nthr, rows_per_cta, d);
        else if (depth == 2) store_kernel<2><<<sms, 128, 140 * 1024>>>(tm, iters, nthr, rows_per_cta, d);
        else store_kernel<4><<<sms, 128, 140 * 1024>>>(tm, iters, nthr, rows_per_cta, d);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      }
      const double bytes = 16384.0 * iters * nthr;
      printf("%d issuing thread(s), %d tile(s) in flight each: %.0f cycles per 16 KB store, %.1f B/clk/SM, chip %.0f GB/s\n", nthr, depth,
             (double)h / iters, bytes / h, bytes * sms / (ms * 1e-3) / 1e9);
    }
  return 0;
}
