// Micro-benchmark: TMA tile-load throughput per SM as a function of the box row width (64 B vs 128 B rows, same bytes per
// box), source resident in L2.  One CTA per SM, one issuing thread, a ring of `depth` boxes in flight.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../neuralsampleid_b200/csrc -I../../include tma_rate.cu -o tma_rate -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace grafp;

__global__ void __launch_bounds__(256, 1) tma_kernel(int nthr, int mode, const __grid_constant__ CUtensorMap tm, int box_bytes, int iters, int depth,
                                                    int ncols_boxes, int nrow_boxes, int box_cols, int box_rows, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_all[8][16];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) { for (int t = 0; t < 8; ++t) for (int i = 0; i < depth; ++i) mbar_init(&full_all[t][i], 1); fence_barrier_init(); }
  __syncthreads();
  const int me = mode == 0 ? (int)threadIdx.x : ((threadIdx.x & 31) == 0 ? (int)(threadIdx.x >> 5) : 999);
  if (me < nthr) {
    uint64_t* full = full_all[me];
    smem += (size_t)me * depth * box_bytes;
    const long long t0 = clock64();
    uint32_t seed = blockIdx.x * 7919u + me * 104729u;
    for (int it = 0; it < iters + depth; ++it) {
      const int s = it % depth;
      if (it >= depth) { while (!mbar_try_wait(&full[s], ((it / depth) - 1) & 1u)) { } }
      if (it < iters) {
        seed = seed * 1664525u + 1013904223u;
        const int cb = (seed >> 8) % ncols_boxes, rb = (seed >> 16) % nrow_boxes;
        mbar_arrive_expect_tx(&full[s], box_bytes);
        tma_load_2d(smem + (size_t)s * box_bytes, &tm, cb * box_cols, rb * box_rows, &full[s]);
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && me == 0) out[0] = (unsigned long long)(t1 - t0);
  }
}

__global__ void timer_cost_kernel(unsigned long long* out) {
  unsigned long long acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < 1000; ++i) acc += global_timer_ns();
  const long long t1 = clock64();
  out[0] = (unsigned long long)(t1 - t0); out[1] = acc;
}

int main() {
  {
    unsigned long long* d2; cudaMalloc(&d2, 16); unsigned long long h2[2];
    timer_cost_kernel<<<1, 1>>>(d2); cudaDeviceSynchronize(); cudaMemcpy(h2, d2, 16, cudaMemcpyDeviceToHost);
    printf("%%globaltimer read: %.1f cycles each\n", h2[0] / 1000.0);
  }
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const int64_t rows = 4096, cols = 4096;              // 16-bit elements: 32 MB, L2-resident
  void* g; cudaMalloc(&g, rows * cols * 2); cudaMemset(g, 0, rows * cols * 2);
  unsigned long long* d; cudaMalloc(&d, 8);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int brows : {128})
  for (int rowb : {64})
  for (int mode : {0, 1})
  for (int nthr : {1, 2, 4}) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)(rowb / 2), (cuuint32_t)brows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = rowb * brows;
    for (int depth : {4}) {
      if (depth * box_bytes > 190 * 1024) continue;
      const int iters = 4000;
      unsigned long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        tma_kernel<<<sms, 256, 200 * 1024>>>(nthr, mode, tm, box_bytes, iters, depth, (int)(cols / (rowb / 2)), (int)(rows / brows), rowb / 2, brows, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      }
      printf("%d issuing %s, box %3d rows of %3d B: %.1f cycles per box per issuer, %.1f B/clk/SM\n", nthr, mode ? "warps" : "lanes of one warp", brows, rowb,
             (double)h / iters, (double)box_bytes * iters * nthr / h);
    }
  }
  return 0;
}
