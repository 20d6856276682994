// NT-Xent: fused similarity tile GEMM + masked online log-sum-exp + positive pick (forward) and
// the matching fused backward.  The (n x n) similarity matrix never exists in memory: a CTA
// owns 32 rows and streams 32-column tiles of z through shared memory.
// Reference: simclr/ntxent.py:5-30 (a = z z^T / tau; row i: log-softmax over the 2B-1
// off-diagonal entries, pick partner i^1; loss = -mean).
#include "common.cuh"

namespace grafp {

constexpr int NTX_T = 32;        // tile edge
constexpr int NTX_MAXD = 256;

__global__ void __launch_bounds__(256)
ntxent_fwd_kernel(const float* __restrict__ z, int n, int D, float tau, int row0, int rows,
                  float* __restrict__ lse_out, float* __restrict__ loss_out) {
  extern __shared__ float sm[];
  const int ld = D + 1;
  float* zi = sm;                 // 32 x ld
  float* zj = zi + NTX_T * ld;    // 32 x ld
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;   // ty: 0..7, rows ty*4..ty*4+3
  const int r0 = row0 + blockIdx.x * NTX_T;
  const int rend = row0 + rows;

  for (int q = tid; q < NTX_T * D; q += 256) {
    const int r = q / D, c = q - r * D;
    zi[r * ld + c] = (r0 + r < rend) ? z[(size_t)(r0 + r) * D + c] : 0.0f;
  }
  float mx[4], sum[4], pos[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { mx[i] = -INFINITY; sum[i] = 0.0f; pos[i] = 0.0f; }

  for (int j0 = 0; j0 < n; j0 += NTX_T) {
    __syncthreads();
    for (int q = tid; q < NTX_T * D; q += 256) {
      const int r = q / D, c = q - r * D;
      zj[r * ld + c] = (j0 + r < n) ? z[(size_t)(j0 + r) * D + c] : 0.0f;
    }
    __syncthreads();
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < D; ++c) {
      const float b = zj[tx * ld + c];
#pragma unroll
      for (int i = 0; i < 4; ++i) dot[i] = fmaf(zi[(ty * 4 + i) * ld + c], b, dot[i]);
    }
    const int j = j0 + tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = r0 + ty * 4 + i;
      const float a = __fdiv_rn(dot[i], tau);
      const bool valid = (j < n) && (j != gi);
      if (j == (gi ^ 1)) pos[i] = a;
      const float av = valid ? a : -INFINITY;
      const float tmax = warp_max(av);
      const float nm = fmaxf(mx[i], tmax);
      float e = valid ? expf(a - nm) : 0.0f;
      e = warp_sum(e);
      if (nm != -INFINITY) sum[i] = sum[i] * expf(mx[i] - nm) + e;
      mx[i] = nm;
    }
  }
  float part = 0.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = r0 + ty * 4 + i;
    const float p = warp_sum(pos[i]);        // exactly one lane (or none) holds the positive
    if (gi < rend) {
      const float lse = mx[i] + logf(sum[i]);
      if (tx == 0) {
        lse_out[gi - row0] = lse;
        part += -(p - lse) / (float)n;
      }
    }
  }
  if (tx == 0 && part != 0.0f) atomicAdd(loss_out, part);
}

__global__ void __launch_bounds__(256)
ntxent_bwd_kernel(const float* __restrict__ z, const float* __restrict__ lse_all, int n, int D,
                  float tau, int row0, int rows, const float* __restrict__ grad_loss,
                  float* __restrict__ dz) {
  extern __shared__ float sm[];
  const int ld = D + 1;
  float* zi = sm;                     // 32 x ld
  float* zj = zi + NTX_T * ld;        // 32 x ld
  float* coef = zj + NTX_T * ld;      // 32 x 33
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int r0 = row0 + blockIdx.x * NTX_T;
  const int rend = row0 + rows;
  for (int q = tid; q < NTX_T * D; q += 256) {
    const int r = q / D, c = q - r * D;
    zi[r * ld + c] = (r0 + r < rend) ? z[(size_t)(r0 + r) * D + c] : 0.0f;
  }
  // accumulation mapping: thread -> row ar = tid / 8, channels (tid % 8) + 8 * t
  const int ar = tid >> 3, ac = tid & 7;
  float acc[NTX_MAXD / 8];
#pragma unroll
  for (int t = 0; t < NTX_MAXD / 8; ++t) acc[t] = 0.0f;
  float lse_i[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = r0 + ty * 4 + i;
    lse_i[i] = gi < rend ? lse_all[gi] : 0.0f;
  }
  for (int j0 = 0; j0 < n; j0 += NTX_T) {
    __syncthreads();
    for (int q = tid; q < NTX_T * D; q += 256) {
      const int r = q / D, c = q - r * D;
      zj[r * ld + c] = (j0 + r < n) ? z[(size_t)(j0 + r) * D + c] : 0.0f;
    }
    __syncthreads();
    float dot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < D; ++c) {
      const float b = zj[tx * ld + c];
#pragma unroll
      for (int i = 0; i < 4; ++i) dot[i] = fmaf(zi[(ty * 4 + i) * ld + c], b, dot[i]);
    }
    const int j = j0 + tx;
    const float lse_j = j < n ? lse_all[j] : 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = r0 + ty * 4 + i;
      float cf = 0.0f;
      if (j < n && j != gi && gi < rend) {
        const float a = __fdiv_rn(dot[i], tau);
        cf = expf(a - lse_i[i]) + expf(a - lse_j) - (j == (gi ^ 1) ? 2.0f : 0.0f);
      }
      coef[(ty * 4 + i) * 33 + tx] = cf;
    }
    __syncthreads();
    for (int jj = 0; jj < NTX_T; ++jj) {
      const float cf = coef[ar * 33 + jj];
#pragma unroll
      for (int t = 0; t < NTX_MAXD / 8; ++t) {
        const int c = ac + 8 * t;
        if (c < D) acc[t] = fmaf(cf, zj[jj * ld + c], acc[t]);
      }
    }
  }
  const int gi = r0 + ar;
  if (gi < rend) {
    const float s = grad_loss[0] / ((float)n * tau);
#pragma unroll
    for (int t = 0; t < NTX_MAXD / 8; ++t) {
      const int c = ac + 8 * t;
      if (c < D) dz[(size_t)(gi - row0) * D + c] = acc[t] * s;
    }
  }
}

}  // namespace grafp

using namespace grafp;

extern "C" {

int grafp_ntxent_fwd(const float* z, int n, int D, float tau, int row0, int rows, float* lse_out,
                     float* loss_out, void* stream) {
  GRAFP_REQUIRE(z && lse_out && loss_out, "ntxent_fwd: null pointer");
  GRAFP_REQUIRE(n >= 2 && n % 2 == 0, "ntxent_fwd: n=%d must be even and >= 2", n);
  GRAFP_REQUIRE(D > 0 && D <= NTX_MAXD, "ntxent_fwd: D=%d out of range (<= %d)", D, NTX_MAXD);
  GRAFP_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= n && row0 % 2 == 0 && rows % 2 == 0,
                "ntxent_fwd: bad row range");
  GRAFP_REQUIRE(tau > 0.0f, "ntxent_fwd: tau must be positive");
  if (rows == 0) return 0;
  const size_t smem = sizeof(float) * 2 * NTX_T * (D + 1);
  cudaFuncSetAttribute(ntxent_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ntxent_fwd_kernel<<<(rows + NTX_T - 1) / NTX_T, 256, smem, as_stream(stream)>>>(
      z, n, D, tau, row0, rows, lse_out, loss_out);
  return check_launch("ntxent_fwd");
}

int grafp_ntxent_bwd(const float* z, const float* lse_all, int n, int D, float tau, int row0,
                     int rows, const float* grad_loss, float* dz, void* stream) {
  GRAFP_REQUIRE(z && lse_all && grad_loss && dz, "ntxent_bwd: null pointer");
  GRAFP_REQUIRE(n >= 2 && n % 2 == 0, "ntxent_bwd: n=%d must be even and >= 2", n);
  GRAFP_REQUIRE(D > 0 && D <= NTX_MAXD, "ntxent_bwd: D=%d out of range (<= %d)", D, NTX_MAXD);
  GRAFP_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= n, "ntxent_bwd: bad row range");
  if (rows == 0) return 0;
  const size_t smem = sizeof(float) * (2 * NTX_T * (D + 1) + NTX_T * 33);
  cudaFuncSetAttribute(ntxent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ntxent_bwd_kernel<<<(rows + NTX_T - 1) / NTX_T, 256, smem, as_stream(stream)>>>(
      z, lse_all, n, D, tau, row0, rows, grad_loss, dz);
  return check_launch("ntxent_bwd");
}

}  // extern "C"
