// fp32 FFMA GEMM with the fused conv epilogue (scale/shift = folded bias+BN, activation,
// residual).  This is the exact-fp32 engine: used for shapes the tcgen05 engine does not take
// (k not a multiple of 32, tiny n) and as the on-device cross-check of the tensor-core engines.
//
// y[m, g*n + j] = act(scale * sum_k A_g[m, k] * W[g*n + j, k] + shift) + residual
// A_g[m, :] = [ a1[m, g*k1 : (g+1)*k1] | a2[m, g*k2 : (g+1)*k2] ]             (see grafp.h)
#include "common.cuh"

namespace grafp {

struct GemmP {
  const float* a1; int64_t lda1; int k1;
  const float* a2; int64_t lda2; int k2;
  const float* w; int64_t ldw;
  const float* scale; const float* shift;
  const float* residual; int64_t ldr;
  float* y; int64_t ldy;
  float* row_sumsq;
  int64_t m; int n; int act; float act_param;
  int tap3_nodes;
};

// BM x BN tile, BK = 16, 256 threads, TM x TN micro-tile per thread.
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmP p) {
  constexpr int BK = 16;
  constexpr int NT = 256;
  static_assert((BM / TM) * (BN / TN) == NT, "thread tiling");
  constexpr int A_LD = (BM * BK / 4) / NT;     // float4 loads of A per thread per k-tile
  constexpr int W_LD = (BN * BK / 4 + NT - 1) / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];

  const int tid = threadIdx.x;
  const int g = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.k1 + p.k2;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  float4 ra[A_LD], rw[W_LD];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      const int q = tid + l * NT;              // float4 id in the BM x (BK/4) tile
      const int r = q / (BK / 4), kq = (q % (BK / 4)) * 4;
      const int64_t m = m0 + r;
      const int k = k0 + kq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < p.m && k < K) {
        if (p.tap3_nodes > 0) {
          // Downsample: output row (b, j) <- input nodes 2j-1, 2j, 2j+1; k1 = 3*Cin
          const int cin = p.k1 / 3;
          const int64_t j = m % p.tap3_nodes;
          if (!(j == 0 && k < cin)) {
            const float* src = p.a1 + (2 * m - 1) * (int64_t)cin + k;
            v = *reinterpret_cast<const float4*>(src);
          }
        } else if (k < p.k1) {
          v = *reinterpret_cast<const float4*>(p.a1 + m * p.lda1 + (int64_t)g * p.k1 + k);
        } else {
          v = *reinterpret_cast<const float4*>(p.a2 + m * p.lda2 + (int64_t)g * p.k2 + (k - p.k1));
        }
      }
      ra[l] = v;
    }
#pragma unroll
    for (int l = 0; l < W_LD; ++l) {
      const int q = tid + l * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < BN * BK / 4) {
        const int r = q / (BK / 4), kq = (q % (BK / 4)) * 4;
        const int nn = n0 + r, k = k0 + kq;
        if (nn < p.n && k < K)
          v = *reinterpret_cast<const float4*>(p.w + ((int64_t)g * p.n + nn) * p.ldw + k);
      }
      rw[l] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      const int q = tid + l * NT;
      const int r = q / (BK / 4), kq = (q % (BK / 4)) * 4;
      As[kq + 0][r] = ra[l].x; As[kq + 1][r] = ra[l].y;
      As[kq + 2][r] = ra[l].z; As[kq + 3][r] = ra[l].w;
    }
#pragma unroll
    for (int l = 0; l < W_LD; ++l) {
      const int q = tid + l * NT;
      if (q < BN * BK / 4) {
        const int r = q / (BK / 4), kq = (q % (BK / 4)) * 4;
        Ws[kq + 0][r] = rw[l].x; Ws[kq + 1][r] = rw[l].y;
        Ws[kq + 2][r] = rw[l].z; Ws[kq + 3][r] = rw[l].w;
      }
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < K; k0 += BK) {
    store_tiles();
    __syncthreads();
    if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Ws[kk][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= p.m) continue;
    float rowsq = 0.0f;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      const int nn = n0 + tx * TN + j;
      if (nn >= p.n) continue;
      const int64_t col = (int64_t)g * p.n + nn;
      float v[4] = {acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]};
      const bool full = (nn + 3 < p.n);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (nn + q < p.n) {
          const float sc = p.scale ? p.scale[col + q] : 1.0f;
          const float sh = p.shift ? p.shift[col + q] : 0.0f;
          float t = fmaf(v[q], sc, sh);
          t = apply_act(t, p.act, p.act_param);
          if (p.residual) t += p.residual[m * p.ldr + col + q];
          v[q] = t;
          rowsq = fmaf(t, t, rowsq);
        }
      }
      float* dst = p.y + m * p.ldy + col;
      if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (nn + q < p.n) dst[q] = v[q];
      }
    }
    if (p.row_sumsq) atomicAdd(p.row_sumsq + m, rowsq);
  }
}

int gemm_simt_launch(const grafp_gemm_args& a, cudaStream_t st) {
  GemmP p;
  p.a1 = a.a1; p.lda1 = a.lda1; p.k1 = a.k1;
  p.a2 = a.a2; p.lda2 = a.lda2; p.k2 = a.k2;
  p.w = a.w; p.ldw = a.ldw; p.scale = a.scale; p.shift = a.shift;
  p.residual = a.residual; p.ldr = a.ldr; p.y = a.y; p.ldy = a.ldy; p.row_sumsq = a.row_sumsq;
  p.m = a.m; p.n = a.n; p.act = a.act; p.act_param = a.act_param; p.tap3_nodes = a.tap3_nodes;
  const int64_t mt = (a.m + 127) / 128;
  GRAFP_REQUIRE(mt <= 0x7fffffff, "gemm: m too large");
  if (a.n <= 32) {
    dim3 grid((unsigned)mt, (a.n + 31) / 32, a.groups);
    gemm_simt_kernel<128, 32, 4, 4><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid((unsigned)mt, (a.n + 63) / 64, a.groups);
    gemm_simt_kernel<128, 64, 8, 4><<<grid, 256, 0, st>>>(p);
  }
  return check_launch("gemm_simt");
}

}  // namespace grafp
