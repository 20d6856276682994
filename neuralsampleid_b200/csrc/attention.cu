// Cross-attention pooling for the re-ranker (reference downstream.py:30-79, CrossAttentionClassifier):
//   out[p, h*Dh + d] = mean_i  sum_j softmax_j( scale * <Q[p,i,h,:], K[p,j,h,:]> ) * V[p,j,h,d]
// i.e. nn.MultiheadAttention's per-head attention followed by the classifier's mean over the query
// nodes -- taken BEFORE the output projection (a linear map commutes with the mean), so the out-proj
// GEMM runs on P rows instead of P*N.  Since mean_i (P_i V) = (mean_i P_i) V, the (N x Dh) per-query
// outputs are never formed either.  One CTA per (pair, head); everything lives in shared memory; exact
// fp32.  The reference calls the classifier once per candidate from a Python loop (eval_hr.py:125-135).
#include "common.cuh"

namespace grafp {

constexpr int MHA_THREADS = 256;

__global__ void __launch_bounds__(MHA_THREADS)
mha_pool_kernel(const float* __restrict__ q, int64_t ldq, const float* __restrict__ k, int64_t ldk,
                const float* __restrict__ v, int64_t ldv, int Nq, int Nk, int H, int Dh, float scale,
                float* __restrict__ out, int64_t ldo) {
  extern __shared__ __align__(16) float sm[];
  const int DP = Dh + 4;                       // padded row: conflict-free 128-bit reads across rows
  float* Qs = sm;                              // [Nq][DP]
  float* Ks = Qs + (size_t)Nq * DP;            // [Nk][DP]
  float* Vs = Ks + (size_t)Nk * DP;            // [Nk][DP]
  float* S = Vs + (size_t)Nk * DP;             // [Nq][Nk + 1]
  float* pbar = S + (size_t)Nq * (Nk + 1);     // [Nk]
  const int p = blockIdx.x / H, h = blockIdx.x - p * H;
  const int tid = threadIdx.x;
  const int d4 = Dh >> 2;
  for (int i = tid; i < Nq * d4; i += MHA_THREADS) {
    const int r = i / d4, c = (i - r * d4) * 4;
    *reinterpret_cast<float4*>(Qs + r * DP + c) =
        __ldg(reinterpret_cast<const float4*>(q + ((int64_t)p * Nq + r) * ldq + h * Dh + c));
  }
  for (int i = tid; i < Nk * d4; i += MHA_THREADS) {
    const int r = i / d4, c = (i - r * d4) * 4;
    *reinterpret_cast<float4*>(Ks + r * DP + c) =
        __ldg(reinterpret_cast<const float4*>(k + ((int64_t)p * Nk + r) * ldk + h * Dh + c));
    *reinterpret_cast<float4*>(Vs + r * DP + c) =
        __ldg(reinterpret_cast<const float4*>(v + ((int64_t)p * Nk + r) * ldv + h * Dh + c));
  }
  __syncthreads();
  // scores
  for (int e = tid; e < Nq * Nk; e += MHA_THREADS) {
    const int i = e / Nk, j = e - i * Nk;
    const float4* qa = reinterpret_cast<const float4*>(Qs + i * DP);
    const float4* kb = reinterpret_cast<const float4*>(Ks + j * DP);
    float acc = 0.0f;
    for (int c = 0; c < d4; ++c) {
      const float4 a = qa[c], b = kb[c];
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    S[i * (Nk + 1) + j] = acc * scale;
  }
  __syncthreads();
  // row softmax: one warp per row
  const int warp = tid >> 5, lane = tid & 31;
  for (int i = warp; i < Nq; i += MHA_THREADS / 32) {
    float* row = S + i * (Nk + 1);
    float mx = -INFINITY;
    for (int j = lane; j < Nk; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.0f;
    for (int j = lane; j < Nk; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < Nk; j += 32) row[j] *= inv;
  }
  __syncthreads();
  // mean attention weight of every key over the queries
  for (int j = tid; j < Nk; j += MHA_THREADS) {
    float acc = 0.0f;
    for (int i = 0; i < Nq; ++i) acc += S[i * (Nk + 1) + j];
    pbar[j] = acc / (float)Nq;
  }
  __syncthreads();
  for (int d = tid; d < Dh; d += MHA_THREADS) {
    float acc = 0.0f;
    for (int j = 0; j < Nk; ++j) acc = fmaf(pbar[j], Vs[j * DP + d], acc);
    out[(int64_t)p * ldo + h * Dh + d] = acc;
  }
}

}  // namespace grafp

using namespace grafp;

extern "C" int grafp_mha_pool_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v,
                                  int64_t ldv, int P, int Nq, int Nk, int H, int Dh, float scale, float* out,
                                  int64_t ldo, void* stream) {
  GRAFP_REQUIRE(P <= 0 || (q && k && v && out), "mha_pool: null pointer");
  GRAFP_REQUIRE(P >= 0 && Nq > 0 && Nk > 0 && H > 0 && Dh > 0 && Dh % 4 == 0, "mha_pool: bad sizes");
  GRAFP_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0, "mha_pool: row strides must be multiples of 4");
  GRAFP_REQUIRE(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                "mha_pool: q / k / v must be 16-byte aligned");
  if (P == 0) return 0;
  const size_t smem = ((size_t)(Nq + 2 * Nk) * (Dh + 4) + (size_t)Nq * (Nk + 1) + Nk) * sizeof(float);
  GRAFP_REQUIRE(smem <= 200 * 1024, "mha_pool: (Nq=%d, Nk=%d, Dh=%d) does not fit shared memory", Nq, Nk, Dh);
  GRAFP_REQUIRE((int64_t)P * H <= 2147483647LL, "mha_pool: too many (pair, head) blocks");
  cudaFuncSetAttribute(mha_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  mha_pool_kernel<<<(unsigned)(P * H), MHA_THREADS, smem, as_stream(stream)>>>(q, ldq, k, ldk, v, ldv, Nq, Nk, H, Dh,
                                                                              scale, out, ldo);
  return check_launch("mha_pool");
}
