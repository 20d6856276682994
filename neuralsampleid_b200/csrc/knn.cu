// Dense dilated kNN graph, exact fp32 engine.
//
// One CTA = (graph b, tile of R rows).  Phase 0 computes the F.normalize denominators and
// squared norms of all N nodes of the graph; phase 1 forms the R x N distance rows in shared
// memory from channel-chunked register-tiled dot products (the N x N matrix never reaches
// HBM); phase 2 is a warp-per-row top-(k*d) extraction in ascending (distance, index) order
// that emits every d-th rank.  Arithmetic follows the reference's association order:
// dist = (sq_i + (-2 * <xn_i, xn_j>)) + sq_j   (encoder/gcn_lib/torch_edge.py:16-18).
#include <float.h>
#include "common.cuh"

namespace grafp {

constexpr int KNN_CK = 32;          // channel chunk
constexpr int KNN_CT = 64;          // column tile
constexpr int KNN_LDS = KNN_CK + 4; // padded chunk row stride (floats)

template <int TR>   // rows per thread; R = 16 * TR rows per CTA
__global__ void __launch_bounds__(256) knn_kernel(const float* __restrict__ x, int N, int C,
                                                  int kk, int d, int k, int normalize,
                                                  int32_t* __restrict__ idx_out,
                                                  float* __restrict__ dist_out) {
  constexpr int R = 16 * TR;
  extern __shared__ __align__(16) float sm[];
  const int npad = ((N + KNN_CT - 1) / KNN_CT) * KNN_CT;
  const int dld = npad + 16;
  float* s_den = sm;                       // npad
  float* s_sq = s_den + npad;              // npad
  float* s_xi = s_sq + npad;               // R * KNN_LDS
  float* s_xj = s_xi + R * KNN_LDS;        // KNN_CT * KNN_LDS
  float* s_dist = s_xj + KNN_CT * KNN_LDS; // R * dld

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, i0 = blockIdx.x * R;
  const float* xb = x + (size_t)b * N * C;

  // ---- phase 0: denominators and squared norms of the (normalised) nodes ----
  for (int n = warp; n < npad; n += 8) {
    float den = 1.0f, sq = 0.0f;
    if (n < N) {
      const float* xr = xb + (size_t)n * C;
      float s = 0.0f;
      for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
      }
      s = warp_sum(s);
      if (normalize) {
        den = fmaxf(sqrtf(s), 1e-12f);
        float t = 0.0f;
        for (int c = lane * 4; c < C; c += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xr + c);
          const float a0 = __fdiv_rn(v.x, den), a1 = __fdiv_rn(v.y, den);
          const float a2 = __fdiv_rn(v.z, den), a3 = __fdiv_rn(v.w, den);
          t = fmaf(a0, a0, t); t = fmaf(a1, a1, t); t = fmaf(a2, a2, t); t = fmaf(a3, a3, t);
        }
        sq = warp_sum(t);
      } else {
        sq = s;
      }
    }
    if (lane == 0) { s_den[n] = den; s_sq[n] = sq; }
  }
  __syncthreads();

  // ---- phase 1: distance rows ----
  const int tx = tid & 15, ty = tid >> 4;
  for (int j0 = 0; j0 < npad; j0 += KNN_CT) {
    float acc[TR][4];
#pragma unroll
    for (int i = 0; i < TR; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;

    for (int c0 = 0; c0 < C; c0 += KNN_CK) {
      // stage the chunk, normalised: rows of the tile and rows of the column tile
      for (int q = tid; q < (R + KNN_CT) * (KNN_CK / 4); q += 256) {
        const int r = q / (KNN_CK / 4), c = (q % (KNN_CK / 4)) * 4;
        const bool is_i = r < R;
        const int node = is_i ? i0 + r : j0 + (r - R);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (node < N && c0 + c < C) {
          v = *reinterpret_cast<const float4*>(xb + (size_t)node * C + c0 + c);
          if (normalize) {
            const float den = s_den[node];
            v.x = __fdiv_rn(v.x, den); v.y = __fdiv_rn(v.y, den);
            v.z = __fdiv_rn(v.z, den); v.w = __fdiv_rn(v.w, den);
          }
        }
        float* dst = is_i ? s_xi + r * KNN_LDS + c : s_xj + (r - R) * KNN_LDS + c;
        *reinterpret_cast<float4*>(dst) = v;
      }
      __syncthreads();
#pragma unroll
      for (int c = 0; c < KNN_CK; c += 4) {
        float4 a[TR], bb[4];
#pragma unroll
        for (int i = 0; i < TR; ++i)
          a[i] = *reinterpret_cast<const float4*>(s_xi + (ty + 16 * i) * KNN_LDS + c);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          bb[q] = *reinterpret_cast<const float4*>(s_xj + (tx + 16 * q) * KNN_LDS + c);
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float s = acc[i][q];
            s = fmaf(a[i].x, bb[q].x, s); s = fmaf(a[i].y, bb[q].y, s);
            s = fmaf(a[i].z, bb[q].z, s); s = fmaf(a[i].w, bb[q].w, s);
            acc[i][q] = s;
          }
      }
      __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TR; ++i) {
      const int r = ty + 16 * i;
      const int gi = i0 + r;
      const float sqi = gi < N ? s_sq[gi] : 0.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + tx + 16 * q;
        float dv = INFINITY;
        if (j < N) dv = __fadd_rn(__fadd_rn(sqi, -2.0f * acc[i][q]), s_sq[j]);
        s_dist[r * dld + j] = dv;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: ascending (distance, index) extraction, every d-th rank emitted ----
  for (int r = warp; r < R; r += 8) {
    const int gi = i0 + r;
    if (gi >= N) break;
    const float* dr = s_dist + r * dld;
    float pd = -INFINITY;
    int pj = -1;
    for (int round = 0; round < kk; ++round) {
      float bd = INFINITY;
      int bj = 0x7fffffff;
      for (int j = lane; j < npad; j += 32) {
        const float v = dr[j];
        const bool after = (v > pd) || (v == pd && j > pj);
        const bool better = (v < bd) || (v == bd && j < bj);
        if (after && better) { bd = v; bj = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
      }
      if (bj >= N) { bj = gi; }     // NaN rows: fall back to the centre itself
      pd = bd; pj = bj;
      if (lane == 0 && (round % d) == 0) {
        const size_t o = ((size_t)b * N + gi) * k + round / d;
        idx_out[o] = bj;
        if (dist_out) dist_out[o] = bd;
      }
    }
  }
}

template <int TR>
static int knn_launch(const float* x, int B, int N, int C, int kk, int d, int k, int normalize,
                      int32_t* idx, float* dist, cudaStream_t st) {
  constexpr int R = 16 * TR;
  const int npad = ((N + KNN_CT - 1) / KNN_CT) * KNN_CT;
  const size_t smem = sizeof(float) * ((size_t)2 * npad + (size_t)(R + KNN_CT) * KNN_LDS +
                                       (size_t)R * (npad + 16));
  GRAFP_REQUIRE(smem <= 220 * 1024, "knn: N=%d needs %zu B of shared memory", N, smem);
  cudaFuncSetAttribute(knn_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((N + R - 1) / R, nb);
    knn_kernel<TR><<<grid, 256, smem, st>>>(x + (size_t)b0 * N * C, N, C, kk, d, k, normalize,
                                            idx + (size_t)b0 * N * k,
                                            dist ? dist + (size_t)b0 * N * k : nullptr);
    if (int rc = check_launch("knn")) return rc;
  }
  return 0;
}

}  // namespace grafp

using namespace grafp;

namespace grafp {
int knn_tc_supported(int B, int N, int C, int kk);
size_t knn_tc_workspace_bytes(int B, int N);
int knn_tc_launch(const float* x, int B, int N, int C, int kk, int d, int k, int normalize,
                  const float* row_sumsq, int32_t* idx, float* dist, float* workspace, cudaStream_t st);
// large graphs / long lists (knn_big.cu): N in {256, 512, ..., 2048}, k*d <= 64
int knn_big_supported(int B, int N, int C, int kk);
size_t knn_big_workspace_bytes(int B, int N, int C);
int knn_big_launch(const float* x, int B, int N, int C, int kk, int d, int k, int normalize, int32_t* idx,
                   float* dist, void* workspace, cudaStream_t st);
}  // namespace grafp

extern "C" size_t grafp_knn_workspace_bytes(int B, int N, int C, int k, int dilation) {
  if (B <= 0 || N <= 0 || C <= 0 || k <= 0 || dilation <= 0) return 0;
  if (knn_tc_supported(B, N, C, k * dilation)) return knn_tc_workspace_bytes(B, N);
  if (knn_big_supported(B, N, C, k * dilation)) return knn_big_workspace_bytes(B, N, C);
  return 0;
}

extern "C" int grafp_knn_fwd(const float* x, int B, int N, int C, int k, int dilation,
                             int normalize, int engine, const float* row_sumsq, int32_t* idx_out,
                             float* dist_out, void* workspace, size_t workspace_bytes, void* stream) {
  GRAFP_REQUIRE(B >= 0 && N > 0 && C > 0 && k > 0 && dilation > 0, "knn: bad sizes");
  GRAFP_REQUIRE(B == 0 || (x && idx_out), "knn: null pointer");
  GRAFP_REQUIRE(C % 4 == 0, "knn: C=%d must be a multiple of 4", C);
  const int kk = k * dilation;
  GRAFP_REQUIRE(kk <= N, "knn: k*dilation=%d exceeds the %d nodes of a graph", kk, N);
  GRAFP_REQUIRE(N <= 2048, "knn: N=%d above the 2048-node limit of the dense kernel", N);
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const bool tc_ok = knn_tc_supported(B, N, C, kk) && workspace &&
                     workspace_bytes >= knn_tc_workspace_bytes(B, N);
  // (un-normalised rows could leave the fp16 range of the large-graph kernel's operand planes: exact SIMT kernel)
  const bool big_ok = !tc_ok && normalize && knn_big_supported(B, N, C, kk) && workspace &&
                      workspace_bytes >= knn_big_workspace_bytes(B, N, C);
  if (engine == GRAFP_ENGINE_TC_3XTF32) {
    GRAFP_REQUIRE(tc_ok || big_ok, "knn: the tcgen05 engines need (N in {16..128 | 128, 256}, C %% 32 == 0, k*d <= 16) or "
                                   "(N in {256, 512, ..., 2048}, C %% 16 == 0, k*d <= 64), and a workspace of "
                                   "grafp_knn_workspace_bytes()");
  }
  if ((engine == GRAFP_ENGINE_AUTO || engine == GRAFP_ENGINE_TC_3XTF32) && tc_ok)
    return knn_tc_launch(x, B, N, C, kk, dilation, k, normalize, row_sumsq, idx_out, dist_out,
                         static_cast<float*>(workspace), st);
  if ((engine == GRAFP_ENGINE_AUTO || engine == GRAFP_ENGINE_TC_3XTF32) && big_ok)
    return knn_big_launch(x, B, N, C, kk, dilation, k, normalize, idx_out, dist_out, workspace, st);
  GRAFP_REQUIRE(engine == GRAFP_ENGINE_AUTO || engine == GRAFP_ENGINE_SIMT, "knn: unknown engine %d", engine);
  if (N <= 16) return knn_launch<1>(x, B, N, C, kk, dilation, k, normalize, idx_out, dist_out, st);
  if (N <= 32) return knn_launch<2>(x, B, N, C, kk, dilation, k, normalize, idx_out, dist_out, st);
  if (N <= 512) return knn_launch<4>(x, B, N, C, kk, dilation, k, normalize, idx_out, dist_out, st);
  if (N <= 1024) return knn_launch<2>(x, B, N, C, kk, dilation, k, normalize, idx_out, dist_out, st);
  return knn_launch<1>(x, B, N, C, kk, dilation, k, normalize, idx_out, dist_out, st);
}
