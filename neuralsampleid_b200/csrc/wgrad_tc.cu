// Weight-gradient GEMM on the tensor cores (sm_100a):  dW[n, k] += sum_m dY[m, n] * A[m, k]
// (the backward of every 1x1 conv / Linear of the train step, reference train.py:70 loss.backward()).
//
// The contraction runs over the ROWS m of two row-major fp32 matrices, so both tcgen05 operands are "MN-major": the
// shared-memory tile of an operand is a stack of 128-byte rows, each row one contraction index m holding 64
// consecutive bf16 output indices -- the MN-major SWIZZLE_128B atom (8 rows x 128 B, 16-byte units XOR-swizzled by the
// row) that the shared-memory descriptor describes with LBO = stride between 64-column chunks and SBO = stride between
// 8-row groups.  One k-step of kind::f16 (K = 16) consumes two 8-row groups.  (kind::tf32 takes MN-major operands only
// in the 32-byte-atom swizzle that this TMA path does not produce: a first tf32 version computed zeros.)
//
//   grid      one CTA per (output tile 128 n x BN k, group, split of the m range)
//   TMA       per stage (32 rows of m): fp32 boxes of 32 columns x 32 rows: 4 of dY, BN/32 of A (the a1 / a2 source
//             the k tile belongs to)
//   transform fp32 -> bf16 hi = bf16(v), lo = bf16(v - hi), re-laid as 64-column MN-major chunks: the error-compensated
//             "bf16x3" split (<= 3 * 2^-18 per product, unbiased; bf16 keeps the fp32 range that gradients need)
//   MMA       D[128 x BN] += dY_lo^T A_hi + dY_hi^T A_lo + dY_hi^T A_hi, fp32 accumulate in TMEM
//   epilogue  the tile's partial goes to the workspace; wgrad_reduce_kernel adds the splits in a fixed order
//             (deterministic: no atomics anywhere)
// Shapes: n a multiple of 32, k1, k2 multiples of 64 (per group), any m; the Downsample form (tap3), the stem (k = 8) and
// the stage-2 MRConv (k = 32 per source and group) stay on the SIMT kernel of train.cu.
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace grafp {

constexpr int WG_THREADS = 320;      // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue, warps 6-9 transform
constexpr int WG_XF_THREADS = 128;
constexpr int WG_BKM = 32;           // rows of m per stage
constexpr int WG_MAX_STAGES = 4;
constexpr uint32_t WG_CHUNK_BYTES = 32 * 128;            // one box: 32 rows of 128 B (32 fp32 or 64 bf16 columns)
constexpr uint32_t WG_DY_RAW = 4 * WG_CHUNK_BYTES;       // fp32: 128 dY columns
constexpr uint32_t WG_DY_OP = 2 * WG_CHUNK_BYTES;        // bf16: 128 dY columns

struct WgradTcParams {
  int n, k1, k2, groups, bn;
  int64_t m;
  int rows_per_split, splits, stages;
  float* part;                 // (splits, groups * n, k1 + k2)
  uint32_t tmem_cols;
};

// MN-major, SWIZZLE_128B shared-memory matrix descriptor: LBO = byte stride between 128-byte-wide chunks of the
// M/N dimension, SBO = byte stride between groups of 8 contraction rows
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 with bf16 operands, both MN-major (bits 15 / 16)
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmA2, const WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t xf_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const uint32_t a_raw_bytes = (uint32_t)(p.bn / 32) * WG_CHUNK_BYTES;       // fp32 boxes of A
  const uint32_t a_op_bytes = (uint32_t)(p.bn / 64) * WG_CHUNK_BYTES;        // bf16 chunks of A
  const uint32_t raw_bytes = WG_DY_RAW + a_raw_bytes;
  const uint32_t op_bytes = WG_DY_OP + a_op_bytes;                           // one of hi / lo
  const uint32_t stage_bytes = raw_bytes + 2u * op_bytes;    // [dY raw | A raw | dY_hi | A_hi | dY_lo | A_lo]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto raw = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto dy_hi = [&](int s) { return smem + (size_t)s * stage_bytes + raw_bytes; };
  auto a_hi = [&](int s) { return smem + (size_t)s * stage_bytes + raw_bytes + WG_DY_OP; };
  auto dy_lo = [&](int s) { return smem + (size_t)s * stage_bytes + raw_bytes + op_bytes; };
  auto a_lo = [&](int s) { return smem + (size_t)s * stage_bytes + raw_bytes + op_bytes + WG_DY_OP; };

  // tile decode: split fastest, then k tile, n tile, group
  const int K = p.k1 + p.k2;
  const int tiles_k = K / p.bn, tiles_n = (p.n + 127) / 128;
  int id = blockIdx.x;
  const int split = id % p.splits; id /= p.splits;
  const int kt = id % tiles_k; id /= tiles_k;
  const int nt = id % tiles_n;
  const int g = id / tiles_n;
  const int kcol = kt * p.bn;                              // column inside the group's [a1 | a2] concatenation
  const bool from_a2 = kcol >= p.k1;
  const int acol = from_a2 ? g * p.k2 + (kcol - p.k1) : g * p.k1 + kcol;     // column inside the source matrix
  const int ncol = g * p.n + nt * 128;                     // first dY column of the tile
  const int64_t m_begin = (int64_t)split * p.rows_per_split;
  const int64_t m_end = m_begin + p.rows_per_split < p.m ? m_begin + p.rows_per_split : p.m;
  const int nst = m_end > m_begin ? (int)((m_end - m_begin + WG_BKM - 1) / WG_BKM) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xf_bar[s], WG_XF_THREADS);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mA = from_a2 ? &tmA2 : &tmA1;
      for (int it = 0; it < nst; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], raw_bytes);
        // rows beyond m_end inside the last stage belong to the next split: the box is clipped by loading at most
        // the rows of this split -- the tensor map of a split covers [0, m_end) rows (see the host side), rows past
        // it are zero-filled by the TMA unit
        const int r0 = (int)(m_begin + (int64_t)it * WG_BKM);
        for (int c = 0; c < 4; ++c)
          tma_load_2d(raw(s) + c * WG_CHUNK_BYTES, &tmDy, ncol + 32 * c, r0, &full_bar[s]);
        for (int c = 0; c < p.bn / 32; ++c)
          tma_load_2d(raw(s) + WG_DY_RAW + c * WG_CHUNK_BYTES, mA, acol + 32 * c, r0, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_mn(TC_BM, p.bn);
      for (int it = 0; it < nst; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1u;
        mbar_wait(&xf_bar[s], ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < WG_BKM / 16; ++ks) {
          const uint32_t off = (uint32_t)ks * 2048u;          // 16 contraction rows x 128 B
          const uint64_t dah = umma_desc_mn_sw128(smem_u32(dy_hi(s)) + off, WG_CHUNK_BYTES, 1024);
          const uint64_t dal = umma_desc_mn_sw128(smem_u32(dy_lo(s)) + off, WG_CHUNK_BYTES, 1024);
          const uint64_t dbh = umma_desc_mn_sw128(smem_u32(a_hi(s)) + off, WG_CHUNK_BYTES, 1024);
          const uint64_t dbl = umma_desc_mn_sw128(smem_u32(a_lo(s)) + off, WG_CHUNK_BYTES, 1024);
          umma_bf16(tmem_base, dal, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          umma_bf16(tmem_base, dah, dbl, idesc, 1u);
          umma_bf16(tmem_base, dah, dbh, idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&done_bar);
    }
  } else if (warp >= 6) {
    // ===== transform: fp32 boxes -> bf16 hi / lo MN-major chunks =====
    // A destination 16-byte unit ud of row r in 64-column chunk c64 holds columns 8 ud .. 8 ud + 7: fp32 units
    // 2 (ud & 3), 2 (ud & 3) + 1 of row r in the 32-column box 2 c64 + (ud >> 2).  Both sides are 128B-swizzled:
    // unit u of row r lives at unit u ^ (r & 7).
    const int t = threadIdx.x - 192;
    const int nunits = (int)(op_bytes / 16);
    for (int it = 0; it < nst; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1u;
      mbar_wait(&full_bar[s], ph);
      if (it >= S) mbar_wait(&empty_bar[s], ((it / S) & 1u) ^ 1u);      // the MMAs that read this operand slot retired
      const uint8_t* src = raw(s);
      uint8_t* hi = dy_hi(s);
      uint8_t* lo = dy_lo(s);
      const int valid = (int)(m_end - (m_begin + (int64_t)it * WG_BKM));     // rows of this stage inside the split
      for (int q = t; q < nunits; q += WG_XF_THREADS) {
        const int c64 = q >> 8, r = (q >> 3) & 31, ud = q & 7;
        const int box = 2 * c64 + (ud >> 2), us = 2 * (ud & 3);
        const uint8_t* sb = src + (size_t)box * WG_CHUNK_BYTES + r * 128;
        float4 v0 = *reinterpret_cast<const float4*>(sb + ((us ^ (r & 7)) << 4));
        float4 v1 = *reinterpret_cast<const float4*>(sb + (((us + 1) ^ (r & 7)) << 4));
        if (r >= valid) { v0 = make_float4(0.f, 0.f, 0.f, 0.f); v1 = v0; }
        const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t hp[4], lp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          const float2 hf = __bfloat1622float2(h);
          const __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
          hp[i] = *reinterpret_cast<const uint32_t*>(&h);
          lp[i] = *reinterpret_cast<const uint32_t*>(&l);
        }
        const size_t doff = (size_t)c64 * WG_CHUNK_BYTES + r * 128 + ((ud ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(hi + doff) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(lo + doff) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&xf_bar[s]);
    }
  } else {
    // ===== epilogue (warps 2..5): TMEM -> the split's partial tile in the workspace =====
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                        // output row inside the tile = dY column
    const bool row_ok = nt * 128 + r < p.n;
    float* dst = p.part + ((size_t)split * p.groups * p.n + (size_t)g * p.n + (size_t)nt * 128 + r) * (size_t)K + kcol;
    if (nst > 0) {
      mbar_wait(&done_bar, 0);
      tc_fence_after();
    }
    const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16);
    for (int c = 0; c < p.bn; c += 32) {
      float v[32];
      if (nst > 0) {
        tmem_ld16_nowait(tacc + (uint32_t)c, v);
        tmem_ld16_nowait(tacc + (uint32_t)c + 16u, v + 16);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.0f;
      }
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          *reinterpret_cast<float4*>(dst + c + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// dw[i] += sum_s part[s][i], splits added in a fixed order
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, int splits, int64_t count, int K, float* __restrict__ dw, int64_t ldw) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= count) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(part + (size_t)s * count + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float* d = dw + (i / K) * ldw + (i % K);
  float4 o = *reinterpret_cast<float4*>(d);
  o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
  *reinterpret_cast<float4*>(d) = o;
}

static int wg_pick_bn(int k1, int k2) {
  for (int bn = 128; bn >= 64; bn >>= 1)
    if (k1 % bn == 0 && k2 % bn == 0) return bn;
  return 0;
}

static void wg_plan(int64_t m, int n, int k1, int k2, int groups, int* bn, int* splits, int* rows_per_split) {
  *bn = wg_pick_bn(k1, k2);
  const int64_t tiles = (int64_t)((n + 127) / 128) * ((k1 + k2) / (*bn ? *bn : 32)) * groups;
  int64_t s = (2 * (int64_t)sm_count() + tiles - 1) / tiles;        // about two waves of CTAs
  const int64_t max_s = (m + 255) / 256;                             // at least 256 rows (8 stages) per split
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  int64_t rps = ((m + s - 1) / s + WG_BKM - 1) / WG_BKM * WG_BKM;
  if (rps < WG_BKM) rps = WG_BKM;
  *rows_per_split = (int)rps;
  *splits = (int)((m + rps - 1) / rps);
  if (*splits < 1) *splits = 1;
}

int wgrad_tc_supported(int64_t m, int n, int k1, int k2, int groups, int tap3_nodes, int64_t ldy, int64_t lda1,
                       int64_t lda2, int64_t ldw) {
  if (tap3_nodes > 0 || m < 1 || groups < 1) return 0;
  if (n % 32 != 0 || k1 % 64 != 0 || k2 % 64 != 0 || k1 < 64) return 0;
  if (wg_pick_bn(k1, k2) == 0) return 0;
  if (ldy % 4 != 0 || lda1 % 4 != 0 || (k2 && lda2 % 4 != 0) || ldw % 4 != 0 || (k1 + k2) % 4 != 0) return 0;
  return 1;
}

size_t wgrad_tc_workspace_bytes(int64_t m, int n, int k1, int k2, int groups) {
  int bn, splits, rps;
  wg_plan(m, n, k1, k2, groups, &bn, &splits, &rps);
  return (size_t)splits * groups * n * (size_t)(k1 + k2) * sizeof(float);
}

// fp32 row-major (rows, cols), box = 32 columns x 32 rows, 128B swizzle
static int wg_make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = tc_encode_fn();
  GRAFP_REQUIRE(fn, "tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)WG_BKM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled(wgrad) failed (%d)", (int)r);
  return 0;
}

int wgrad_tc_launch(const float* dy, int64_t ldy, const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2,
                    int k2, int64_t m, int n, int groups, float* dw, int64_t ldw, float* workspace, cudaStream_t st) {
  WgradTcParams p;
  wg_plan(m, n, k1, k2, groups, &p.bn, &p.splits, &p.rows_per_split);
  p.n = n; p.k1 = k1; p.k2 = k2; p.groups = groups; p.m = m; p.part = workspace;
  uint32_t cols = 32;
  while ((int)cols < p.bn) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage_bytes = (size_t)WG_DY_RAW + (size_t)(p.bn / 32) * WG_CHUNK_BYTES +
                             2 * ((size_t)WG_DY_OP + (size_t)(p.bn / 64) * WG_CHUNK_BYTES);
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  p.stages = stages;
  const size_t smem = stage_bytes * stages + 1024;
  CUtensorMap mDy, mA1, mA2;
  if (int rc = wg_make_map(&mDy, dy, m, (int64_t)groups * n, ldy)) return rc;
  if (int rc = wg_make_map(&mA1, a1, m, (int64_t)groups * k1, lda1)) return rc;
  if (k2 > 0) {
    if (int rc = wg_make_map(&mA2, a2, m, (int64_t)groups * k2, lda2)) return rc;
  } else {
    mA2 = mA1;
  }
  const int K = k1 + k2;
  const int64_t grid = (int64_t)((n + 127) / 128) * (K / p.bn) * groups * p.splits;
  cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  wgrad_tc_kernel<<<(unsigned)grid, WG_THREADS, smem, st>>>(mDy, mA1, mA2, p);
  if (int rc = check_launch("wgrad_tc")) return rc;
  const int64_t count = (int64_t)groups * n * K;
  wgrad_reduce_kernel<<<(unsigned)((count / 4 + 255) / 256), 256, 0, st>>>(workspace, p.splits, count, K, dw, ldw);
  return check_launch("wgrad_reduce");
}

}  // namespace grafp
