// Dense dilated kNN graph on the tensor cores (sm_100a).
//
// Pairwise distance = a 3xTF32 tcgen05 Gram-tile contraction fed by TMA; the 128 x N distance
// tile lives only in TMEM; the epilogue is a thread-per-row (TMEM lane = node) streaming top-(k*d)
// insertion that emits every d-th rank.  The N x N matrix never reaches shared or global memory.
//
//   prepass   rinv[m] = 1 / max(||x_m||, 1e-12), sq[m] = sum_c (x_mc * rinv_m)^2    (one read of x)
//   tile      rows = 128 consecutive nodes; columns = the nodes of the rows' graph (N >= 128) or
//             the same 128 nodes (N < 128, block-diagonal mask in the epilogue)
//   transform both operand stages are scaled by rinv (F.normalize) and split hi/lo in shared memory
//   MMA       D = Xn_rows * Xn_cols^T, 3 kind::tf32 passes, fp32 accumulate in TMEM (double-buffered)
//   epilogue  dist = (sq_i + (-2 * dot)) + sq_j  (reference association, torch_edge.py:16-18),
//             ascending (distance, index) insertion, lowest index first on exact ties
//
// Warp roles (576 threads): warp 0 TMA, warp 1 MMA issuer + TMEM allocator, warps 2-9 epilogue,
// warps 10-17 transform.  The epilogue is two groups of four warps (one warp per TMEM lane quadrant
// each); group g owns the tiles whose accumulator sits in TMEM buffer g, so the groups never exchange
// data and two independent instruction streams share every scheduler.
//
// Selection.  A sorted-insertion select network costs ~20 ALU-pipe instructions per distance and the
// ALU pipe issues a warp instruction every other cycle: ncu showed the first version ALU-bound (73 %
// pipe utilisation, 483 us for the N = 256 stage).  The epilogue therefore selects in two passes over
// the TMEM tile:
//   pass 1  minimum of every group of bn/16 columns (one FMNMX per distance); the (k*d)-th smallest
//           group minimum tau bounds the (k*d)-th smallest distance of the row from above
//   pass 2  distances are recomputed (bit-identical) and the few with dist <= tau are appended, in
//           column order, to a short per-row candidate list in shared memory (one FSETP + a predicated
//           store per distance)
//   final   exact (distance, lowest index first) insertion over the candidates only
// A row whose list overflows (mass ties), or a shape with fewer column groups than k*d, takes the
// full select-network scan instead, so the result is always the exact one.
#include <stdlib.h>
#include "tc_common.cuh"

namespace grafp {

constexpr int KT_THREADS = 576;
constexpr int KT_MAX_STAGES = 6;
constexpr int KT_XF_THREADS = 256;
constexpr int KT_EPI_THREADS = 256;

struct KnnTcParams {
  int N, C, kk, d, k;
  int64_t M;
  int bn;                 // columns per tile: N (N >= 128) or 128
  int kbk;                // k-block width in fp32 elements: 32 (128-byte rows, 128B swizzle) or 16 (64B)
  int thresh;             // two-pass threshold selection (needs k*d <= column groups of a graph)
  int stages;
  const float* rinv; const float* sq;   // prepass outputs (rinv holds the normalize DENOMINATOR), or (rinv == nullptr) sq = raw sum of squares
  int normalize;
  int32_t* idx; float* dist;
  uint32_t tmem_cols;
};

// lanes_per_row = min(32, C/4) lanes cooperate on one node, 32/lanes_per_row nodes per warp
__global__ void __launch_bounds__(256)
knn_rownorm_kernel(const float* __restrict__ x, int64_t M, int C, int normalize, int lanes_per_row,
                   float* __restrict__ rinv, float* __restrict__ sq) {
  const int rows_per_warp = 32 / lanes_per_row;
  const int lane = threadIdx.x & 31;
  const int sub = lane / lanes_per_row, sl = lane % lanes_per_row;
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t row = warp_id * rows_per_warp + sub;
  const bool ok = row < M;
  const float* xr = x + (ok ? row : 0) * C;
  float s = 0.0f;
  if (ok)
    for (int c = sl * 4; c < C; c += lanes_per_row * 4) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
  for (int o = lanes_per_row >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float ri = 1.0f, q = s;          // ri: the F.normalize denominator max(||x||, 1e-12) (1 when not normalising)
  if (normalize) {
    ri = fmaxf(sqrtf(s), 1e-12f);
    float t = 0.0f;
    if (ok)
      for (int c = sl * 4; c < C; c += lanes_per_row * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float a0 = __fdiv_rn(v.x, ri), a1 = __fdiv_rn(v.y, ri), a2 = __fdiv_rn(v.z, ri), a3 = __fdiv_rn(v.w, ri);
        t = fmaf(a0, a0, t); t = fmaf(a1, a1, t); t = fmaf(a2, a2, t); t = fmaf(a3, a3, t);
      }
    for (int o = lanes_per_row >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    q = t;
  }
  if (ok && sl == 0) { rinv[row] = ri; sq[row] = q; }
}

template <int KMAX>
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmCols, const KnnTcParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[KT_MAX_STAGES];
  __shared__ __align__(8) uint64_t xf_bar[KT_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[KT_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_sq[2][2][256];   // [group][tile parity]: squared norms of the column set

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  // 256-column tiles use 16-wide k-blocks: a 32-wide stage is 64 KB, only two fit, and the TMA -> transform
  // -> MMA chain of one tile could not overlap the next tile's loads (measured 11 k cycles per tile, all latency)
  const uint32_t b_bytes = (uint32_t)p.bn * p.kbk * 4;
  // The 128 rows of a tile are a subset of its column set (same graph), so one stage holds only
  // the column operand [hi | lo]; the row operand is a 1024-aligned window into it.
  const uint32_t stage_bytes = 2u * b_bytes;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto b_hi = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto b_lo = [&](int s) { return smem + (size_t)s * stage_bytes + b_bytes; };

  const int nkb = p.C / p.kbk;
  const int64_t total_tiles = (p.M + TC_BM - 1) / TC_BM;
  auto col_start = [&](int64_t m0) -> int64_t { return p.N >= TC_BM ? (m0 / p.N) * p.N : m0; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmCols);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xf_bar[s], KT_XF_THREADS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();            // nothing above reads or writes a tensor (common.cuh)

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * TC_BM, c0 = col_start(m0);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full_bar[s], b_bytes);
          tma_load_2d(b_hi(s), &tmCols, kb * p.kbk, (int)c0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, p.bn);
      uint32_t it = 0, ti = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
        mbar_wait(&tmem_empty_bar[buf], tph ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn;
        const int64_t m0 = tile * TC_BM;
        const uint32_t row_off = (uint32_t)(m0 - col_start(m0)) * (uint32_t)(p.kbk * 4);   // 0 or 128 rows
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(&xf_bar[s], ph);
          tc_fence_after();
          const bool w64 = p.kbk == 16;                  // 64-byte operand rows: 64B swizzle descriptors
          const uint64_t dah = w64 ? umma_desc_sw64(smem_u32(b_hi(s)) + row_off) : umma_desc_sw128(smem_u32(b_hi(s)) + row_off);
          const uint64_t dal = w64 ? umma_desc_sw64(smem_u32(b_lo(s)) + row_off) : umma_desc_sw128(smem_u32(b_lo(s)) + row_off);
          const uint64_t dbh = w64 ? umma_desc_sw64(smem_u32(b_hi(s))) : umma_desc_sw128(smem_u32(b_hi(s)));
          const uint64_t dbl = w64 ? umma_desc_sw64(smem_u32(b_lo(s))) : umma_desc_sw128(smem_u32(b_lo(s)));
          const int ksteps = p.kbk / 8;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);
            umma_tf32(tacc, dal + koff, dbh + koff, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tacc, dah + koff, dbl + koff, idesc, 1u);
            umma_tf32(tacc, dah + koff, dbh + koff, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else if (warp >= 10) {
    // ===== transform (256 threads): scale rows by rinv (F.normalize), split tf32 hi / lo =====
    // hi = top 19 bits of v, lo = v - hi (exact); each thread owns the same rows in every k-block
    const int t = threadIdx.x - 320;
    const int per = (int)(b_bytes / 16) / KT_XF_THREADS;          // float4 per thread per stage: 4 or 8
    uint32_t it = 0;
    // the row scales of a tile come from global memory: they are fetched one tile ahead so their latency
    // hides behind the previous tile's k-loop (the profile showed the transform warps stalled ~45 % of
    // the time on these loads at every tile start, with the MMA starving behind them)
    // Only the raw loads are issued ahead; the sqrt / reciprocal that consume them run at the next tile's
    // start, when the data has long arrived.
    auto load_raw = [&](int64_t tile, float (&dst)[8]) {
      const int64_t c0 = col_start(tile * TC_BM);
      const int rsh = p.kbk == 16 ? 2 : 3;             // float4 per operand row: 4 or 8
      const float* src = p.rinv ? p.rinv : p.sq;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t node = c0 + (t >> rsh) + (KT_XF_THREADS >> rsh) * i;
        dst[i] = (tile < total_tiles && i < per && node < p.M) ? __ldg(src + node) : 0.0f;
      }
    };
    float ri_next[8];
    load_raw(blockIdx.x, ri_next);
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      // F.normalize divides (torch_edge.py:281): x / den is reproduced as the correctly rounded quotient from the
      // reciprocal and one residual correction (q = x*r; q += (x - q*den)*r), not as the 1-ulp-off product x*r
      float ri[8], dn[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float s0 = ri_next[i];                   // the denominator itself (prepass) or the raw sum of squares
        dn[i] = p.rinv ? s0 : (p.normalize ? fmaxf(sqrtf(s0), 1e-12f) : 1.0f);
        ri[i] = __frcp_rn(dn[i]);
      }
      load_raw(tile + gridDim.x, ri_next);
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1u;
        mbar_wait(&full_bar[s], ph);
        float4* bh = reinterpret_cast<float4*>(b_hi(s));
        float4* bl = reinterpret_cast<float4*>(b_lo(s));
#pragma unroll
        for (int half4 = 0; half4 < 2; ++half4) {
          if (half4 * 4 < per) {
            float4 v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = bh[t + KT_XF_THREADS * (half4 * 4 + i)];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float rr = ri[half4 * 4 + i], dd = dn[half4 * 4 + i];
              float x0 = v[i].x * rr, x1 = v[i].y * rr, x2 = v[i].z * rr, x3 = v[i].w * rr;
              x0 = fmaf(fmaf(-x0, dd, v[i].x), rr, x0);
              x1 = fmaf(fmaf(-x1, dd, v[i].y), rr, x1);
              x2 = fmaf(fmaf(-x2, dd, v[i].z), rr, x2);
              x3 = fmaf(fmaf(-x3, dd, v[i].w), rr, x3);
              float4 h, l;
              h.x = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
              h.y = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
              h.z = __uint_as_float(__float_as_uint(x2) & 0xFFFFE000u);
              h.w = __uint_as_float(__float_as_uint(x3) & 0xFFFFE000u);
              l.x = x0 - h.x; l.y = x1 - h.y; l.z = x2 - h.z; l.w = x3 - h.w;
              bh[t + KT_XF_THREADS * (half4 * 4 + i)] = h;
              bl[t + KT_XF_THREADS * (half4 * 4 + i)] = l;
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&xf_bar[s]);
      }
    }
  } else {
    // ===== epilogue (warps 2..9): per-row top-(k*d) over the TMEM distance tile =====
    constexpr int L = KMAX + 8;                  // candidate slots per row (+8 overflow sink slots)
    const int ew = warp - 2;
    const uint32_t grp = (uint32_t)(ew >> 2);    // owns the tiles of TMEM buffer `grp`
    const int et = (ew & 3) * 32 + lane;         // 0..127 inside the group
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    float2* lst = reinterpret_cast<float2*>(smem + (size_t)S * stage_bytes) + (size_t)grp * (L + 8) * TC_BM + r;
    const uint32_t lst_addr = smem_u32(lst), lst_sink = lst_addr + (uint32_t)L * TC_BM * 8u;
    const bool full_tile = p.N >= p.bn;          // every column of the tile belongs to the row's graph
    const int gshift = p.bn == 256 ? 4 : 3;      // log2(columns per group): 16 groups per tile
    // squared norms (own row + two staged columns) are fetched one of MY tiles ahead (raw loads only)
    auto sq_final = [&](float s0) {
      if (p.rinv || !p.normalize) return s0;
      const float ri = __frcp_rn(fmaxf(sqrtf(s0), 1e-12f));
      return s0 * ri * ri;
    };
    auto load_sq_raw = [&](int64_t tile, float (&dst)[3]) {
      const int64_t m0 = tile * TC_BM, c0 = col_start(m0);
      const bool ok = tile < total_tiles;
      dst[0] = (ok && m0 + r < p.M) ? __ldg(p.sq + m0 + r) : 0.0f;
      dst[1] = (ok && c0 + et < p.M) ? __ldg(p.sq + c0 + et) : 0.0f;
      dst[2] = (ok && p.bn > 128 && c0 + 128 + et < p.M) ? __ldg(p.sq + c0 + 128 + et) : 0.0f;
    };
    float sq_next[3];
    load_sq_raw((int64_t)blockIdx.x + (int64_t)grp * gridDim.x, sq_next);
    uint32_t ti = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      if ((ti & 1u) != grp) continue;
      const int64_t m0 = tile * TC_BM, c0 = col_start(m0);
      const float sq_now[3] = {sq_final(sq_next[0]), sq_final(sq_next[1]), sq_final(sq_next[2])};
      load_sq_raw(tile + 2 * (int64_t)gridDim.x, sq_next);
      const uint32_t buf = grp, tph = (ti >> 1) & 1u;
      const int64_t grow = m0 + r;
      const bool row_ok = grow < p.M;
      const int64_t gs = row_ok ? (grow / p.N) * p.N : c0;     // first node of this row's graph
      const int lo_col = (int)(gs - c0);                       // its first column inside the tile
      const unsigned ncols = row_ok ? (unsigned)p.N : 0u;      // columns [lo_col, lo_col + N) are its graph
      const float sqi = row_ok ? sq_now[0] : 0.0f;
      // stage the column set's squared norms
      float* ssq = s_sq[grp][(ti >> 1) & 1u];
      ssq[et] = sq_now[1];
      if (p.bn > 128) ssq[et + 128] = sq_now[2];
      named_bar_sync(1 + (int)grp, 128);
      const float4* sqv = reinterpret_cast<const float4*>(ssq);
      float bd[KMAX];
      int bj[KMAX];
#pragma unroll
      for (int t = 0; t < KMAX; ++t) { bd[t] = INFINITY; bj[t] = (int)(grow - gs); }
      // Packed tiles (N < 128: 128 / N graphs per tile, block-diagonal): a warp's 32 rows belong to at most two
      // graphs (one when N >= 32), whose columns lie inside ONE aligned window of max(N, 32) columns: only those
      // 32-column chunks are scanned (warp-uniform bounds) -- the first version scanned all 128 columns of every row
      // and masked 50-87 % of them away.
      int c_begin = 0, c_end = p.bn;
      if (!full_tile) {
        const int w = p.N >= 32 ? p.N : 32;
        c_begin = ((quad * 32) / w) * w;
        c_end = c_begin + w;
      }
      mbar_wait(&tmem_full_bar[buf], tph);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn + ((uint32_t)(quad * 32) << 16);
      float v[32];
      // distances of one 32-column chunk, in place: (sq_i + (-2 dot)) + sq_j (reference association)
      auto load_dist = [&](int c) {
        tmem_ld16_nowait(tacc + (uint32_t)c, v);
        tmem_ld16_nowait(tacc + (uint32_t)c + 16u, v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 32; q4 += 4) {
          const float4 s4 = sqv[(c + q4) >> 2];                          // shared-memory broadcast
          v[q4 + 0] = __fadd_rn(fmaf(v[q4 + 0], -2.0f, sqi), s4.x);
          v[q4 + 1] = __fadd_rn(fmaf(v[q4 + 1], -2.0f, sqi), s4.y);
          v[q4 + 2] = __fadd_rn(fmaf(v[q4 + 2], -2.0f, sqi), s4.z);
          v[q4 + 3] = __fadd_rn(fmaf(v[q4 + 3], -2.0f, sqi), s4.w);
        }
      };
      // Branch-free sorted insertion (a select network): rows of a warp are independent, so a
      // data-dependent branch would be taken by some lane at almost every column and serialise the
      // warp.  Strict '<' keeps the earlier (lower index) entry ahead on exact ties.
      auto insert = [&](float dv, int jl) {
        bool lt[KMAX];
#pragma unroll
        for (int t = 0; t < KMAX; ++t) lt[t] = dv < bd[t];
#pragma unroll
        for (int t = KMAX - 1; t > 0; --t) {
          bd[t] = lt[t - 1] ? bd[t - 1] : (lt[t] ? dv : bd[t]);
          bj[t] = lt[t - 1] ? bj[t - 1] : (lt[t] ? jl : bj[t]);
        }
        bd[0] = lt[0] ? dv : bd[0];
        bj[0] = lt[0] ? jl : bj[0];
      };
      auto full_scan = [&]() {
        for (int c = c_begin; c < c_end; c += 32) {
          load_dist(c);
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int jl = c + q - lo_col;
            insert(((unsigned)jl < ncols) ? v[q] : INFINITY, jl);         // outside this row's graph: never
          }
        }
      };
      if (p.thresh) {
        // ---- pass 1: tau = (k*d)-th smallest of the 16 column-group minima ----
        float tb[KMAX];
#pragma unroll
        for (int t = 0; t < KMAX; ++t) tb[t] = INFINITY;
        auto push_min = [&](float x) {                  // value-only sorted insertion (min / max chain)
#pragma unroll
          for (int t = 0; t < KMAX; ++t) {
            const float lo = fminf(tb[t], x);
            x = fmaxf(tb[t], x);
            tb[t] = lo;
          }
        };
        for (int c = c_begin; c < c_end; c += 32) {
          load_dist(c);
          float m8[4];
#pragma unroll
          for (int b8 = 0; b8 < 4; ++b8) {
            float m = fminf(fminf(fminf(v[8 * b8], v[8 * b8 + 1]), fminf(v[8 * b8 + 2], v[8 * b8 + 3])),
                            fminf(fminf(v[8 * b8 + 4], v[8 * b8 + 5]), fminf(v[8 * b8 + 6], v[8 * b8 + 7])));
            if (!full_tile && !((unsigned)(c + 8 * b8 - lo_col) < ncols)) m = INFINITY;
            m8[b8] = m;
          }
          if (gshift == 4) {
            push_min(fminf(m8[0], m8[1]));
            push_min(fminf(m8[2], m8[3]));
          } else {
#pragma unroll
            for (int b8 = 0; b8 < 4; ++b8) push_min(m8[b8]);
          }
        }
        float tau = INFINITY;
#pragma unroll
        for (int t = 0; t < KMAX; ++t)
          if (t == p.kk - 1) tau = tb[t];
        // ---- pass 2: append every distance <= tau (column order) to the row's candidate list ----
        // Predicated PTX (a compare, a predicated 64-bit shared store, a predicated pointer bump): as C++
        // the compiler emitted a divergent branch per distance (45 cycles each in the profile).  The
        // write pointer is clamped into the 8-slot sink once per 8 columns; reaching the sink = overflow.
        uint32_t wp = lst_addr;
        for (int c = c_begin; c < c_end; c += 32) {
          load_dist(c);
#pragma unroll
          for (int b8 = 0; b8 < 4; ++b8) {
            const float tq = (full_tile || (unsigned)(c + 8 * b8 - lo_col) < ncols) ? tau : -INFINITY;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int q = 8 * b8 + u;
              // the bump goes through a fresh register (selp + add): bumping the store's own address
              // register in place made every add wait for the store to release it (short-scoreboard
              // stall, ~30 cycles per distance in the profile)
              uint32_t bump;
              asm volatile(
                  "{\n\t.reg .pred p;\n\t"
                  "setp.le.f32 p, %2, %3;\n\t"
                  "@p st.shared.v2.b32 [%1], {%4, %5};\n\t"
                  "selp.u32 %0, 1024, 0, p;\n\t}"
                  : "=r"(bump)
                  : "r"(wp), "f"(v[q]), "f"(tq), "r"(__float_as_uint(v[q])), "r"(c + q)
                  : "memory");
              wp += bump;
            }
            wp = min(wp, lst_sink);
          }
        }
        const int cnt = (int)((wp - lst_addr) >> 10);   // == L: the list may have overflowed
        if (__any_sync(0xffffffffu, cnt >= L)) {
          full_scan();                                  // mass ties: exact fallback for this warp
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[buf]);
        } else {
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[buf]);            // TMEM reads done: the accumulator may be overwritten
          for (int i = 0; i < cnt; ++i) {               // divergent trip count, a handful per row
            const float2 e = lst[i * TC_BM];
            insert(e.x, __float_as_int(e.y) - lo_col);
          }
        }
      } else {
        full_scan();
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[buf]);
      }
      if (row_ok) {
#pragma unroll
        for (int t = 0; t < KMAX; ++t) {
          if (t < p.kk && (t % p.d) == 0) {
            const int64_t o = grow * p.k + t / p.d;
            p.idx[o] = bj[t];
            if (p.dist) p.dist[o] = bd[t];
          }
        }
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

int knn_tc_supported(int B, int N, int C, int kk) {
  if (kk > 16) return 0;
  if (C % TC_BK != 0) return 0;   // (16-wide k-blocks are an internal choice for 256-column tiles)
  if (N > 256) return 0;
  if (N >= TC_BM) return N % TC_BM == 0;
  return N >= 16 && TC_BM % N == 0;
}

// rinv[M] | pad | sq[M] | pad : the pads let the epilogue read sq in whole float4 groups
size_t knn_tc_workspace_bytes(int B, int N) { return ((size_t)B * N + 128) * 2 * sizeof(float); }

template <int KMAX>
static int knn_tc_launch_t(const CUtensorMap& mc, KnnTcParams p, int grid, cudaStream_t st) {
  // dynamic shared memory: operand stages, then the two groups' candidate lists
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, knn_tc_kernel<KMAX>);
  const size_t stage_bytes = 2 * (size_t)p.bn * p.kbk * 4;
  const size_t list_bytes = 2 * (size_t)(KMAX + 8 + 8) * TC_BM * sizeof(float2);
  int stages = (int)((227 * 1024 - fa.sharedSizeBytes - 2048 - list_bytes) / stage_bytes);
  if (stages > KT_MAX_STAGES) stages = KT_MAX_STAGES;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const size_t smem = stage_bytes * stages + list_bytes + 1024;
  cudaFuncSetAttribute(knn_tc_kernel<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_ex(knn_tc_kernel<KMAX>, dim3(grid), dim3(KT_THREADS), smem, st, 0, mc, p);
  return check_launch("knn_tc");
}

int knn_tc_launch(const float* x, int B, int N, int C, int kk, int d, int k, int normalize,
                  const float* row_sumsq, int32_t* idx, float* dist, float* workspace, cudaStream_t st) {
  const int64_t M = (int64_t)B * N;
  float* rinv = workspace;
  float* sq = workspace + M + 128;
  KnnTcParams p;
  p.normalize = normalize;
  if (row_sumsq) {
    p.rinv = nullptr; p.sq = row_sumsq;
  } else {
  int lpr = 32;
  while (lpr > 1 && lpr * 4 > C) lpr >>= 1;               // power of two, <= C/4
  const int rows_per_block = 8 * (32 / lpr);
  knn_rownorm_kernel<<<(unsigned)((M + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(
      x, M, C, normalize, lpr, rinv, sq);
  if (int rc = check_launch("knn_rownorm")) return rc;
    p.rinv = rinv; p.sq = sq;
  }
  p.N = N; p.C = C; p.kk = kk; p.d = d; p.k = k; p.M = M;
  p.bn = N >= TC_BM ? N : TC_BM;
  p.idx = idx; p.dist = dist;
  {
    static int no_thresh = -1;
    if (no_thresh < 0) { const char* e = getenv("GRAFP_KNN_NO_THRESH"); no_thresh = e ? atoi(e) : 0; }
    const int groups_per_graph = (N < p.bn ? N : p.bn) / (p.bn / 16);   // column groups of one graph
    p.thresh = (!no_thresh && kk <= groups_per_graph) ? 1 : 0;
  }
  uint32_t cols = 32;
  while ((int)cols < 2 * p.bn) cols <<= 1;
  p.tmem_cols = cols;
  p.kbk = p.bn > 128 ? 16 : TC_BK;
  CUtensorMap mc;
  if (p.kbk == 16) {
    if (int rc = tc_make_map_2d_bk16(&mc, x, M, C, C, p.bn)) return rc;
  } else if (int rc = tc_make_map_2d(&mc, x, M, C, C, p.bn)) {
    return rc;
  }
  const int64_t tiles = (M + TC_BM - 1) / TC_BM;
  int grid = sm_count();
  if (tiles < grid) grid = (int)tiles;
  if (kk <= 4) return knn_tc_launch_t<4>(mc, p, grid, st);
  if (kk <= 8) return knn_tc_launch_t<8>(mc, p, grid, st);
  return knn_tc_launch_t<16>(mc, p, grid, st);
}

}  // namespace grafp
