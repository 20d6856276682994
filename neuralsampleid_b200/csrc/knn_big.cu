// Dense dilated kNN graph on the tensor cores for LARGE graphs / long lists (sm_100a):
// N in {256, 512, ..., 2048} nodes per graph, k * dilation <= 64 (BASELINE configs[4]: k in {9, 16, 32} x
// N in {256, 512, 1024, 2048}, dilation on).  knn_tc.cu covers N <= 256 with k * dilation <= 16.
//
//   prepass    one read of x: den = max(||x||, 1e-12), xn = x / den (F.normalize, exact division), sq = sum xn^2, and
//              the IEEE-half operand planes hi = f16(256 xn), lo = f16(256 xn - hi) written once to the workspace (4
//              bytes per element, like the fp32 input; the "f16x3" split of gemm_tc.cu: ~3 * 2^-24 per product), so
//              the main kernel needs no conversion stage and its L2 -> SM operand traffic -- the limiter of the first
//              tf32-plane version -- is halved
// Work unit = 128 consecutive rows (nodes) of one graph; its N columns are swept in tiles of 256:
//   TMA        per k-block of 32 channels: hi / lo of the unit's 128 rows and of the tile's 256 column nodes (64B swizzle)
//   MMA        D[128 x 256] = 2^16 Xn_rows * Xn_cols^T, 3 kind::f16 passes, fp32 accumulate in TMEM (2 x 256 columns:
//              the next tile's contraction overlaps this tile's selection)
//   select     thread-per-row (TMEM lane = node); the two epilogue groups take the two 128-column halves of every
//              tile.  The N x N matrix never leaves TMEM.
//
// Selection (exact), in the DOT-PRODUCT domain: for rows of one graph dist_ij = (sq_i + (-2 dot_ij)) + sq_j
// (torch_edge.py:16-18) is monotone in dot_ij up to the spread of the sq_j (1 +- 3e-7 for normalised rows), so both
// sweeps work on the raw TMEM values with conservative bounds and the exact distance is only formed for the few
// candidates.  Every tile is contracted TWICE per unit (the MMA is cheap next to the selection):
//   sweep 1  maximum dot of every group of g in {4, 8, 16} of the thread's columns -> an upper bound of the group's
//            smallest distance, pushed (two at a time, packed bf16x2 rounded UP) into sorted lists of ceil(kk/4) entries.
//            A row has four lists (two per thread, two threads per row) over disjoint column groups, so
//            tau = max of their ceil(kk/4)-th smallest entries has at least kk distances under it.
//   sweep 2  every column with dot >= theta(tau) is a candidate: (dot, column) goes to a per-thread shared-memory
//            staging area with one predicated store, and at the end of the tile the thread forms the exact distances
//            of its handful of candidates and appends them to its list in an L2-resident per-CTA scratch.
//   final    at the end of the unit one warp per row bitonic-sorts the row's candidates by (distance, index) and
//            emits ranks 0, d, 2d, ...
// A list can hold all of the thread's columns and a tile whose staging area overflows (mass ties) is redone straight
// into the list, so the result is always the exact one.
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace grafp {

constexpr int KB_THREADS = 320;       // warp 0 TMA, warp 1 MMA + TMEM allocator, warps 2-9 epilogue
constexpr int KB_STAGES_MAX = 3;
constexpr int KB_EPI_THREADS = 256;
constexpr int KB_BN = 256;            // columns per tile
constexpr int KB_KBK = 32;            // channels per k-block (64-byte fp16 operand rows, 64B swizzle)
constexpr uint32_t KB_A_BYTES = TC_BM * KB_KBK * 2;    // 8 KB
constexpr uint32_t KB_B_BYTES = KB_BN * KB_KBK * 2;    // 16 KB
constexpr float KB_PRESCALE = 256.0f;                  // operands are 256 xn: the fp16 lo parts stay normal numbers
constexpr float KB_UNSCALE = 1.0f / 65536.0f;          // TMEM holds 2^16 dot
constexpr uint32_t KB_STAGE_BYTES = 2 * (KB_A_BYTES + KB_B_BYTES);   // [A_hi | A_lo | B_hi | B_lo] = 48 KB
constexpr int KB_CAP = 24;            // staged candidates per thread per tile (+ 8 overflow sink slots)
constexpr uint32_t KB_SLOT_BYTES = KB_EPI_THREADS * 8;               // one staging slot of every epilogue thread
constexpr uint32_t KB_STAGING_BYTES = (KB_CAP + 8) * KB_SLOT_BYTES;  // 64 KB

struct KnnBigParams {
  int N, C, kk, d, k;
  int64_t M;
  int stages;
  int gsel;               // columns per group of sweep 1: 4, 8 or 16
  const float* sq;        // (M) squared norms of the normalised rows
  const float2* sq_minmax; // (B) per graph: (min, max) of sq
  int32_t* idx; float* dist;
  float2* scratch;        // per CTA: [2 unit parities][128 rows][2 halves][N/2] (distance, local column)
};

__device__ __forceinline__ uint32_t bf16_up_bits(float x) {
  // bf16 >= x in the high 16 bits (round toward +inf): the threshold may only be over-estimated
  const uint32_t b = __float_as_uint(x);
  return (x >= 0.0f) ? (b + 0xFFFFu) & 0xFFFF0000u : b & 0xFFFF0000u;
}

// ---- final exact selection: one warp per row --------------------------------------------------------------
// A candidate is ordered by the 64-bit key (order-preserving image of the fp32 distance, local column index):
// ascending keys = ascending (distance, lowest index first), the library's tie convention.
// a list entry is (dot, local column): the exact distance in the reference's association (torch_edge.py:16-18),
// dist = (sq_i + (-2 dot)) + sq_j, is formed here, where the sq_j gathers of a whole batch are in flight together
__device__ __forceinline__ float2 knn_cand_dist(float2 e, float sqi, const float* __restrict__ sq_graph) {
  return make_float2(__fadd_rn(fmaf(e.x * KB_UNSCALE, -2.0f, sqi), __ldg(sq_graph + __float_as_int(e.y))), e.y);
}
__device__ __forceinline__ uint64_t knn_key(float2 e) {
  const uint32_t b = __float_as_uint(e.x + 0.0f);                 // -0 -> +0
  const uint32_t s = b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
  return ((uint64_t)s << 32) | (uint32_t)__float_as_int(e.y);
}
__device__ __forceinline__ float knn_key_dist(uint64_t k) {
  const uint32_t s = (uint32_t)(k >> 32);
  return __uint_as_float(s ^ ((s & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu));
}
constexpr uint64_t KNN_KEY_PAD = ~0ull;

__device__ __forceinline__ void knn_emit(const KnnBigParams& p, int64_t row, int self, int rank, uint64_t key) {
  if (rank < p.kk && (rank % p.d) == 0) {
    const int64_t o = row * p.k + rank / p.d;
    const bool empty = key == KNN_KEY_PAD;                        // NaN rows: fall back to the centre itself
    p.idx[o] = empty ? self : (int)(uint32_t)key;
    if (p.dist) p.dist[o] = empty ? INFINITY : knn_key_dist(key);
  }
}

// Bitonic sort of the 32 * E candidate keys of each of R rows at once (element q * 32 + lane of a row in register q;
// the R independent networks are interleaved instruction by instruction, which hides the shuffle and L2 latencies
// that a row-at-a-time sort exposes -- the first version spent most of the kernel there).  The two half lists are
// concatenated, the padding sorts last.  Ranks 0, d, 2d, ... are written out.
template <int E, int R>
__device__ __noinline__ void sort_rows(const float2* scratch_par, const int (*cnt)[TC_BM], int row0, int halfN, int lane,
                                          const KnnBigParams& p, int64_t m0, int64_t gs) {
  uint64_t key[R][E];
  {
    float2 cd[R][E];
    float sqi[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int row = row0 + rr;
      const int c0 = cnt[0][row], c1 = cnt[1][row];
      const float2* l0 = scratch_par + ((size_t)row * 2u) * (size_t)halfN;
      const float2* l1 = l0 + halfN;
      sqi[rr] = __ldg(p.sq + m0 + row);
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int i = q * 32 + lane;
        cd[rr][q] = make_float2(0.0f, __int_as_float(-1));
        if (i < c0 + c1) cd[rr][q] = __ldcg(i < c0 ? l0 + i : l1 + (i - c0));
      }
    }
    float sqj[R][E];
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const int j = __float_as_int(cd[rr][q].y);
        sqj[rr][q] = j >= 0 ? __ldg(p.sq + gs + j) : 0.0f;
      }
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const float2 dj = make_float2(__fadd_rn(fmaf(cd[rr][q].x * KB_UNSCALE, -2.0f, sqi[rr]), sqj[rr][q]), cd[rr][q].y);
        key[rr][q] = __float_as_int(cd[rr][q].y) >= 0 ? knn_key(dj) : KNN_KEY_PAD;
      }
  }
  // The stage loops stay ROLLED (runtime k, j): fully unrolled, the networks of all the instantiations were 128 k SASS
  // instructions and the kernel stalled on instruction fetch a quarter of the time.  Only the register pairings of the
  // in-lane stages (partner register q ^ dq) are static.
#pragma unroll 1
  for (int k = 2; k <= 32 * E; k <<= 1) {
    // in-lane stages j = 32 * dq >= 32 (both elements in this lane's registers), largest distance first
#pragma unroll
    for (int dq = E / 2; dq >= 1; dq >>= 1) {
      if (dq * 32 < k) {
#pragma unroll
        for (int q = 0; q < E; ++q) {
          if ((q & dq) == 0) {
            const bool asc = (((q * 32) & k) == 0);              // bit k of the element index lives in q here
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
              const uint64_t a = key[rr][q], b = key[rr][q | dq];
              const bool sw = asc ? (a > b) : (a < b);
              key[rr][q] = sw ? b : a;
              key[rr][q | dq] = sw ? a : b;
            }
          }
        }
      }
    }
    // cross-lane stages j < 32
#pragma unroll 1
    for (int j = min(k >> 1, 16); j > 0; j >>= 1) {
      const bool lower = (lane & j) == 0;
#pragma unroll
      for (int q = 0; q < E; ++q) {
        const bool asc = (((q * 32 + lane) & k) == 0);
        const bool take_min = (asc == lower);
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const uint64_t mine = key[rr][q];
          const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, j);
          const uint64_t mn = mine < other ? mine : other, mx = mine < other ? other : mine;
          key[rr][q] = take_min ? mn : mx;
        }
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < R; ++rr) {
    const int64_t orow = m0 + row0 + rr;
#pragma unroll
    for (int q = 0; q < E; ++q) knn_emit(p, orow, (int)(orow - gs), q * 32 + lane, key[rr][q]);
  }
}

// more than 512 candidates (mass ties, e.g. an all-zero input): kk rounds of warp-wide minimum extraction
// straight from the lists
__device__ __noinline__ void extract_row(const float2* l0, int c0, const float2* l1, int c1, int lane,
                                         const KnnBigParams& p, int64_t row, int self) {
  const float sqi = __ldg(p.sq + row);
  const float* sq_graph = p.sq + (row - self);
  uint64_t prev = 0;
  bool first = true;
  for (int rank = 0; rank < p.kk; ++rank) {
    uint64_t best = KNN_KEY_PAD;
    for (int i = lane; i < c0 + c1; i += 32) {
      const uint64_t k = knn_key(knn_cand_dist(i < c0 ? __ldcg(l0 + i) : __ldcg(l1 + (i - c0)), sqi, sq_graph));
      if ((first || k > prev) && k < best) best = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    if (lane == 0) knn_emit(p, row, self, rank, best);
    prev = best;
    first = false;
  }
}

template <int KH>
__global__ void __launch_bounds__(KB_THREADS, 1)
knn_big_kernel(const __grid_constant__ CUtensorMap tmRows, const __grid_constant__ CUtensorMap tmCols,
               const KnnBigParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[KB_STAGES_MAX];
  __shared__ __align__(8) uint64_t empty_bar[KB_STAGES_MAX];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_cnt[2][2][TC_BM];                  // [unit parity][half][row]: candidates appended
  __shared__ float s_tau[2][TC_BM];                   // [half][row]: the two threads of a row exchange their bounds

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem;                                               // (KB_CAP + 8) slots x 256 threads x 8 B
  uint8_t* stage0 = smem + KB_STAGING_BYTES;
  auto a_hi = [&](int s) { return stage0 + (size_t)s * KB_STAGE_BYTES; };
  auto a_lo = [&](int s) { return stage0 + (size_t)s * KB_STAGE_BYTES + KB_A_BYTES; };
  auto b_hi = [&](int s) { return stage0 + (size_t)s * KB_STAGE_BYTES + 2 * KB_A_BYTES; };
  auto b_lo = [&](int s) { return stage0 + (size_t)s * KB_STAGE_BYTES + 2 * KB_A_BYTES + KB_B_BYTES; };

  const int nkb = p.C / KB_KBK;
  const int T = p.N / KB_BN;                               // column tiles per unit
  const int SU = T == 1 ? 1 : 2 * T;                       // MMA steps per unit (a single tile stays in TMEM for both sweeps)
  const int64_t units = p.M / TC_BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmRows);
    tma_prefetch_desc(&tmCols);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], KB_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer: hi / lo planes of the rows and of the column tile =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int64_t m0 = u * TC_BM, gs = (m0 / p.N) * p.N;
        for (int su = 0; su < SU; ++su) {
          const int c0 = (int)(gs + (int64_t)(su % T) * KB_BN);
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full_bar[s], KB_STAGE_BYTES);
            tma_load_3d(a_hi(s), &tmRows, kb * KB_KBK, (int)m0, 0, &full_bar[s]);
            tma_load_3d(a_lo(s), &tmRows, kb * KB_KBK, (int)m0, 1, &full_bar[s]);
            tma_load_3d(b_hi(s), &tmCols, kb * KB_KBK, c0, 0, &full_bar[s]);
            tma_load_3d(b_lo(s), &tmCols, kb * KB_KBK, c0, 1, &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(TC_BM, KB_BN);
      uint32_t it = 0, st = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        for (int su = 0; su < SU; ++su, ++st) {
          const uint32_t buf = st & 1u, tph = (st >> 1) & 1u;
          mbar_wait(&tmem_empty_bar[buf], tph ^ 1u);
          tc_fence_after();
          const uint32_t tacc = tmem_base + buf * (uint32_t)KB_BN;
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1u;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint64_t dah = umma_desc_sw64(smem_u32(a_hi(s))), dal = umma_desc_sw64(smem_u32(a_lo(s)));
            const uint64_t dbh = umma_desc_sw64(smem_u32(b_hi(s))), dbl = umma_desc_sw64(smem_u32(b_lo(s)));
#pragma unroll
            for (int ks = 0; ks < KB_KBK / 16; ++ks) {              // UMMA_K = 16 halves = 32 B
              const uint64_t koff = (uint64_t)((ks * 16 * 2) >> 4);
              umma_bf16(tacc, dal + koff, dbh + koff, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
              umma_bf16(tacc, dah + koff, dbl + koff, idesc, 1u);
              umma_bf16(tacc, dah + koff, dbh + koff, idesc, 1u);
            }
            umma_commit(&empty_bar[s]);
          }
          umma_commit(&tmem_full_bar[buf]);
        }
      }
    }
  } else {
    // ===== epilogue (warps 2..9): thread-per-row selection, half `grp` of every tile's columns =====
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int quad = warp & 3;                     // the TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;
    const int e = grp * 128 + r;                   // 0..255 over the epilogue threads
    const int halfN = p.N >> 1;
    const int hq = (p.kk + 3) >> 2;                // entries per threshold list: four lists per row (two per thread)
    float2* cta_scratch = p.scratch + (size_t)blockIdx.x * 2u * TC_BM * 2u * (size_t)halfN;
    const uint32_t stg_addr = smem_u32(staging) + (uint32_t)e * 8u;
    const uint32_t stg_sink = stg_addr + (uint32_t)KB_CAP * KB_SLOT_BYTES;
    const float2* stg = reinterpret_cast<const float2*>(staging) + e;
    uint32_t st = 0, un = 0;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++un) {
      const int64_t m0 = u * TC_BM, gs = (m0 / p.N) * p.N;
      const int64_t grow = m0 + r;
      const float sqi = __ldg(p.sq + grow);
      const uint32_t par = un & 1u;
      float2* my_list = cta_scratch + (((size_t)par * TC_BM + r) * 2u + grp) * (size_t)halfN;
      float2* wp = my_list;
      // spread of the graph's squared norms (1 +- a few ulp for normalised rows; from the prepass): the dot-domain
      // bounds below
      const float2 mm = __ldg(p.sq_minmax + m0 / p.N);
      const float sqmin = mm.x, sqmax = mm.y;
      const float slack = 2e-6f * fmaxf(1.0f, sqi + sqmax);
      const float ub0 = (sqi + sqmax) + slack;        // dist_j <= ub0 - 2 dot_j for every column j of the graph
      __nv_bfloat162 tb[KH];
      {
        const uint32_t inf2 = 0x7F807F80u;
#pragma unroll
        for (int i = 0; i < KH; ++i) tb[i] = *reinterpret_cast<const __nv_bfloat162*>(&inf2);
      }
      float v[32];
      float theta = -INFINITY;
      for (int su = 0; su < 2 * T; ++su) {
        const int t = su % T;
        const bool sweep2 = su >= T;
        const uint32_t buf = st & 1u, tph = (st >> 1) & 1u;
        if (!(T == 1 && sweep2)) {                    // a single tile is kept in TMEM for both sweeps
          mbar_wait(&tmem_full_bar[buf], tph);
          tc_fence_after();
        }
        const uint32_t tacc = tmem_base + buf * (uint32_t)KB_BN + (uint32_t)(grp * 128) + ((uint32_t)(quad * 32) << 16);
        auto load_dots = [&](int c) {
          tmem_ld16_nowait(tacc + (uint32_t)c, v);
          tmem_ld16_nowait(tacc + (uint32_t)c + 16u, v + 16);
          tmem_ld_wait();
        };
        if (!sweep2) {
          // ---- sweep 1: per-group maximum dot -> upper bound of the group's smallest distance -> threshold lists ----
          auto push2 = [&](float da, float db) {
            const float a = fmaf(da, -2.0f * KB_UNSCALE, ub0), b = fmaf(db, -2.0f * KB_UNSCALE, ub0);
            const uint32_t pk = __byte_perm(bf16_up_bits(a), bf16_up_bits(b), 0x7632);
            __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int i = 0; i < KH; ++i) {
              const __nv_bfloat162 lo = __hmin2(tb[i], x);
              x = __hmax2(tb[i], x);
              tb[i] = lo;
            }
          };
          for (int c = 0; c < 128; c += 32) {
            load_dots(c);
            float m4[8];
#pragma unroll
            for (int b = 0; b < 8; ++b)
              m4[b] = fmaxf(fmaxf(v[4 * b], v[4 * b + 1]), fmaxf(v[4 * b + 2], v[4 * b + 3]));
            if (p.gsel == 4) {
              push2(m4[0], m4[1]); push2(m4[2], m4[3]); push2(m4[4], m4[5]); push2(m4[6], m4[7]);
            } else if (p.gsel == 8) {
              push2(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
              push2(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7]));
            } else {
              push2(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
            }
          }
          if (t == T - 1) {
            float mine = INFINITY;
#pragma unroll
            for (int i = 0; i < KH; ++i)
              if (i == hq - 1) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(&tb[i]);
                mine = fmaxf(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
              }
            s_tau[grp][r] = mine;
            named_bar_sync(1, KB_EPI_THREADS);
            const float tau = fmaxf(mine, s_tau[grp ^ 1][r]);
            // dist_j <= tau  =>  dot_j >= (sq_i + sq_j - tau) / 2 - rounding  >=  theta
            theta = (((sqi + sqmin) - tau) * 0.5f - slack) * (1.0f / KB_UNSCALE);     // in the 2^16-scaled domain
          }
        } else {
          // ---- sweep 2: stage (dot, column) of every column with dot >= theta, then append exact distances ----
          const int jl0 = t * KB_BN + grp * 128;
          uint32_t sp = stg_addr;
          for (int c = 0; c < 128; c += 32) {
            load_dots(c);
#pragma unroll
            for (int b8 = 0; b8 < 4; ++b8) {
#pragma unroll
              for (int q8 = 0; q8 < 8; ++q8) {
                const int q = 8 * b8 + q8;
                uint32_t bump;
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "setp.ge.f32 p, %2, %3;\n\t"
                    "@p st.shared.v2.b32 [%1], {%4, %5};\n\t"
                    "selp.u32 %0, %6, 0, p;\n\t}"
                    : "=r"(bump)
                    : "r"(sp), "f"(v[q]), "f"(theta), "r"(__float_as_uint(v[q])), "r"(jl0 + c + q), "n"(KB_SLOT_BYTES)
                    : "memory");
                sp += bump;
              }
              sp = min(sp, stg_sink);
            }
          }
          const int cnt = (int)((sp - stg_addr) / KB_SLOT_BYTES);
          if (__any_sync(0xffffffffu, cnt >= KB_CAP)) {
            // a staging area overflowed (mass ties): this warp redoes the tile straight into the lists
            for (int c = 0; c < 128; c += 32) {
              load_dots(c);
#pragma unroll
              for (int q = 0; q < 32; ++q)
                if (v[q] >= theta) *wp++ = make_float2(v[q], __int_as_float(jl0 + c + q));
            }
          } else {
            for (int i = 0; i < cnt; ++i) *wp++ = stg[(size_t)i * KB_EPI_THREADS];
          }
        }
        if (T > 1 || sweep2) {                        // done with this accumulator
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[buf]);
          ++st;
        }
      }
      // ---- end of the unit: publish my count; then the eight warps sort 16 rows each (warp per row) ----
      s_cnt[par][grp][r] = (int)(wp - my_list);
      __threadfence_block();
      named_bar_sync(1, KB_EPI_THREADS);
      {
        const float2* sp_par = cta_scratch + (size_t)par * TC_BM * 2u * (size_t)halfN;
        const int (*cn)[TC_BM] = s_cnt[par];
        for (int g4 = 0; g4 < 4; ++g4) {              // my 16 rows, four at a time
          const int row0 = ew * 16 + g4 * 4;
          int cmax = 0;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) cmax = max(cmax, cn[0][row0 + rr] + cn[1][row0 + rr]);
          if (cmax <= 32) sort_rows<1, 4>(sp_par, cn, row0, halfN, lane, p, m0, gs);
          else if (cmax <= 64) sort_rows<2, 4>(sp_par, cn, row0, halfN, lane, p, m0, gs);
          else if (cmax <= 128) {
            sort_rows<4, 2>(sp_par, cn, row0, halfN, lane, p, m0, gs);
            sort_rows<4, 2>(sp_par, cn, row0 + 2, halfN, lane, p, m0, gs);
          } else {
            for (int rr = row0; rr < row0 + 4; ++rr) {
              const int c = cn[0][rr] + cn[1][rr];
              if (c <= 256) sort_rows<8, 1>(sp_par, cn, rr, halfN, lane, p, m0, gs);
              else if (c <= 512) sort_rows<16, 1>(sp_par, cn, rr, halfN, lane, p, m0, gs);
              else {
                const float2* l0 = sp_par + ((size_t)rr * 2u) * (size_t)halfN;
                extract_row(l0, cn[0][rr], l0 + halfN, cn[1][rr], lane, p, m0 + rr, (int)(m0 + rr - gs));
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// One read of x: normalise (exact division), squared norms, tf32 hi / lo operand planes.
// lanes_per_row = min(32, C/4) lanes cooperate on one node.
__global__ void __launch_bounds__(256)
knn_big_prep_kernel(const float* __restrict__ x, int64_t M, int C, int normalize, int lanes_per_row,
                    float* __restrict__ sq, __half* __restrict__ hi, __half* __restrict__ lo) {
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
  const float den = normalize ? fmaxf(sqrtf(s), 1e-12f) : 1.0f;
  float t = 0.0f;
  if (ok)
    for (int c = sl * 4; c < C; c += lanes_per_row * 4) {
      const float4 v = *reinterpret_cast<const float4*>(xr + c);
      float4 a;
      a.x = __fdiv_rn(v.x, den); a.y = __fdiv_rn(v.y, den); a.z = __fdiv_rn(v.z, den); a.w = __fdiv_rn(v.w, den);
      t = fmaf(a.x, a.x, t); t = fmaf(a.y, a.y, t); t = fmaf(a.z, a.z, t); t = fmaf(a.w, a.w, t);
      const float s0 = a.x * KB_PRESCALE, s1 = a.y * KB_PRESCALE, s2 = a.z * KB_PRESCALE, s3 = a.w * KB_PRESCALE;
      const __half2 h01 = __floats2half2_rn(s0, s1), h23 = __floats2half2_rn(s2, s3);
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(s0 - f01.x, s1 - f01.y), l23 = __floats2half2_rn(s2 - f23.x, s3 - f23.y);
      uint2 hv, lv;
      hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
      lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(hi + row * C + c) = hv;
      *reinterpret_cast<uint2*>(lo + row * C + c) = lv;
    }
  for (int o = lanes_per_row >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (ok && sl == 0) sq[row] = normalize ? t : s;
}

// one warp per graph: (min, max) of its squared norms
__global__ void __launch_bounds__(256)
knn_big_minmax_kernel(const float* __restrict__ sq, int B, int N, float2* __restrict__ out) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (g >= B) return;
  float lo = INFINITY, hi = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    const float q = sq[(size_t)g * N + j];
    lo = fminf(lo, q); hi = fmaxf(hi, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) out[g] = make_float2(lo, hi);
}

int knn_big_supported(int B, int N, int C, int kk) {
  (void)B;
  return N >= KB_BN && N <= 2048 && N % KB_BN == 0 && C % KB_KBK == 0 && kk >= 1 && kk <= 64 && kk <= N;
}

// sq[M] | pad | (min, max)[M / 256 at most] | hi plane (M, C) | lo plane (M, C) | per-CTA candidate scratch
static size_t knn_big_sq_bytes(int64_t M) {
  return (((size_t)(M + 128) * sizeof(float) + 255) & ~(size_t)255) + (((size_t)(M / KB_BN + 1) * sizeof(float2) + 255) & ~(size_t)255);
}
static size_t knn_big_plane_bytes(int64_t M, int C) { return ((size_t)M * C * sizeof(__half) + 255) & ~(size_t)255; }

size_t knn_big_workspace_bytes(int B, int N, int C) {
  const int64_t M = (int64_t)B * N;
  return knn_big_sq_bytes(M) + 2 * knn_big_plane_bytes(M, C) +
         (size_t)sm_count() * 2 * TC_BM * 2 * (size_t)(N / 2) * sizeof(float2);
}

// fp16 (planes, rows, cols), element strides (plane_stride, ld, 1); box = (32 cols = 64 B, box_rows, 1), 64B swizzle
static int knn_big_make_map(CUtensorMap* map, const __half* base, int64_t rows, int64_t cols, int64_t plane_stride,
                            int box_rows) {
  EncodeTiledFn fn = tc_encode_fn();
  GRAFP_REQUIRE(fn, "tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)KB_KBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled(knn_big planes) failed (%d)", (int)r);
  return 0;
}

template <int KH>
static int knn_big_launch_t(const CUtensorMap& mr, const CUtensorMap& mc, KnnBigParams p, int grid, cudaStream_t st) {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, knn_big_kernel<KH>);
  int stages = (int)((227 * 1024 - fa.sharedSizeBytes - 2048 - KB_STAGING_BYTES) / KB_STAGE_BYTES);
  if (stages > KB_STAGES_MAX) stages = KB_STAGES_MAX;
  if (stages < 2) return fail("knn_big: not enough shared memory for two operand stages");
  p.stages = stages;
  const size_t smem = (size_t)KB_STAGE_BYTES * stages + KB_STAGING_BYTES + 1024;
  cudaFuncSetAttribute(knn_big_kernel<KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  knn_big_kernel<KH><<<grid, KB_THREADS, smem, st>>>(mr, mc, p);
  return check_launch("knn_big");
}

int knn_big_launch(const float* x, int B, int N, int C, int kk, int d, int k, int normalize, int32_t* idx,
                   float* dist, void* workspace, cudaStream_t st) {
  const int64_t M = (int64_t)B * N;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* sq = reinterpret_cast<float*>(ws);
  __half* hi = reinterpret_cast<__half*>(ws + knn_big_sq_bytes(M));
  __half* lo = reinterpret_cast<__half*>(ws + knn_big_sq_bytes(M) + knn_big_plane_bytes(M, C));
  int lpr = 32;
  while (lpr > 1 && lpr * 4 > C) lpr >>= 1;
  const int rows_per_block = 8 * (32 / lpr);
  knn_big_prep_kernel<<<(unsigned)((M + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(x, M, C, normalize, lpr, sq,
                                                                                            hi, lo);
  if (int rc = check_launch("knn_big_prep")) return rc;
  float2* mm = reinterpret_cast<float2*>(ws + (((size_t)(M + 128) * sizeof(float) + 255) & ~(size_t)255));
  knn_big_minmax_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(sq, B, N, mm);
  if (int rc = check_launch("knn_big_minmax")) return rc;
  KnnBigParams p;
  p.sq_minmax = mm;
  p.N = N; p.C = C; p.kk = kk; p.d = d; p.k = k; p.M = M;
  p.sq = sq; p.idx = idx; p.dist = dist;
  p.scratch = reinterpret_cast<float2*>(ws + knn_big_sq_bytes(M) + 2 * knn_big_plane_bytes(M, C));
  // columns per group of sweep 1: the largest of 16 / 8 / 4 that still feeds each of a thread's two threshold lists
  // (N/2 columns per thread -> N / (4 g) group bounds per list) four times the ceil(kk/4) entries it keeps
  const int hq = (kk + 3) / 4;
  p.gsel = (N / 64 >= 4 * hq) ? 16 : (N / 32 >= 4 * hq) ? 8 : 4;
  {
    static int g_env = -1;
    if (g_env < 0) { const char* e = getenv("GRAFP_KNN_BIG_G"); g_env = e ? atoi(e) : 0; }
    if (g_env == 4 || g_env == 8 || g_env == 16) p.gsel = g_env;
  }
  CUtensorMap mr, mc;
  const int64_t plane_stride = (int64_t)(knn_big_plane_bytes(M, C) / sizeof(__half));
  if (int rc = knn_big_make_map(&mr, hi, M, C, plane_stride, TC_BM)) return rc;
  if (int rc = knn_big_make_map(&mc, hi, M, C, plane_stride, KB_BN)) return rc;
  const int64_t units = M / TC_BM;
  int grid = sm_count();
  if (units < grid) grid = (int)units;
  if (kk <= 16) return knn_big_launch_t<4>(mr, mc, p, grid, st);
  if (kk <= 32) return knn_big_launch_t<8>(mr, mc, p, grid, st);
  return knn_big_launch_t<16>(mr, mc, p, grid, st);
}

}  // namespace grafp
