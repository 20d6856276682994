// Dense dilated kNN graph on the tensor cores for LARGE graphs / long lists (sm_100a):
// N in {256, 512, ..., 2048} nodes per graph, k * dilation <= 64 (BASELINE configs[4]: k in {9, 16, 32} x
// N in {256, 512, 1024, 2048}, dilation on).  knn_tc.cu covers N <= 256 with k * dilation <= 16.
//
// Work unit = 128 consecutive rows (nodes) of one graph; its N columns are swept in tiles of 256:
//   TMA        per k-block of 16 channels: the unit's 128 rows and the tile's 256 column nodes (fp32, 64B swizzle)
//   transform  x / max(||x||, 1e-12) (F.normalize as a correctly rounded division), tf32 hi / lo split, in place
//   MMA        D[128 x 256] = Xn_rows * Xn_cols^T, 3 kind::tf32 passes, fp32 accumulate in TMEM (2 x 256 columns:
//              the tile i+1 contraction overlaps the tile i selection)
//   select     thread-per-row (TMEM lane = node); the two epilogue groups take the two 128-column halves of every
//              tile.  The N x N matrix never leaves TMEM.
//
// Selection (exact).  A row's k*d best are found by threshold filtering, as in knn_tc.cu, but the threshold is carried
// ACROSS the column tiles and the candidate lists live in an L2-resident per-CTA scratch (global memory) because they
// can be long:
//   pass 1  minimum of every group of g in {4, 8, 16} of the thread's 128 columns; the minima are pushed, two at a time
//           as a packed bf16x2 (rounded UP, so the bound stays valid), into a sorted pair-list of ceil(kk/2) entries:
//           tau = max of the two lists' ceil(kk/2)-th smallest >= the thread's kk-th smallest distance so far
//   pass 2  distances are recomputed bit-identically, dist = (sq_i + (-2 dot)) + sq_j (torch_edge.py:16-18), and every
//           dist <= tau is appended (predicated store) to the thread's list: a superset of the row's kk best among this
//           thread's columns, because tau only decreases from tile to tile
//   final   at the end of the unit one thread per row inserts both halves' candidates into an exact (distance, index)
//           select network, 32 ranks per round, and emits ranks 0, d, 2d, ...
// A list can hold all of the thread's columns, so nothing overflows (an all-ties input degrades to a full sort).
#include <stdlib.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace grafp {

constexpr int KB_THREADS = 576;       // warp 0 TMA, warp 1 MMA + TMEM allocator, warps 2-9 epilogue, 10-17 transform
constexpr int KB_STAGES_MAX = 4;
constexpr int KB_XF_THREADS = 256;
constexpr int KB_EPI_THREADS = 256;
constexpr int KB_BN = 256;            // columns per tile
constexpr int KB_KBK = 16;            // channels per k-block (64-byte operand rows, 64B swizzle)
constexpr uint32_t KB_A_BYTES = TC_BM * KB_KBK * 4;    // 8 KB
constexpr uint32_t KB_B_BYTES = KB_BN * KB_KBK * 4;    // 16 KB
constexpr uint32_t KB_STAGE_BYTES = 2 * (KB_A_BYTES + KB_B_BYTES);   // [A_hi | A_lo | B_hi | B_lo] = 48 KB

struct KnnBigParams {
  int N, C, kk, d, k;
  int64_t M;
  int stages;
  int gsel;               // columns per minimum group: 4, 8 or 16
  const float* den;       // (M) F.normalize denominators (1 when not normalising)
  const float* sq;        // (M) squared norms of the normalised rows
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

// Bitonic sort of the 32 * E candidate keys of one row held E per lane (element q * 32 + lane in register q);
// the two half lists are concatenated, the padding sorts last.  Ranks 0, d, 2d, ... are written out.
template <int E>
__device__ __forceinline__ void sort_row(const float2* l0, int c0, const float2* l1, int c1, int lane,
                                         const KnnBigParams& p, int64_t row, int self) {
  uint64_t key[E];
#pragma unroll
  for (int q = 0; q < E; ++q) {
    const int i = q * 32 + lane;
    key[q] = KNN_KEY_PAD;
    if (i < c0) key[q] = knn_key(__ldcg(l0 + i));
    else if (i < c0 + c1) key[q] = knn_key(__ldcg(l1 + (i - c0)));
  }
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dq = j >> 5;
#pragma unroll
        for (int q = 0; q < E; ++q) {
          if ((q & dq) == 0) {
            const bool asc = (((q * 32) & k) == 0);              // bit k of the element index lives in q here
            const uint64_t a = key[q], b = key[q | dq];
            const bool sw = asc ? (a > b) : (a < b);
            key[q] = sw ? b : a;
            key[q | dq] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < E; ++q) {
          const uint64_t mine = key[q];
          const uint64_t other = __shfl_xor_sync(0xffffffffu, mine, j);
          const bool asc = (((q * 32 + lane) & k) == 0);
          const bool lower = (lane & j) == 0;
          const bool take_min = (asc == lower);
          const uint64_t mn = mine < other ? mine : other, mx = mine < other ? other : mine;
          key[q] = take_min ? mn : mx;
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < E; ++q) knn_emit(p, row, self, q * 32 + lane, key[q]);
}

// more than 512 candidates (mass ties, e.g. an all-zero input): kk rounds of warp-wide minimum extraction
// straight from the lists
__device__ __noinline__ void extract_row(const float2* l0, int c0, const float2* l1, int c1, int lane,
                                         const KnnBigParams& p, int64_t row, int self) {
  uint64_t prev = 0;
  bool first = true;
  for (int rank = 0; rank < p.kk; ++rank) {
    uint64_t best = KNN_KEY_PAD;
    for (int i = lane; i < c0 + c1; i += 32) {
      const uint64_t k = knn_key(i < c0 ? __ldcg(l0 + i) : __ldcg(l1 + (i - c0)));
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

template <int KH, int KMAX>
__global__ void __launch_bounds__(KB_THREADS, 1)
knn_big_kernel(const __grid_constant__ CUtensorMap tmRows, const __grid_constant__ CUtensorMap tmCols,
               const KnnBigParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[KB_STAGES_MAX];
  __shared__ __align__(8) uint64_t xf_bar[KB_STAGES_MAX];
  __shared__ __align__(8) uint64_t empty_bar[KB_STAGES_MAX];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_sq[2][KB_BN];      // squared norms of the tile's column nodes, per TMEM buffer
  __shared__ int s_cnt[2][2][TC_BM];                  // [unit parity][half][row]: candidates appended
  __shared__ float s_tau[2][TC_BM];                   // [half][row]: the two threads of a row exchange their bounds

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto a_hi = [&](int s) { return smem + (size_t)s * KB_STAGE_BYTES; };
  auto a_lo = [&](int s) { return smem + (size_t)s * KB_STAGE_BYTES + KB_A_BYTES; };
  auto b_hi = [&](int s) { return smem + (size_t)s * KB_STAGE_BYTES + 2 * KB_A_BYTES; };
  auto b_lo = [&](int s) { return smem + (size_t)s * KB_STAGE_BYTES + 2 * KB_A_BYTES + KB_B_BYTES; };

  const int nkb = p.C / KB_KBK;
  const int T = p.N / KB_BN;                               // column tiles per unit
  // Every tile is contracted twice per unit when there is more than one: sweep 1 only derives the row's threshold,
  // sweep 2 collects the candidates under it (the MMA is cheap next to the selection; a single tile stays in TMEM
  // for both passes)
  const int SU = T == 1 ? 1 : 2 * T;                       // MMA steps per unit
  const int64_t units = p.M / TC_BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmRows);
    tma_prefetch_desc(&tmCols);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xf_bar[s], KB_XF_THREADS);
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
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int64_t m0 = u * TC_BM, gs = (m0 / p.N) * p.N;
        for (int su = 0; su < SU; ++su) {
          const int t = su % T;
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1u;
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full_bar[s], KB_A_BYTES + KB_B_BYTES);
            tma_load_2d(a_hi(s), &tmRows, kb * KB_KBK, (int)m0, &full_bar[s]);
            tma_load_2d(b_hi(s), &tmCols, kb * KB_KBK, (int)(gs + (int64_t)t * KB_BN), &full_bar[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, KB_BN);
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
            mbar_wait(&xf_bar[s], ph);
            tc_fence_after();
            const uint64_t dah = umma_desc_sw64(smem_u32(a_hi(s))), dal = umma_desc_sw64(smem_u32(a_lo(s)));
            const uint64_t dbh = umma_desc_sw64(smem_u32(b_hi(s))), dbl = umma_desc_sw64(smem_u32(b_lo(s)));
#pragma unroll
            for (int ks = 0; ks < KB_KBK / 8; ++ks) {
              const uint64_t koff = (uint64_t)((ks * 8 * 4) >> 4);
              umma_tf32(tacc, dal + koff, dbh + koff, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
              umma_tf32(tacc, dah + koff, dbl + koff, idesc, 1u);
              umma_tf32(tacc, dah + koff, dbh + koff, idesc, 1u);
            }
            umma_commit(&empty_bar[s]);
          }
          umma_commit(&tmem_full_bar[buf]);
        }
      }
    }
  } else if (warp >= 10) {
    // ===== transform (256 threads): normalise (exact division), tf32 hi / lo split, in place =====
    // A stage holds 384 operand rows of 4 float4 each: physical float4 q belongs to row q >> 2 (the 64B swizzle only
    // permutes the chunks inside a row).  Thread t owns float4 t + 256 i: A rows (t >> 2) + 64 i (i = 0, 1), B rows
    // (t >> 2) + 64 (i - 2) (i = 2..5) in every k-block.
    const int t = threadIdx.x - 320;
    const int rq = t >> 2;
    uint32_t it = 0;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
      const int64_t m0 = u * TC_BM, gs = (m0 / p.N) * p.N;
      float dn[6], ri[6];
      dn[0] = __ldg(p.den + m0 + rq);
      dn[1] = __ldg(p.den + m0 + rq + 64);
      for (int su = 0; su < SU; ++su) {
        const int64_t c0 = gs + (int64_t)(su % T) * KB_BN;
#pragma unroll
        for (int i = 2; i < 6; ++i) dn[i] = __ldg(p.den + c0 + rq + 64 * (i - 2));
#pragma unroll
        for (int i = 0; i < 6; ++i) ri[i] = __frcp_rn(dn[i]);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          mbar_wait(&full_bar[s], ph);
          float4* hi = reinterpret_cast<float4*>(a_hi(s));            // A_hi then (after A_lo) B_hi: handled per part
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            // part offsets in float4 units: A_hi [0, 512), B_hi [1024, 2048); lo = hi + part size
            float4* src = (i < 2) ? hi + (t + 256 * i) : hi + 1024 + (t + 256 * (i - 2));
            float4* dst = (i < 2) ? src + 512 : src + 1024;
            const float4 v = *src;
            const float rr = ri[i], dd = dn[i];
            float x0 = v.x * rr, x1 = v.y * rr, x2 = v.z * rr, x3 = v.w * rr;
            x0 = fmaf(fmaf(-x0, dd, v.x), rr, x0);
            x1 = fmaf(fmaf(-x1, dd, v.y), rr, x1);
            x2 = fmaf(fmaf(-x2, dd, v.z), rr, x2);
            x3 = fmaf(fmaf(-x3, dd, v.w), rr, x3);
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
            h.y = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
            h.z = __uint_as_float(__float_as_uint(x2) & 0xFFFFE000u);
            h.w = __uint_as_float(__float_as_uint(x3) & 0xFFFFE000u);
            l.x = x0 - h.x; l.y = x1 - h.y; l.z = x2 - h.z; l.w = x3 - h.w;
            *src = h;
            *dst = l;
          }
          fence_proxy_async_smem();
          mbar_arrive(&xf_bar[s]);
        }
      }
    }
  } else {
    // ===== epilogue (warps 2..9): thread-per-row threshold selection, half `grp` of every tile's columns =====
    const int ew = warp - 2;
    const int grp = ew >> 2;
    const int quad = warp & 3;                     // the TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;
    const int e = grp * 128 + r;                   // my slot when staging the column norms
    const int halfN = p.N >> 1;
    const int hq = (p.kk + 3) >> 2;                // entries per threshold list: four lists per row (two per thread)
    float2* cta_scratch = p.scratch + (size_t)blockIdx.x * 2u * TC_BM * 2u * (size_t)halfN;
    uint32_t st = 0, un = 0;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++un) {
      const int64_t m0 = u * TC_BM, gs = (m0 / p.N) * p.N;
      const int64_t grow = m0 + r;
      const float sqi = __ldg(p.sq + grow);
      const uint32_t par = un & 1u;
      float2* my_list = cta_scratch + (((size_t)par * TC_BM + r) * 2u + grp) * (size_t)halfN;
      float2* wp = my_list;
      __nv_bfloat162 tb[KH];
      {
        const uint32_t inf2 = 0x7F807F80u;
#pragma unroll
        for (int i = 0; i < KH; ++i) tb[i] = *reinterpret_cast<const __nv_bfloat162*>(&inf2);
      }
      float v[32];
      float tau = INFINITY;
      for (int su = 0; su < 2 * T; ++su) {
        const int t = su % T;
        const bool sweep2 = su >= T;
        const bool fresh = !(T == 1 && sweep2);      // a single tile is kept in TMEM for both passes
        const uint32_t buf = st & 1u, tph = (st >> 1) & 1u;
        if (fresh) {
          s_sq[buf][e] = __ldg(p.sq + gs + (int64_t)t * KB_BN + e);
          named_bar_sync(1, KB_EPI_THREADS);
          mbar_wait(&tmem_full_bar[buf], tph);
          tc_fence_after();
        }
        const float4* sqv = reinterpret_cast<const float4*>(&s_sq[buf][grp * 128]);
        const uint32_t tacc = tmem_base + buf * (uint32_t)KB_BN + (uint32_t)(grp * 128) + ((uint32_t)(quad * 32) << 16);
        auto load_dist = [&](int c) {
          tmem_ld16_nowait(tacc + (uint32_t)c, v);
          tmem_ld16_nowait(tacc + (uint32_t)c + 16u, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int q4 = 0; q4 < 32; q4 += 4) {
            const float4 s4 = sqv[(c + q4) >> 2];
            v[q4 + 0] = __fadd_rn(fmaf(v[q4 + 0], -2.0f, sqi), s4.x);
            v[q4 + 1] = __fadd_rn(fmaf(v[q4 + 1], -2.0f, sqi), s4.y);
            v[q4 + 2] = __fadd_rn(fmaf(v[q4 + 2], -2.0f, sqi), s4.z);
            v[q4 + 3] = __fadd_rn(fmaf(v[q4 + 3], -2.0f, sqi), s4.w);
          }
        };
        if (!sweep2) {
          // ---- sweep 1: group minima -> threshold lists ----
          auto push2 = [&](float a, float b) {
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
            load_dist(c);
            float m4[8];
#pragma unroll
            for (int b = 0; b < 8; ++b)
              m4[b] = fminf(fminf(v[4 * b], v[4 * b + 1]), fminf(v[4 * b + 2], v[4 * b + 3]));
            if (p.gsel == 4) {
              push2(m4[0], m4[1]); push2(m4[2], m4[3]); push2(m4[4], m4[5]); push2(m4[6], m4[7]);
            } else if (p.gsel == 8) {
              push2(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
              push2(fminf(m4[4], m4[5]), fminf(m4[6], m4[7]));
            } else {
              push2(fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3])), fminf(fminf(m4[4], m4[5]), fminf(m4[6], m4[7])));
            }
          }
          if (t == T - 1) {
            // the row's threshold: each of its four lists (two per thread, two threads per row) holds the minima of
            // disjoint column groups, so the max of their ceil(kk/4)-th smallest entries has >= kk distances under it
            float mine = INFINITY;
#pragma unroll
            for (int i = 0; i < KH; ++i)
              if (i == hq - 1) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(&tb[i]);
                mine = fmaxf(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
              }
            s_tau[grp][r] = mine;
            named_bar_sync(1, KB_EPI_THREADS);
            tau = fmaxf(mine, s_tau[grp ^ 1][r]);
          }
        } else {
          // ---- sweep 2: append every distance <= tau to my candidate list (global scratch, L2 resident) ----
          const int jl0 = t * KB_BN + grp * 128;
          for (int c = 0; c < 128; c += 32) {
            load_dist(c);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              uint32_t bump;
              asm volatile(
                  "{\n\t.reg .pred p;\n\t"
                  "setp.le.f32 p, %2, %3;\n\t"
                  "@p st.global.v2.b32 [%1], {%4, %5};\n\t"
                  "selp.u32 %0, 1, 0, p;\n\t}"
                  : "=r"(bump)
                  : "l"(wp), "f"(v[q]), "f"(tau), "r"(__float_as_uint(v[q])), "r"(jl0 + c + q)
                  : "memory");
              wp += bump;
            }
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
      for (int rr = ew * 16; rr < ew * 16 + 16; ++rr) {
        const int c0 = s_cnt[par][0][rr], c1 = s_cnt[par][1][rr];
        const float2* l0 = cta_scratch + (((size_t)par * TC_BM + rr) * 2u) * (size_t)halfN;
        const float2* l1 = l0 + halfN;
        const int64_t orow = m0 + rr;
        const int oself = (int)(orow - gs);
        const int c = c0 + c1;
        if (c <= 32) sort_row<1>(l0, c0, l1, c1, lane, p, orow, oself);
        else if (c <= 64) sort_row<2>(l0, c0, l1, c1, lane, p, orow, oself);
        else if (c <= 128) sort_row<4>(l0, c0, l1, c1, lane, p, orow, oself);
        else if (c <= 256) sort_row<8>(l0, c0, l1, c1, lane, p, orow, oself);
        else if (c <= 512) sort_row<16>(l0, c0, l1, c1, lane, p, orow, oself);
        else extract_row(l0, c0, l1, c1, lane, p, orow, oself);
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

__global__ void __launch_bounds__(256)
knn_big_rownorm_kernel(const float* __restrict__ x, int64_t M, int C, int normalize, int lanes_per_row,
                       float* __restrict__ den_out, float* __restrict__ sq) {
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
  float den = 1.0f, q = s;
  if (normalize) {
    den = fmaxf(sqrtf(s), 1e-12f);
    float t = 0.0f;
    if (ok)
      for (int c = sl * 4; c < C; c += lanes_per_row * 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float a0 = __fdiv_rn(v.x, den), a1 = __fdiv_rn(v.y, den), a2 = __fdiv_rn(v.z, den), a3 = __fdiv_rn(v.w, den);
        t = fmaf(a0, a0, t); t = fmaf(a1, a1, t); t = fmaf(a2, a2, t); t = fmaf(a3, a3, t);
      }
    for (int o = lanes_per_row >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    q = t;
  }
  if (ok && sl == 0) { den_out[row] = den; sq[row] = q; }
}

int knn_big_supported(int B, int N, int C, int kk) {
  (void)B;
  return N >= KB_BN && N <= 2048 && N % KB_BN == 0 && C % KB_KBK == 0 && kk >= 1 && kk <= 64 && kk <= N;
}

static size_t knn_big_norm_floats(int64_t M) { return (size_t)(M + 128) * 2; }

// den[M] | pad | sq[M] | pad | per-CTA candidate scratch
size_t knn_big_workspace_bytes(int B, int N) {
  const int64_t M = (int64_t)B * N;
  const size_t norms = (knn_big_norm_floats(M) * sizeof(float) + 255) & ~(size_t)255;
  return norms + (size_t)sm_count() * 2 * TC_BM * 2 * (size_t)(N / 2) * sizeof(float2);
}

template <int KH, int KMAX>
static int knn_big_launch_t(const CUtensorMap& mr, const CUtensorMap& mc, KnnBigParams p, int grid, cudaStream_t st) {
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, knn_big_kernel<KH, KMAX>);
  int stages = (int)((227 * 1024 - fa.sharedSizeBytes - 2048) / KB_STAGE_BYTES);
  if (stages > KB_STAGES_MAX) stages = KB_STAGES_MAX;
  if (stages < 2) return fail("knn_big: not enough shared memory for two operand stages");
  p.stages = stages;
  const size_t smem = (size_t)KB_STAGE_BYTES * stages + 1024;
  cudaFuncSetAttribute(knn_big_kernel<KH, KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  knn_big_kernel<KH, KMAX><<<grid, KB_THREADS, smem, st>>>(mr, mc, p);
  return check_launch("knn_big");
}

int knn_big_launch(const float* x, int B, int N, int C, int kk, int d, int k, int normalize, int32_t* idx,
                   float* dist, void* workspace, cudaStream_t st) {
  const int64_t M = (int64_t)B * N;
  float* den = static_cast<float*>(workspace);
  float* sq = den + M + 128;
  const size_t norms = (knn_big_norm_floats(M) * sizeof(float) + 255) & ~(size_t)255;
  int lpr = 32;
  while (lpr > 1 && lpr * 4 > C) lpr >>= 1;
  const int rows_per_block = 8 * (32 / lpr);
  knn_big_rownorm_kernel<<<(unsigned)((M + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(x, M, C, normalize, lpr,
                                                                                               den, sq);
  if (int rc = check_launch("knn_big_rownorm")) return rc;
  KnnBigParams p;
  p.N = N; p.C = C; p.kk = kk; p.d = d; p.k = k; p.M = M;
  p.den = den; p.sq = sq; p.idx = idx; p.dist = dist;
  p.scratch = reinterpret_cast<float2*>(static_cast<uint8_t*>(workspace) + norms);
  // columns per minimum group: the largest of 16 / 8 / 4 that still feeds each of a thread's two threshold lists
  // (N/2 columns per thread -> N / (4 g) minima per list) four times the ceil(kk/4) entries it keeps
  const int hq = (kk + 3) / 4;
  p.gsel = (N / 64 >= 4 * hq) ? 16 : (N / 32 >= 4 * hq) ? 8 : 4;
  {
    static int g_env = -1;
    if (g_env < 0) { const char* e = getenv("GRAFP_KNN_BIG_G"); g_env = e ? atoi(e) : 0; }
    if (g_env == 4 || g_env == 8 || g_env == 16) p.gsel = g_env;
  }
  CUtensorMap mr, mc;
  if (int rc = tc_make_map_2d_bk16(&mr, x, M, C, C, TC_BM)) return rc;
  if (int rc = tc_make_map_2d_bk16(&mc, x, M, C, C, KB_BN)) return rc;
  const int64_t units = M / TC_BM;
  int grid = sm_count();
  if (units < grid) grid = (int)units;
  if (kk <= 16) return knn_big_launch_t<4, 0>(mr, mc, p, grid, st);
  if (kk <= 32) return knn_big_launch_t<8, 0>(mr, mc, p, grid, st);
  return knn_big_launch_t<16, 0>(mr, mc, p, grid, st);
}

}  // namespace grafp
