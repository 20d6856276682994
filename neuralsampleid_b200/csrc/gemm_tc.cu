// tcgen05 GEMM engine (sm_100a): persistent, TMA-fed, TMEM-accumulated, warp-specialised.
//
//   y[128 x BN tile] = epilogue( A[128 x K] * W[BN x K]^T )      A, W fp32 row-major (K-major)
//
// Precision modes
//   passes = 1 : one kind::tf32 MMA per k-step on the raw fp32 operands (TF32 accuracy)
//   passes = 3 : error-compensated "3xTF32": A and W are split hi = tf32(v), lo = tf32(v - hi);
//                D += A_lo*W_hi + A_hi*W_lo + A_hi*W_hi  -> fp32-class accuracy.
//                W is pre-split once per weight version (stacked [hi; lo] matrix,
//                grafp_split_tf32); A is split in shared memory by the transform warps between
//                the TMA arrival and the MMA issue, so activations cross HBM once, as plain fp32.
//
// One persistent CTA per SM walks the tile list with n fastest: the CTAs running at the same time
// share the same 128 rows of A (read from HBM once, L2 hits for the other column tiles) while the
// weights, a few MB, stay L2-resident.  Warp roles (480 threads):
//   warp 0      TMA producer (A from one of two sources or the 3-tap Downsample view, W hi/lo)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-9   epilogue, two groups of four warps (one per TMEM lane quadrant): group h takes the
//               32-column chunks of parity h: TMEM -> registers -> scale/shift/activation/residual
//               -> its own 128B-swizzled staging tile -> TMA store (coalesced, clipped at the edge).
//               One warp per scheduler could not hide the TMEM / L2 / store latencies of a chunk
//               (ncu: 17 % issue utilisation, 2.1 k cycles per chunk); two groups interleave them.
//   warps 10-13 transform: hi / lo split of the A stage (tf32 in place, or fp32 -> bf16 tiles)
//   warp 14     second TMA producer (weight tiles) of the bf16 engines
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.  Pipelines: full[s] (TMA -> transform), xf[s] (transform -> MMA),
// empty[s] (MMA commit -> TMA), tmem_full[b] (MMA commit -> epilogue), tmem_empty[b].
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace grafp {

constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 480;
constexpr int TC_RAW_MAX = 8;                   // max depth of the fp32 A staging ring (bf16 engines)
constexpr int TC_XF_THREADS = 128;
constexpr int TC_EPI_THREADS = 256;
constexpr int TC_STORE_BYTES = TC_BM * 32 * 4;   // one 128 x 32 fp32 staging tile

struct TcParams {
  int k1, k2;          // per-group k extents of the two A sources (multiples of 32)
  int n;               // per-group output columns
  int bn;              // tile width (multiple of 32, divides n)
  int n_total;         // groups * n (row offset of the lo copy inside the stacked W)
  int groups;
  int64_t m;
  int stages;
  int tap3_rows;       // > 0: Downsample form, output rows per graph (= input nodes / 2)
  int tap3_cin;
  const float* scale; const float* shift;
  const float* residual; int64_t ldr;
  float* row_sumsq;
  int act; float act_param;
  float unscale;       // multiplies the epilogue scale: undoes the host's power-of-two weight pre-scaling (f16x3)
  uint32_t tmem_cols;
  int y_split;         // 0: fp32 output; 2: bf16 [hi ; lo] planes; 1: bf16 hi plane only (1-pass engine)
  int res_tma;         // the shortcut tile comes by TMA into the fp32 staging tile (coalesced) instead of per-thread row loads
  int y_both;          // with y_split: ALSO write the fp32 output (tmY fp32 map + tmYs split map, two staging tiles)
  int raw;             // depth of the fp32 A ring (bf16 engines with an fp32 A operand)
  // fused max-relative aggregation: the second A source is not read but computed by the transform warps,
  // a2[m, c] = max_t (a1[graph(m) + idx[m, t], c] - a1[m, c])  (bf16 engines, fp32 A)
  const int32_t* gat_idx; const float* gat_x; int64_t gat_ld; int gat_n, gat_k;
};

// v = act(v * scale + shift) + residual over one 32-column chunk of a row; scale/shift come from
// shared memory as 128-bit broadcast reads (every thread of the warp reads the same 32 columns), the
// residual through `res(q)`: the staged TMA tile (shared memory) or the row in global memory
// The first 16 columns' scale / shift (sc0, sh0) were fetched BEFORE the TMEM load: under the MMA's operand traffic a
// shared-memory read takes > 100 cycles and the profile showed this loop waiting on the short scoreboard (the whole
// scale / shift / activation step cost 1 300-2 000 cycles per chunk); the second half is fetched while the first computes.
template <int ACT>
__device__ __forceinline__ void epi_apply(float (&v)[32], const float4 (&sc0)[4], const float4 (&sh0)[4], const float* scale,
                                          const float* shift, float act_param) {
  float4 sc1[4], sh1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sc1[i] = *reinterpret_cast<const float4*>(scale + 16 + 4 * i);
    sh1[i] = *reinterpret_cast<const float4*>(shift + 16 + 4 * i);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = 4 * i;
    v[q + 0] = apply_act(fmaf(v[q + 0], sc0[i].x, sh0[i].x), ACT, act_param);
    v[q + 1] = apply_act(fmaf(v[q + 1], sc0[i].y, sh0[i].y), ACT, act_param);
    v[q + 2] = apply_act(fmaf(v[q + 2], sc0[i].z, sh0[i].z), ACT, act_param);
    v[q + 3] = apply_act(fmaf(v[q + 3], sc0[i].w, sh0[i].w), ACT, act_param);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = 16 + 4 * i;
    v[q + 0] = apply_act(fmaf(v[q + 0], sc1[i].x, sh1[i].x), ACT, act_param);
    v[q + 1] = apply_act(fmaf(v[q + 1], sc1[i].y, sh1[i].y), ACT, act_param);
    v[q + 2] = apply_act(fmaf(v[q + 2], sc1[i].z, sh1[i].z), ACT, act_param);
    v[q + 3] = apply_act(fmaf(v[q + 3], sc1[i].w, sh1[i].w), ACT, act_param);
  }
}
// v += shortcut row piece (32 floats: the staged TMA tile in shared memory, 128B-swizzled, or global memory)
__device__ __forceinline__ void epi_add_smem(float (&v)[32], const uint8_t* row_base, int r) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(row_base + ((q ^ (r & 7)) << 4));
    v[4 * q + 0] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    if ((q & 3) == 3) asm volatile("" ::: "memory");      // at most four loads (16 registers) in flight
  }
}
__device__ __forceinline__ void epi_add_global(float (&v)[32], const float* row) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *reinterpret_cast<const float4*>(row + 4 * q);
    v[4 * q + 0] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
    if (q & 1) asm volatile("" ::: "memory");      // two loads in flight, not eight: this rare path must not cost the common ones registers
  }
}


// 16-bit operand packing of the split engines: element 0 in the low half.  kF16: IEEE half with a
// saturating conversion (a value beyond 65504 becomes hi = 65504, lo = the rest, instead of inf);
// else bfloat16.
template <bool kF16>
__device__ __forceinline__ uint32_t pack16(float a, float b) {
  if (kF16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
template <bool kF16>
__device__ __forceinline__ float2 unpack16(uint32_t v) {
  if (kF16) return __half22float2(*reinterpret_cast<const __half2*>(&v));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

// kCluster == 2: CTA pairs (thread-block cluster 2x1) work on two vertically adjacent 128-row tiles
// of the same column tile; each CTA fetches HALF of the weight tile and TMA-multicasts it into both
// CTAs' shared memory, halving the per-SM weight traffic out of L2 (the limiter of the 1-CTA form).
// A stage may be refilled only when BOTH CTAs have consumed it: every MMA commit is multicast to
// the pair's `empty` barriers (count 2).
//
// kBf16: the MMA operands are bf16 (kind::f16, twice the tf32 rate, half the operand bytes):
//   passes = 3 : "bf16x3"  a = a1 + a2 (+ 2^-18 |a|), w = w1 + w2;  D += a2*w1 + a1*w2 + a1*w1
//                (per-product error <= 3 * 2^-18; the fp32-parity engine under the power cap)
//   passes = 1 : plain bf16 operands, fp32 accumulate (reduced precision, stated separately)
// A still arrives from HBM as fp32 (TMA, 128B swizzle); the transform warps convert it into
// 64-byte-row bf16 tiles (64B swizzle); W is pre-split on the host into stacked bf16 [w1 ; w2].
//
// kASplit (bf16 engines): the A operand arrives already split, as the bf16 (2, M, K) [hi ; lo] planes a
// previous GEMM's epilogue wrote (p.y_split): warp 0 TMA-loads the hi / lo tiles straight into a ring of A operand
// tiles (as many as shared memory holds beside three W stages: the A stream comes from HBM, W from L2), there is no
// fp32 ring and no transform, the MMA waits on raw_full[sa] and full[s] and frees both with its commits.
// kGather: fused max-relative aggregation of the second A source (opt-in, see TcParams::gat_idx); a separate
// instantiation so the default kernels carry none of its registers or code.
// kF16 (with kBf16, which selects the 16-bit kind::f16 operand path): the operand pair is IEEE fp16 instead of
// bfloat16 -- "f16x3": hi carries 11 significant bits, lo the next 11, so a = a1 + a2 to 2^-24 |a| (an absolute
// floor of 2^-25 where lo is subnormal) and the dropped a2*w2 term is 2^-24: fp32-class products at the bf16 MMA
// rate.  The weights are pre-scaled by a power of two on the host (p.unscale undoes it in the epilogue) so that
// their lo parts stay normal numbers.
#ifdef TC_TRACE
// debug build (GRAFP_NVCC_EXTRA=-DTC_TRACE): where block 0's MMA warp spends its cycles (scripts/gemm_trace.py)
__device__ unsigned long long g_tc_trace[24];
#define TC_ACC(var, stmt) do { const long long c0_ = clock64(); stmt; var += clock64() - c0_; } while (0)
#define TC_CLK(name) const long long name = clock64()
#define TC_SINCE(var, name) var += clock64() - name
#else
#define TC_ACC(var, stmt) do { stmt; } while (0)
#define TC_CLK(name) do { } while (0)
#define TC_SINCE(var, name) do { } while (0)
#endif

template <int kPasses, int kCluster, bool kBf16, bool kASplit, bool kGather, bool kF16 = false, bool kPair = false>
__global__ void __launch_bounds__(TC_THREADS, 1)      // 128 registers: 4 warps per scheduler x 32 x 128 = its 16 K registers
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY,
               const __grid_constant__ CUtensorMap tmYs, const __grid_constant__ CUtensorMap tmR,
               const TcParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t xf_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t raw_full_bar[TC_RAW_MAX];
  __shared__ __align__(8) uint64_t raw_empty_bar[TC_RAW_MAX];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t res_bar[3];          // shortcut tile landed in my group's staging tile
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_scale[2][256];   // folded scale / shift of the tile's columns
  __shared__ __align__(16) float s_shift[2][256];
  __shared__ float s_rowsq[2][TC_BM];

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int S = p.stages;
  constexpr uint32_t kOpRow = kBf16 ? 64u : 128u;             // operand bytes per row per k-block
  constexpr uint32_t kAop = TC_BM * kOpRow;                   // one A operand tile: 8 KB / 16 KB
  constexpr uint32_t kNP = kPasses == 3 ? 2u : 1u;            // operand copies (hi [+ lo])
  // kPair (with kCluster == 2, kASplit, 3 passes of kind::f16): ONE tcgen05.mma.cta_group::2 per k-slice computes the pair's
  // 256 x bn tile; each CTA keeps its 128 rows of A and its bn/2 rows of W (half the W bytes per k-block: twice the k-blocks
  // in flight in the same shared memory), the leader (rank 0) issues, every barrier the MMA waits on lives in the leader.
  static_assert(!kPair || (kCluster == 2 && kASplit && kBf16 && kPasses == 3 && !kGather), "pair MMA: pre-split 16-bit operands only");
  const uint32_t b_bytes = (uint32_t)(kPair ? p.bn / 2 : p.bn) * kOpRow;
  // kASplit: the pre-split A tiles have their own ring (p.raw slots of hi [+ lo], filled by warp 0 straight from HBM,
  // freed by the MMA commits) so the HBM-latency-bound A stream runs deeper than the L2-resident W stream; the operand
  // stages then hold W only.
  const uint32_t stage_bytes = kASplit ? kNP * b_bytes : kNP * (kAop + b_bytes);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* store_buf = smem;                                  // 2 (4 in dual-output mode) x 16 KB staging tiles
  // epilogue groups of 128 threads: warps 2..9, plus the transform warps 10..13 (idle: A arrives pre-split) in pair mode,
  // where the main loop is fast enough for the epilogue of short-k tiles to become the limiter
  constexpr int kEpiGroups = kPair ? 3 : 2;
  constexpr int kEpiThreads = kEpiGroups * 128;
  const uint32_t n_store = (p.y_both ? 2u : 1u) * kEpiGroups;
  // tf32: operand stages [A raw = hi | A_lo (3x) | B_hi | B_lo (3x)]
  // bf16: a separate ring of p.raw fp32 A tiles (freed as soon as the transform has read them, so the
  //       HBM-latency-bound A loads run up to p.raw k-blocks ahead of the MMA), then operand stages
  //       [A_hi | A_lo (3x) | B_hi | B_lo (3x)]
  uint8_t* raw0 = smem + n_store * TC_STORE_BYTES;
  const int RAW = p.raw;
  uint8_t* stage0 = raw0 + (kBf16 ? RAW * (kASplit ? kNP * kAop : (uint32_t)TC_A_BYTES) : 0);
  auto a_raw = [&](int s) { return kBf16 ? raw0 + (size_t)s * TC_A_BYTES : stage0 + (size_t)s * stage_bytes; };
  auto a_hi = [&](int s) { return kASplit ? raw0 + (size_t)s * (kNP * kAop) : stage0 + (size_t)s * stage_bytes; };
  auto a_lo = [&](int s) { return a_hi(s) + kAop; };
  auto b_hi = [&](int s) { return stage0 + (size_t)s * stage_bytes + (kASplit ? 0u : kNP * kAop); };
  auto b_lo = [&](int s) { return b_hi(s) + b_bytes; };

  const int nkb = (p.k1 + p.k2) / TC_BK;
  const int64_t tiles_m = (p.m + TC_BM - 1) / TC_BM;
  const int tiles_n = p.n / p.bn;
  // work units: (row-tile group of kCluster tiles) x column tile x group; a cluster walks units,
  // CTA `crank` of the cluster takes row tile  unit_row * kCluster + crank
  const uint32_t crank = kCluster == 2 ? cluster_ctarank() : 0u;
  const int64_t first_unit = blockIdx.x / kCluster, unit_step = gridDim.x / kCluster;
  const int64_t total_tiles = ((tiles_m + kCluster - 1) / kCluster) * tiles_n * p.groups;   // units

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmY);
    tma_prefetch_desc(&tmYs);
    tma_prefetch_desc(&tmR);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xf_bar[s], TC_XF_THREADS);
      mbar_init(&empty_bar[s], kPair ? 1 : kCluster);
    }
    for (int r = 0; r < TC_RAW_MAX; ++r) {
      mbar_init(&raw_full_bar[r], 1);
      mbar_init(&raw_empty_bar[r], kASplit ? 1 : TC_XF_THREADS);
    }
    for (int b = 0; b < 3; ++b) mbar_init(&res_bar[b], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], (kPair ? 2 : 1) * kEpiThreads / 32);      // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) { if (kPair) tmem_alloc_pair(&tmem_base_s, p.tmem_cols); else tmem_alloc(&tmem_base_s, p.tmem_cols); }
  tc_fence_before();
  if (kCluster == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();            // nothing above reads or writes a tensor (common.cuh)

  if (warp == 0 || warp == 14) {
    // ===== TMA producers =====
    // tf32: warp 0 issues A and W of a k-block together.  bf16: warp 0 streams the fp32 A tiles through
    // the raw ring (gated by the transform), warp 14 streams the W tiles into the operand stages
    // (gated by the MMA commits), so A prefetch depth is not tied to the MMA's progress.
    const bool do_a = warp == 0 && !kASplit, do_w = kBf16 ? warp == 14 : warp == 0;
    if (kASplit && warp == 0 && lane == 0) {
      // pre-split A: hi / lo planes of my rows, k-block by k-block, into the A ring
      uint32_t ra = 0;
      for (int64_t tile = first_unit; tile < total_tiles; tile += unit_step) {
        const int64_t rest = tile / tiles_n;
        const int g = (int)(rest % p.groups);
        const int m0 = (int)(((rest / p.groups) * kCluster + crank) * TC_BM);
        for (int kb = 0; kb < nkb; ++kb, ++ra) {
          const int r = ra % RAW;
          mbar_wait(&raw_empty_bar[r], ((ra / RAW) & 1u) ^ 1u);
          if (kPair) {
            if (crank == 0) mbar_arrive_expect_tx(&raw_full_bar[r], 2 * kNP * kAop);      // both CTAs' tiles
            const uint32_t lb = mapa_rank(smem_u32(&raw_full_bar[r]), 0);
            tma_load_3d_pair(a_hi(r), &tmA1, g * p.k1 + kb * TC_BK, m0, 0, lb);      // hi and lo planes in one box
            continue;
          }
          mbar_arrive_expect_tx(&raw_full_bar[r], kNP * kAop);
          tma_load_3d(a_hi(r), &tmA1, g * p.k1 + kb * TC_BK, m0, 0, &raw_full_bar[r]);   // hi [and lo] planes in one box
        }
      }
    }
    if (lane == 0 && (do_a || do_w)) {
      uint32_t it = 0, ra = 0;                  // ra: fp32 A tiles issued into the raw ring
      for (int64_t tile = first_unit; tile < total_tiles; tile += unit_step) {
        const int nt = (int)(tile % tiles_n);
        const int64_t rest = tile / tiles_n;
        const int g = (int)(rest % p.groups);
        const int64_t mt = (rest / p.groups) * kCluster + crank;
        const int m0 = (int)(mt * TC_BM), n0 = nt * p.bn;      // mt may be one past the last tile: all OOB
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          const int k = kb * TC_BK;
          uint64_t* abar;
          uint8_t* adst;
          const bool gathered = kGather && k >= p.k1;   // built by the transform
          if (kBf16) {
            const int r = ra % RAW;
            abar = &raw_full_bar[r];
            adst = a_raw(r);
            if (do_a && !gathered) {
              mbar_wait(&raw_empty_bar[r], ((ra / RAW) & 1u) ^ 1u);
              mbar_arrive_expect_tx(abar, TC_A_BYTES);
              ++ra;
            }
            if (do_w) {
              mbar_wait(&empty_bar[s], ph ^ 1u);
              if (!kPair) mbar_arrive_expect_tx(&full_bar[s], kNP * b_bytes);
              else if (crank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * kNP * b_bytes);   // both halves of the W tile
            }
          } else {
            abar = &full_bar[s];
            adst = a_raw(s);
            mbar_wait(&empty_bar[s], ph ^ 1u);
            mbar_arrive_expect_tx(&full_bar[s], TC_A_BYTES + kNP * b_bytes);
          }
          if (do_a && !gathered) {
            if (p.tap3_rows > 0) {
              // Downsample: W columns are [tap0 | tap1 | tap2]; taps 1,2 = the (M, 2*Cin) view of
              // the input, tap 0 = the same view shifted one output row up inside each graph
              // (3-D map, out-of-range row -1 zero-filled by TMA)
              if (k < p.tap3_cin) {
                const int b0 = m0 / p.tap3_rows, j0 = m0 % p.tap3_rows;
                tma_load_3d(adst, &tmA2, p.tap3_cin + k, j0 - 1, b0, abar);
              } else {
                tma_load_2d(adst, &tmA1, k - p.tap3_cin, m0, abar);
              }
            } else if (k < p.k1) {
              tma_load_2d(adst, &tmA1, g * p.k1 + k, m0, abar);
            } else {
              tma_load_2d(adst, &tmA2, g * p.k2 + (k - p.k1), m0, abar);
            }
          }
          if (do_w && kPair) {
            // my half of the W tile into MY shared memory, counted on the leader's barrier
            const int half = p.bn / 2;
            const uint32_t lb = mapa_rank(smem_u32(&full_bar[s]), 0);
            tma_load_3d_pair(b_hi(s), &tmW, k, g * p.n + n0 + (int)crank * half, 0, lb);
          } else if (do_w) {
            if (kBf16) {
              // 16-bit engines: tmW is (k, n_total, planes); without multicast one box carries the hi and the lo tile
              if (kCluster == 2) {
                // my half of the weight tile, multicast to both CTAs of the pair (tmW box = bn/2 rows per plane): the hi
                // half-tile lands at b_hi + off, the lo half-tile b_bytes further
                const int half = p.bn / 2;
                const uint32_t off = crank * (uint32_t)half * kOpRow;
                // (my hi half-tile and my lo half-tile are a whole tile apart in shared memory: two requests, the map's
                // box holds one plane in this mode)
                tma_load_3d_mc(b_hi(s) + off, &tmW, k, g * p.n + n0 + (int)crank * half, 0, &full_bar[s], 0x3);
                if (kPasses == 3)
                  tma_load_3d_mc(b_lo(s) + off, &tmW, k, g * p.n + n0 + (int)crank * half, 1, &full_bar[s], 0x3);
              } else {
                tma_load_3d(b_hi(s), &tmW, k, g * p.n + n0, 0, &full_bar[s]);
              }
            } else if (kCluster == 2) {
              const int half = p.bn / 2;
              const uint32_t off = crank * (uint32_t)half * kOpRow;
              tma_load_2d_mc(b_hi(s) + off, &tmW, k, g * p.n + n0 + (int)crank * half, &full_bar[s], 0x3);
              if (kPasses == 3)
                tma_load_2d_mc(b_lo(s) + off, &tmW, k, p.n_total + g * p.n + n0 + (int)crank * half,
                               &full_bar[s], 0x3);
            } else {
              tma_load_2d(b_hi(s), &tmW, k, g * p.n + n0, &full_bar[s]);
              if (kPasses == 3) tma_load_2d(b_lo(s), &tmW, k, p.n_total + g * p.n + n0, &full_bar[s]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the warp stays converged, one elected lane issues (tc_common.cuh) =====
    if (!kPair || crank == 0) {
      constexpr int kMmaM = kPair ? 2 * TC_BM : TC_BM;
      const uint32_t idesc = kBf16 ? (kF16 ? umma_idesc_f16(kMmaM, p.bn) : umma_idesc_bf16(kMmaM, p.bn))
                                   : umma_idesc_tf32(kMmaM, p.bn);
      constexpr uint32_t kHi = kBf16 ? UMMA_HI_SW64 : UMMA_HI_SW128;
      const uint32_t d_stage0 = umma_desc_lo(smem_u32(stage0)), d_stage = stage_bytes >> 4;
      const uint32_t d_alo = kAop >> 4, d_bhi = kASplit ? 0u : (kNP * kAop) >> 4, d_blo = d_bhi + (b_bytes >> 4);
      const uint32_t d_ring0 = umma_desc_lo(smem_u32(raw0)), d_ring = (kNP * kAop) >> 4;     // kASplit: the A ring
      uint32_t s = 0, ph = 0, ti = 0, sa = 0, pha = 0;
#ifdef TC_TRACE
      long long cy_e = 0, cy_w = 0, cy_a = 0; const long long cy_t0 = clock64();
#endif
      for (int64_t tile = first_unit; tile < total_tiles; tile += unit_step, ++ti) {
        const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
        TC_ACC(cy_e, mbar_wait(&tmem_empty_bar[buf], tph ^ 1u));          // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn;
        for (int kb = 0; kb < nkb; ++kb) {
          if (kBf16) TC_ACC(cy_w, mbar_wait(&full_bar[s], ph));             // W tiles landed (A comes via xf)
          if (!kASplit) TC_ACC(cy_a, mbar_wait((kPasses == 3 || kBf16) ? &xf_bar[s] : &full_bar[s], ph));
          else TC_ACC(cy_a, mbar_wait(&raw_full_bar[sa], pha));
          tc_fence_after();
          const uint32_t dst = d_stage0 + s * d_stage;
          const uint32_t dah = kASplit ? d_ring0 + sa * d_ring : dst, dal = dah + d_alo, dbh = dst + d_bhi, dbl = dst + d_blo;
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < (kBf16 ? TC_BK / 16 : TC_BK / 8); ++k) {   // UMMA_K = 32 B of the operand row
              const uint32_t koff = 2u * k;
              const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
              if (kPair) {
                umma_f16_lh_pair(tacc, dal + koff, dbh + koff, kHi, idesc, acc);
                umma_f16_lh_pair(tacc, dah + koff, dbl + koff, kHi, idesc, 1u);
                umma_f16_lh_pair(tacc, dah + koff, dbh + koff, kHi, idesc, 1u);
              } else if (kBf16) {
                if (kPasses == 3) {
                  umma_f16_lh(tacc, dal + koff, dbh + koff, kHi, idesc, acc);
                  umma_f16_lh(tacc, dah + koff, dbl + koff, kHi, idesc, 1u);
                  umma_f16_lh(tacc, dah + koff, dbh + koff, kHi, idesc, 1u);
                } else {
                  umma_f16_lh(tacc, dah + koff, dbh + koff, kHi, idesc, acc);
                }
              } else {
                if (kPasses == 3) {
                  umma_tf32_lh(tacc, dal + koff, dbh + koff, kHi, idesc, acc);
                  umma_tf32_lh(tacc, dah + koff, dbl + koff, kHi, idesc, 1u);
                  umma_tf32_lh(tacc, dah + koff, dbh + koff, kHi, idesc, 1u);
                } else {
                  umma_tf32_lh(tacc, dah + koff, dbh + koff, kHi, idesc, acc);
                }
              }
            }
            // smem slot reusable once these MMAs retire (in both CTAs of a pair)
            if (kPair) {                                             // every commit reaches both CTAs of the pair
              umma_commit_pair(&empty_bar[s]);
              umma_commit_pair(&raw_empty_bar[sa]);
              if (kb == nkb - 1) umma_commit_pair(&tmem_full_bar[buf]);
            } else {
              if (kCluster == 2) umma_commit_mc(&empty_bar[s], 0x3); else umma_commit(&empty_bar[s]);
              if (kASplit) umma_commit(&raw_empty_bar[sa]);             // my A tile (not shared with the pair)
              if (kb == nkb - 1) umma_commit(&tmem_full_bar[buf]);       // accumulator complete
            }
          }
          __syncwarp();
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
          if (kASplit && ++sa == (uint32_t)RAW) { sa = 0; pha ^= 1u; }
        }
      }
#ifdef TC_TRACE
      if (blockIdx.x == 0 && lane == 0) {
        g_tc_trace[0] = clock64() - cy_t0; g_tc_trace[1] = cy_e; g_tc_trace[2] = cy_w; g_tc_trace[3] = cy_a;
        g_tc_trace[4] = ti; g_tc_trace[5] = (unsigned long long)nkb; g_tc_trace[6] = (unsigned long long)S; g_tc_trace[7] = (unsigned long long)RAW;
      }
#endif
    }
  } else if (!kPair && warp >= 10 && warp < 14) {
    // ===== transform (warps 10..13, 128 threads): build the MMA A operand(s) from the fp32 stage =====
    // tf32: hi = v with the 13 low mantissa bits cleared (exactly representable in tf32), lo = v - hi
    //       (exact in fp32, |lo| < 2^-10 |v|; the tensor core reads its top 19 bits), in place.
    // bf16: a1 = bf16(v), a2 = bf16(v - a1) written as 64-byte rows (64B swizzle: 16-byte chunk c of
    //       row r lives at chunk c ^ ((r >> 1) & 3)); the source float4 sits at swizzled chunk q & 7 of
    //       its 128-byte row, i.e. logical k-chunk (q & 7) ^ (r & 7).
    if ((kPasses == 3 || kBf16) && !kASplit) {
      constexpr int PER = TC_A_BYTES / 16 / TC_XF_THREADS;
      const int t = threadIdx.x - 320;
      uint32_t it = 0, ra = 0;
      for (int64_t tile = first_unit; tile < total_tiles; tile += unit_step) {
        const int64_t rest = tile / tiles_n;
        const int g = (int)(rest % p.groups);
        const int64_t m0 = ((rest / p.groups) * kCluster + crank) * TC_BM;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          const int k = kb * TC_BK;
          const bool gathered = kGather && k >= p.k1;
          float4 v[PER];
          int rs = s;
          if (gathered) {
            // Fused gather + max-relative (reference torch_vertex.py:21-29): this k-block of the second A
            // source is max_t (y[j_t] - y[i]) over the row's neighbour list, read straight from y (the rows of
            // one graph are a contiguous, L2-resident 64 KB block that this kernel streams anyway).  Identical
            // arithmetic to mr_aggregate_staged_kernel, so the result is bit-identical to the unfused route.
            // v[i] holds LOGICAL chunk q & 7 of row q >> 3 (the store below un-swizzles with r & 7).
            const int kc = g * p.k2 + (k - p.k1);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
              const int q = t + TC_XF_THREADS * i;
              const int r = q >> 3;
              const int lc = (q & 7) ^ (r & 7);                        // logical chunk this slot must hold
              const int64_t row = m0 + r;
              float4 best = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < p.m) {
                const int64_t gb = (row / p.gat_n) * p.gat_n;
                const float* col = p.gat_x + kc + lc * 4;
                const float4 xi = __ldg(reinterpret_cast<const float4*>(col + row * p.gat_ld));
                const int32_t* nb = p.gat_idx + row * p.gat_k;
                best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                for (int tt = 0; tt < p.gat_k; ++tt) {
                  const int j = __ldg(nb + tt);
                  const float4 xj = __ldg(reinterpret_cast<const float4*>(col + (gb + j) * p.gat_ld));
                  const float dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z, dw = xj.w - xi.w;
                  if (dx > best.x) best.x = dx;
                  if (dy > best.y) best.y = dy;
                  if (dz > best.z) best.z = dz;
                  if (dw > best.w) best.w = dw;
                }
              }
              v[i] = best;
            }
          } else {
            rs = kBf16 ? (int)(ra % RAW) : s;
            mbar_wait(kBf16 ? &raw_full_bar[rs] : &full_bar[s], kBf16 ? (ra / RAW) & 1u : ph);
            const float4* raw = reinterpret_cast<const float4*>(a_raw(rs));
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] = raw[t + TC_XF_THREADS * i];
          }
          if (kBf16) {
            // raw tile is in registers: its slot may be refilled.  The refill is an async-proxy (TMA) write
            // after generic-proxy reads: without the proxy fence the write can overtake the reads.
            if (!gathered) {
              fence_proxy_async_smem();
              mbar_arrive(&raw_empty_bar[rs]);
              ++ra;
            }
            mbar_wait(&empty_bar[s], ph ^ 1u);          // operand stage s free (MMAs of k-block it - S retired)
            uint8_t* hi = a_hi(s);
            uint8_t* lo = a_lo(s);
#pragma unroll
            for (int i = 0; i < PER; ++i) {
              const int q = t + TC_XF_THREADS * i;
              const int r = q >> 3;
              const int lc = (q & 7) ^ (r & 7);                          // logical 4-float chunk 0..7
              const uint32_t dst = (uint32_t)r * 64u + ((uint32_t)((lc >> 1) ^ ((r >> 1) & 3)) << 4) +
                                   ((uint32_t)(lc & 1) << 3);
              uint2 hv;
              hv.x = pack16<kF16>(v[i].x, v[i].y);
              hv.y = pack16<kF16>(v[i].z, v[i].w);
              *reinterpret_cast<uint2*>(hi + dst) = hv;
              if (kPasses == 3) {
                const float2 f01 = unpack16<kF16>(hv.x), f23 = unpack16<kF16>(hv.y);
                uint2 lv;
                lv.x = pack16<kF16>(v[i].x - f01.x, v[i].y - f01.y);
                lv.y = pack16<kF16>(v[i].z - f23.x, v[i].w - f23.y);
                *reinterpret_cast<uint2*>(lo + dst) = lv;
              }
            }
          } else {
            float4* hi = reinterpret_cast<float4*>(a_hi(s));
            float4* lo = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
            for (int i = 0; i < PER; ++i) {
              float4 h, l;
              h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u);
              h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u);
              h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u);
              h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u);
              l.x = v[i].x - h.x; l.y = v[i].y - h.y; l.z = v[i].z - h.z; l.w = v[i].w - h.w;
              hi[t + TC_XF_THREADS * i] = h;
              lo[t + TC_XF_THREADS * i] = l;
            }
          }
          fence_proxy_async_smem();             // generic-proxy writes -> visible to the MMA (async proxy)
          mbar_arrive(&xf_bar[s]);
        }
      }
    }
  } else if (warp >= 2 && warp < 2 + 4 * kEpiGroups) {
    // ===== epilogue: groups of 128 threads, group h owns the 32-column chunks h, h + G, h + 2G, ... =====
    const int ew = warp - 2;                     // 0..7 (0..11 with three groups)
    const int half = ew >> 2;                    // my group
    const int et = (ew & 3) * 32 + lane;         // 0..127 inside the group
    const int e256 = ew * 32 + lane;             // 0..255 over the first two groups (they stage scale / shift)
    const int quad = warp & 3;                   // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;              // row inside the tile
    const bool store_thread = (et == 0);
    uint8_t* sb = store_buf + half * TC_STORE_BYTES;
    uint8_t* sb32 = store_buf + (kEpiGroups + half) * TC_STORE_BYTES;   // dual-output mode: the fp32 tile beside the split tiles
    // hand an accumulator back to the MMA thread (pair mode: the leader's, from both CTAs)
    auto tmem_release = [&](uint64_t* bar) {
      __syncwarp();
      if (lane == 0) { if (kPair && crank != 0) mbar_arrive_cluster(mapa_rank(smem_u32(bar), 0)); else mbar_arrive(bar); }
    };
    uint32_t ti = 0;
    uint32_t res_phase = 0;                       // shortcut tiles my group has consumed
#ifdef TC_TRACE
    long long ec[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long ec_t0 = clock64();
#endif
    for (int64_t tile = first_unit; tile < total_tiles; tile += unit_step, ++ti) {
      const int nt = (int)(tile % tiles_n);
      const int64_t rest = tile / tiles_n;
      const int g = (int)(rest % p.groups);
      const int64_t mt = (rest / p.groups) * kCluster + crank;
      const int m0 = (int)(mt * TC_BM), n0 = nt * p.bn;
      const uint32_t buf = ti & 1u, tph = (ti >> 1) & 1u;
      const int64_t row = (int64_t)m0 + r;
      const bool row_ok = row < p.m;
      const int64_t col0 = (int64_t)g * p.n + n0;
      // stage this tile's scale / shift (overlaps the tile's main loop) and pull the residual lines of
      // my chunks towards L2 so the loads below do not pay the HBM latency
      float* ssc = s_scale[ti & 1u];
      float* ssh = s_shift[ti & 1u];
      if (e256 < p.bn) {
        ssc[e256] = (p.scale ? __ldg(p.scale + col0 + e256) : 1.0f) * p.unscale;
        ssh[e256] = p.shift ? __ldg(p.shift + col0 + e256) : 0.0f;
      }
      const float* res_row = (p.residual && row_ok) ? p.residual + row * p.ldr + col0 : nullptr;
      if (res_row)
        for (int c = half * 32; c < p.bn; c += 32 * kEpiGroups)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(res_row + c));
      TC_ACC(ec[0], named_bar_sync(3, kEpiThreads));
      TC_ACC(ec[1], mbar_wait(&tmem_full_bar[buf], tph));
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn + ((uint32_t)(quad * 32) << 16);
      float rowsq = 0.0f;
      if (half * 32 >= p.bn) {                    // a narrow tile: nothing for this group to read
        tc_fence_before();
        tmem_release(&tmem_empty_bar[buf]);
      }
      for (int c = half * 32; c < p.bn; c += 32 * kEpiGroups) {
        if (p.res_tma) {
          // The shortcut tile (128 rows x 32 fp32) comes by TMA into the fp32 staging tile my group writes its result to:
          // one coalesced request instead of 8 row-strided 16-byte loads per thread (measured: 5 000+ cycles per chunk,
          // the longest step of the epilogue).  Every thread later reads its own row of it and overwrites the same bytes.
          if (store_thread) {
            bulk_wait_group_read<0>();                      // my group's previous store has read the staging tiles
            mbar_arrive_expect_tx(&res_bar[half], TC_STORE_BYTES);
            tma_load_2d(p.y_both ? sb32 : sb, &tmR, (int)(col0 + c), m0, &res_bar[half]);
          }
        }
        float4 sc0[4], sh0[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          sc0[i] = *reinterpret_cast<const float4*>(ssc + c + 4 * i);
          sh0[i] = *reinterpret_cast<const float4*>(ssh + c + 4 * i);
        }
        float v[32];
        TC_ACC(ec[2], { tmem_ld16_nowait(tacc + (uint32_t)c, v);
        tmem_ld16_nowait(tacc + (uint32_t)c + 16u, v + 16);
        tmem_ld_wait(); });
#ifdef TC_TRACE
        ec[11] += 1;
#endif
        if (c + 32 * kEpiGroups >= p.bn) {        // my last read of this accumulator: hand it back
          tc_fence_before();
          tmem_release(&tmem_empty_bar[buf]);
        }
        if (p.res_tma) {
          mbar_wait(&res_bar[half], res_phase & 1u);
          ++res_phase;
        }

        switch (p.act) {        // one specialised, branch-free instance per activation
          case GRAFP_ACT_NONE:  epi_apply<GRAFP_ACT_NONE>(v, sc0, sh0, ssc + c, ssh + c, p.act_param); break;
          case GRAFP_ACT_RELU:  epi_apply<GRAFP_ACT_RELU>(v, sc0, sh0, ssc + c, ssh + c, p.act_param); break;
          case GRAFP_ACT_LEAKY: epi_apply<GRAFP_ACT_LEAKY>(v, sc0, sh0, ssc + c, ssh + c, p.act_param); break;
          case GRAFP_ACT_GELU:  epi_apply<GRAFP_ACT_GELU>(v, sc0, sh0, ssc + c, ssh + c, p.act_param); break;
          default:              epi_apply<GRAFP_ACT_ELU>(v, sc0, sh0, ssc + c, ssh + c, p.act_param); break;
        }
        if (p.res_tma) epi_add_smem(v, (p.y_both ? sb32 : sb) + r * 128, r);
        else if (res_row) epi_add_global(v, res_row + c);      // (rare: split-only output or unaligned shortcut)
        if (p.row_sumsq) {
#pragma unroll
          for (int q = 0; q < 32; ++q) rowsq = fmaf(v[q], v[q], rowsq);
        }
        if (!p.res_tma) {       // (with a TMA shortcut tile the wait on its barrier has already ordered us after the store)
          TC_ACC(ec[3], { if (store_thread) bulk_wait_group_read<0>(); });      // my group's previous store has read `sb`
          TC_ACC(ec[4], named_bar_sync(half < 2 ? 1 + half : 5, 128));
        }
        TC_CLK(ec_c0);
        if (p.y_split) {
          // bf16 [hi ; lo] planes: exactly the operand pair a consuming bf16x3 GEMM would derive from the
          // fp32 value (hi = bf16(v), lo = bf16(v - hi)); two 128 x 64 B tiles, 64B swizzle
          uint32_t hp[16], lp[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            hp[q] = pack16<kF16>(v[2 * q], v[2 * q + 1]);
            const float2 f = unpack16<kF16>(hp[q]);
            lp[q] = pack16<kF16>(v[2 * q] - f.x, v[2 * q + 1] - f.y);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t off = (uint32_t)r * 64u + ((uint32_t)(q ^ ((r >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(sb + off) = make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
            if (p.y_split == 2)
              *reinterpret_cast<uint4*>(sb + TC_STORE_BYTES / 2 + off) =
                  make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
          }
        }
        if (!p.y_split || p.y_both) {
          uint8_t* dst = p.y_both ? sb32 : sb;
#pragma unroll
          for (int q = 0; q < 8; ++q)                     // 128B swizzle: 16-byte chunk q -> q ^ (row & 7)
            *reinterpret_cast<float4*>(dst + r * 128 + ((q ^ (r & 7)) << 4)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        TC_SINCE(ec[5], ec_c0);
        TC_ACC(ec[6], fence_proxy_async_smem());
        TC_ACC(ec[7], named_bar_sync(half < 2 ? 1 + half : 5, 128));
        TC_CLK(ec_c1);
        if (store_thread) {
          if (p.y_split) {
            tma_store_3d(&tmYs, sb, (int)(col0 + c), m0, 0);      // one box: the hi tile and (3-pass engines) the lo tile
            if (p.y_both) tma_store_2d(&tmY, sb32, (int)(col0 + c), m0);
          } else {
            tma_store_2d(&tmY, sb, (int)(col0 + c), m0);
          }
          bulk_commit_group();
        }
        TC_SINCE(ec[8], ec_c1);
      }
      if (p.row_sumsq) {
        // deterministic: group 1 hands its partial to group 0 (fixed order), one atomic per row and
        // column tile (at most two column tiles: fp32 addition of two terms commutes)
        if (half >= 1) s_rowsq[half - 1][r] = rowsq;
        named_bar_sync(4, kEpiThreads);
        if (half == 0 && row_ok) {
          float t = rowsq + (p.bn > 32 ? s_rowsq[0][r] : 0.0f);
          if (kEpiGroups == 3 && p.bn > 64) t += s_rowsq[1][r];
          atomicAdd(p.row_sumsq + row, t);
        }
      }
    }
    if (store_thread) bulk_wait_group_all();
#ifdef TC_TRACE
    if (blockIdx.x == 0 && warp == 2 && lane == 0) {
      g_tc_trace[8] = clock64() - ec_t0;
      for (int i = 0; i < 9; ++i) g_tc_trace[9 + i] = ec[i];
      g_tc_trace[18] = ec[11];
    }
#endif
  }
  tc_fence_before();
  if (kCluster == 2) cluster_sync_all(); else __syncthreads();   // the peer may still signal my barriers
  if (warp == 1) {
    tc_fence_after();
    if (kPair) tmem_dealloc_pair(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

#ifdef TC_TRACE
extern "C" int grafp_debug_tc_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, grafp::g_tc_trace, sizeof(unsigned long long) * 24);
}
#endif
// launches of the CTA-pair kernel so far (debug / test hook, not part of the ABI in include/grafp.h)
static long long g_pair_launches = 0;
extern "C" long long grafp_debug_pair_launches() { return g_pair_launches; }
// ---- host side -------------------------------------------------------------------------
EncodeTiledFn tc_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int tc_make_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                   int box_rows) {
  EncodeTiledFn fn = tc_encode_fn();
  GRAFP_REQUIRE(fn, "tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled(2d) failed (%d)", (int)r);
  return 0;
}

// fp32 row-major, box = (16 cols = 64 B, box_rows), 64B swizzle (kNN's narrow k-blocks)
int tc_make_map_2d_bk16(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows) {
  EncodeTiledFn fn = tc_encode_fn();
  GRAFP_REQUIRE(fn, "tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled(2d bk16) failed (%d)", (int)r);
  return 0;
}

int tc_make_map_2d_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld,
                        int box_rows) {
  EncodeTiledFn fn = tc_encode_fn();
  GRAFP_REQUIRE(fn, "tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled(bf16) failed (%d)", (int)r);
  return 0;
}

// bf16 (planes, rows, cols) with element strides (plane_stride, ld, 1); box = (32 cols, box_rows, 1), 64B swizzle
int tc_make_map_3d_bf16(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int64_t planes,
                        int64_t ld, int64_t plane_stride, int box_rows, int box_planes) {
  // box_planes = 2: ONE request moves the hi and the lo tile (they are adjacent in shared memory).  A TMA request costs
  // its issuing thread ~400 cycles whatever its size (scripts/micro/tma_rate.cu): two requests per operand and k-block
  // made the producers, not the tensor pipe, the pace-setters of the 3-pass GEMMs.
  EncodeTiledFn fn = tc_encode_fn();
  GRAFP_REQUIRE(fn, "tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled(3d bf16) failed (%d)", (int)r);
  return 0;
}

int tc_make_map_3d(CUtensorMap* map, const float* base, int64_t d0, int64_t d1, int64_t d2,
                   int64_t s1, int64_t s2, int box1, int box2) {
  EncodeTiledFn fn = tc_encode_fn();
  GRAFP_REQUIRE(fn, "tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)s1 * 4, (cuuint64_t)s2 * 4};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box1, (cuuint32_t)box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "tc: cuTensorMapEncodeTiled(3d) failed (%d)", (int)r);
  return 0;
}

static int pick_bn(int n) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("GRAFP_TC_BN");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0 && forced <= 256 && forced % 32 == 0 && n % forced == 0) return forced;
  for (int bn = 256; bn >= 32; bn -= 32)
    if (n % bn == 0) return bn;
  return 0;
}

int gemm_tc_supported(const grafp_gemm_args& a) {
  if (a.n % 32 != 0 || pick_bn(a.n) == 0) return 0;
  if (a.m < 1) return 0;
  if (a.tap3_nodes > 0) {
    const int cin = a.k1 / 3;
    if (a.groups != 1 || a.k2 != 0 || cin % TC_BK != 0 || a.lda1 != cin) return 0;
    const int r = a.tap3_nodes;                       // output rows per graph
    if (!((r >= TC_BM && r % TC_BM == 0) || (r < TC_BM && TC_BM % r == 0))) return 0;
  } else {
    if (a.k1 % TC_BK != 0 || a.k2 % TC_BK != 0) return 0;
    if (a.a1_split) {
      if (a.k2 != 0 || a.lda1s % 8 != 0 || (reinterpret_cast<uintptr_t>(a.a1_split) & 15)) return 0;
    } else if ((a.lda1 * 4) % 16 != 0) {
      return 0;
    }
    if (a.k2 && !a.a2_gather_idx && (a.lda2 * 4) % 16 != 0) return 0;
  }
  if (a.a1_split && a.tap3_nodes > 0) return 0;
  if ((a.ldw * 4) % 16 != 0 || a.ldw % 8 != 0) return 0;
  if (a.y_split) {
    if (a.ldys % 8 != 0 || (reinterpret_cast<uintptr_t>(a.y_split) & 15) || a.row_sumsq) return 0;
  }
  if (a.y && (a.ldy % 4 != 0 || (reinterpret_cast<uintptr_t>(a.y) & 15))) return 0;
  if (!a.y && !a.y_split) return 0;
  if (a.residual && (a.ldr % 4 != 0 || (reinterpret_cast<uintptr_t>(a.residual) & 15))) return 0;
  return 1;
}

// tf32 engines: passes == 3 reads a.w_split = stacked fp32 [W_hi ; W_lo] (2 * groups * n rows).
// bf16 engines (fmt 1): a.w_split_bf16 = stacked bf16 [w1 ; w2]; passes == 1 uses only w1.
// f16x3 (fmt 2): a.w_split_f16 = stacked fp16 split of w * 2^s, a.w_f16_unscale = 2^-s.
int gemm_tc_launch(const grafp_gemm_args& a, int passes, int fmt, cudaStream_t st) {
  const int bf16 = fmt != 0;
  const bool f16 = fmt == 2;
  const int bn = pick_bn(a.n);
  const int n_total = a.groups * a.n;
  CUtensorMap mA1, mA2, mW, mY, mYs, mR;
  TcParams p;
  p.tap3_rows = 0; p.tap3_cin = 0;
  if (a.tap3_nodes > 0) {
    const int cin = a.k1 / 3, r = a.tap3_nodes;
    const int64_t graphs = a.m / r;
    // plain view: (M, 2*Cin); shifted view: (graphs, r, 2*Cin)
    if (int rc = tc_make_map_2d(&mA1, a.a1, a.m, 2 * cin, 2 * cin, TC_BM)) return rc;
    const int box1 = r >= TC_BM ? TC_BM : r, box2 = r >= TC_BM ? 1 : TC_BM / r;
    if (int rc = tc_make_map_3d(&mA2, a.a1, 2 * cin, r, graphs, 2 * cin, (int64_t)r * 2 * cin, box1, box2))
      return rc;
    p.tap3_rows = r; p.tap3_cin = cin;
  } else {
    if (a.a1_split) {
      GRAFP_REQUIRE(bf16, "gemm_tc: a split-bf16 A operand needs a bf16 engine");
      if (int rc = tc_make_map_3d_bf16(&mA1, a.a1_split, (int64_t)a.groups * a.k1, a.m, passes == 3 ? 2 : 1, a.lda1s,
                                       a.m * a.lda1s, TC_BM, passes == 3 ? 2 : 1))
        return rc;
    } else if (int rc = tc_make_map_2d(&mA1, a.a1, a.m, (int64_t)a.groups * a.k1, a.lda1, TC_BM)) {
      return rc;
    }
    if (a.k2 > 0 && !a.a2_gather_idx) {
      if (int rc = tc_make_map_2d(&mA2, a.a2, a.m, (int64_t)a.groups * a.k2, a.lda2, TC_BM)) return rc;
    } else {
      mA2 = mA1;
    }
  }
  const int64_t tiles_m = (a.m + TC_BM - 1) / TC_BM;
  static int mc_env = -1;
  if (mc_env < 0) { const char* e = getenv("GRAFP_TC_CLUSTER"); mc_env = e ? atoi(e) : 2; }
  // CTA pairs with weight multicast pay off only for full-width tiles (measured on B200: bn = 256 3-7 % faster,
  // bn <= 128 10-30 % slower than independent CTAs -- the pair's lock-step costs more than the L2 traffic saved)
  // ... and only when every CTA has several tiles to amortise the pair's lock-step over (the train step's GEMMs have
  // one or two tiles per SM: there the clustered launch measured 25.6 us per call against 15.7 us un-clustered)
  const int64_t all_tiles = tiles_m * (a.n / bn) * a.groups;
  static int min_waves = -1;
  if (min_waves < 0) { const char* e = getenv("GRAFP_TC_CLUSTER_MIN_WAVES"); min_waves = e ? atoi(e) : 4; }
  const int cluster = (mc_env == 2 && tiles_m >= 2 && bn >= 256 && sm_count() % 2 == 0 &&
                       all_tiles >= min_waves * (int64_t)sm_count()) ? 2 : 1;
  const bool y_both = a.y_split && a.y;                    // dual output: fp32 and split
  if (a.y_split) {
    if (int rc = tc_make_map_3d_bf16(&mYs, a.y_split, n_total, a.m, passes == 3 ? 2 : 1, a.ldys, a.m * a.ldys, TC_BM, passes == 3 ? 2 : 1)) return rc;
  }
  if (a.y) {
    if (int rc = tc_make_map_2d(&mY, a.y, a.m, n_total, a.ldy, TC_BM)) return rc;
  }
  if (!a.y) mY = mYs;
  if (!a.y_split) mYs = mY;
  // shortcut by TMA: needs the fp32 staging tile (fp32 or dual output) and a TMA-addressable residual matrix
  static int res_tma_env = -1;
  if (res_tma_env < 0) { const char* e = getenv("GRAFP_TC_RES_TMA"); res_tma_env = e ? atoi(e) : 1; }
  const bool res_tma = res_tma_env != 0 && a.residual && a.y && a.ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(a.residual) & 15) == 0;
  if (res_tma) {
    if (int rc = tc_make_map_2d(&mR, a.residual, a.m, n_total, a.ldr, TC_BM)) return rc;
  } else {
    mR = mY;
  }
  p.y_both = y_both ? 1 : 0;
  p.res_tma = res_tma ? 1 : 0;
  p.y_split = a.y_split ? (passes == 3 ? 2 : 1) : 0;     // the 1-pass engine carries the hi plane only
  p.gat_idx = a.a2_gather_idx; p.gat_x = a.a1; p.gat_ld = a.lda1; p.gat_n = a.a2_gather_nodes; p.gat_k = a.a2_gather_k;
  const bool asplit = a.a1_split != nullptr;
  p.k1 = a.k1; p.k2 = a.k2; p.n = a.n; p.bn = bn; p.n_total = n_total; p.groups = a.groups; p.m = a.m;
  p.scale = a.scale; p.shift = a.shift; p.residual = a.residual; p.ldr = a.ldr;
  p.row_sumsq = a.row_sumsq;
  p.act = a.act; p.act_param = a.act_param;
  p.unscale = f16 ? a.w_f16_unscale : 1.0f;
  uint32_t cols = 32;
  while ((int)cols < 2 * bn) cols <<= 1;
  p.tmem_cols = cols;
  const size_t np = passes == 3 ? 2 : 1;
  const size_t a_slot = np * (size_t)TC_BM * 64;                          // pre-split A tile: hi [+ lo]
  // CTA-pair MMA (cta_group::2) for the pre-split f16x3 GEMMs on 256-wide tiles: half the W bytes per CTA and k-block
  static int pair_env = -1;
  if (pair_env < 0) { const char* e = getenv("GRAFP_TC_PAIR"); pair_env = e ? atoi(e) : 1; }
  // Measured (scripts/gemm_trace.py, M = 262 144 / 131 072): k-block time 1 200 -> 830-990 cycles for k >= 512; at k = 256 the
  // tile's main loop becomes shorter than its epilogue and the pair is 4 % slower than two independent CTAs.
  static int pair_min_k = -1;
  if (pair_min_k < 0) { const char* e = getenv("GRAFP_TC_PAIR_MIN_K"); pair_min_k = e ? atoi(e) : 512; }
  const bool pair = pair_env != 0 && cluster == 2 && asplit && f16 && passes == 3 && a.groups == 1 && !a.row_sumsq &&
                    a.k1 + a.k2 >= pair_min_k;
  if (bf16) {
    // the stacked [hi ; lo] weight matrix as (k, n_total, 2 planes): one box = the hi and the lo tile of a k-block
    if (int rc = tc_make_map_3d_bf16(&mW, f16 ? a.w_split_f16 : a.w_split_bf16, a.k1 + a.k2, n_total, passes == 3 ? 2 : 1, a.ldw,
                                     (int64_t)n_total * a.ldw, cluster == 2 ? bn / 2 : bn,
                                     (passes == 3 && (cluster == 1 || pair)) ? 2 : 1))
      return rc;
  } else if (int rc = tc_make_map_2d(&mW, passes == 3 ? a.w_split : a.w, (int64_t)n_total * (passes == 3 ? 2 : 1),
                                     a.k1 + a.k2, a.ldw, cluster == 2 ? bn / 2 : bn))
    return rc;
  const size_t stage_bytes = bf16 ? (asplit ? np * (size_t)(pair ? bn / 2 : bn) * 64 : np * ((size_t)TC_BM * 64 + (size_t)bn * 64))
                                  : np * (TC_A_BYTES + (size_t)bn * TC_BK * 4);
  // Shared-memory plan.  The fp32 A ring covers the HBM latency (a memory-bound shape needs ~75 KB in flight
  // per SM to stream at the HBM rate), the operand stages only the transform -> MMA hand-off: narrow tiles
  // (small stages) get a deep ring and three operand stages, 256-wide tiles keep three ring slots.
  const size_t n_store = (y_both ? 2 : 1) * (pair ? 3 : 2);        // staging tiles: one (two) per epilogue group
  const size_t budget = 220 * 1024 - 1024 - n_store * TC_STORE_BYTES;      // 227 KB minus ~6 KB static
  int raw = 0, stages;
  if (bf16 && !asplit) {
    static int raw_env = -1;
    if (raw_env < 0) { const char* e = getenv("GRAFP_TC_RAW"); raw_env = e ? atoi(e) : 0; }
    raw = 3;
    if (3 * stage_bytes + 4 * TC_A_BYTES <= budget) {
      raw = (int)((budget - 3 * stage_bytes) / TC_A_BYTES);
      if (raw > TC_RAW_MAX) raw = TC_RAW_MAX;
    }
    if (raw_env >= 2 && raw_env <= TC_RAW_MAX) raw = raw_env;
  }
  size_t ring_bytes = (size_t)raw * TC_A_BYTES;
  int asplit_ws = 3;
  if (asplit) {
    // W stages cover the L2 latency (three, two when the tile is so wide that the A ring would starve), the rest of
    // shared memory is A tiles in flight from HBM
    static int ws_env = -1;
    if (ws_env < 0) { const char* e = getenv("GRAFP_TC_ASPLIT_WSTAGES"); ws_env = e ? atoi(e) : 0; }
    int ws = ws_env >= 1 && ws_env <= TC_MAX_STAGES ? ws_env : (pair ? 5 : 3);
    while (ws > 1 && budget < ws * stage_bytes + 3 * a_slot) --ws;
    raw = (int)((budget - ws * stage_bytes) / a_slot);
    if (raw > TC_RAW_MAX) raw = TC_RAW_MAX;
    if (raw < 1) raw = 1;
    ring_bytes = (size_t)raw * a_slot;
    asplit_ws = ws;
  }
  p.raw = raw > 0 ? raw : 1;
  const size_t fixed_bytes = n_store * TC_STORE_BYTES + ring_bytes;
  const int nkb = (a.k1 + a.k2) / TC_BK;
  stages = (int)((220 * 1024 - fixed_bytes - 1024) / stage_bytes);
  if (asplit && raw < TC_RAW_MAX && stages > asplit_ws) stages = asplit_ws;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 1) stages = 1;
  p.stages = stages;
  (void)nkb;
  const size_t smem = stage_bytes * stages + fixed_bytes + 1024;
  const int64_t units = ((tiles_m + cluster - 1) / cluster) * (a.n / bn) * a.groups;
  int grid = sm_count() / cluster;
  if (units < grid) grid = (int)units;
  grid *= cluster;
  using KernFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, TcParams);
  KernFn kern;
  if (f16) {
    GRAFP_REQUIRE(passes == 3 && !a.a2_gather_idx, "gemm_tc: the fp16 operand format exists as the 3-pass f16x3 engine only");
    if (pair) { kern = gemm_tc_kernel<3, 2, true, true, false, true, true>; ++g_pair_launches; }
    else if (asplit) kern = cluster == 2 ? gemm_tc_kernel<3, 2, true, true, false, true> : gemm_tc_kernel<3, 1, true, true, false, true>;
    else kern = cluster == 2 ? gemm_tc_kernel<3, 2, true, false, false, true> : gemm_tc_kernel<3, 1, true, false, false, true>;
  } else if (a.a2_gather_idx)
    kern = passes == 3 ? (cluster == 2 ? gemm_tc_kernel<3, 2, true, false, true> : gemm_tc_kernel<3, 1, true, false, true>)
                       : (cluster == 2 ? gemm_tc_kernel<1, 2, true, false, true> : gemm_tc_kernel<1, 1, true, false, true>);
  else if (asplit)
    kern = passes == 3 ? (cluster == 2 ? gemm_tc_kernel<3, 2, true, true, false> : gemm_tc_kernel<3, 1, true, true, false>)
                       : (cluster == 2 ? gemm_tc_kernel<1, 2, true, true, false> : gemm_tc_kernel<1, 1, true, true, false>);
  else if (bf16)
    kern = passes == 3 ? (cluster == 2 ? gemm_tc_kernel<3, 2, true, false, false> : gemm_tc_kernel<3, 1, true, false, false>)
                       : (cluster == 2 ? gemm_tc_kernel<1, 2, true, false, false> : gemm_tc_kernel<1, 1, true, false, false>);
  else
    kern = passes == 3 ? (cluster == 2 ? gemm_tc_kernel<3, 2, false, false, false> : gemm_tc_kernel<3, 1, false, false, false>)
                       : (cluster == 2 ? gemm_tc_kernel<1, 2, false, false, false> : gemm_tc_kernel<1, 1, false, false, false>);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaError_t le = launch_ex(kern, dim3(grid), dim3(TC_THREADS), smem, st, cluster, mA1, mA2, mW, mY, mYs, mR, p);
  if (le != cudaSuccess) return fail("gemm_tc launch: %s", cudaGetErrorString(le));
  return check_launch("gemm_tc");
}

// hi = tf32_rna(w), lo = tf32_rna(w - hi): out is (2*rows, cols) = [hi ; lo]
__global__ void split_tf32_kernel(const float* __restrict__ w, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = w[i];
    const float h = to_tf32(v);
    out[i] = h;
    out[n + i] = to_tf32(v - h);
  }
}

// w1 = bf16(w), w2 = bf16(w - w1): out is bf16 (2*rows, cols) = [w1 ; w2]
__global__ void split_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = w[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = h;
    out[n + i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// w1 = f16(w * prescale), w2 = f16(w * prescale - w1): out is fp16 (2*rows, cols) = [w1 ; w2]
__global__ void split_f16_kernel(const float* __restrict__ w, __half* __restrict__ out, int64_t n, float prescale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = w[i] * prescale;
    const float c = fminf(fmaxf(v, -65504.0f), 65504.0f);
    const __half h = __float2half_rn(c);
    out[i] = h;
    out[n + i] = __float2half_rn(fminf(fmaxf(v - __half2float(h), -65504.0f), 65504.0f));
  }
}

}  // namespace grafp

using namespace grafp;

extern "C" int grafp_split_f16(const float* w, int64_t count, float prescale, void* out_f16_hi_lo, void* stream) {
  GRAFP_REQUIRE(count >= 0 && (count == 0 || (w && out_f16_hi_lo)), "split_f16: bad arguments");
  GRAFP_REQUIRE(prescale > 0.0f, "split_f16: prescale must be positive");
  if (count == 0) return 0;
  split_f16_kernel<<<(unsigned)((count + 255) / 256), 256, 0, as_stream(stream)>>>(
      w, static_cast<__half*>(out_f16_hi_lo), count, prescale);
  return check_launch("split_f16");
}

extern "C" int grafp_split_bf16(const float* w, int64_t count, void* out_bf16_hi_lo, void* stream) {
  GRAFP_REQUIRE(count >= 0 && (count == 0 || (w && out_bf16_hi_lo)), "split_bf16: bad arguments");
  if (count == 0) return 0;
  split_bf16_kernel<<<(unsigned)((count + 255) / 256), 256, 0, as_stream(stream)>>>(
      w, static_cast<__nv_bfloat16*>(out_bf16_hi_lo), count);
  return check_launch("split_bf16");
}

extern "C" int grafp_split_tf32(const float* w, int64_t count, float* out_hi_lo, void* stream) {
  GRAFP_REQUIRE(count >= 0 && (count == 0 || (w && out_hi_lo)), "split_tf32: bad arguments");
  if (count == 0) return 0;
  split_tf32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, as_stream(stream)>>>(w, out_hi_lo, count);
  return check_launch("split_tf32");
}
