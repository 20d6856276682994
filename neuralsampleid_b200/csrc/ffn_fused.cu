// Fused node FFN (sm_100a):  y = x + s2 * (act(s1 * (x W1^T) + t1) W2^T) + t2   in ONE kernel, the hidden tile on chip.
// Reference: FFN.forward (encoder/graph_encoder.py:82-89) with eval-mode BatchNorm folded into (s, t).
//
// The two 1x1-conv GEMMs of the FFN are HBM-bound at C <= 128 (stages 1-2 of size 't'): the 4C-wide hidden tensor
// (1.07 GB at 4096 segments) is written by fc1 and read back by fc2.  Here a persistent CTA owns 128 rows at a time:
//   TMA        x rows (fp32, 128B swizzle) through a small raw ring; weight k-blocks (pre-split fp16 hi/lo) through
//              a ring, in exactly the order the MMA warp consumes them -- or, when a tile's k-blocks all fit (C = 64),
//              loaded once and resident for the kernel's lifetime (wres)
//   transform  x -> fp16 hi / lo operand k-blocks (the f16x3 split of gemm_tc.cu), resident for the whole tile
//   GEMM 1     for every chunk of 64 hidden columns: acc1[j & 1] (TMEM, 64 columns) = x W1_j^T, 3 kind::f16 passes
//   epilogue 1 TMEM -> registers -> s1 / t1 / activation -> fp16 hi / lo pairs -> tcgen05.st into TMEM (lane = row,
//              two k-elements per 32-bit column: the layout of an A operand in TMEM).  FF_HTMEM = 0: into shared memory,
//              in the K-major 64B-swizzled operand layout instead
//   GEMM 2     acc2 (TMEM, C columns) += h_j W2[:, chunk j]^T, A operand from TMEM
//   epilogue 2 raw accumulators -> per-warp staging block (transpose) -> s2 / t2 + the shortcut x (L2 hit), coalesced -> y
// TMEM: acc1 2 x 64 | acc2 2 x C | h 2 x (32 hi + 32 lo) columns = 512 at C = 128.
// GEMM 1 of chunk j+1 is issued before GEMM 2 of chunk j, so the tensor pipe works while epilogue 1 converts.
// The arithmetic (operand values, accumulation order over k) is that of the two gemm_tc.cu launches it replaces:
// the result is bit-identical (tests/test_gpu_kernels.py::test_ffn_fused_bit_exact).
//
// MR = true is the same skeleton for the tail of the Grapher (encoder/gcn_lib/torch_vertex.py:24-34 MRConv2d.forward
// + :183-195 Grapher.forward): out = res + s2 * (act(s1 * ([x | m] W1^T) + t1) W2^T) + t2, where W1 is the grouped
// (groups = 4) BasicConv over the channel-interleaved [x, max-relative m] and W2 is fc2.  The 2C-wide MRConv output
// stays on chip.  A chunk of 64 hidden columns only reads x[:, 32j:32j+32) and m[:, 32j:32j+32) (C <= 128: the chunk
// is one group at C = 128, two block-diagonal groups at C = 64), so the A operand is a k-block RING here (every k-block
// is consumed once), not a tile-resident operand.
#include <cuda_fp16.h>
#include "tc_common.cuh"

namespace grafp {

#ifndef FF_WAIT
#define FF_WAIT mbar_wait
#endif
// The hidden chunk (GEMM 2's A operand) lives in TMEM (tcgen05.st by epilogue 1, tcgen05.mma with A from TMEM): no
// shared-memory round trip and 64 KB of shared memory back for the rings.  -DFF_HTMEM=0: the shared-memory form.
#ifndef FF_HTMEM
#define FF_HTMEM 1
#endif
// w0 TMA x, w1 MMA + TMEM, w2 TMA W, w4-7 epilogue 2, w8-11 transform, w12-27 epilogue 1 (the activation is the
// instruction-heaviest stage: 16 warps = {chunk parity} x {32-column half} x {TMEM lane quarter})
constexpr int FF_THREADS = 896;
constexpr int FF_HC = 64;            // hidden columns per chunk
constexpr int FF_RAW = 2;            // fp32 x k-blocks in flight
constexpr int FF_WMAX = 16;          // weight ring slots
constexpr int FF_AMAX = 6;           // A operand ring slots (MR mode)
constexpr uint32_t FF_KB_BYTES = TC_BM * 64;          // one 128-row fp16 k-block (32 columns): 8 KB
constexpr size_t FF_HOP_BYTES = FF_HTMEM ? 0 : 2 * 2 * 2 * FF_KB_BYTES;      // hidden operand in shared memory
constexpr size_t FF_STAGE_BYTES = 4 * 2048;           // epilogue 2 staging: 32 rows x 16 columns fp32 per warp
constexpr size_t FF_SMEM_BUDGET = 216 * 1024;         // dynamic shared memory (the 227 KB limit less ~6 KB static + alignment)

#ifdef FF_TRACE
__device__ unsigned long long g_ff_trace[3 * 1024];
#define FF_T(region, id)                                                                          \
  do {                                                                                            \
    if (blockIdx.x == 0 && tr_n < 1023) g_ff_trace[(region) * 1024 + tr_n++] = (global_timer_ns() << 8) | (id); \
  } while (0)
#define FF_ACC(var, stmt) do { const long long c0_ = clock64(); stmt; var += clock64() - c0_; } while (0)
#define FF_DECL(...) long long __VA_ARGS__
#define FF_OUT(slot, val) do { if (blockIdx.x == 0 && lane == 0) g_ff_trace[slot] = (unsigned long long)(val); } while (0)
#else
#define FF_DECL(...) do { } while (0)
#define FF_OUT(slot, val) do { } while (0)
#define FF_T(region, id) do { } while (0)
#define FF_ACC(var, stmt) do { stmt; } while (0)
#endif

struct FfnParams {
  int C, Hd;                 // channels, hidden width
  int64_t M;
  int wslots; uint32_t wslot_bytes;
  int wres;                  // the weights of a whole tile fit the ring: loaded once, resident for the kernel's lifetime
  const float* scale1; const float* shift1; float unscale1;
  const float* scale2; const float* shift2; float unscale2;
  int act; float act_param;
  const float* x; int64_t ldx;       // shortcut (FFN: the same tensor as the A operand; MR: the Grapher's input)
  float* y; int64_t ldy;
  int aslots;                        // MR mode: A operand ring slots
};

// one arrival per warp (the barriers count warps): 32 serialised arrivals on one barrier word cost more than the sync
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint32_t ff_pack(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float2 ff_unpack(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}

constexpr int FF_HMAX = 512;         // widest hidden layer (its folded scale / shift are staged in shared memory)

// sc = folded scale * weight un-scale (one fp32 product, formed once at kernel start), sh = folded shift: shared memory
template <int ACT>
__device__ __forceinline__ void ff_act32(float (&v)[32], const float* sc, const float* sh, float act_param) {
#pragma unroll
  for (int q = 0; q < 32; q += 4) {
    const float4 s4 = *reinterpret_cast<const float4*>(sc + q);
    const float4 t4 = *reinterpret_cast<const float4*>(sh + q);
    v[q + 0] = apply_act(fmaf(v[q + 0], s4.x, t4.x), ACT, act_param);
    v[q + 1] = apply_act(fmaf(v[q + 1], s4.y, t4.y), ACT, act_param);
    v[q + 2] = apply_act(fmaf(v[q + 2], s4.z, t4.z), ACT, act_param);
    v[q + 3] = apply_act(fmaf(v[q + 3], s4.w, t4.w), ACT, act_param);
  }
}

template <bool MR>
__global__ void __launch_bounds__(FF_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const FfnParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t raw_full[FF_RAW], raw_empty[FF_RAW];
  __shared__ __align__(8) uint64_t xop_full, xop_free;
  // MR mode: A k-block ring, each slot: TMA lands fp32 (a_raw) -> converted IN PLACE to hi / lo (a_full) -> MMAs retire (a_empty)
  __shared__ __align__(8) uint64_t a_raw[FF_AMAX], a_full[FF_AMAX], a_empty[FF_AMAX];
  __shared__ __align__(8) uint64_t w_full[FF_WMAX], w_empty[FF_WMAX];
  __shared__ __align__(8) uint64_t acc1_full[2], acc1_empty[2];
  __shared__ __align__(8) uint64_t h_full[2], h_empty[2];
  __shared__ __align__(8) uint64_t acc2_full[2], acc2_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_sc1[FF_HMAX], s_sh1[FF_HMAX], s_sc2[128], s_sh2[128];

  const int warp = warp_idx_uniform(), lane = threadIdx.x & 31;
  const int nkb1 = MR ? 2 : p.C / 32;        // k-blocks of GEMM 1 (MR: per chunk, one of x and one of m)
  const int nch = p.Hd / FF_HC;              // hidden chunks
  const int na = MR ? p.aslots : nkb1;       // A operand k-block slots
  const int nraw = MR ? 2 * nch : nkb1;      // fp32 A k-blocks per tile
  const int64_t tiles = (p.M + TC_BM - 1) / TC_BM;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // [ x operand: nkb1 x (hi 8 KB | lo 8 KB) ][ h operand: 2 buffers x 2 k-blocks x (hi | lo) ][ raw ring ][ W ring ]
  uint8_t* xop = smem;
  uint8_t* hop = xop + (size_t)na * 2 * FF_KB_BYTES;
  uint8_t* rawb = hop + (FF_HTMEM ? 0 : 2 * 2 * 2 * FF_KB_BYTES);
  uint8_t* wring = rawb + (MR ? 0 : FF_RAW * TC_A_BYTES);      // MR: no separate raw ring
  uint8_t* stage = wring + (size_t)p.wslots * p.wslot_bytes;   // epilogue 2: 4 warps x 2 KB
  auto x_hi = [&](int kb) { return xop + (size_t)kb * 2 * FF_KB_BYTES; };
  auto h_hi = [&](int buf, int kb) { return hop + ((size_t)buf * 2 + kb) * 2 * FF_KB_BYTES; };
  auto w_slot = [&](int s) { return wring + (size_t)s * p.wslot_bytes; };
  const uint32_t w1_bytes = 2u * FF_HC * 64u;            // hi + lo of one W1 k-block (64 rows x 64 B)
  const uint32_t w2_bytes = 2u * (uint32_t)p.C * 64u;    // hi + lo of one W2 k-block (C rows x 64 B)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    if (MR) tma_prefetch_desc(&tmX2);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int i = 0; i < FF_RAW; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&raw_empty[i], 4); }
    mbar_init(&xop_full, 4);
    mbar_init(&xop_free, 1);
    for (int i = 0; i < FF_AMAX; ++i) { mbar_init(&a_raw[i], 1); mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < FF_WMAX; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 8);
      mbar_init(&h_full[i], 8); mbar_init(&h_empty[i], 1);
      mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], 4);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.Hd; i += FF_THREADS) { s_sc1[i] = p.scale1[i] * p.unscale1; s_sh1[i] = p.shift1[i]; }
  for (int i = threadIdx.x; i < p.C; i += FF_THREADS) { s_sc2[i] = p.scale2[i] * p.unscale2; s_sh2[i] = p.shift2[i]; }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t t_acc1 = tmem_base;                      // 2 x 64 columns
  const uint32_t t_acc2 = tmem_base + 128;                // 2 x C columns (C <= 128)
  const uint32_t t_h = tmem_base + 384;                   // FF_HTMEM: 2 buffers x [hi 32 | lo 32] columns of packed fp16 pairs
  pdl_wait();            // above: barriers, TMEM, the folded (constant) scale / shift vectors -- no tensor access (common.cuh)

  if (warp == 0) {
    // ===== TMA: x rows of my tiles, k-block by k-block, into the raw ring =====
    if (lane == 0) {
      uint32_t r = 0;
      FF_DECL(cy_e = 0, cy_t0 = clock64());
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x)
        for (int kb = 0; kb < nraw; ++kb, ++r) {
          if (MR) {
            // chunk j = kb >> 1 reads columns [32 j, 32 j + 32) of x (even kb) and of m (odd kb)
            const int s = (int)(r % (uint32_t)na);
            FF_ACC(cy_e, mbar_wait(&a_empty[s], ((r / (uint32_t)na) & 1u) ^ 1u));
            mbar_arrive_expect_tx(&a_raw[s], TC_A_BYTES);
            tma_load_2d(x_hi(s), (kb & 1) ? &tmX2 : &tmX, (kb >> 1) * 32, (int)(tile * TC_BM), &a_raw[s]);
            continue;
          }
          const int s = r % FF_RAW;
          mbar_wait(&raw_empty[s], ((r / FF_RAW) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&raw_full[s], TC_A_BYTES);
          tma_load_2d(rawb + (size_t)s * TC_A_BYTES, &tmX, kb * 32, (int)(tile * TC_BM), &raw_full[s]);
        }
      FF_OUT(8, clock64() - cy_t0); FF_OUT(9, cy_e);
    }
  } else if (warp == 2) {
    // ===== TMA: weight k-blocks in the MMA's consumption order =====
    if (lane == 0) {
      uint32_t w = 0;
      int tr_n = 0; (void)tr_n;
      auto load_w1 = [&](int j) {
        for (int kb = 0; kb < nkb1; ++kb, ++w) {
          const int s = w % p.wslots;
          mbar_wait(&w_empty[s], ((w / p.wslots) & 1u) ^ 1u);
          FF_T(2, 30);
          mbar_arrive_expect_tx(&w_full[s], w1_bytes);
          tma_load_3d(w_slot(s), &tmW1, kb * 32, j * FF_HC, 0, &w_full[s]);       // hi and lo tile in one request
        }
      };
      auto load_w2 = [&](int j) {
        for (int kb = 0; kb < FF_HC / 32; ++kb, ++w) {
          const int s = w % p.wslots;
          mbar_wait(&w_empty[s], ((w / p.wslots) & 1u) ^ 1u);
          FF_T(2, 31);
          mbar_arrive_expect_tx(&w_full[s], w2_bytes);
          tma_load_3d(w_slot(s), &tmW2, j * FF_HC + kb * 32, 0, 0, &w_full[s]);
        }
      };
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        for (int j = 0; j < nch; ++j) {
          load_w1(j);
          if (j >= 1) load_w2(j - 1);
        }
        load_w2(nch - 1);
        if (p.wres) break;          // one tile's worth of k-blocks is all of W1 and W2: they stay in their slots
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane issues (tc_common.cuh) =====
    {
      const uint32_t idesc1 = umma_idesc_f16(TC_BM, FF_HC), idesc2 = umma_idesc_f16(TC_BM, p.C);
      const uint32_t d_x0 = umma_desc_lo(smem_u32(xop)), d_h0 = umma_desc_lo(smem_u32(hop)), d_w0 = umma_desc_lo(smem_u32(wring));
      constexpr uint32_t d_kb = (2 * FF_KB_BYTES) >> 4, d_lo = FF_KB_BYTES >> 4;     // k-block stride, hi -> lo plane
      const uint32_t d_wslot = p.wslot_bytes >> 4, d_w1lo = (FF_HC * 64) >> 4, d_w2lo = ((uint32_t)p.C * 64u) >> 4;
      uint32_t ws = 0, wph = 0, c1 = 0, hcnt = 0, ti = 0;     // W slot + phase, GEMM-1 chunks, GEMM-2 chunks, tiles
      uint32_t as = 0, aph = 0;                               // MR: A ring slot + phase
      long long cy_w = 0, cy_h = 0, cy_a = 0, cy_x = 0, cy_2 = 0; const long long cy_t0 = clock64(); (void)cy_t0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
        const uint32_t a2 = t_acc2 + (ti & 1u) * (uint32_t)p.C;
        auto gemm = [&](uint32_t tacc, uint32_t da0, int nkb, uint32_t d_wlo, uint32_t idesc, bool fresh, uint64_t* done) {
          for (int kb = 0; kb < nkb; ++kb) {
            if (!p.wres || ti == 0) FF_ACC(cy_w, FF_WAIT(&w_full[ws], wph));
            tc_fence_after();
            const uint32_t dah = da0 + kb * d_kb, dal = dah + d_lo, dbh = d_w0 + ws * d_wslot, dbl = dbh + d_wlo;
            if (elect_one()) {
#pragma unroll
              for (uint32_t k = 0; k < 4; k += 2) {
                umma_f16_lh(tacc, dal + k, dbh + k, UMMA_HI_SW64, idesc, (!fresh || kb > 0 || k > 0) ? 1u : 0u);
                umma_f16_lh(tacc, dah + k, dbl + k, UMMA_HI_SW64, idesc, 1u);
                umma_f16_lh(tacc, dah + k, dbh + k, UMMA_HI_SW64, idesc, 1u);
              }
              if (!p.wres) umma_commit(&w_empty[ws]);
              if (kb == nkb - 1) umma_commit(done);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.wslots) { ws = 0; wph ^= 1u; }
          }
        };
        auto gemm1 = [&](int j) {
          const uint32_t b = c1 & 1u;
          FF_ACC(cy_a, FF_WAIT(&acc1_empty[b], ((c1 >> 1) & 1u) ^ 1u));
          tc_fence_after();
          if (MR) {
            // the chunk's two A k-blocks come off the ring and are released with the weight slot
            const uint32_t tacc = t_acc1 + b * FF_HC;
            for (int kb = 0; kb < 2; ++kb) {
              if (!p.wres || ti == 0) FF_ACC(cy_w, FF_WAIT(&w_full[ws], wph));
              FF_ACC(cy_x, FF_WAIT(&a_full[as], aph));
              tc_fence_after();
              const uint32_t dah = d_x0 + as * d_kb, dal = dah + d_lo, dbh = d_w0 + ws * d_wslot, dbl = dbh + d_w1lo;
              if (elect_one()) {
#pragma unroll
                for (uint32_t k = 0; k < 4; k += 2) {
                  umma_f16_lh(tacc, dal + k, dbh + k, UMMA_HI_SW64, idesc1, (kb > 0 || k > 0) ? 1u : 0u);
                  umma_f16_lh(tacc, dah + k, dbl + k, UMMA_HI_SW64, idesc1, 1u);
                  umma_f16_lh(tacc, dah + k, dbh + k, UMMA_HI_SW64, idesc1, 1u);
                }
                if (!p.wres) umma_commit(&w_empty[ws]);
                umma_commit(&a_empty[as]);
                if (kb == 1) umma_commit(&acc1_full[b]);
              }
              __syncwarp();
              if (++ws == (uint32_t)p.wslots) { ws = 0; wph ^= 1u; }
              if (++as == (uint32_t)na) { as = 0; aph ^= 1u; }
            }
          } else {
            gemm(t_acc1 + b * FF_HC, d_x0, nkb1, d_w1lo, idesc1, true, &acc1_full[b]);
          }
          ++c1;
        };
        auto gemm2 = [&](int j) {
          const uint32_t b = hcnt & 1u;
          FF_ACC(cy_h, FF_WAIT(&h_full[b], (hcnt >> 1) & 1u));
          if (j == 0) FF_ACC(cy_2, FF_WAIT(&acc2_empty[ti & 1u], ((ti >> 1) & 1u) ^ 1u));
          tc_fence_after();
#if FF_HTMEM
          for (int kb = 0; kb < FF_HC / 32; ++kb) {
            if (!p.wres || ti == 0) FF_ACC(cy_w, FF_WAIT(&w_full[ws], wph));
            tc_fence_after();
            const uint32_t tah = t_h + b * 64u + (uint32_t)kb * 16u, tal = tah + 32u;
            const uint32_t dbh = d_w0 + ws * d_wslot, dbl = dbh + d_w2lo;
            if (elect_one()) {
#pragma unroll
              for (uint32_t k = 0; k < 2; ++k) {
                umma_f16_ts(a2, tal + 8u * k, dbh + 2u * k, UMMA_HI_SW64, idesc2, (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
                umma_f16_ts(a2, tah + 8u * k, dbl + 2u * k, UMMA_HI_SW64, idesc2, 1u);
                umma_f16_ts(a2, tah + 8u * k, dbh + 2u * k, UMMA_HI_SW64, idesc2, 1u);
              }
              if (!p.wres) umma_commit(&w_empty[ws]);
              if (kb == FF_HC / 32 - 1) umma_commit(&h_empty[b]);
            }
            __syncwarp();
            if (++ws == (uint32_t)p.wslots) { ws = 0; wph ^= 1u; }
          }
#else
          gemm(a2, d_h0 + b * 2 * d_kb, FF_HC / 32, d_w2lo, idesc2, j == 0, &h_empty[b]);
#endif
          ++hcnt;
        };
        if (!MR) {
          FF_ACC(cy_x, FF_WAIT(&xop_full, ti & 1u));
          tc_fence_after();
        }
        for (int j = 0; j < nch; ++j) {
          gemm1(j);
          if (!MR && j == nch - 1) {       // every GEMM 1 of this tile issued: x operand reusable when they retire
            if (elect_one()) umma_commit(&xop_free);
            __syncwarp();
          }
          if (j >= 1) gemm2(j - 1);
        }
        gemm2(nch - 1);
        if (elect_one()) umma_commit(&acc2_full[ti & 1u]);
        __syncwarp();
      }
#ifdef FF_TRACE
      if (blockIdx.x == 0 && lane == 0) {
        g_ff_trace[0] = clock64() - cy_t0; g_ff_trace[1] = cy_w; g_ff_trace[2] = cy_h; g_ff_trace[3] = cy_a;
        g_ff_trace[4] = cy_x; g_ff_trace[5] = cy_2; g_ff_trace[6] = ti;
      }
#endif
    }
  } else if (warp >= 8 && warp < 12) {
    // ===== transform: fp32 x k-block (128B-swizzled rows) -> fp16 hi / lo operand k-block (64B-swizzled rows) =====
    const int t = threadIdx.x - 256;
    uint32_t r = 0, ti = 0;
    int tr_n = 1 << 20; (void)tr_n;
    FF_DECL(cy_r = 0, cy_t0 = clock64());
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      for (int kb = 0; kb < nraw; ++kb, ++r) {
        const int s = r % FF_RAW;
        const int slot = MR ? (int)(r % (uint32_t)na) : kb;
        if (MR) FF_ACC(cy_r, mbar_wait(&a_raw[slot], (r / (uint32_t)na) & 1u));
        else FF_ACC(cy_r, mbar_wait(&raw_full[s], (r / FF_RAW) & 1u));
        FF_T(2, 20);
        const float4* raw = reinterpret_cast<const float4*>(MR ? x_hi(slot) : rawb + (size_t)s * TC_A_BYTES);
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = raw[t + 128 * i];
        if (MR) {
          named_bar_sync(1, 128);                 // in place: every transform thread has read before anyone writes
        } else {
          fence_proxy_async_smem();
          warp_arrive(&raw_empty[s], lane);
          if (kb == 0 && ti > 0) mbar_wait(&xop_free, (ti - 1) & 1u);      // the previous tile's GEMM 1s have retired
        }
        FF_T(2, 21);
        uint8_t* hi = x_hi(slot);
        uint8_t* lo = hi + FF_KB_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int q = t + 128 * i;
          const int row = q >> 3;
          const int lc = (q & 7) ^ (row & 7);                          // logical 4-float chunk 0..7
          const uint32_t dst = (uint32_t)row * 64u + ((uint32_t)((lc >> 1) ^ ((row >> 1) & 3)) << 4) + ((uint32_t)(lc & 1) << 3);
          uint2 hv, lv;
          hv.x = ff_pack(v[i].x, v[i].y);
          hv.y = ff_pack(v[i].z, v[i].w);
          const float2 f01 = ff_unpack(hv.x), f23 = ff_unpack(hv.y);
          lv.x = ff_pack(v[i].x - f01.x, v[i].y - f01.y);
          lv.y = ff_pack(v[i].z - f23.x, v[i].w - f23.y);
          *reinterpret_cast<uint2*>(hi + dst) = hv;
          *reinterpret_cast<uint2*>(lo + dst) = lv;
        }
        if (MR) {
          fence_proxy_async_smem();
          warp_arrive(&a_full[slot], lane);
        }
      }
      if (!MR) {
        fence_proxy_async_smem();
        warp_arrive(&xop_full, lane);
      }
      FF_T(2, 22);
    }
    if (warp == 8) { FF_OUT(10, clock64() - cy_t0); FF_OUT(11, cy_r); }
  } else if (warp >= 12) {
    // ===== epilogue 1: hidden chunk -> s1 / t1 / activation -> fp16 hi / lo operand of GEMM 2 =====
    // Warp (parity, half, quad) converts columns [32 half, 32 half + 32) -- k-block `half` of the h operand -- of the
    // chunks whose accumulator buffer is `parity`, for TMEM lanes [32 quad, 32 quad + 32).
    const int quad = warp & 3;
    const int sub = (warp - 12) >> 2;
    const uint32_t mine = sub & 1, half = sub >> 1;
    const int row = quad * 32 + lane;
    uint32_t c1 = 0;
    int tr_n = (warp == 12 && lane == 0) ? 0 : 1 << 20; (void)tr_n;
    FF_DECL(cy_f = 0, cy_he = 0, cy_ld = 0, cy_act = 0, cy_pk = 0, cy_fn = 0, cy_t0 = clock64());
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      for (int j = 0; j < nch; ++j, ++c1) {
        const uint32_t b = c1 & 1u;
        if (b != mine) continue;
        const float* sc = s_sc1 + j * FF_HC + half * 32;
        const float* sh = s_sh1 + j * FF_HC + half * 32;
        FF_ACC(cy_f, FF_WAIT(&acc1_full[b], (c1 >> 1) & 1u));
        tc_fence_after();
        FF_T(1, 10);
        const uint32_t tacc = t_acc1 + b * FF_HC + half * 32u + ((uint32_t)(quad * 32) << 16);
        float v[32];
        FF_ACC(cy_ld, tmem_ld16_nowait(tacc, v);
        tmem_ld16_nowait(tacc + 16u, v + 16);
        tmem_ld_wait();
        tc_fence_before();
        warp_arrive(&acc1_empty[b], lane));
        FF_T(1, 11);
        FF_DECL(c_a0 = clock64());
        switch (p.act) {
          case GRAFP_ACT_NONE:  ff_act32<GRAFP_ACT_NONE>(v, sc, sh, p.act_param); break;
          case GRAFP_ACT_RELU:  ff_act32<GRAFP_ACT_RELU>(v, sc, sh, p.act_param); break;
          case GRAFP_ACT_LEAKY: ff_act32<GRAFP_ACT_LEAKY>(v, sc, sh, p.act_param); break;
          case GRAFP_ACT_GELU:  ff_act32<GRAFP_ACT_GELU>(v, sc, sh, p.act_param); break;
          default:              ff_act32<GRAFP_ACT_ELU>(v, sc, sh, p.act_param); break;
        }
        FF_T(1, 12);
#ifdef FF_TRACE
        cy_act += clock64() - c_a0;
#endif
        FF_ACC(cy_he, FF_WAIT(&h_empty[b], ((c1 >> 1) & 1u) ^ 1u));  // GEMM 2 of chunk c1 - 2 has read this buffer
        FF_T(1, 13);
        FF_DECL(c_p0 = clock64());
#if FF_HTMEM
        tc_fence_after();
        // this warp's 32 hidden columns = 16 packed columns of the hi plane and 16 of the lo plane, rows = its TMEM lanes
        const uint32_t th = t_h + b * 64u + half * 16u + ((uint32_t)(quad * 32) << 16);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t hp[8], lp[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            hp[e] = ff_pack(v[16 * q + 2 * e], v[16 * q + 2 * e + 1]);
            const float2 f = ff_unpack(hp[e]);
            lp[e] = ff_pack(v[16 * q + 2 * e] - f.x, v[16 * q + 2 * e + 1] - f.y);
          }
          tmem_st8(th + 8u * q, hp);
          tmem_st8(th + 32u + 8u * q, lp);
        }
#ifdef FF_TRACE
        cy_pk += clock64() - c_p0;
#endif
        FF_ACC(cy_fn, tmem_st_wait();
        tc_fence_before();
        warp_arrive(&h_full[b], lane));
#else
        uint8_t* hi = h_hi(b, half);
        uint8_t* lo = hi + FF_KB_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t hp[4], lp[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hp[e] = ff_pack(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
            const float2 f = ff_unpack(hp[e]);
            lp[e] = ff_pack(v[8 * q + 2 * e] - f.x, v[8 * q + 2 * e + 1] - f.y);
          }
          const uint32_t off = (uint32_t)row * 64u + ((uint32_t)(q ^ ((row >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(hi + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
          *reinterpret_cast<uint4*>(lo + off) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
        }
#ifdef FF_TRACE
        cy_pk += clock64() - c_p0;
#endif
        FF_ACC(cy_fn, fence_proxy_async_smem();
        warp_arrive(&h_full[b], lane));
#endif
        FF_T(1, 14);
      }
    }
    if (warp == 12) {
      FF_OUT(12, clock64() - cy_t0); FF_OUT(13, cy_f); FF_OUT(14, cy_he);
      FF_OUT(17, cy_ld); FF_OUT(18, cy_act); FF_OUT(19, cy_pk); FF_OUT(20, cy_fn);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===== epilogue 2: output tile -> s2 / t2 + shortcut -> y =====
    // A thread's TMEM lane is one tile row, but a row-per-thread global access touches 32 lines per instruction and
    // the L1 pipe (shared with the operand conversions) becomes the limiter.  So every 32 x 16 block goes through a
    // 2 KB per-warp staging block (64 B rows, 16 B chunks XOR-swizzled: conflict-free both ways) and the shortcut loads
    // / output stores are issued with 4 lanes per row (64 B runs, 8 rows per instruction).
    const int quad = warp & 3;
    uint8_t* stg = stage + quad * 2048;
    const int lr = lane >> 2, lq = lane & 3;                 // coalesced form: row within a group of 8, 16 B chunk
    uint32_t ti = 0;
    FF_DECL(cy_f2 = 0, cy_tm = 0, cy_st = 0, cy_out = 0, cy_pf = 0, cy_t0 = clock64());
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++ti) {
      const int64_t row0 = tile * TC_BM + quad * 32;         // first row of this warp's 32
      if (MR) {
        // the shortcut is a tensor nobody else on this SM reads: bring this warp's rows of the NEXT tile into L2 now
        const int64_t nrow0 = row0 + (int64_t)gridDim.x * TC_BM;
        const int lpr = p.C / 32;                            // 128 B lines per row
        for (int l = lane; l < 32 * lpr; l += 32) {
          const int64_t rr = nrow0 + l / lpr;
          if (rr < p.M) prefetch_l2(p.x + rr * p.ldx + (l % lpr) * 32);
        }
      }
      FF_ACC(cy_f2, FF_WAIT(&acc2_full[ti & 1u], (ti >> 1) & 1u));
      tc_fence_after();
      const uint32_t tacc = t_acc2 + (ti & 1u) * (uint32_t)p.C + ((uint32_t)(quad * 32) << 16);
      for (int c = 0; c < p.C; c += 16) {
        float4 res[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int64_t rr = row0 + it * 8 + lr;
          res[it] = rr < p.M ? __ldg(reinterpret_cast<const float4*>(p.x + rr * p.ldx + c + 4 * lq)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float4 s4 = *reinterpret_cast<const float4*>(s_sc2 + c + 4 * lq);
        const float4 t4 = *reinterpret_cast<const float4*>(s_sh2 + c + 4 * lq);
        float v[16];
        FF_ACC(cy_tm, tmem_ld16_nowait(tacc + (uint32_t)c, v);
        tmem_ld_wait());
        if (c + 16 >= p.C) {
          tc_fence_before();
          warp_arrive(&acc2_empty[ti & 1u], lane);
        }
        FF_DECL(c_s0 = clock64());
        // raw accumulators through the staging block; scale / shift are applied after the transpose, where a thread
        // owns 4 fixed columns (2 shared-memory loads per block instead of 8 on the TMEM-load critical path)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(stg + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) =
              make_float4(v[4 * q + 0], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
#ifdef FF_TRACE
        cy_st += clock64() - c_s0;
        const long long c_o0 = clock64();
#endif
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 8 + lr;
          const int64_t rr = row0 + rl;
          float4 o = *reinterpret_cast<const float4*>(stg + rl * 64 + ((lq ^ ((rl >> 1) & 3)) << 4));
          o.x = fmaf(o.x, s4.x, t4.x) + res[it].x;
          o.y = fmaf(o.y, s4.y, t4.y) + res[it].y;
          o.z = fmaf(o.z, s4.z, t4.z) + res[it].z;
          o.w = fmaf(o.w, s4.w, t4.w) + res[it].w;
          if (rr < p.M) *reinterpret_cast<float4*>(p.y + rr * p.ldy + c + 4 * lq) = o;
        }
        __syncwarp();
#ifdef FF_TRACE
        cy_out += clock64() - c_o0;
#endif
      }
    }
    if (warp == 4) { FF_OUT(15, clock64() - cy_t0); FF_OUT(16, cy_f2); FF_OUT(21, cy_tm); FF_OUT(22, cy_st); FF_OUT(23, cy_out); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace grafp

using namespace grafp;

#ifdef FF_TRACE
extern "C" int grafp_debug_ffn_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ff_trace, sizeof(unsigned long long) * 3 * 1024);
}
#endif

extern "C" int grafp_ffn_fused_supported(int64_t M, int C, int Hd) {
  return M >= 1 && (C == 64 || C == 128) && Hd % FF_HC == 0 && Hd >= FF_HC && Hd <= FF_HMAX;
}

extern "C" int grafp_ffn_fused_fwd(const float* x, int64_t ldx, int64_t M, int C, int Hd, const void* w1_split_f16,
                                   int64_t ldw1, float w1_unscale, const float* scale1, const float* shift1, int act,
                                   float act_param, const void* w2_split_f16, int64_t ldw2, float w2_unscale,
                                   const float* scale2, const float* shift2, float* y, int64_t ldy, void* stream) {
  GRAFP_REQUIRE(grafp_ffn_fused_supported(M, C, Hd), "ffn_fused: needs C in {64, 128} and a hidden width that is a multiple of 64, at most 512");
  GRAFP_REQUIRE(x && y && w1_split_f16 && w2_split_f16 && scale1 && shift1 && scale2 && shift2, "ffn_fused: null pointer");
  GRAFP_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ldw1 % 8 == 0 && ldw2 % 8 == 0 && w1_unscale > 0.0f && w2_unscale > 0.0f,
                "ffn_fused: bad strides / scales");
  GRAFP_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w1_split_f16) |
                  reinterpret_cast<uintptr_t>(w2_split_f16)) & 15) == 0, "ffn_fused: operands must be 16-byte aligned");
  GRAFP_REQUIRE(act >= GRAFP_ACT_NONE && act <= GRAFP_ACT_ELU, "ffn_fused: unsupported activation %d", act);
  CUtensorMap mX, mW1, mW2;
  if (int rc = tc_make_map_2d(&mX, x, M, C, ldx, TC_BM)) return rc;
  // weights as (k, rows, 2 planes): one TMA request brings the hi and the lo tile of a k-block (a request costs the issuing
  // thread ~400 cycles whatever its size, and the weight stream is 16-48 k-blocks per 128-row tile)
  if (int rc = tc_make_map_3d_bf16(&mW1, w1_split_f16, C, Hd, 2, ldw1, (int64_t)Hd * ldw1, FF_HC, 2)) return rc;
  if (int rc = tc_make_map_3d_bf16(&mW2, w2_split_f16, Hd, C, 2, ldw2, (int64_t)C * ldw2, C, 2)) return rc;
  FfnParams p;
  p.C = C; p.Hd = Hd; p.M = M;
  p.scale1 = scale1; p.shift1 = shift1; p.unscale1 = w1_unscale;
  p.scale2 = scale2; p.shift2 = shift2; p.unscale2 = w2_unscale;
  p.act = act; p.act_param = act_param;
  p.x = x; p.ldx = ldx; p.y = y; p.ldy = ldy;
  const uint32_t w1b = 2u * FF_HC * 64u, w2b = 2u * (uint32_t)C * 64u;
  p.wslot_bytes = w1b > w2b ? w1b : w2b;
  const size_t fixed = (size_t)(C / 32) * 2 * FF_KB_BYTES + FF_HOP_BYTES + FF_RAW * TC_A_BYTES + FF_STAGE_BYTES;
  int slots = (int)((FF_SMEM_BUDGET - fixed) / p.wslot_bytes);
  if (slots > FF_WMAX) slots = FF_WMAX;
  GRAFP_REQUIRE(slots >= 2, "ffn_fused: not enough shared memory");
  const int kb_per_tile = (Hd / FF_HC) * (C / 32 + FF_HC / 32);       // W1 + W2 k-blocks a tile consumes
  p.wres = kb_per_tile <= slots;
  if (p.wres) slots = kb_per_tile;
  p.wslots = slots;
  p.aslots = 0;
  const size_t smem = fixed + (size_t)slots * p.wslot_bytes + 1024;
  const int64_t tiles = (M + TC_BM - 1) / TC_BM;
  int grid = sm_count();
  if (tiles < grid) grid = (int)tiles;
  cudaFuncSetAttribute(ffn_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_ex(ffn_fused_kernel<false>, dim3(grid), dim3(FF_THREADS), smem, as_stream(stream), 0, mX, mX, mW1, mW2, p);
  return check_launch("ffn_fused");
}

extern "C" int grafp_mrconv_fc2_fused_supported(int64_t M, int C) {
  return M >= 1 && (C == 64 || C == 128);
}

extern "C" int grafp_mrconv_fc2_fused_fwd(const float* x, int64_t ldx, const float* m, int64_t ldm, int64_t M, int C,
                                          const void* w1_chunked_f16, int64_t ldw1, float w1_unscale, const float* scale1,
                                          const float* shift1, int act, float act_param, const void* w2_split_f16,
                                          int64_t ldw2, float w2_unscale, const float* scale2, const float* shift2,
                                          const float* res, int64_t ldr, float* y, int64_t ldy, void* stream) {
  GRAFP_REQUIRE(grafp_mrconv_fc2_fused_supported(M, C), "mrconv_fc2_fused: needs C in {64, 128}");
  GRAFP_REQUIRE(x && m && res && y && w1_chunked_f16 && w2_split_f16 && scale1 && shift1 && scale2 && shift2,
                "mrconv_fc2_fused: null pointer");
  GRAFP_REQUIRE(ldx % 4 == 0 && ldm % 4 == 0 && ldr % 4 == 0 && ldy % 4 == 0 && ldw1 % 8 == 0 && ldw2 % 8 == 0 &&
                w1_unscale > 0.0f && w2_unscale > 0.0f, "mrconv_fc2_fused: bad strides / scales");
  GRAFP_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(res) |
                  reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w1_chunked_f16) |
                  reinterpret_cast<uintptr_t>(w2_split_f16)) & 15) == 0, "mrconv_fc2_fused: operands must be 16-byte aligned");
  GRAFP_REQUIRE(act >= GRAFP_ACT_NONE && act <= GRAFP_ACT_ELU, "mrconv_fc2_fused: unsupported activation %d", act);
  const int Hd = 2 * C;
  CUtensorMap mX, mM, mW1, mW2;
  if (int rc = tc_make_map_2d(&mX, x, M, C, ldx, TC_BM)) return rc;
  if (int rc = tc_make_map_2d(&mM, m, M, C, ldm, TC_BM)) return rc;
  // W1 (2 planes, 2C rows, 64 columns): rows of chunk j = [64 j, 64 j + 64), columns [x part 32 | m part 32]
  if (int rc = tc_make_map_3d_bf16(&mW1, w1_chunked_f16, 64, Hd, 2, ldw1, (int64_t)Hd * ldw1, FF_HC, 2)) return rc;
  if (int rc = tc_make_map_3d_bf16(&mW2, w2_split_f16, Hd, C, 2, ldw2, (int64_t)C * ldw2, C, 2)) return rc;
  FfnParams p;
  p.C = C; p.Hd = Hd; p.M = M;
  p.scale1 = scale1; p.shift1 = shift1; p.unscale1 = w1_unscale;
  p.scale2 = scale2; p.shift2 = shift2; p.unscale2 = w2_unscale;
  p.act = act; p.act_param = act_param;
  p.x = res; p.ldx = ldr; p.y = y; p.ldy = ldy;
  const uint32_t w1b = 2u * FF_HC * 64u, w2b = 2u * (uint32_t)C * 64u;
  p.wslot_bytes = w1b > w2b ? w1b : w2b;
  p.aslots = FF_HTMEM ? 6 : 5;
  const size_t fixed = (size_t)p.aslots * 2 * FF_KB_BYTES + FF_HOP_BYTES + FF_STAGE_BYTES;
  int slots = (int)((FF_SMEM_BUDGET - fixed) / p.wslot_bytes);
  if (slots > FF_WMAX) slots = FF_WMAX;
  GRAFP_REQUIRE(slots >= 2, "mrconv_fc2_fused: not enough shared memory");
  const int kb_per_tile = (Hd / FF_HC) * (2 + FF_HC / 32);
  p.wres = kb_per_tile <= slots;
  if (p.wres) slots = kb_per_tile;
  p.wslots = slots;
  const size_t smem = fixed + (size_t)slots * p.wslot_bytes + 1024;
  const int64_t tiles = (M + TC_BM - 1) / TC_BM;
  int grid = sm_count();
  if (tiles < grid) grid = (int)tiles;
  cudaFuncSetAttribute(ffn_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  launch_ex(ffn_fused_kernel<true>, dim3(grid), dim3(FF_THREADS), smem, as_stream(stream), 0, mX, mM, mW1, mW2, p);
  return check_launch("mrconv_fc2_fused");
}
