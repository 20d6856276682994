// tcgen05 GEMM engine (sm_100a): TMA-fed, TMEM-accumulated, warp-specialised.
//
//   y[128 x BN tile] = epilogue( A[128 x K] * W[BN x K]^T )      A, W fp32 row-major (K-major)
//
// Precision modes
//   passes = 1 : one kind::tf32 MMA per k-step on the raw fp32 operands (TF32 accuracy)
//   passes = 3 : error-compensated "3xTF32": A and W are split hi = tf32(v), lo = tf32(v - hi);
//                D += A_lo*W_hi + A_hi*W_lo + A_hi*W_hi  -> fp32-class accuracy (~1e-6 rel).
//                W is pre-split on the host (stacked [hi; lo] matrix, done once per weight
//                version); A is split in shared memory by the transform warps between the TMA
//                arrival and the MMA issue, so activations cross HBM once, as plain fp32.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA
// issuer, warps 2-5 = A-split transform during the main loop, then the epilogue (TMEM -> regs ->
// scale/shift/activation/residual -> global), one TMEM lane quadrant per warp.
// Pipelines: full[s] (TMA -> transform), xf[s] (transform -> MMA), empty[s] (MMA commit -> TMA),
// tmem_full (last MMA commit -> epilogue).
#include <cuda.h>
#include <mutex>
#include "common.cuh"

namespace grafp {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // fp32 elements = 128 B = one swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_MAX_STAGES = 6;
constexpr int TC_THREADS = 192;

struct TcParams {
  int k1, k2;          // per-group k extents of the two A sources (multiples of 32)
  int n;               // per-group output columns
  int bn;              // tile width (multiple of 16, divides n)
  int n_total;         // groups * n (row offset of the lo copy inside the stacked W)
  int64_t m;
  int stages;
  const float* scale; const float* shift;
  const float* residual; int64_t ldr;
  float* y; int64_t ldy;
  int act; float act_param;
  uint32_t tmem_cols;
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows *
// 128 B = 1024) | version=1 [46,48) | layout_type=2 (SWIZZLE_128B) [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor: D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2, both
// K-major, N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

template <int kPasses>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t xf_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.z;
  const int m0 = blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * p.bn;
  const int S = p.stages;
  const uint32_t b_bytes = (uint32_t)p.bn * TC_BK * 4;
  const uint32_t stage_bytes = (kPasses == 3 ? 2u : 1u) * (TC_A_BYTES + b_bytes);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  // stage layout: [A_hi | A_lo (3x) | B_hi | B_lo (3x)]
  auto a_hi = [&](int s) { return smem + (size_t)s * stage_bytes; };
  auto a_lo = [&](int s) { return smem + (size_t)s * stage_bytes + TC_A_BYTES; };
  auto b_hi = [&](int s) { return smem + (size_t)s * stage_bytes + (kPasses == 3 ? 2 : 1) * TC_A_BYTES; };
  auto b_lo = [&](int s) { return b_hi(s) + b_bytes; };

  const int nkb = (p.k1 + p.k2) / TC_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&xf_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], TC_A_BYTES + (kPasses == 3 ? 2u : 1u) * b_bytes);
        const int k = kb * TC_BK;
        if (k < p.k1) tma_load_2d(a_hi(s), &tmA1, g * p.k1 + k, m0, &full_bar[s]);
        else          tma_load_2d(a_hi(s), &tmA2, g * p.k2 + (k - p.k1), m0, &full_bar[s]);
        tma_load_2d(b_hi(s), &tmW, k, g * p.n + n0, &full_bar[s]);
        if (kPasses == 3) tma_load_2d(b_lo(s), &tmW, k, p.n_total + g * p.n + n0, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(TC_BM, p.bn);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(kPasses == 3 ? &xf_bar[s] : &full_bar[s], ph);
        tc_fence_after();
        const uint64_t dah = umma_desc_sw128(smem_u32(a_hi(s)));
        const uint64_t dbh = umma_desc_sw128(smem_u32(b_hi(s)));
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {
          const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);       // 32 B per k-step
          const uint32_t first = (kb > 0 || k > 0) ? 1u : 0u;
          if (kPasses == 3) {
            const uint64_t dal = umma_desc_sw128(smem_u32(a_lo(s)));
            const uint64_t dbl = umma_desc_sw128(smem_u32(b_lo(s)));
            umma_tf32(tmem_base, dal + koff, dbh + koff, idesc, first);
            umma_tf32(tmem_base, dah + koff, dbl + koff, idesc, 1u);
            umma_tf32(tmem_base, dah + koff, dbh + koff, idesc, 1u);
          } else {
            umma_tf32(tmem_base, dah + koff, dbh + koff, idesc, first);
          }
        }
        umma_commit(&empty_bar[s]);           // smem slot reusable once these MMAs retire
      }
      umma_commit(&tmem_full_bar);            // accumulator complete
    }
  } else {
    // ===== transform (A split) then epilogue: warps 2..5, 128 threads =====
    const int t = threadIdx.x - 64;
    if (kPasses == 3) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(&full_bar[s], ph);
        float4* hi = reinterpret_cast<float4*>(a_hi(s));
        float4* lo = reinterpret_cast<float4*>(a_lo(s));
#pragma unroll
        for (int i = 0; i < TC_A_BYTES / 16 / 128; ++i) {
          const float4 v = hi[t + 128 * i];
          float4 h, l;
          h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
          l.x = to_tf32(v.x - h.x); l.y = to_tf32(v.y - h.y);
          l.z = to_tf32(v.z - h.z); l.w = to_tf32(v.w - h.w);
          hi[t + 128 * i] = h;
          lo[t + 128 * i] = l;
        }
        fence_proxy_async_smem();             // generic-proxy writes -> visible to the MMA (async proxy)
        mbar_arrive(&xf_bar[s]);
      }
    }
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    const int quad = warp & 3;                // TMEM lane quadrant this warp may access
    const int64_t row = (int64_t)m0 + quad * 32 + lane;
    const bool row_ok = row < p.m;
    const int64_t col0 = (int64_t)g * p.n + n0;
    for (int c = 0; c < p.bn; c += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, v);
      if (row_ok) {
        float* dst = p.y + row * p.ldy + col0 + c;
        const float* res = p.residual ? p.residual + row * p.ldr + col0 + c : nullptr;
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t col = col0 + c + q + j;
            const float sc = p.scale ? __ldg(p.scale + col) : 1.0f;
            const float sh = p.shift ? __ldg(p.shift + col) : 0.0f;
            o[j] = apply_act(fmaf(v[q + j], sc, sh), p.act, p.act_param);
          }
          if (res) {
            const float4 r4 = *reinterpret_cast<const float4*>(res + q);
            o[0] += r4.x; o[1] += r4.y; o[2] += r4.z; o[3] += r4.w;
          }
          *reinterpret_cast<float4*>(dst + q) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---- host side -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// 2-D fp32 row-major (rows, cols) with row stride ld elements; box = (32 cols, box_rows), 128B swizzle
static int make_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                       int box_rows) {
  EncodeTiledFn fn = encode_fn();
  GRAFP_REQUIRE(fn, "gemm_tc: cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GRAFP_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

static int pick_bn(int n) {
  for (int bn = 256; bn >= 16; bn -= 16)
    if (n % bn == 0) return bn;
  return 0;
}

int gemm_tc_supported(const grafp_gemm_args& a) {
  if (a.tap3_nodes > 0) return 0;
  if (a.k1 % TC_BK != 0 || a.k2 % TC_BK != 0) return 0;
  if (a.n % 16 != 0 || pick_bn(a.n) == 0) return 0;
  if (a.m < 1) return 0;
  if ((a.lda1 * 4) % 16 != 0 || (a.k2 && (a.lda2 * 4) % 16 != 0) || (a.ldw * 4) % 16 != 0) return 0;
  if (a.ldy % 4 != 0 || (reinterpret_cast<uintptr_t>(a.y) & 15)) return 0;
  if (a.residual && (a.ldr % 4 != 0 || (reinterpret_cast<uintptr_t>(a.residual) & 15))) return 0;
  return 1;
}

// For passes == 3 the weight pointer must address the stacked [W_hi ; W_lo] matrix
// (2 * groups * n rows); grafp_split_tf32 builds it.
int gemm_tc_launch(const grafp_gemm_args& a, int passes, cudaStream_t st) {
  const int bn = pick_bn(a.n);
  const int n_total = a.groups * a.n;
  CUtensorMap mA1, mA2, mW;
  if (int rc = make_map_2d(&mA1, a.a1, a.m, (int64_t)a.groups * a.k1, a.lda1, TC_BM)) return rc;
  if (a.k2 > 0) {
    if (int rc = make_map_2d(&mA2, a.a2, a.m, (int64_t)a.groups * a.k2, a.lda2, TC_BM)) return rc;
  } else {
    mA2 = mA1;
  }
  if (int rc = make_map_2d(&mW, passes == 3 ? a.w_split : a.w, (int64_t)n_total * (passes == 3 ? 2 : 1), a.k1 + a.k2, a.ldw, bn))
    return rc;
  TcParams p;
  p.k1 = a.k1; p.k2 = a.k2; p.n = a.n; p.bn = bn; p.n_total = n_total; p.m = a.m;
  p.scale = a.scale; p.shift = a.shift; p.residual = a.residual; p.ldr = a.ldr;
  p.y = a.y; p.ldy = a.ldy; p.act = a.act; p.act_param = a.act_param;
  uint32_t cols = 32;
  while ((int)cols < bn) cols <<= 1;
  p.tmem_cols = cols;
  const size_t stage_bytes = (size_t)(passes == 3 ? 2 : 1) * (TC_A_BYTES + (size_t)bn * TC_BK * 4);
  const int nkb = (a.k1 + a.k2) / TC_BK;
  int stages = (int)((200 * 1024) / stage_bytes);
  // keep two CTAs per SM resident when the tile is small enough (epilogue/mainloop overlap)
  if (stage_bytes * 2 <= 100 * 1024 && cols <= 256) stages = (int)((100 * 1024) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages > nkb) stages = nkb;
  if (stages < 1) stages = 1;
  p.stages = stages;
  const size_t smem = stage_bytes * stages + 1024;
  const int64_t mt = (a.m + TC_BM - 1) / TC_BM;
  GRAFP_REQUIRE(mt <= 0x7fffffff, "gemm_tc: m too large");
  dim3 grid((unsigned)mt, a.n / bn, a.groups);
  if (passes == 3) {
    cudaFuncSetAttribute(gemm_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_tc_kernel<3><<<grid, TC_THREADS, smem, st>>>(mA1, mA2, mW, p);
  } else {
    cudaFuncSetAttribute(gemm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    gemm_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(mA1, mA2, mW, p);
  }
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

}  // namespace grafp

using namespace grafp;

extern "C" int grafp_split_tf32(const float* w, int64_t count, float* out_hi_lo, void* stream) {
  GRAFP_REQUIRE(w && out_hi_lo && count >= 0, "split_tf32: bad arguments");
  if (count == 0) return 0;
  split_tf32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, as_stream(stream)>>>(w, out_hi_lo, count);
  return check_launch("split_tf32");
}
