// grafp_gemm_fwd: argument validation and engine dispatch.
#include "common.cuh"

namespace grafp {
int gemm_simt_launch(const grafp_gemm_args& a, cudaStream_t st);
int gemm_tc_supported(const grafp_gemm_args& a);
int gemm_tc_launch(const grafp_gemm_args& a, int passes, int fmt, cudaStream_t st);
}  // namespace grafp

using namespace grafp;

extern "C" int grafp_gemm_fwd(const grafp_gemm_args* args, void* stream) {
  GRAFP_REQUIRE(args, "gemm: null args");
  const grafp_gemm_args& a = *args;
  GRAFP_REQUIRE(a.m >= 0 && a.n > 0 && a.groups > 0 && a.k1 > 0 && a.k2 >= 0, "gemm: bad sizes");
  GRAFP_REQUIRE(a.m == 0 || ((a.a1 || a.a1_split) && a.w && (a.y || a.y_split)), "gemm: null pointer");
  GRAFP_REQUIRE(a.m == 0 || (a.k2 == 0) == (a.a2 == nullptr && a.a2_gather_idx == nullptr), "gemm: a2/k2 mismatch");
  GRAFP_REQUIRE(a.k1 % 4 == 0 && a.k2 % 4 == 0, "gemm: k1=%d, k2=%d must be multiples of 4", a.k1,
                a.k2);
  GRAFP_REQUIRE(a.k2 == 0 || a.k1 % 16 == 0, "gemm: dual-source needs k1 %% 16 == 0 (k1=%d)", a.k1);
  GRAFP_REQUIRE(a.lda1 % 4 == 0 && a.lda2 % 4 == 0 && a.ldw % 4 == 0, "gemm: strides must be multiples of 4");
  GRAFP_REQUIRE((reinterpret_cast<uintptr_t>(a.a1) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0 &&
                (reinterpret_cast<uintptr_t>(a.a2) & 15) == 0, "gemm: operands must be 16-byte aligned");
  if (a.tap3_nodes > 0) {
    GRAFP_REQUIRE(a.groups == 1 && a.k2 == 0 && a.k1 % 12 == 0, "gemm: tap3 needs groups=1, k2=0, k1=3*Cin");
    GRAFP_REQUIRE(a.m % a.tap3_nodes == 0, "gemm: tap3 m must be a multiple of tap3_nodes");
  }
  GRAFP_REQUIRE(a.act >= GRAFP_ACT_NONE && a.act <= GRAFP_ACT_SIGMOID, "gemm: unknown activation %d", a.act);
  GRAFP_REQUIRE(a.act != GRAFP_ACT_SIGMOID || a.engine == GRAFP_ENGINE_SIMT || a.engine == GRAFP_ENGINE_AUTO,
                "gemm: the sigmoid epilogue exists on the fp32 SIMT engine only");
  if (a.m == 0) return 0;
  cudaStream_t st = as_stream(stream);
  // 16-bit operand format of the split engines: 2 = fp16 (f16x3), 1 = bf16
  const bool f16_ok = a.w_split_f16 && a.w_f16_unscale > 0.0f;
  if (a.a2_gather_idx) {
    GRAFP_REQUIRE(!a.a2 && !a.a1_split && a.a1 && a.k2 == a.k1 && a.tap3_nodes == 0,
                  "gemm: a2_gather needs fp32 a1, a2 == NULL, k2 == k1, no tap3");
    GRAFP_REQUIRE(a.a2_gather_nodes > 0 && a.a2_gather_k > 0 && a.m % a.a2_gather_nodes == 0,
                  "gemm: a2_gather needs m to be a multiple of a2_gather_nodes");
    GRAFP_REQUIRE(a.engine == GRAFP_ENGINE_TC_BF16X3 || a.engine == GRAFP_ENGINE_TC_BF16,
                  "gemm: a2_gather needs a bf16 tensor-core engine (engine=%d)", a.engine);
    GRAFP_REQUIRE(a.w_split_bf16 && gemm_tc_supported(a), "gemm: a2_gather needs w_split_bf16 and a tcgen05-supported shape");
    return gemm_tc_launch(a, a.engine == GRAFP_ENGINE_TC_BF16 ? 1 : 3, 1, st);
  }
  if (a.a1_split || a.y_split) {
    // split 16-bit activations exist only on the 16-bit tensor-core engines: fail loudly, never convert
    GRAFP_REQUIRE(a.engine == GRAFP_ENGINE_AUTO || a.engine == GRAFP_ENGINE_TC_BF16X3 ||
                      a.engine == GRAFP_ENGINE_TC_BF16 || a.engine == GRAFP_ENGINE_TC_F16X3,
                  "gemm: split activations need a 16-bit tensor-core engine (engine=%d)", a.engine);
    GRAFP_REQUIRE(gemm_tc_supported(a), "gemm: split activations need a tcgen05-supported shape");
    if (a.engine == GRAFP_ENGINE_TC_F16X3 || (a.engine == GRAFP_ENGINE_AUTO && f16_ok)) {
      GRAFP_REQUIRE(f16_ok, "gemm: the f16x3 engine needs w_split_f16 (grafp_split_f16) and w_f16_unscale");
      return gemm_tc_launch(a, 3, 2, st);
    }
    GRAFP_REQUIRE(a.w_split_bf16, "gemm: split-bf16 activations need w_split_bf16");
    return gemm_tc_launch(a, a.engine == GRAFP_ENGINE_TC_BF16 ? 1 : 3, 1, st);
  }
  switch (a.engine) {
    case GRAFP_ENGINE_SIMT:
      return gemm_simt_launch(a, st);
    case GRAFP_ENGINE_TC_3XTF32:
      GRAFP_REQUIRE(gemm_tc_supported(a), "gemm: shape not supported by the tcgen05 engine");
      GRAFP_REQUIRE(a.w_split, "gemm: TC_3XTF32 needs w_split (grafp_split_tf32)");
      return gemm_tc_launch(a, 3, 0, st);
    case GRAFP_ENGINE_TC_TF32:
      GRAFP_REQUIRE(gemm_tc_supported(a), "gemm: shape not supported by the tcgen05 engine");
      return gemm_tc_launch(a, 1, 0, st);
    case GRAFP_ENGINE_TC_BF16X3:
    case GRAFP_ENGINE_TC_BF16:
      GRAFP_REQUIRE(gemm_tc_supported(a), "gemm: shape not supported by the tcgen05 engine");
      GRAFP_REQUIRE(a.w_split_bf16, "gemm: the bf16 engines need w_split_bf16 (grafp_split_bf16)");
      return gemm_tc_launch(a, a.engine == GRAFP_ENGINE_TC_BF16X3 ? 3 : 1, 1, st);
    case GRAFP_ENGINE_TC_F16X3:
      GRAFP_REQUIRE(gemm_tc_supported(a), "gemm: shape not supported by the tcgen05 engine");
      GRAFP_REQUIRE(f16_ok, "gemm: the f16x3 engine needs w_split_f16 (grafp_split_f16) and w_f16_unscale");
      return gemm_tc_launch(a, 3, 2, st);
    case GRAFP_ENGINE_AUTO:
      if (a.act == GRAFP_ACT_SIGMOID) return gemm_simt_launch(a, st);
      if (f16_ok && gemm_tc_supported(a)) return gemm_tc_launch(a, 3, 2, st);
      if (a.w_split_bf16 && gemm_tc_supported(a)) return gemm_tc_launch(a, 3, 1, st);
      if (a.w_split && gemm_tc_supported(a)) return gemm_tc_launch(a, 3, 0, st);
      return gemm_simt_launch(a, st);
    default:
      return fail("gemm: unknown engine %d", a.engine);
  }
}

extern "C" int grafp_gemm_tc_supported(const grafp_gemm_args* args) {
  return args ? gemm_tc_supported(*args) : 0;
}
