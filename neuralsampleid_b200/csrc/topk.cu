// Row-wise k-smallest selection for the exact fingerprint search (the reference's
// eval.py builds a FAISS index over the (n, 128) fingerprint memmap and calls index.search(q, k_probe),
// eval.py:37-151, 306; index type 'l2' = IndexFlatL2 is the exact one).  The distance matrix
//   Y[q, j] = |d_j|^2 - 2 <q, d_j>
// comes from the tcgen05 GEMM engine (database chunk as the weight operand, |d|^2 as the per-column shift);
// these kernels stream it once:
//   topk_rows_kernel   one warp per (row, column split): lanes read coalesced float4s; the warp keeps the k
//                      smallest in one (value, index) slot per lane and a threshold tau = current k-th
//                      smallest, so after the first few hundred columns almost every element costs one compare
//   topk_merge_kernel  one warp per row: merges the partial lists of every split / database chunk, adds |q|^2
//                      and emits the k results sorted ascending (ties: lower index first)
#include "common.cuh"

namespace grafp {

constexpr int TOPK_WARPS = 8;

// warp-wide list: lane l < k owns one slot; insert x (if x < tau) by replacing the slot that holds tau
struct WarpList {
  float v;          // my slot (+inf when empty or lane >= k)
  long long i;
  float tau;        // max over the k slots (warp-uniform)
};

__device__ __forceinline__ float warp_max_all(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

__device__ __forceinline__ void list_insert(WarpList& L, float x, long long idx, int lane, int k) {
  // warp-uniform x, idx.  Precondition: x < L.tau.
  const unsigned holders = __ballot_sync(0xffffffffu, lane < k && L.v == L.tau);
  const int victim = __ffs(holders) - 1;
  if (lane == victim) { L.v = x; L.i = idx; }
  L.tau = warp_max_all(lane < k ? L.v : -INFINITY);
}

// drain the candidates (x < tau) the lanes hold, one at a time (lowest lane first)
__device__ __forceinline__ void list_offer(WarpList& L, float x, long long idx, int lane, int k) {
  unsigned pending = __ballot_sync(0xffffffffu, x < L.tau);
  while (pending) {
    const int src = __ffs(pending) - 1;
    const float cx = __shfl_sync(0xffffffffu, x, src);
    const long long ci = __shfl_sync(0xffffffffu, idx, src);
    if (cx < L.tau) list_insert(L, cx, ci, lane, k);
    if (lane == src) x = INFINITY;                       // consumed
    pending = __ballot_sync(0xffffffffu, x < L.tau);
  }
}

__global__ void __launch_bounds__(TOPK_WARPS * 32)
topk_rows_kernel(const float* __restrict__ y, int64_t ldy, int rows, int64_t cols, int64_t col_offset, int k,
                 int splits, float* __restrict__ part_val, long long* __restrict__ part_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * TOPK_WARPS + (threadIdx.x >> 5);
  if (w >= (int64_t)rows * splits) return;
  const int row = (int)(w / splits), sp = (int)(w - (int64_t)row * splits);
  // column range of this split, in whole float4s (cols and ldy are multiples of 4)
  const int64_t c4 = cols >> 2;
  const int64_t per = (c4 + splits - 1) / splits;
  const int64_t b4 = per * sp, e4 = (b4 + per < c4) ? b4 + per : c4;
  const float4* yr = reinterpret_cast<const float4*>(y + (int64_t)row * ldy);
  WarpList L;
  L.v = INFINITY; L.i = -1; L.tau = INFINITY;
  for (int64_t q0 = b4; q0 < e4; q0 += 32) {               // warp-uniform trip count (ballots inside)
    const int64_t q = q0 + lane;
    float4 t = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    if (q < e4) t = __ldcs(yr + q);
    const long long base = col_offset + (q << 2);
    if (__any_sync(0xffffffffu, fminf(fminf(t.x, t.y), fminf(t.z, t.w)) < L.tau)) {
      list_offer(L, t.x, base, lane, k);
      list_offer(L, t.y, base + 1, lane, k);
      list_offer(L, t.z, base + 2, lane, k);
      list_offer(L, t.w, base + 3, lane, k);
    }
  }
  if (lane < k) {
    const int64_t o = ((int64_t)row * splits + sp) * k + lane;
    part_val[o] = L.v;
    part_idx[o] = L.i;
  }
}

// (value, index) lexicographic "a before b"
__device__ __forceinline__ bool lex_less(float av, long long ai, float bv, long long bi) {
  return av < bv || (av == bv && ai < bi);
}

__global__ void __launch_bounds__(TOPK_WARPS * 32)
topk_merge_kernel(const float* __restrict__ part_val, const long long* __restrict__ part_idx, int rows, int parts,
                  int k, const float* __restrict__ row_add, float* __restrict__ out_val,
                  long long* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * TOPK_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t n = (int64_t)parts * k;
  const float* pv = part_val + (int64_t)row * n;
  const long long* pi = part_idx + (int64_t)row * n;
  WarpList L;
  L.v = INFINITY; L.i = -1; L.tau = INFINITY;
  for (int64_t q0 = 0; q0 < n; q0 += 32) {                  // warp-uniform trip count (ballots inside)
    const int64_t q = q0 + lane;
    float x = INFINITY;
    long long idx = -1;
    if (q < n) { x = pv[q]; idx = pi[q]; }
    if (idx < 0) x = INFINITY;                                 // empty partial slot
    list_offer(L, x, idx, lane, k);
  }
  // sort the k slots ascending by (value, index): rank of my slot = number of slots before it
  // empty slots of the list sort after every real entry, lanes beyond the list after those: ranks of the k
  // list lanes are then exactly 0..k-1
  float mv = lane < k ? L.v : INFINITY;
  long long mi = (lane < k && L.i >= 0) ? L.i : (0x7fffffffffffff00LL + lane + (lane < k ? 0 : 64));
  int rank = 0;
  for (int s = 0; s < 32; ++s) {
    const float ov = __shfl_sync(0xffffffffu, mv, s);
    const long long oi = __shfl_sync(0xffffffffu, mi, s);
    if (s != lane && lex_less(ov, oi, mv, mi)) ++rank;
  }
  if (lane < k) {
    const float add = row_add ? row_add[row] : 0.0f;
    out_val[(int64_t)row * k + rank] = (L.i >= 0) ? L.v + add : INFINITY;
    out_idx[(int64_t)row * k + rank] = L.i;
  }
}

// out[m] = sum_c x[m, c]^2
__global__ void row_sumsq_kernel(const float* __restrict__ x, int64_t M, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  float s = 0.0f;
  for (int c = lane; c < D; c += 32) { const float v = x[row * D + c]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// Song-level match score of eval.py:322-331: for every candidate start id, the mean over the query sequence of the
// inner products with the database rows that follow it:  out[c] = mean_{t < len} <q[t], db[cand[c] + t]>,
// len = min(sl, n - cand[c]).  One warp per candidate (the reference does this in a Python loop per candidate).
__global__ void __launch_bounds__(256)
sequence_score_kernel(const float* __restrict__ q, int sl, int D, const float* __restrict__ db, int64_t n,
                      const long long* __restrict__ cand, int nc, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= nc) return;
  const long long cid = cand[c];
  float acc = 0.0f;
  int len = 0;
  if (cid >= 0 && cid < n) {
    len = (int)((n - cid < (long long)sl) ? (n - cid) : (long long)sl);
    for (int t = 0; t < len; ++t) {
      const float* qr = q + (int64_t)t * D;
      const float* dr = db + (cid + t) * D;
      for (int d = lane; d < D; d += 32) acc = fmaf(qr[d], dr[d], acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) out[c] = len > 0 ? acc / (float)len : 0.0f;
}

}  // namespace grafp

using namespace grafp;

extern "C" {

int grafp_topk_rows_fwd(const float* y, int64_t ldy, int rows, int64_t cols, int64_t col_offset, int k, int splits,
                        float* part_val, int64_t* part_idx, void* stream) {
  GRAFP_REQUIRE(rows <= 0 || (y && part_val && part_idx), "topk_rows: null pointer");
  GRAFP_REQUIRE(rows >= 0 && cols > 0 && k >= 1 && k <= 32 && splits >= 1, "topk_rows: bad sizes (1 <= k <= 32)");
  GRAFP_REQUIRE(cols % 4 == 0 && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
                "topk_rows: cols / ldy must be multiples of 4 and y 16-byte aligned");
  if (rows == 0) return 0;
  const int64_t warps = (int64_t)rows * splits;
  const int64_t blocks = (warps + TOPK_WARPS - 1) / TOPK_WARPS;
  GRAFP_REQUIRE(blocks <= 2147483647LL, "topk_rows: too many blocks");
  topk_rows_kernel<<<(unsigned)blocks, TOPK_WARPS * 32, 0, as_stream(stream)>>>(
      y, ldy, rows, cols, col_offset, k, splits, part_val, reinterpret_cast<long long*>(part_idx));
  return check_launch("topk_rows");
}

int grafp_topk_merge_fwd(const float* part_val, const int64_t* part_idx, int rows, int parts, int k,
                         const float* row_add, float* out_val, int64_t* out_idx, void* stream) {
  GRAFP_REQUIRE(rows <= 0 || (part_val && part_idx && out_val && out_idx), "topk_merge: null pointer");
  GRAFP_REQUIRE(rows >= 0 && parts >= 1 && k >= 1 && k <= 32, "topk_merge: bad sizes (1 <= k <= 32)");
  if (rows == 0) return 0;
  topk_merge_kernel<<<(unsigned)((rows + TOPK_WARPS - 1) / TOPK_WARPS), TOPK_WARPS * 32, 0, as_stream(stream)>>>(
      part_val, reinterpret_cast<const long long*>(part_idx), rows, parts, k, row_add, out_val,
      reinterpret_cast<long long*>(out_idx));
  return check_launch("topk_merge");
}

int grafp_sequence_score_fwd(const float* q, int sl, int D, const float* db, int64_t n, const int64_t* cand, int nc,
                             float* out, void* stream) {
  GRAFP_REQUIRE(nc <= 0 || (q && db && cand && out), "sequence_score: null pointer");
  GRAFP_REQUIRE(sl > 0 && D > 0 && n >= 0 && nc >= 0, "sequence_score: bad sizes");
  if (nc == 0) return 0;
  sequence_score_kernel<<<(unsigned)((nc + 7) / 8), 256, 0, as_stream(stream)>>>(
      q, sl, D, db, n, reinterpret_cast<const long long*>(cand), nc, out);
  return check_launch("sequence_score");
}

int grafp_row_sumsq(const float* x, int64_t M, int D, float* out, void* stream) {
  GRAFP_REQUIRE(M <= 0 || (x && out), "row_sumsq: null pointer");
  GRAFP_REQUIRE(M >= 0 && D > 0, "row_sumsq: bad sizes");
  if (M == 0) return 0;
  row_sumsq_kernel<<<(unsigned)((M + 7) / 8), 256, 0, as_stream(stream)>>>(x, M, D, out);
  return check_launch("row_sumsq");
}

}  // extern "C"
