// Neighbour gather + max-relative aggregation (the HBM-bound stage of the path).
//
// Node-major layout makes one graph a single contiguous N*C*4-byte span, so the staged kernel
// is a persistent CTA per SM that pulls whole graphs into shared memory with 1-D bulk async
// copies (TMA engine, mbarrier completion) through a multi-stage ring, gathers neighbours out
// of shared memory with conflict-free 128-bit loads, and streams the result back with
// coalesced 128-bit stores: every byte of x crosses HBM once, every byte of m once.
// Algorithmic bytes per graph-layer: 2*N*C*4 + 4*N*k.
#include "common.cuh"

namespace grafp {

constexpr int AGG_THREADS = 512;
constexpr int AGG_MAX_STAGES = 4;

__device__ __forceinline__ void mr_item(const float* __restrict__ sx, const int32_t* __restrict__ nb,
                                        int k, int C, int node, int c4, float4& out,
                                        uint32_t& arg) {
  const float4 xi = *reinterpret_cast<const float4*>(sx + (size_t)node * C + c4 * 4);
  float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  uint32_t ax = 0, ay = 0, az = 0, aw = 0;
  for (int t = 0; t < k; ++t) {
    const int j = __ldg(nb + t);
    const float4 xj = *reinterpret_cast<const float4*>(sx + (size_t)j * C + c4 * 4);
    const float dx = xj.x - xi.x, dy = xj.y - xi.y, dz = xj.z - xi.z, dw = xj.w - xi.w;
    if (dx > best.x) { best.x = dx; ax = t; }
    if (dy > best.y) { best.y = dy; ay = t; }
    if (dz > best.z) { best.z = dz; az = t; }
    if (dw > best.w) { best.w = dw; aw = t; }
  }
  out = best;
  arg = ax | (ay << 8) | (az << 16) | (aw << 24);
}

// One stage = [ x of one graph : N*C floats | its neighbour lists : N*k int32 ], both fetched with
// 1-D bulk copies on the same mbarrier.  Each thread works on AGG_UNROLL items (one item = one node x
// 4 channels) at a time so 4 independent idx -> gather -> max chains are in flight.
constexpr int AGG_UNROLL = 4;

// kArg == false (inference): max_t (x_j - x_i) is computed as (max_t x_j) - x_i.  Round-to-nearest subtraction is
// monotone in x_j, so fl(max_t x_j - x_i) == max_t fl(x_j - x_i) BIT FOR BIT, and the inner loop is one 128-bit
// shared-memory load + 4 FMNMX per neighbour instead of 4 subtractions + 4 compare / select pairs + the arg-max
// bookkeeping (the profile showed the kernel issue-bound, not HBM-bound, once k grows: 0.93 TB/s at k = 32).
template <bool kArg>
__global__ void __launch_bounds__(AGG_THREADS, 1)
mr_aggregate_staged_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int B,
                           int N, int C, int k, int stages, uint32_t stage_bytes, int idx_in_smem,
                           float* __restrict__ m, uint32_t* __restrict__ arg_out) {
  pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[AGG_MAX_STAGES];
  const uint32_t graph_floats = (uint32_t)N * C;
  const uint32_t graph_bytes = graph_floats * 4u;
  const uint32_t idx_bytes = (uint32_t)N * k * 4u;
  const int tid = threadIdx.x;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();            // nothing above reads or writes a tensor (common.cuh)

  auto issue = [&](int s, int g) {
    unsigned char* dst = smem_raw + (size_t)s * stage_bytes;
    mbar_arrive_expect_tx(&full[s], graph_bytes + (idx_in_smem ? idx_bytes : 0u));
    bulk_g2s(dst, x + (size_t)g * graph_floats, graph_bytes, &full[s]);
    if (idx_in_smem) bulk_g2s(dst + graph_bytes, idx + (size_t)g * N * k, idx_bytes, &full[s]);
  };

  const int first = blockIdx.x, step = gridDim.x;
  if (tid == 0)
    for (int s = 0; s < stages; ++s)
      if (first + s * step < B) issue(s, first + s * step);

  const int c4n = C >> 2;
  const int items = N * c4n;
  int s = 0;
  uint32_t phase = 0;
  for (int g = first; g < B; g += step) {
    mbar_wait(&full[s], phase);
    const float* gx = reinterpret_cast<const float*>(smem_raw + (size_t)s * stage_bytes);
    const int32_t* gidx = idx_in_smem ? reinterpret_cast<const int32_t*>(smem_raw + (size_t)s * stage_bytes + graph_bytes)
                                      : idx + (size_t)g * N * k;
    float4* gm = reinterpret_cast<float4*>(m + (size_t)g * graph_floats);
    uint32_t* ga = arg_out ? arg_out + (size_t)g * items : nullptr;
    for (int it0 = tid; it0 < items; it0 += AGG_UNROLL * AGG_THREADS) {
      int node[AGG_UNROLL], off[AGG_UNROLL];
      float4 xi[AGG_UNROLL], best[AGG_UNROLL];
      uint32_t arg[AGG_UNROLL];
#pragma unroll
      for (int u = 0; u < AGG_UNROLL; ++u) {
        int it = it0 + u * AGG_THREADS;
        if (it >= items) it = it0;                      // tail: recompute item it0, store masked below
        node[u] = it / c4n;
        off[u] = (it - node[u] * c4n) * 4;
        xi[u] = *reinterpret_cast<const float4*>(gx + (size_t)node[u] * C + off[u]);
        best[u] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        arg[u] = 0u;
      }
      if (kArg) {
        for (int t = 0; t < k; ++t) {
#pragma unroll
          for (int u = 0; u < AGG_UNROLL; ++u) {
            const int j = gidx[node[u] * k + t];
            const float4 xj = *reinterpret_cast<const float4*>(gx + (size_t)j * C + off[u]);
            const float dx = xj.x - xi[u].x, dy = xj.y - xi[u].y, dz = xj.z - xi[u].z, dw = xj.w - xi[u].w;
            if (dx > best[u].x) { best[u].x = dx; arg[u] = (arg[u] & 0xFFFFFF00u) | (uint32_t)t; }
            if (dy > best[u].y) { best[u].y = dy; arg[u] = (arg[u] & 0xFFFF00FFu) | ((uint32_t)t << 8); }
            if (dz > best[u].z) { best[u].z = dz; arg[u] = (arg[u] & 0xFF00FFFFu) | ((uint32_t)t << 16); }
            if (dw > best[u].w) { best[u].w = dw; arg[u] = (arg[u] & 0x00FFFFFFu) | ((uint32_t)t << 24); }
          }
        }
      } else {
        // running maximum of the neighbour rows themselves (fmaxf: a NaN neighbour never wins, as in the
        // compare-and-select form)
        for (int t = 0; t < k; ++t) {
#pragma unroll
          for (int u = 0; u < AGG_UNROLL; ++u) {
            const int j = gidx[node[u] * k + t];
            const float4 xj = *reinterpret_cast<const float4*>(gx + (size_t)j * C + off[u]);
            best[u].x = fmaxf(best[u].x, xj.x);
            best[u].y = fmaxf(best[u].y, xj.y);
            best[u].z = fmaxf(best[u].z, xj.z);
            best[u].w = fmaxf(best[u].w, xj.w);
          }
        }
#pragma unroll
        for (int u = 0; u < AGG_UNROLL; ++u) {
          best[u].x -= xi[u].x; best[u].y -= xi[u].y; best[u].z -= xi[u].z; best[u].w -= xi[u].w;
        }
      }
#pragma unroll
      for (int u = 0; u < AGG_UNROLL; ++u) {
        const int it = it0 + u * AGG_THREADS;
        if (it < items) {
          __stcs(gm + it, best[u]);
          if (ga) ga[it] = arg[u];
        }
      }
    }
    fence_proxy_async_smem();             // generic-proxy reads of stage s before its async-proxy refill
    __syncthreads();                      // every warp is done reading stage s
    if (tid == 0) {
      const int gn = g + stages * step;
      if (gn < B) issue(s, gn);
    }
    if (++s == stages) { s = 0; phase ^= 1u; }
  }
}

// Direct (un-staged) form for graphs that do not fit a shared-memory stage: neighbours are
// gathered straight from global memory (L2-resident: a graph was just written by fc1).
__global__ void __launch_bounds__(256)
mr_aggregate_direct_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int N,
                           int C, int k, float* __restrict__ m, uint32_t* __restrict__ arg_out) {
  const int g = blockIdx.y;
  const int c4n = C >> 2;
  const int items = N * c4n;
  const float* gx = x + (size_t)g * N * C;
  const int32_t* gidx = idx + (size_t)g * N * k;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int node = it / c4n, c4 = it - node * c4n;
    float4 v; uint32_t a = 0;
    if (arg_out) {
      mr_item(gx, gidx + (size_t)node * k, k, C, node, c4, v, a);
    } else {
      // (max_t x_j) - x_i == max_t (x_j - x_i) bit for bit (monotone rounding); the k row loads are independent
      const int32_t* nb = gidx + (size_t)node * k;
      const float4 xi = *reinterpret_cast<const float4*>(gx + (size_t)node * C + c4 * 4);
      float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      int t = 0;
      for (; t + 4 <= k; t += 4) {
        float4 xj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xj[u] = __ldg(reinterpret_cast<const float4*>(gx + (size_t)__ldg(nb + t + u) * C + c4 * 4));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          best.x = fmaxf(best.x, xj[u].x); best.y = fmaxf(best.y, xj[u].y);
          best.z = fmaxf(best.z, xj[u].z); best.w = fmaxf(best.w, xj[u].w);
        }
      }
      for (; t < k; ++t) {
        const float4 xj = __ldg(reinterpret_cast<const float4*>(gx + (size_t)__ldg(nb + t) * C + c4 * 4));
        best.x = fmaxf(best.x, xj.x); best.y = fmaxf(best.y, xj.y);
        best.z = fmaxf(best.z, xj.z); best.w = fmaxf(best.w, xj.w);
      }
      v = make_float4(best.x - xi.x, best.y - xi.y, best.z - xi.z, best.w - xi.w);
    }
    reinterpret_cast<float4*>(m + (size_t)g * N * C)[it] = v;
    if (arg_out) arg_out[(size_t)g * items + it] = a;
  }
}

// batched_index_select: out (B, C, N, k) <- x (B*N, C)[idx]
__global__ void index_select_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                    int N, int C, int k, float* __restrict__ out) {
  const int g = blockIdx.y;
  const size_t total = (size_t)C * N * k;
  const float* gx = x + (size_t)g * N * C;
  const int32_t* gidx = idx + (size_t)g * N * k;
  float* go = out + (size_t)g * total;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(o % k);
    const int n = (int)((o / k) % N);
    const int c = (int)(o / ((size_t)k * N));
    go[o] = gx[(size_t)gidx[(size_t)n * k + t] * C + c];
  }
}

// backward, deterministic: dx[j*] += dm, dx[n] -= dm.  One WARP per (graph, slice of 32 channels): lane = channel, the
// N source nodes are visited in order and each lane adds into its own column of an N x 32 shared-memory accumulator
// (bank = lane: conflict-free whatever the destinations), so every sum has a fixed order -- no atomics.  (The atomic
// form below left the gradients of a train step reproducible only to ~1e-4 run to run: hub nodes collect hundreds of
// mixed-sign terms.)
__global__ void mr_aggregate_bwd_det_kernel(const float* __restrict__ dm, const int32_t* __restrict__ idx,
                                            const uint8_t* __restrict__ arg, int64_t units, int N, int C, int k,
                                            float* __restrict__ dx) {
  extern __shared__ __align__(16) float acc_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t unit = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (unit >= units) return;
  const int slices = C / 32;
  const int64_t g = unit / slices;
  const int c = (int)(unit % slices) * 32 + lane;
  float* acc = acc_all + (size_t)warp * N * 32;
  for (int n = 0; n < N; ++n) acc[n * 32 + lane] = 0.0f;
  __syncwarp();
  const size_t gbase = (size_t)g * N * C;
  const int32_t* gidx = idx + (size_t)g * N * k;
  for (int n0 = 0; n0 < N; n0 += 4) {
    float gv[4];
    int j[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                      // the loads of four sources are independent
      const int n = n0 + u;
      const bool ok = n < N;
      const size_t e = gbase + (size_t)(ok ? n : 0) * C + c;
      gv[u] = ok ? dm[e] : 0.0f;
      j[u] = ok ? gidx[(size_t)n * k + arg[e]] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = n0 + u;
      if (n < N) {
        acc[j[u] * 32 + lane] += gv[u];
        acc[n * 32 + lane] -= gv[u];
      }
    }
  }
  __syncwarp();
  for (int n = 0; n < N; ++n) dx[gbase + (size_t)n * C + c] += acc[n * 32 + lane];
}

// backward: dx[j*] += dm, dx[n] -= dm, accumulated per graph in shared memory when it fits.
__global__ void __launch_bounds__(256)
mr_aggregate_bwd_kernel(const float* __restrict__ dm, const int32_t* __restrict__ idx,
                        const uint8_t* __restrict__ arg, int N, int C, int k, int cs, int use_smem,
                        float* __restrict__ dx) {
  // One CTA = (graph, slice of `cs` channels): the scatter never crosses channels, so a graph is cut into C / cs
  // independent slices with N * cs accumulators each -- enough CTAs to fill the machine at the train step's 32-graph
  // batches (one CTA per graph left 116 of 148 SMs idle and took ~100 us per call).
  // fp64 accumulators: a hub node collects hundreds of mixed-sign terms in an order that varies from run to run;
  // in fp32 that left a train step's gradients reproducible only to ~1e-4, in fp64 the order is invisible after the
  // final rounding to fp32 (except on exact rounding ties)
  extern __shared__ __align__(16) double acc[];
  const int g = blockIdx.x, c0 = blockIdx.y * cs;
  const size_t gbase = (size_t)g * N * C;
  const int total = N * cs;
  if (use_smem) {
    for (int i = threadIdx.x; i < total; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int n = i / cs, cc = i - n * cs;
    const size_t e = gbase + (size_t)n * C + c0 + cc;
    const float gval = dm[e];
    const int j = idx[((size_t)g * N + n) * k + arg[e]];
    if (use_smem) {
      atomicAdd(&acc[j * cs + cc], (double)gval);
      atomicAdd(&acc[i], -(double)gval);
    } else {
      atomicAdd(&dx[gbase + (size_t)j * C + c0 + cc], gval);
      atomicAdd(&dx[e], -gval);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int n = i / cs, cc = i - n * cs;
      dx[gbase + (size_t)n * C + c0 + cc] += (float)acc[i];
    }
  }
}

// ---------------------------------------------------------------------------------------
// Generalised neighbour reductions: the aggregation halves of the reference's other GraphConv2d
// variants (encoder/gcn_lib/torch_vertex.py:37-89), evaluated per NODE instead of per edge:
//   GRAFP_NBR_MAX       out = max_k x_j                                   (GraphSAGE, :66-67)
//   GRAFP_NBR_SUM_SELF  out = (1 + eps) * x_i + sum_k x_j                 (GINConv2d, :86-87)
//   GRAFP_NBR_EDGE_MAX  out = max_k act(scale * (x_j - x_i) + shift)      (EdgeConv2d, :50-51, applied to
//                       P = W x: the 1x1 conv commutes with the gather, so the k-times larger edge tensor
//                       and its GEMM are never formed; the activation is applied per edge, so any
//                       activation -- monotone or not -- is exact)
// Same staging as the max-relative kernel: whole graphs in shared memory via 1-D bulk copies.
// ---------------------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ float4 nbr_item(const float* __restrict__ sx, const int32_t* __restrict__ nb, int k,
                                           int C, int node, int c0, float4 sc, float4 sh, int act,
                                           float act_param, float self_w) {
  const float4 xi = *reinterpret_cast<const float4*>(sx + (size_t)node * C + c0);
  float4 r;
  if (MODE == GRAFP_NBR_SUM_SELF) r = make_float4(0.f, 0.f, 0.f, 0.f);
  else r = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int t = 0; t < k; ++t) {
    const int j = nb[t];
    const float4 xj = *reinterpret_cast<const float4*>(sx + (size_t)j * C + c0);
    if (MODE == GRAFP_NBR_MAX) {
      r.x = fmaxf(r.x, xj.x); r.y = fmaxf(r.y, xj.y); r.z = fmaxf(r.z, xj.z); r.w = fmaxf(r.w, xj.w);
    } else if (MODE == GRAFP_NBR_SUM_SELF) {
      r.x += xj.x; r.y += xj.y; r.z += xj.z; r.w += xj.w;
    } else {
      const float ex = apply_act(fmaf(xj.x - xi.x, sc.x, sh.x), act, act_param);
      const float ey = apply_act(fmaf(xj.y - xi.y, sc.y, sh.y), act, act_param);
      const float ez = apply_act(fmaf(xj.z - xi.z, sc.z, sh.z), act, act_param);
      const float ew = apply_act(fmaf(xj.w - xi.w, sc.w, sh.w), act, act_param);
      r.x = fmaxf(r.x, ex); r.y = fmaxf(r.y, ey); r.z = fmaxf(r.z, ez); r.w = fmaxf(r.w, ew);
    }
  }
  if (MODE == GRAFP_NBR_SUM_SELF) {
    r.x = fmaf(self_w, xi.x, r.x); r.y = fmaf(self_w, xi.y, r.y);
    r.z = fmaf(self_w, xi.z, r.z); r.w = fmaf(self_w, xi.w, r.w);
  }
  return r;
}

struct NbrParams {
  const float* x; const int32_t* idx; int B, N, C, k;
  const float* scale; const float* shift; int act; float act_param; const float* eps;
  float* out; int64_t ldo;
};

template <int MODE>
__global__ void __launch_bounds__(AGG_THREADS, 1)
nbr_reduce_staged_kernel(const NbrParams p, int stages, uint32_t stage_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[AGG_MAX_STAGES];
  const int N = p.N, C = p.C, k = p.k;
  const uint32_t graph_bytes = (uint32_t)N * C * 4u, idx_bytes = (uint32_t)N * k * 4u;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();
  auto issue = [&](int s, int g) {
    unsigned char* dst = smem_raw + (size_t)s * stage_bytes;
    mbar_arrive_expect_tx(&full[s], graph_bytes + idx_bytes);
    bulk_g2s(dst, p.x + (size_t)g * N * C, graph_bytes, &full[s]);
    bulk_g2s(dst + graph_bytes, p.idx + (size_t)g * N * k, idx_bytes, &full[s]);
  };
  const int first = blockIdx.x, step = gridDim.x;
  if (tid == 0)
    for (int s = 0; s < stages; ++s)
      if (first + s * step < p.B) issue(s, first + s * step);
  const float self_w = (MODE == GRAFP_NBR_SUM_SELF) ? 1.0f + (p.eps ? __ldg(p.eps) : 0.0f) : 0.0f;
  const int c4n = C >> 2, items = N * c4n;
  int s = 0;
  uint32_t phase = 0;
  for (int g = first; g < p.B; g += step) {
    mbar_wait(&full[s], phase);
    const float* gx = reinterpret_cast<const float*>(smem_raw + (size_t)s * stage_bytes);
    const int32_t* gidx = reinterpret_cast<const int32_t*>(smem_raw + (size_t)s * stage_bytes + graph_bytes);
    float* go = p.out + (size_t)g * N * p.ldo;
    for (int it = tid; it < items; it += AGG_THREADS) {
      const int node = it / c4n, c0 = (it - node * c4n) * 4;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == GRAFP_NBR_EDGE_MAX) {
        if (p.scale) sc = __ldg(reinterpret_cast<const float4*>(p.scale + c0));
        if (p.shift) sh = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
      }
      const float4 r = nbr_item<MODE>(gx, gidx + node * k, k, C, node, c0, sc, sh, p.act, p.act_param, self_w);
      __stcs(reinterpret_cast<float4*>(go + (size_t)node * p.ldo + c0), r);
    }
    fence_proxy_async_smem();             // generic-proxy reads of stage s before its async-proxy refill
    __syncthreads();
    if (tid == 0) {
      const int gn = g + stages * step;
      if (gn < p.B) issue(s, gn);
    }
    if (++s == stages) { s = 0; phase ^= 1u; }
  }
}

// un-staged form (graphs that do not fit a stage, or unaligned sizes): gathers from global / L2
template <int MODE>
__global__ void __launch_bounds__(256)
nbr_reduce_direct_kernel(const NbrParams p, int g0) {
  const int g = g0 + blockIdx.y;
  const int N = p.N, C = p.C, k = p.k;
  const int c4n = C >> 2, items = N * c4n;
  const float* gx = p.x + (size_t)g * N * C;
  const int32_t* gidx = p.idx + (size_t)g * N * k;
  float* go = p.out + (size_t)g * N * p.ldo;
  const float self_w = (MODE == GRAFP_NBR_SUM_SELF) ? 1.0f + (p.eps ? __ldg(p.eps) : 0.0f) : 0.0f;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
    const int node = it / c4n, c0 = (it - node * c4n) * 4;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE == GRAFP_NBR_EDGE_MAX) {
      if (p.scale) sc = __ldg(reinterpret_cast<const float4*>(p.scale + c0));
      if (p.shift) sh = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
    }
    const float4 r = nbr_item<MODE>(gx, gidx + (size_t)node * k, k, C, node, c0, sc, sh, p.act, p.act_param, self_w);
    *reinterpret_cast<float4*>(go + (size_t)node * p.ldo + c0) = r;
  }
}

template <int MODE>
static int nbr_reduce_launch(const NbrParams& p, cudaStream_t st) {
  const size_t graph_bytes = (size_t)p.N * p.C * 4, idx_bytes = (size_t)p.N * p.k * 4;
  const size_t budget = 216 * 1024, stage_bytes = graph_bytes + idx_bytes;
  if (stage_bytes <= budget / 2 && graph_bytes % 16 == 0 && idx_bytes % 16 == 0) {
    int stages = (int)(budget / stage_bytes);
    if (stages > AGG_MAX_STAGES) stages = AGG_MAX_STAGES;
    int grid = sm_count();
    if (grid > p.B) grid = p.B;
    const size_t smem = (size_t)stages * stage_bytes;
    cudaFuncSetAttribute(nbr_reduce_staged_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nbr_reduce_staged_kernel<MODE><<<grid, AGG_THREADS, smem, st>>>(p, stages, (uint32_t)stage_bytes);
    return check_launch("nbr_reduce_staged");
  }
  for (int b0 = 0; b0 < p.B; b0 += 65535) {
    const int nb = p.B - b0 < 65535 ? p.B - b0 : 65535;
    const int items = p.N * (p.C / 4);
    dim3 grid((items + 255) / 256 > 64 ? 64 : (items + 255) / 256, nb);
    nbr_reduce_direct_kernel<MODE><<<grid, 256, 0, st>>>(p, b0);
    if (int rc = check_launch("nbr_reduce_direct")) return rc;
  }
  return 0;
}

}  // namespace grafp

using namespace grafp;

extern "C" {

int grafp_mr_aggregate_fwd(const float* x, const int32_t* idx, int B, int N, int C, int k,
                           float* m, uint8_t* arg_out, void* stream) {
  GRAFP_REQUIRE(B <= 0 || (x && idx && m), "mr_aggregate: null pointer");
  GRAFP_REQUIRE(B >= 0 && N > 0 && C > 0 && k > 0 && k <= 255, "mr_aggregate: bad sizes");
  GRAFP_REQUIRE(C % 4 == 0, "mr_aggregate: C=%d must be a multiple of 4", C);
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const size_t graph_bytes = (size_t)N * C * 4;
  const size_t budget = 216 * 1024;
  const size_t idx_bytes = (size_t)N * k * 4;
  const int idx_in_smem = (idx_bytes % 16 == 0) ? 1 : 0;
  const size_t stage_bytes = graph_bytes + (idx_in_smem ? idx_bytes : 0);
  if (stage_bytes <= budget / 2 && graph_bytes % 16 == 0) {
    int stages = (int)(budget / stage_bytes);
    if (stages > AGG_MAX_STAGES) stages = AGG_MAX_STAGES;
    int grid = sm_count();
    if (grid > B) grid = B;
    const size_t smem = (size_t)stages * stage_bytes;
    if (arg_out) {
      cudaFuncSetAttribute(mr_aggregate_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_ex(mr_aggregate_staged_kernel<true>, dim3(grid), dim3(AGG_THREADS), smem, st, 0,
                x, idx, B, N, C, k, stages, (uint32_t)stage_bytes, idx_in_smem, m, reinterpret_cast<uint32_t*>(arg_out));
    } else {
      cudaFuncSetAttribute(mr_aggregate_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_ex(mr_aggregate_staged_kernel<false>, dim3(grid), dim3(AGG_THREADS), smem, st, 0,
                x, idx, B, N, C, k, stages, (uint32_t)stage_bytes, idx_in_smem, m, (uint32_t*)nullptr);
    }
    return check_launch("mr_aggregate_staged");
  }
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    const int items = N * (C / 4);
    dim3 grid((items + 255) / 256 > 64 ? 64 : (items + 255) / 256, nb);
    mr_aggregate_direct_kernel<<<grid, 256, 0, st>>>(
        x + (size_t)b0 * N * C, idx + (size_t)b0 * N * k, N, C, k, m + (size_t)b0 * N * C,
        arg_out ? reinterpret_cast<uint32_t*>(arg_out) + (size_t)b0 * items : nullptr);
    if (int rc = check_launch("mr_aggregate_direct")) return rc;
  }
  return 0;
}

int grafp_mr_aggregate_bwd(const float* dm, const int32_t* idx, const uint8_t* arg, int B, int N,
                           int C, int k, float* dx, void* stream) {
  GRAFP_REQUIRE(dm && idx && arg && dx, "mr_aggregate_bwd: null pointer");
  GRAFP_REQUIRE(B >= 0 && N > 0 && C > 0 && k > 0, "mr_aggregate_bwd: bad sizes");
  if (B == 0) return 0;
  static int det_env = -1;
  if (det_env < 0) { const char* e = getenv("GRAFP_AGG_BWD_SEQUENTIAL"); det_env = e ? atoi(e) : 0; }
  if (det_env && C % 32 == 0 && (size_t)N * 128 <= 200 * 1024) {
    // strictly ordered form (opt-in: ~5x slower): one warp per (graph, 32-channel slice), sources visited in node order
    int wpb = 4;
    while (wpb > 1 && (size_t)wpb * N * 128 > 200 * 1024) wpb >>= 1;
    const int64_t units = (int64_t)B * (C / 32);
    const size_t smem = (size_t)wpb * N * 128;
    cudaFuncSetAttribute(mr_aggregate_bwd_det_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mr_aggregate_bwd_det_kernel<<<(unsigned)((units + wpb - 1) / wpb), wpb * 32, smem, as_stream(stream)>>>(
        dm, idx, arg, units, N, C, k, dx);
    return check_launch("mr_aggregate_bwd_det");
  }
  int cs = C;
  while (cs > 16 && cs % 2 == 0) cs >>= 1;             // channel slice: 16 (or the odd factor left of C)
  if (C % cs != 0) cs = C;
  const size_t bytes = (size_t)N * cs * sizeof(double);
  const int use_smem = bytes <= 200 * 1024;
  cudaFuncSetAttribute(mr_aggregate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int b0 = 0; b0 < B; b0 += 65535) {               // (gridDim.x is the graph index; y the channel slice)
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid(nb, C / cs);
    mr_aggregate_bwd_kernel<<<grid, 256, use_smem ? bytes : 0, as_stream(stream)>>>(
        dm + (size_t)b0 * N * C, idx + (size_t)b0 * N * k, arg + (size_t)b0 * N * C, N, C, k, cs, use_smem,
        dx + (size_t)b0 * N * C);
  }
  return check_launch("mr_aggregate_bwd");
}

int grafp_nbr_reduce_fwd(const float* x, const int32_t* idx, int B, int N, int C, int k, int mode,
                         const float* scale, const float* shift, int act, float act_param,
                         const float* eps, float* out, int64_t ldo, void* stream) {
  GRAFP_REQUIRE(B <= 0 || (x && idx && out), "nbr_reduce: null pointer");
  GRAFP_REQUIRE(B >= 0 && N > 0 && C > 0 && k > 0 && k <= 255, "nbr_reduce: bad sizes");
  GRAFP_REQUIRE(C % 4 == 0 && ldo % 4 == 0 && ldo >= C, "nbr_reduce: C=%d and ldo=%lld must be multiples of 4, ldo >= C",
                C, (long long)ldo);
  GRAFP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                "nbr_reduce: x / out must be 16-byte aligned");
  GRAFP_REQUIRE(act >= GRAFP_ACT_NONE && act <= GRAFP_ACT_ELU, "nbr_reduce: unknown activation %d", act);
  if (B == 0) return 0;
  NbrParams p{x, idx, B, N, C, k, scale, shift, act, act_param, eps, out, ldo};
  cudaStream_t st = as_stream(stream);
  switch (mode) {
    case GRAFP_NBR_MAX:      return nbr_reduce_launch<GRAFP_NBR_MAX>(p, st);
    case GRAFP_NBR_SUM_SELF: return nbr_reduce_launch<GRAFP_NBR_SUM_SELF>(p, st);
    case GRAFP_NBR_EDGE_MAX: return nbr_reduce_launch<GRAFP_NBR_EDGE_MAX>(p, st);
    default: return fail("nbr_reduce: unknown mode %d", mode);
  }
}

int grafp_index_select(const float* x, const int32_t* idx, int B, int N, int C, int k,
                       float* out_bcnk, void* stream) {
  GRAFP_REQUIRE(x && idx && out_bcnk, "index_select: null pointer");
  GRAFP_REQUIRE(B >= 0 && N > 0 && C > 0 && k > 0, "index_select: bad sizes");
  if (B == 0) return 0;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    const size_t total = (size_t)C * N * k;
    dim3 grid((unsigned)((total + 255) / 256 > 256 ? 256 : (total + 255) / 256), nb);
    index_select_kernel<<<grid, 256, 0, as_stream(stream)>>>(
        x + (size_t)b0 * N * C, idx + (size_t)b0 * N * k, N, C, k, out_bcnk + (size_t)b0 * total);
    if (int rc = check_launch("index_select")) return rc;
  }
  return 0;
}

}  // extern "C"
