// Error plumbing + the small layout / reduction / wrapper-stage kernels.
#include <stdarg.h>
#include "common.cuh"

namespace grafp {

thread_local char g_err[512] = {0};
std::atomic<int64_t> g_launches{0};

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
  return 0;
}

int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GRAFP_PDL"); v = e ? (atoi(e) != 0) : 0; }
  return v;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

// ---------------------------------------------------------------------------------------
// (B, C, N) <-> (B*N, C): 32x32 smem-tiled transpose, coalesced on both sides.
// ---------------------------------------------------------------------------------------
__global__ void transpose_batched_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                         int R, int S) {
  // src: (B, R, S) -> dst: (B, S, R)
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* s = src + (size_t)b * R * S;
  float* d = dst + (size_t)b * R * S;
  const int r0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = s0 + threadIdx.x;
    if (r < R && c < S) tile[i][threadIdx.x] = s[(size_t)r * S + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = s0 + i, r = r0 + threadIdx.x;
    if (r < R && c < S) d[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

// (B, C, N) -> (B*N, C) with a per-node row added: dst[b, n, c] = src[b, c, n] + pos[n, c]
__global__ void transpose_add_kernel(const float* __restrict__ src, const float* __restrict__ pos,
                                     float* __restrict__ dst, int R, int S) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* s = src + (size_t)b * R * S;
  float* d = dst + (size_t)b * R * S;
  const int r0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = s0 + threadIdx.x;
    if (r < R && c < S) tile[i][threadIdx.x] = s[(size_t)r * S + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = s0 + i, r = r0 + threadIdx.x;          // c: node, r: channel
    if (r < R && c < S) d[(size_t)c * R + r] = tile[threadIdx.x][i] + (pos ? pos[(size_t)c * R + r] : 0.0f);
  }
}

static int transpose_batched(const float* src, float* dst, int B, int R, int S, cudaStream_t st) {
  if (B == 0 || R == 0 || S == 0) return 0;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((S + 31) / 32, (R + 31) / 32, nb), block(32, 8);
    transpose_batched_kernel<<<grid, block, 0, st>>>(src + (size_t)b0 * R * S,
                                                    dst + (size_t)b0 * R * S, R, S);
    if (int rc = check_launch("transpose")) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// mean over nodes: (B*N, C) -> (B, C)
// ---------------------------------------------------------------------------------------
__global__ void node_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int N,
                                 int C) {
  const int b = blockIdx.x;
  const float* xb = x + (size_t)b * N * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.0f;
    for (int n = 0; n < N; ++n) s += xb[(size_t)n * C + c];
    out[(size_t)b * C + c] = s / (float)N;
  }
}

// ---------------------------------------------------------------------------------------
// row-wise L2 normalise (F.normalize): v / max(||v||, eps)
// ---------------------------------------------------------------------------------------
__global__ void l2norm_rows_kernel(const float* __restrict__ z, float* __restrict__ out,
                                   int64_t M, int D, float eps) {
  const int warps = blockDim.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* zr = z + row * D;
  float s = 0.0f;
  for (int c = lane; c < D; c += 32) s = fmaf(zr[c], zr[c], s);
  s = warp_sum(s);
  const float den = fmaxf(sqrtf(s), eps);
  for (int c = lane; c < D; c += 32) out[row * D + c] = __fdiv_rn(zr[c], den);
}

// ---------------------------------------------------------------------------------------
// peak extractor: one CTA per segment.  min/max reduce, then the (pb x pf)/stride patch conv
// over the 3-plane (t-ramp, f-ramp, normalised spec) image, ReLU, node-major output.
// ---------------------------------------------------------------------------------------
__global__ void peak_extract_kernel(const float* __restrict__ spec, const float* __restrict__ w,
                                    const float* __restrict__ bias, int n_mels, int n_frames,
                                    int F, int pb, int pf, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* s_spec = sm;                                   // n_mels * n_frames
  float* s_w = s_spec + n_mels * n_frames;              // F * 3 * pb * pf
  __shared__ float s_red[64];
  __shared__ float s_mn, s_mx;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int hw = n_mels * n_frames;
  const float* sp = spec + (size_t)b * hw;
  float mn = INFINITY, mx = -INFINITY;
  bool has_nan = false;                                 // torch.min / torch.max propagate NaN
  for (int i = tid; i < hw; i += nt) {
    float v = sp[i];
    s_spec[i] = v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
    has_nan |= (v != v);
  }
  if (has_nan) mx = INFINITY, mn = -INFINITY;           // inf - (-inf) below: every normalised value NaN
  for (int i = tid; i < F * 3 * pb * pf; i += nt) s_w[i] = w[i];
  mn = -warp_max(-mn);
  mx = warp_max(mx);
  if ((tid & 31) == 0) { s_red[tid >> 5] = mn; s_red[32 + (tid >> 5)] = mx; }
  __syncthreads();
  if (tid == 0) {
    float a = s_red[0], c = s_red[32];
    for (int i = 1; i < (nt >> 5); ++i) { a = fminf(a, s_red[i]); c = fmaxf(c, s_red[32 + i]); }
    s_mn = a; s_mx = c;
  }
  __syncthreads();
  const float lo = s_mn, den = s_mx - s_mn;
  const int gh = n_mels / pb, gw = n_frames / pf, nodes = gh * gw;
  const float tstep = n_frames > 1 ? 1.0f / (float)(n_frames - 1) : 0.0f;
  const float fstep = n_mels > 1 ? 1.0f / (float)(n_mels - 1) : 0.0f;
  const int pp = pb * pf;
  for (int o = tid; o < nodes * F; o += nt) {
    const int node = o / F, f = o - node * F;
    const int gy = node / gw, gx = node - gy * gw;
    const float* wf = s_w + f * 3 * pp;
    float acc = 0.0f;
    for (int ch = 0; ch < 3; ++ch) {
      for (int i = 0; i < pb; ++i) {
        const int y = gy * pb + i;
        for (int j = 0; j < pf; ++j) {
          const int xx = gx * pf + j;
          // torch.linspace(0,1,steps): symmetric evaluation, start + i*step for the first half,
          // end - (steps-1-i)*step for the second (ATen RangeFactories)
          float v;
          if (ch == 0) v = (xx < n_frames / 2) ? xx * tstep : 1.0f - (n_frames - 1 - xx) * tstep;
          else if (ch == 1) v = (y < n_mels / 2) ? y * fstep : 1.0f - (n_mels - 1 - y) * fstep;
          else v = __fdiv_rn(s_spec[y * n_frames + xx] - lo, den);
          acc = fmaf(v, wf[ch * pp + i * pf + j], acc);
        }
      }
    }
    acc += bias[f];
    out[((size_t)b * nodes + node) * F + f] = acc < 0.0f ? 0.0f : acc;   // NaN propagates
  }
}


// Fast path (pf % 4 == 0, F in {4, 8, 16}, nodes <= 1024): persistent CTAs, one THREAD per node.
// The (t-ramp, f-ramp) part of every output is the same for every segment: its partial sum -- the first 2*pb*pf terms
// of the reference's accumulation order -- is formed once per CTA into a table and each segment only adds the
// pb*pf spectrogram terms, in the same order: bit-identical to peak_extract_kernel.  The spectrogram is normalised
// once per element (the per-output form divided every element F times), a thread reads its patch with 128-bit loads
// and writes its node's F outputs as one contiguous run.
template <int F>
__global__ void __launch_bounds__(256)
peak_extract_nodes_kernel(const float* __restrict__ spec, const float* __restrict__ w, const float* __restrict__ bias,
                          int B, int n_mels, int n_frames, int pb, int pf, float* __restrict__ out) {
  extern __shared__ float sm[];
  const int hw = n_mels * n_frames, pp = pb * pf;
  const int gh = n_mels / pb, gw = n_frames / pf, nodes = gh * gw;
  float* s_spec = sm;                                   // hw
  float* s_w = s_spec + hw;                             // F * 3 * pp
  float* s_pre = s_w + F * 3 * pp;                      // nodes * F: ramp partial sums
  __shared__ float s_red[64];
  __shared__ float s_mn, s_mx;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < F * 3 * pp; i += nt) s_w[i] = w[i];
  __syncthreads();
  const float tstep = n_frames > 1 ? 1.0f / (float)(n_frames - 1) : 0.0f;
  const float fstep = n_mels > 1 ? 1.0f / (float)(n_mels - 1) : 0.0f;
  for (int o = tid; o < nodes * F; o += nt) {
    const int node = o / F, f = o - node * F;
    const int gy = node / gw, gx = node - gy * gw;
    const float* wf = s_w + f * 3 * pp;
    float acc = 0.0f;
    for (int ch = 0; ch < 2; ++ch)
      for (int i = 0; i < pb; ++i) {
        const int y = gy * pb + i;
        for (int j = 0; j < pf; ++j) {
          const int xx = gx * pf + j;
          float v;
          if (ch == 0) v = (xx < n_frames / 2) ? xx * tstep : 1.0f - (n_frames - 1 - xx) * tstep;
          else v = (y < n_mels / 2) ? y * fstep : 1.0f - (n_mels - 1 - y) * fstep;
          acc = fmaf(v, wf[ch * pp + i * pf + j], acc);
        }
      }
    s_pre[o] = acc;
  }
  float bs[F];
#pragma unroll
  for (int f = 0; f < F; ++f) bs[f] = bias[f];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const float4* sp4 = reinterpret_cast<const float4*>(spec + (size_t)b * hw);
    float mn = INFINITY, mx = -INFINITY;
    bool has_nan = false;                               // torch.min / torch.max propagate NaN
    __syncthreads();                                    // the previous segment's patch reads are done
    for (int i = tid; i < hw / 4; i += nt) {
      const float4 v = __ldg(sp4 + i);
      reinterpret_cast<float4*>(s_spec)[i] = v;
      mn = fminf(fminf(mn, v.x), fminf(fminf(v.y, v.z), v.w));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
      has_nan |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    if (has_nan) mx = INFINITY, mn = -INFINITY;         // inf - (-inf) below: every normalised value NaN
    mn = -warp_max(-mn);
    mx = warp_max(mx);
    if ((tid & 31) == 0) { s_red[tid >> 5] = mn; s_red[32 + (tid >> 5)] = mx; }
    __syncthreads();
    if (tid == 0) {
      float a = s_red[0], c = s_red[32];
      for (int i = 1; i < (nt >> 5); ++i) { a = fminf(a, s_red[i]); c = fmaxf(c, s_red[32 + i]); }
      s_mn = a; s_mx = c;
    }
    __syncthreads();
    const float lo = s_mn, den = s_mx - s_mn;
    for (int i = tid; i < hw / 4; i += nt) {            // each thread normalises the elements it staged
      float4 v = reinterpret_cast<float4*>(s_spec)[i];
      v.x = __fdiv_rn(v.x - lo, den); v.y = __fdiv_rn(v.y - lo, den);
      v.z = __fdiv_rn(v.z - lo, den); v.w = __fdiv_rn(v.w - lo, den);
      reinterpret_cast<float4*>(s_spec)[i] = v;
    }
    __syncthreads();
    for (int node = tid; node < nodes; node += nt) {
      const int gy = node / gw, gx = node - gy * gw;
      float acc[F];
#pragma unroll
      for (int f = 0; f < F; ++f) acc[f] = s_pre[node * F + f];
      for (int i = 0; i < pb; ++i) {
        const float* row = s_spec + (gy * pb + i) * n_frames + gx * pf;
        for (int j = 0; j < pf; j += 4) {
          const float4 v = *reinterpret_cast<const float4*>(row + j);
          const float* wj = s_w + 2 * pp + i * pf + j;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const float4 w4 = *reinterpret_cast<const float4*>(wj + f * 3 * pp);      // broadcast
            acc[f] = fmaf(v.x, w4.x, acc[f]);
            acc[f] = fmaf(v.y, w4.y, acc[f]);
            acc[f] = fmaf(v.z, w4.z, acc[f]);
            acc[f] = fmaf(v.w, w4.w, acc[f]);
          }
        }
      }
      float4* o4 = reinterpret_cast<float4*>(out + ((size_t)b * nodes + node) * F);
#pragma unroll
      for (int f = 0; f < F; f += 4) {
        float4 r;
        r.x = acc[f] + bs[f];         r.x = r.x < 0.0f ? 0.0f : r.x;                    // NaN propagates
        r.y = acc[f + 1] + bs[f + 1]; r.y = r.y < 0.0f ? 0.0f : r.y;
        r.z = acc[f + 2] + bs[f + 2]; r.z = r.z < 0.0f ? 0.0f : r.z;
        r.w = acc[f + 3] + bs[f + 3]; r.w = r.w < 0.0f ? 0.0f : r.w;
        o4[f / 4] = r;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Stem: per-node linear with a tiny input width (Conv2d(Cin -> Cout, 1x1) + BN + activation,
// graph_encoder.py:151-153) straight from the reference's (B, Cin, N) layout.  K = Cin <= 16 is
// far below a tensor-core k-block; the op is one streaming write of the (B*N, Cout) output, so a
// graph per CTA is staged in shared memory and every warp store covers whole output rows.
// ---------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256)
stem_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
            const float* __restrict__ shift, int B, int N, int Cout, int nchw, int act, float act_param,
            float* __restrict__ out) {
  extern __shared__ float sm[];
  float* xs = sm;                          // [CIN][N] (k-major)
  float* ws = xs + (size_t)CIN * N;        // [CIN][Cout] (k-major)
  const int tid = threadIdx.x;
  for (int i = tid; i < CIN * Cout; i += 256) {
    const int n = i / CIN, k = i - n * CIN;            // w is (Cout, CIN) row-major
    ws[k * Cout + n] = w[i];
  }
  const int cgs = Cout >> 2;                           // float4 column groups per row
  const int cg = tid % cgs, rl = tid / cgs, rows_in_flight = 256 / cgs;
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) sc = *reinterpret_cast<const float4*>(scale + 4 * cg);
  if (shift) sh = *reinterpret_cast<const float4*>(shift + 4 * cg);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();                                   // previous graph fully consumed (and ws written)
    const float* xb = x + (size_t)b * CIN * N;
    if (nchw) {
      for (int i = tid; i < CIN * N; i += 256) xs[i] = xb[i];
    } else {
      for (int i = tid; i < CIN * N; i += 256) {       // (N, CIN) node-major -> k-major
        const int n = i / CIN, k = i - n * CIN;
        xs[k * N + n] = xb[i];
      }
    }
    __syncthreads();
    float4* ob = reinterpret_cast<float4*>(out + (size_t)b * N * Cout);
    for (int n = rl; n < N; n += rows_in_flight) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < CIN; ++k) {
        const float a = xs[k * N + n];
        const float4 wv = *reinterpret_cast<const float4*>(ws + k * Cout + 4 * cg);
        acc.x = fmaf(a, wv.x, acc.x); acc.y = fmaf(a, wv.y, acc.y);
        acc.z = fmaf(a, wv.z, acc.z); acc.w = fmaf(a, wv.w, acc.w);
      }
      float4 o;
      o.x = apply_act(fmaf(acc.x, sc.x, sh.x), act, act_param);
      o.y = apply_act(fmaf(acc.y, sc.y, sh.y), act, act_param);
      o.z = apply_act(fmaf(acc.z, sc.z, sh.z), act, act_param);
      o.w = apply_act(fmaf(acc.w, sc.w, sh.w), act, act_param);
      stg_stream(ob + (size_t)n * cgs + cg, o);
    }
  }
}

}  // namespace grafp

using namespace grafp;

extern "C" {

int grafp_abi_version(void) { return GRAFP_ABI_VERSION; }
const char* grafp_last_error(void) { return g_err; }
int64_t grafp_launch_count(void) { return g_launches.load(); }

int grafp_nchw_to_nodes(const float* src, float* dst, int B, int C, int N, void* stream) {
  GRAFP_REQUIRE(B >= 0 && C > 0 && N > 0 && (B == 0 || (src && dst)), "nchw_to_nodes: bad arguments");
  return transpose_batched(src, dst, B, C, N, as_stream(stream));
}
int grafp_nodes_to_nchw(const float* src, float* dst, int B, int C, int N, void* stream) {
  GRAFP_REQUIRE(B >= 0 && C > 0 && N > 0 && (B == 0 || (src && dst)), "nodes_to_nchw: bad arguments");
  return transpose_batched(src, dst, B, N, C, as_stream(stream));
}

int grafp_node_mean(const float* x, int B, int N, int C, float* out, void* stream) {
  GRAFP_REQUIRE(B >= 0 && N > 0 && C > 0 && (B == 0 || (x && out)), "node_mean: bad arguments");
  if (B == 0) return 0;
  node_mean_kernel<<<B, C < 256 ? ((C + 31) / 32) * 32 : 256, 0, as_stream(stream)>>>(x, out, N, C);
  return check_launch("node_mean");
}

int grafp_l2_normalize_rows(const float* z, int64_t M, int D, float eps, float* out,
                            void* stream) {
  GRAFP_REQUIRE(M >= 0 && D > 0 && (M == 0 || (z && out)), "l2_normalize_rows: bad arguments");
  if (M == 0) return 0;
  const int warps = 8;
  l2norm_rows_kernel<<<(unsigned)((M + warps - 1) / warps), warps * 32, 0, as_stream(stream)>>>(
      z, out, M, D, eps);
  return check_launch("l2_normalize_rows");
}

int grafp_peak_extract_fwd(const float* spec, const float* w, const float* bias, int B,
                           int n_mels, int n_frames, int F, int pb, int pf, float* out,
                           void* stream) {
  GRAFP_REQUIRE(B == 0 || (spec && w && bias && out), "peak_extract: null pointer");
  GRAFP_REQUIRE(pb > 0 && pf > 0 && n_mels % pb == 0 && n_frames % pf == 0,
                "peak_extract: patch (%d,%d) must tile (%d,%d)", pb, pf, n_mels, n_frames);
  if (B == 0) return 0;
  {
    // fast path: one thread per node over persistent CTAs (bit-identical; GRAFP_PEAK_SIMPLE=1 keeps the simple kernel)
    const char* env = getenv("GRAFP_PEAK_SIMPLE");
    const bool simple = env && atoi(env) != 0;
    const int nodes = (n_mels / pb) * (n_frames / pf), pp = pb * pf;
    const size_t sm2 = ((size_t)n_mels * n_frames + (size_t)F * 3 * pp + (size_t)nodes * F) * sizeof(float);
    const bool aligned = ((reinterpret_cast<uintptr_t>(spec) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (!simple && aligned && pf % 4 == 0 && (n_mels * n_frames) % 4 == 0 && (3 * pp) % 4 == 0 && nodes <= 4096 &&
        (F == 4 || F == 8 || F == 16) && sm2 <= 100 * 1024) {
      int per_sm = (int)((200 * 1024) / (sm2 + 1024));
      if (per_sm > 4) per_sm = 4;
      int grid = sm_count() * per_sm;
      if (grid > B) grid = B;
      auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        kern<<<grid, 256, sm2, as_stream(stream)>>>(spec, w, bias, B, n_mels, n_frames, pb, pf, out);
      };
      if (F == 4) launch(peak_extract_nodes_kernel<4>);
      else if (F == 8) launch(peak_extract_nodes_kernel<8>);
      else launch(peak_extract_nodes_kernel<16>);
      return check_launch("peak_extract_nodes");
    }
  }
  size_t smem = ((size_t)n_mels * n_frames + (size_t)F * 3 * pb * pf) * sizeof(float);
  GRAFP_REQUIRE(smem <= 200 * 1024, "peak_extract: segment too large for shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(peak_extract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         200 * 1024);
    attr_set = true;
  }
  peak_extract_kernel<<<B, 256, smem, as_stream(stream)>>>(spec, w, bias, n_mels, n_frames, F, pb,
                                                          pf, out);
  return check_launch("peak_extract");
}

int grafp_nchw_to_nodes_add(const float* src, const float* pos, float* dst, int B, int C, int N, void* stream) {
  GRAFP_REQUIRE(B == 0 || (src && dst), "nchw_to_nodes_add: null pointer");
  GRAFP_REQUIRE(B >= 0 && C > 0 && N > 0, "nchw_to_nodes_add: bad sizes");
  for (int b0 = 0; b0 < B; b0 += 65535) {
    int nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((N + 31) / 32, (C + 31) / 32, nb), block(32, 8);
    transpose_add_kernel<<<grid, block, 0, as_stream(stream)>>>(src + (size_t)b0 * C * N, pos,
                                                               dst + (size_t)b0 * C * N, C, N);
    if (int rc = check_launch("transpose_add")) return rc;
  }
  return 0;
}

int grafp_stem_fwd(const float* x, const float* w, const float* scale, const float* shift, int B, int Cin,
                   int N, int Cout, int nchw, int act, float act_param, float* out, void* stream) {
  GRAFP_REQUIRE(B == 0 || (x && w && out), "stem: null pointer");
  GRAFP_REQUIRE(Cin == 4 || Cin == 8 || Cin == 16, "stem: Cin=%d (supported: 4, 8, 16; use grafp_gemm_fwd)", Cin);
  GRAFP_REQUIRE(Cout % 4 == 0 && Cout >= 4 && Cout <= 1024 && 256 % (Cout / 4) == 0,
                "stem: Cout=%d must be 4 * a divisor of 256", Cout);
  GRAFP_REQUIRE(act >= GRAFP_ACT_NONE && act <= GRAFP_ACT_ELU, "stem: unknown activation %d", act);
  GRAFP_REQUIRE(N > 0, "stem: N must be positive");
  if (B == 0) return 0;
  const size_t smem = ((size_t)Cin * N + (size_t)Cin * Cout) * sizeof(float);
  GRAFP_REQUIRE(smem <= 96 * 1024, "stem: graph too large for shared memory");
  int grid = sm_count() * 4;
  if (B < grid) grid = B;
  cudaStream_t st = as_stream(stream);
#define GRAFP_STEM_CASE(C)                                                                           \
  case C:                                                                                            \
    cudaFuncSetAttribute(stem_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);    \
    stem_kernel<C><<<grid, 256, smem, st>>>(x, w, scale, shift, B, N, Cout, nchw, act, act_param, out); \
    break;
  switch (Cin) {
    GRAFP_STEM_CASE(4)
    GRAFP_STEM_CASE(8)
    GRAFP_STEM_CASE(16)
  }
#undef GRAFP_STEM_CASE
  return check_launch("stem");
}

}  // extern "C"
