// Kernels of the contrastive train step that are not shared with inference: train-mode BatchNorm
// (batch statistics, running-stat update, fused normalise+activation+shortcut), its backward,
// weight-gradient GEMM, Downsample / mean / normalise / peak-extractor backward, and the fused
// gradient-clip + Adam update.   Reference semantics: nn.BatchNorm2d (eps 1e-5, momentum 0.1,
// biased batch variance for normalisation, unbiased for running_var), train.py:70-75
// (clip_grad_norm_(1.0) then Adam).
#include "common.cuh"

namespace grafp {

__device__ __forceinline__ float act_grad(float z, int act, float p) {
  switch (act) {
    case GRAFP_ACT_RELU:  return z > 0.0f ? 1.0f : 0.0f;
    case GRAFP_ACT_LEAKY: return z > 0.0f ? 1.0f : p;
    case GRAFP_ACT_GELU: {
      const float c = 0.70710678118654752440f, phi = 0.5f * (1.0f + erff(z * c));
      return phi + z * 0.3989422804014327f * expf(-0.5f * z * z);
    }
    case GRAFP_ACT_ELU:   return z > 0.0f ? 1.0f : expf(z);
    default:              return 1.0f;
  }
}

// ---- per-column sums over the rows of an (M, C) matrix, fp64 accumulation ----------------
// mode 0: s0 += x, s1 += x^2                       (BatchNorm batch statistics)
// mode 1: dz = dout * act'(raw*scale+shift);  s0 += dz,  s1 += dz * (raw - mean) * invstd
__global__ void __launch_bounds__(256)
col_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dout, int64_t M, int C,
                  int64_t ld, int64_t ldd, int mode, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ mean,
                  const float* __restrict__ invstd, int act, float act_param, int rows_per_block,
                  double* __restrict__ s0, double* __restrict__ s1) {
  __shared__ double r0[8][33], r1[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  const int64_t m_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m_end = m_begin + rows_per_block < M ? m_begin + rows_per_block : M;
  double a0 = 0.0, a1 = 0.0;
  if (c < C) {
    float sc = 1.f, sh = 0.f, mu = 0.f, is = 1.f;
    if (mode == 1) { sc = scale[c]; sh = shift[c]; mu = mean[c]; is = invstd[c]; }
    for (int64_t m = m_begin + ty; m < m_end; m += 8) {
      const float v = x[m * ld + c];
      if (mode == 0) {
        a0 += (double)v;
        a1 += (double)v * (double)v;
      } else {
        const float dz = dout[m * ldd + c] * act_grad(fmaf(v, sc, sh), act, act_param);
        a0 += (double)dz;
        a1 += (double)dz * (double)((v - mu) * is);
      }
    }
  }
  r0[ty][tx] = a0; r1[ty][tx] = a1;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { a0 += r0[i][tx]; a1 += r1[i][tx]; }
    atomicAdd(&s0[c], a0);
    atomicAdd(&s1[c], a1);
  }
}

// BatchNorm train-mode finalize: batch mean / biased var -> fused (scale, shift) for the affine
// kernel, saved (mean, invstd) for the backward, running-stat update (unbiased var, momentum).
__global__ void bn_finalize_kernel(const double* __restrict__ s0, const double* __restrict__ s1,
                                   int64_t M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                   float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = s0[c] / (double)M;
  double var = s1[c] / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + (double)eps);
  const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
  scale[c] = (float)(g * invstd);
  shift[c] = (float)(b - mean * g * invstd);
  mean_out[c] = (float)mean;
  invstd_out[c] = (float)invstd;
  if (running_mean) {
    const double bias = conv_bias ? (double)conv_bias[c] : 0.0;   // the conv bias is not in `raw`
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * (mean + bias));
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
  }
}

// out = act(x * scale + shift) + residual      (float4 over an (M, C) matrix, C % 4 == 0)
__global__ void affine_act_kernel(const float* __restrict__ x, int64_t M, int C, int64_t ld,
                                  const float* __restrict__ scale, const float* __restrict__ shift,
                                  int act, float act_param, const float* __restrict__ residual,
                                  int64_t ldr, float* __restrict__ out, int64_t ldo) {
  const int c4n = C >> 2;
  const int64_t total = M * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / c4n;
    const int c = (int)(i - m * c4n) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + m * ld + c);
    const float4 sc = scale ? *reinterpret_cast<const float4*>(scale + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 sh = shift ? *reinterpret_cast<const float4*>(shift + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = apply_act(fmaf(v.x, sc.x, sh.x), act, act_param);
    o.y = apply_act(fmaf(v.y, sc.y, sh.y), act, act_param);
    o.z = apply_act(fmaf(v.z, sc.z, sh.z), act, act_param);
    o.w = apply_act(fmaf(v.w, sc.w, sh.w), act, act_param);
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + m * ldr + c);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    *reinterpret_cast<float4*>(out + m * ldo + c) = o;
  }
}

// Train-mode BatchNorm in ONE launch: every block derives the C (scale, shift) pairs from the fp64 column sums into
// shared memory (block 0 also publishes them with the saved mean / invstd and moves the running statistics), then
// applies out = act(x * scale + shift) + residual.  Same arithmetic as bn_finalize_kernel + affine_act_kernel.
__global__ void __launch_bounds__(256)
bn_finalize_apply_kernel(const double* __restrict__ s0, const double* __restrict__ s1, int64_t M, int C,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ conv_bias, float eps, float momentum,
                         float* __restrict__ running_mean, float* __restrict__ running_var,
                         float* __restrict__ scale_out, float* __restrict__ shift_out, float* __restrict__ mean_out,
                         float* __restrict__ invstd_out, const float* __restrict__ x, int64_t ld, int act,
                         float act_param, const float* __restrict__ residual, int64_t ldr, float* __restrict__ out,
                         int64_t ldo) {
  extern __shared__ __align__(16) float s_ss[];        // scale[C] | shift[C]
  float* s_scale = s_ss;
  float* s_shift = s_ss + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = s0[c] / (double)M;
    double var = s1[c] / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const double invstd = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
    const float sc = (float)(g * invstd), sh = (float)(b - mean * g * invstd);
    s_scale[c] = sc;
    s_shift[c] = sh;
    if (blockIdx.x == 0) {
      scale_out[c] = sc;
      shift_out[c] = sh;
      mean_out[c] = (float)mean;
      invstd_out[c] = (float)invstd;
      if (running_mean) {
        const double bias = conv_bias ? (double)conv_bias[c] : 0.0;   // the conv bias is not in `raw`
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * (mean + bias));
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
      }
    }
  }
  __syncthreads();
  const int c4n = C >> 2;
  const int64_t total = M * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / c4n;
    const int c = (int)(i - m * c4n) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + m * ld + c);
    const float4 sc = *reinterpret_cast<const float4*>(s_scale + c);
    const float4 sh = *reinterpret_cast<const float4*>(s_shift + c);
    float4 o;
    o.x = apply_act(fmaf(v.x, sc.x, sh.x), act, act_param);
    o.y = apply_act(fmaf(v.y, sc.y, sh.y), act, act_param);
    o.z = apply_act(fmaf(v.z, sc.z, sh.z), act, act_param);
    o.w = apply_act(fmaf(v.w, sc.w, sh.w), act, act_param);
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + m * ldr + c);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    *reinterpret_cast<float4*>(out + m * ldo + c) = o;
  }
}

// draw = scale * (dz - [bn] (s0/M + xhat * s1/M)),  dz = dout * act'(raw*scale+shift)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ raw,
                                    int64_t M, int C, int64_t ldd, int64_t ld,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    int act, float act_param, int bn, const double* __restrict__ s0,
                                    const double* __restrict__ s1, float* __restrict__ draw,
                                    int64_t ldo, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int64_t total = M * C;
  const double invM = 1.0 / (double)M;
  if (blockIdx.x == 0) {            // parameter gradients (accumulating): dgamma += sum dz*xhat, dbeta += sum dz
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dgamma) dgamma[c] += (float)s1[c];
      if (dbeta) dbeta[c] += (float)s0[c];
    }
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / C;
    const int c = (int)(i - m * C);
    const float v = raw[m * ld + c];
    const float sc = scale[c];
    float dz = dout[m * ldd + c] * act_grad(fmaf(v, sc, shift[c]), act, act_param);
    if (bn) {
      const float xhat = (v - mean[c]) * invstd[c];
      dz = dz - (float)(s0[c] * invM) - xhat * (float)(s1[c] * invM);
    }
    draw[m * ldo + c] = dz * sc;
  }
}

// dgamma += s1, dbeta += s0   (or dbias += s0 when there is no BN)
__global__ void bn_param_grad_kernel(const double* __restrict__ s0, const double* __restrict__ s1,
                                     int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dgamma) dgamma[c] += (float)s1[c];
  if (dbeta) dbeta[c] += (float)s0[c];
}

// tcgen05 engine (wgrad_tc.cu)
int wgrad_tc_supported(int64_t m, int n, int k1, int k2, int groups, int tap3_nodes, int64_t ldy, int64_t lda1,
                       int64_t lda2, int64_t ldw);
size_t wgrad_tc_workspace_bytes(int64_t m, int n, int k1, int k2, int groups);
int wgrad_tc_launch(const float* dy, int64_t ldy, const float* a1, int64_t lda1, int k1, const float* a2, int64_t lda2,
                    int k2, int64_t m, int n, int groups, float* dw, int64_t ldw, float* workspace, cudaStream_t st);

// ---- weight gradient: dW[g*n + j, kk] += sum_m dy[m, g*n + j] * A_g[m, kk] ------------------
struct WgradP {
  const float* dy; int64_t ldy;
  const float* a1; int64_t lda1; int k1;
  const float* a2; int64_t lda2; int k2;
  float* dw; int64_t ldw;
  int64_t m; int n; int tap3_nodes; int rows_per_split;
};

__global__ void __launch_bounds__(256) gemm_wgrad_kernel(const WgradP p, int groups) {
  // CTA tile: 64 (n) x 64 (k) outputs, reduction over a slice of m in chunks of 16 rows
  __shared__ __align__(16) float Ds[16][64 + 4];
  __shared__ __align__(16) float As[16][64 + 4];
  const int tid = threadIdx.x;
  const int g = blockIdx.z % groups, split = blockIdx.z / groups;
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int K = p.k1 + p.k2;
  const int64_t m_begin = (int64_t)split * p.rows_per_split;
  const int64_t m_end = m_begin + p.rows_per_split < p.m ? m_begin + p.rows_per_split : p.m;
  const int tx = tid & 15, ty = tid >> 4;          // outputs n = ty*4.., k = tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const int lr = tid >> 4, lc = (tid & 15) * 4;    // loader: row 0..15, 4 consecutive columns
  for (int64_t mb = m_begin; mb < m_end; mb += 16) {
    const int64_t m = mb + lr;
    float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < m_end) {
      if (n0 + lc < p.n) dv = *reinterpret_cast<const float4*>(p.dy + m * p.ldy + (int64_t)g * p.n + n0 + lc);
      const int k = k0 + lc;
      if (k < K) {
        if (p.tap3_nodes > 0) {
          const int cin = p.k1 / 3;
          const int64_t j = m % p.tap3_nodes;
          if (!(j == 0 && k < cin)) av = *reinterpret_cast<const float4*>(p.a1 + (2 * m - 1) * (int64_t)cin + k);
        } else if (k < p.k1) {
          av = *reinterpret_cast<const float4*>(p.a1 + m * p.lda1 + (int64_t)g * p.k1 + k);
        } else {
          av = *reinterpret_cast<const float4*>(p.a2 + m * p.lda2 + (int64_t)g * p.k2 + (k - p.k1));
        }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Ds[lr][lc]) = dv;
    *reinterpret_cast<float4*>(&As[lr][lc]) = av;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const float4 d4 = *reinterpret_cast<const float4*>(&Ds[r][ty * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[r][tx * 4]);
      const float d[4] = {d4.x, d4.y, d4.z, d4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], a[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nn = n0 + ty * 4 + i;
    if (nn >= p.n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(p.dw + ((int64_t)g * p.n + nn) * p.ldw + k, acc[i][j]);
    }
  }
}

// Downsample input gradient: dX (B*2r, Cin) from dA (B*r, 3*Cin) (taps [-1, 0, +1] of node 2j)
__global__ void tap3_bwd_input_kernel(const float* __restrict__ dA, int64_t rows, int r, int cin,
                                      float* __restrict__ dX) {
  const int64_t total = rows * 2 * cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int64_t node = i / cin;            // global input node index
    const int64_t row = node >> 1;           // output row (b, j)
    const int64_t j = row % r;
    float v;
    if ((node & 1) == 0) {
      v = dA[row * 3 * cin + cin + c];                                  // centre tap of row j
    } else {
      v = dA[row * 3 * cin + 2 * cin + c];                              // +1 tap of row j
      if (j + 1 < r) v += dA[(row + 1) * 3 * cin + c];                  // -1 tap of row j+1
    }
    dX[i] = v;
  }
}

__global__ void node_mean_bwd_kernel(const float* __restrict__ dmean, int B, int N, int C,
                                     float* __restrict__ dx) {
  const int64_t total = (int64_t)B * N * C;
  const float inv = 1.0f / (float)N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t b = i / ((int64_t)N * C);
    dx[i] = dmean[b * C + c] * inv;
  }
}

// F.normalize backward: z = v / den, den = max(||v||, eps):  dv = (dz - z (z . dz)) / den
// (for ||v|| < eps the norm is clamped and dv = dz / eps)
__global__ void l2norm_rows_bwd_kernel(const float* __restrict__ v, const float* __restrict__ dz,
                                       int64_t M, int D, float eps, float* __restrict__ dv) {
  const int warps = blockDim.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * warps + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* vr = v + row * D;
  const float* gr = dz + row * D;
  float s = 0.0f, dot = 0.0f;
  for (int c = lane; c < D; c += 32) { s = fmaf(vr[c], vr[c], s); dot = fmaf(vr[c], gr[c], dot); }
  s = warp_sum(s);
  dot = warp_sum(dot);
  const float nrm = sqrtf(s);
  if (nrm < eps) {
    for (int c = lane; c < D; c += 32) dv[row * D + c] = gr[c] / eps;
  } else {
    const float inv = 1.0f / nrm, k = dot * inv * inv * inv;
    for (int c = lane; c < D; c += 32) dv[row * D + c] = gr[c] * inv - vr[c] * k;
  }
}

// ---- peak extractor backward: dW (F,3,pb,pf), db (F) accumulated over segments -------------
__global__ void peak_extract_bwd_kernel(const float* __restrict__ spec, const float* __restrict__ w,
                                        const float* __restrict__ bias, const float* __restrict__ dout,
                                        int n_mels, int n_frames, int F, int pb, int pf,
                                        float* __restrict__ dw, float* __restrict__ db) {
  extern __shared__ float sm[];
  const int hw = n_mels * n_frames;
  const int gh = n_mels / pb, gw = n_frames / pf, nodes = gh * gw, pp = pb * pf;
  float* s_in = sm;                        // hw   normalised spectrogram
  float* s_w = s_in + hw;                  // F*3*pp
  float* s_dz = s_w + F * 3 * pp;          // nodes*F
  __shared__ float s_red[64];
  __shared__ float s_mn, s_mx;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* sp = spec + (size_t)b * hw;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = tid; i < hw; i += nt) { const float v = sp[i]; s_in[i] = v; mn = fminf(mn, v); mx = fmaxf(mx, v); }
  for (int i = tid; i < F * 3 * pp; i += nt) s_w[i] = w[i];
  mn = -warp_max(-mn); mx = warp_max(mx);
  if ((tid & 31) == 0) { s_red[tid >> 5] = mn; s_red[32 + (tid >> 5)] = mx; }
  __syncthreads();
  if (tid == 0) {
    float a = s_red[0], c = s_red[32];
    for (int i = 1; i < (nt >> 5); ++i) { a = fminf(a, s_red[i]); c = fmaxf(c, s_red[32 + i]); }
    s_mn = a; s_mx = c;
  }
  __syncthreads();
  const float lo = s_mn, den = s_mx - s_mn;
  for (int i = tid; i < hw; i += nt) s_in[i] = __fdiv_rn(s_in[i] - lo, den);
  __syncthreads();
  const float tstep = n_frames > 1 ? 1.0f / (float)(n_frames - 1) : 0.0f;
  const float fstep = n_mels > 1 ? 1.0f / (float)(n_mels - 1) : 0.0f;
  auto input = [&](int ch, int y, int xx) -> float {
    if (ch == 0) return (xx < n_frames / 2) ? xx * tstep : 1.0f - (n_frames - 1 - xx) * tstep;
    if (ch == 1) return (y < n_mels / 2) ? y * fstep : 1.0f - (n_mels - 1 - y) * fstep;
    return s_in[y * n_frames + xx];
  };
  // dz = dout * [pre-activation > 0]
  for (int o = tid; o < nodes * F; o += nt) {
    const int node = o / F, f = o - node * F;
    const int gy = node / gw, gx = node - gy * gw;
    float acc = 0.0f;
    for (int ch = 0; ch < 3; ++ch)
      for (int i = 0; i < pb; ++i)
        for (int j = 0; j < pf; ++j)
          acc = fmaf(input(ch, gy * pb + i, gx * pf + j), s_w[(f * 3 + ch) * pp + i * pf + j], acc);
    acc += bias[f];
    s_dz[o] = acc > 0.0f ? dout[((size_t)b * nodes + node) * F + f] : 0.0f;
  }
  __syncthreads();
  for (int o = tid; o < F * 3 * pp + F; o += nt) {
    float acc = 0.0f;
    if (o < F * 3 * pp) {
      const int f = o / (3 * pp), rem = o - f * 3 * pp, ch = rem / pp, ij = rem - ch * pp;
      const int i = ij / pf, j = ij - i * pf;
      for (int node = 0; node < nodes; ++node) {
        const int gy = node / gw, gx = node - gy * gw;
        acc = fmaf(s_dz[node * F + f], input(ch, gy * pb + i, gx * pf + j), acc);
      }
      atomicAdd(dw + o, acc);
    } else {
      const int f = o - F * 3 * pp;
      for (int node = 0; node < nodes; ++node) acc += s_dz[node * F + f];
      atomicAdd(db + f, acc);
    }
  }
}

// ---- gradient clip + Adam -------------------------------------------------------------------
__global__ void sq_norm_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
  __shared__ double red[32];
  double a = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    a += (double)g[i] * (double)g[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) a += red[i];
    atomicAdd(out, a);
  }
}

__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g,
                                 float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                 float b1, float b2, float eps, float bc1, float bc2, float max_norm,
                                 const double* __restrict__ sqnorm) {
  float coef = 1.0f;
  if (max_norm > 0.0f) {
    const float total = (float)sqrt(*sqnorm);
    coef = fminf(max_norm / (total + 1e-6f), 1.0f);
  }
  const float step = lr / bc1, rs = 1.0f / sqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step * (mi / (sqrtf(vi) * rs + eps));
  }
}

// Graph-capturable form: the step counter, the learning rate and the NaN guard live in device memory
__global__ void adam_step_counter_kernel(int* step, const float* loss_guard) {
  if (loss_guard && isnan(*loss_guard)) return;
  *step += 1;
}

__global__ void adam_clip_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                     float* __restrict__ m, float* __restrict__ v, int64_t n,
                                     const float* __restrict__ lr_dev, float b1, float b2, float eps,
                                     const int* __restrict__ step_dev, float max_norm,
                                     const double* __restrict__ sqnorm, const float* __restrict__ loss_guard) {
  if (loss_guard && isnan(*loss_guard)) return;            // train.py:65-68: skip the batch
  float coef = 1.0f;
  if (max_norm > 0.0f) {
    const float total = (float)sqrt(*sqnorm);
    coef = fminf(max_norm / (total + 1e-6f), 1.0f);
  }
  const int t = *step_dev;
  const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
  const float step = *lr_dev / bc1, rs = 1.0f / sqrtf(bc2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step * (mi / (sqrtf(vi) * rs + eps));
  }
}

__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] += x[i];
}

static inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace grafp

using namespace grafp;

extern "C" {

int grafp_col_stats(const float* x, int64_t M, int C, int64_t ld, double* sum, double* sumsq,
                    void* stream) {
  GRAFP_REQUIRE(M >= 0 && C > 0 && (M == 0 || (x && sum && sumsq)), "col_stats: bad arguments");
  if (M == 0) return 0;
  const int rows = 512;
  dim3 grid((unsigned)((M + rows - 1) / rows), (C + 31) / 32);
  col_reduce_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, nullptr, M, C, ld, 0, 0, nullptr, nullptr,
                                                        nullptr, nullptr, 0, 0.f, rows, sum, sumsq);
  return check_launch("col_stats");
}

int grafp_bn_finalize(const double* sum, const double* sumsq, int64_t M, int C, const float* gamma,
                      const float* beta, const float* conv_bias, float eps, float momentum,
                      float* running_mean, float* running_var, float* scale, float* shift,
                      float* mean, float* invstd, void* stream) {
  GRAFP_REQUIRE(sum && sumsq && scale && shift && mean && invstd && M > 0 && C > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(
      sum, sumsq, M, C, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale, shift,
      mean, invstd);
  return check_launch("bn_finalize");
}

int grafp_bn_finalize_apply(const double* sum, const double* sumsq, int64_t M, int C, const float* gamma,
                            const float* beta, const float* conv_bias, float eps, float momentum,
                            float* running_mean, float* running_var, float* scale, float* shift, float* mean,
                            float* invstd, const float* x, int64_t ld, int act, float act_param,
                            const float* residual, int64_t ldr, float* out, int64_t ldo, void* stream) {
  GRAFP_REQUIRE(sum && sumsq && scale && shift && mean && invstd && x && out && M > 0 && C > 0 && C % 4 == 0 &&
                    C <= 8192, "bn_finalize_apply: bad arguments");
  int64_t blocks = (M * (C / 4) + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();     // every block repeats the C-column finalize
  bn_finalize_apply_kernel<<<(unsigned)blocks, 256, 2 * (size_t)C * sizeof(float), as_stream(stream)>>>(
      sum, sumsq, M, C, gamma, beta, conv_bias, eps, momentum, running_mean, running_var, scale, shift, mean,
      invstd, x, ld, act, act_param, residual, ldr, out, ldo);
  return check_launch("bn_finalize_apply");
}

int grafp_affine_act(const float* x, int64_t M, int C, int64_t ld, const float* scale,
                     const float* shift, int act, float act_param, const float* residual,
                     int64_t ldr, float* out, int64_t ldo, void* stream) {
  GRAFP_REQUIRE(M >= 0 && C > 0 && C % 4 == 0 && (M == 0 || (x && out)), "affine_act: bad arguments");
  if (M == 0) return 0;
  affine_act_kernel<<<grid_for(M * (C / 4)), 256, 0, as_stream(stream)>>>(x, M, C, ld, scale, shift, act,
                                                                        act_param, residual, ldr, out, ldo);
  return check_launch("affine_act");
}

int grafp_bn_bwd_reduce(const float* dout, int64_t ldd, const float* raw, int64_t ld, int64_t M, int C,
                        const float* scale, const float* shift, const float* mean, const float* invstd,
                        int act, float act_param, double* sum_dz, double* sum_dz_xhat, void* stream) {
  GRAFP_REQUIRE(M > 0 && C > 0 && dout && raw && scale && shift && mean && invstd && sum_dz && sum_dz_xhat,
                "bn_bwd_reduce: bad arguments");
  const int rows = 512;
  dim3 grid((unsigned)((M + rows - 1) / rows), (C + 31) / 32);
  col_reduce_kernel<<<grid, 256, 0, as_stream(stream)>>>(raw, dout, M, C, ld, ldd, 1, scale, shift, mean,
                                                        invstd, act, act_param, rows, sum_dz, sum_dz_xhat);
  return check_launch("bn_bwd_reduce");
}

int grafp_bn_bwd_apply(const float* dout, int64_t ldd, const float* raw, int64_t ld, int64_t M, int C,
                       const float* scale, const float* shift, const float* mean, const float* invstd,
                       int act, float act_param, int bn, const double* sum_dz, const double* sum_dz_xhat,
                       float* draw, int64_t ldo, float* dgamma, float* dbeta, void* stream) {
  GRAFP_REQUIRE(M > 0 && C > 0 && dout && raw && scale && shift && mean && invstd && draw,
                "bn_bwd_apply: bad arguments");
  GRAFP_REQUIRE((!dgamma && !dbeta) || (sum_dz && sum_dz_xhat), "bn_bwd_apply: parameter gradients need the column sums");
  bn_bwd_apply_kernel<<<grid_for(M * C), 256, 0, as_stream(stream)>>>(
      dout, raw, M, C, ldd, ld, scale, shift, mean, invstd, act, act_param, bn, sum_dz, sum_dz_xhat, draw, ldo,
      dgamma, dbeta);
  return check_launch("bn_bwd_apply");
}

int grafp_bn_param_grad(const double* sum_dz, const double* sum_dz_xhat, int C, float* dgamma,
                        float* dbeta, void* stream) {
  GRAFP_REQUIRE(sum_dz && sum_dz_xhat && C > 0, "bn_param_grad: bad arguments");
  bn_param_grad_kernel<<<(C + 127) / 128, 128, 0, as_stream(stream)>>>(sum_dz, sum_dz_xhat, C, dgamma, dbeta);
  return check_launch("bn_param_grad");
}

size_t grafp_gemm_wgrad_workspace_bytes(int64_t m, int n, int k1, int k2, int groups, int tap3_nodes) {
  if (m <= 0 || n <= 0 || k1 <= 0 || k2 < 0 || groups <= 0) return 0;
  // (row strides are checked at launch; packed operands are assumed for the query)
  if (!grafp::wgrad_tc_supported(m, n, k1, k2, groups, tap3_nodes, 4, 4, 4, 4)) return 0;
  return grafp::wgrad_tc_workspace_bytes(m, n, k1, k2, groups);
}

int grafp_gemm_wgrad(const float* dy, int64_t ldy, const float* a1, int64_t lda1, int k1,
                     const float* a2, int64_t lda2, int k2, int64_t m, int n, int groups,
                     int tap3_nodes, float* dw, int64_t ldw, int engine, void* workspace, size_t workspace_bytes,
                     void* stream) {
  GRAFP_REQUIRE(dy && a1 && dw && m > 0 && n > 0 && groups > 0 && k1 > 0 && k2 >= 0, "gemm_wgrad: bad arguments");
  GRAFP_REQUIRE(k1 % 4 == 0 && k2 % 4 == 0 && n % 4 == 0 && ldy % 4 == 0, "gemm_wgrad: sizes must be multiples of 4");
  GRAFP_REQUIRE((k2 == 0) == (a2 == nullptr), "gemm_wgrad: a2/k2 mismatch");
  GRAFP_REQUIRE(engine == GRAFP_ENGINE_AUTO || engine == GRAFP_ENGINE_SIMT || engine == GRAFP_ENGINE_TC_3XTF32,
                "gemm_wgrad: engine must be AUTO, SIMT or TC_3XTF32 (got %d)", engine);
  {
    const bool aligned = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(a1) |
                           reinterpret_cast<uintptr_t>(a2) | reinterpret_cast<uintptr_t>(dw)) & 15) == 0;
    const bool tc_ok = aligned && grafp::wgrad_tc_supported(m, n, k1, k2, groups, tap3_nodes, ldy, lda1, lda2, ldw) &&
                       workspace && workspace_bytes >= grafp::wgrad_tc_workspace_bytes(m, n, k1, k2, groups);
    if (engine == GRAFP_ENGINE_TC_3XTF32)
      GRAFP_REQUIRE(tc_ok, "gemm_wgrad: the tcgen05 engine needs n, k1, k2 multiples of 32, no tap3, 16-byte aligned "
                           "operands and a workspace of grafp_gemm_wgrad_workspace_bytes()");
    if (engine != GRAFP_ENGINE_SIMT && tc_ok)
      return grafp::wgrad_tc_launch(dy, ldy, a1, lda1, k1, a2, lda2, k2, m, n, groups, dw, ldw,
                                    static_cast<float*>(workspace), as_stream(stream));
  }
  WgradP p;
  p.dy = dy; p.ldy = ldy; p.a1 = a1; p.lda1 = lda1; p.k1 = k1; p.a2 = a2; p.lda2 = lda2; p.k2 = k2;
  p.dw = dw; p.ldw = ldw; p.m = m; p.n = n; p.tap3_nodes = tap3_nodes;
  const int tiles = ((n + 63) / 64) * ((k1 + k2 + 63) / 64) * groups;
  int64_t splits = (4LL * sm_count() + tiles - 1) / tiles;
  const int64_t max_splits = (m + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t rows = (m + splits - 1) / splits;
  rows = ((rows + 15) / 16) * 16;
  splits = (m + rows - 1) / rows;
  p.rows_per_split = (int)rows;
  GRAFP_REQUIRE(splits * groups <= 65535, "gemm_wgrad: too many splits");
  dim3 grid((n + 63) / 64, (k1 + k2 + 63) / 64, (unsigned)(splits * groups));
  gemm_wgrad_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, groups);
  return check_launch("gemm_wgrad");
}

int grafp_tap3_bwd_input(const float* dA, int64_t rows, int rows_per_graph, int cin, float* dX,
                         void* stream) {
  GRAFP_REQUIRE(dA && dX && rows > 0 && rows_per_graph > 0 && cin > 0 && rows % rows_per_graph == 0,
                "tap3_bwd_input: bad arguments");
  tap3_bwd_input_kernel<<<grid_for(rows * 2 * cin), 256, 0, as_stream(stream)>>>(dA, rows, rows_per_graph, cin, dX);
  return check_launch("tap3_bwd_input");
}

int grafp_node_mean_bwd(const float* dmean, int B, int N, int C, float* dx, void* stream) {
  GRAFP_REQUIRE(B >= 0 && N > 0 && C > 0 && (B == 0 || (dmean && dx)), "node_mean_bwd: bad arguments");
  if (B == 0) return 0;
  node_mean_bwd_kernel<<<grid_for((int64_t)B * N * C), 256, 0, as_stream(stream)>>>(dmean, B, N, C, dx);
  return check_launch("node_mean_bwd");
}

int grafp_l2_normalize_rows_bwd(const float* v, const float* dz, int64_t M, int D, float eps, float* dv,
                                void* stream) {
  GRAFP_REQUIRE(M >= 0 && D > 0 && (M == 0 || (v && dz && dv)), "l2_normalize_rows_bwd: bad arguments");
  if (M == 0) return 0;
  l2norm_rows_bwd_kernel<<<(unsigned)((M + 7) / 8), 256, 0, as_stream(stream)>>>(v, dz, M, D, eps, dv);
  return check_launch("l2_normalize_rows_bwd");
}

int grafp_peak_extract_bwd(const float* spec, const float* w, const float* bias, const float* dout,
                           int B, int n_mels, int n_frames, int F, int pb, int pf, float* dw, float* db,
                           void* stream) {
  GRAFP_REQUIRE(B == 0 || (spec && w && bias && dout && dw && db), "peak_extract_bwd: null pointer");
  GRAFP_REQUIRE(pb > 0 && pf > 0 && n_mels % pb == 0 && n_frames % pf == 0, "peak_extract_bwd: bad patch");
  if (B == 0) return 0;
  const int nodes = (n_mels / pb) * (n_frames / pf);
  const size_t smem = ((size_t)n_mels * n_frames + (size_t)F * 3 * pb * pf + (size_t)nodes * F) * sizeof(float);
  GRAFP_REQUIRE(smem <= 200 * 1024, "peak_extract_bwd: segment too large for shared memory");
  cudaFuncSetAttribute(peak_extract_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  peak_extract_bwd_kernel<<<B, 512, smem, as_stream(stream)>>>(spec, w, bias, dout, n_mels, n_frames, F, pb,
                                                              pf, dw, db);
  return check_launch("peak_extract_bwd");
}

int grafp_sq_norm(const float* g, int64_t n, double* out_accum, void* stream) {
  GRAFP_REQUIRE(n >= 0 && out_accum && (n == 0 || g), "sq_norm: bad arguments");
  if (n == 0) return 0;
  sq_norm_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(g, n, out_accum);
  return check_launch("sq_norm");
}

int grafp_adam_clip_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                         float beta2, float eps, int step, float max_norm, const double* sq_norm,
                         void* stream) {
  GRAFP_REQUIRE(n >= 0 && step >= 1 && (n == 0 || (p && g && m && v)), "adam_clip_step: bad arguments");
  GRAFP_REQUIRE(max_norm <= 0.0f || sq_norm, "adam_clip_step: clipping needs the squared gradient norm");
  if (n == 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adam_clip_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2,
                                                              max_norm, sq_norm);
  return check_launch("adam_clip_step");
}

int grafp_adam_clip_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev,
                             float beta1, float beta2, float eps, int* step_dev, float max_norm,
                             const double* sq_norm, const float* loss_guard, void* stream) {
  GRAFP_REQUIRE(n >= 0 && lr_dev && step_dev && (n == 0 || (p && g && m && v)), "adam_clip_step_dev: bad arguments");
  GRAFP_REQUIRE(max_norm <= 0.0f || sq_norm, "adam_clip_step_dev: clipping needs the squared gradient norm");
  adam_step_counter_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev, loss_guard);
  if (int rc = check_launch("adam_step_counter")) return rc;
  if (n == 0) return 0;
  adam_clip_dev_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr_dev, beta1, beta2, eps,
                                                                  step_dev, max_norm, sq_norm, loss_guard);
  return check_launch("adam_clip_step_dev");
}

int grafp_add_inplace(float* y, const float* x, int64_t n, void* stream) {
  GRAFP_REQUIRE(n >= 0 && (n == 0 || (x && y)), "add_inplace: bad arguments");
  if (n == 0) return 0;
  add_inplace_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>(y, x, n);
  return check_launch("add_inplace");
}

}  // extern "C"
