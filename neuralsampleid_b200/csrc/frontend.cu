// Log-mel front end pieces (reference modules/transformations.py:27-34 builds
// torchaudio MelSpectrogram(n_fft 1024, win 1024, hop 512, 64 mels) + AmplitudeToDB, and :96-104 cuts the
// (T, n_mels) spectrogram into overlapping 128-frame segments).  The two contractions -- frames x DFT basis
// and power x mel filterbank -- run on the GEMM engine; these kernels are the glue around them:
//   frame_window_kernel   centred, reflect-padded framing times the analysis window  -> (T, n_fft)
//   power_kernel          |re|^2 + |im|^2 of the DFT GEMM output                      -> (T, bins padded)
//   amplitude_to_db_kernel  10 log10(max(x, amin))                                    (AmplitudeToDB, power)
//   unfold_segments_kernel  out[s, m, f] = db[s * step + f, m]                        -> (S, n_mels, n_frames)
#include "common.cuh"

namespace grafp {

__global__ void frame_window_kernel(const float* __restrict__ wave, int64_t L, const float* __restrict__ win,
                                    int n_fft, int hop, int64_t T, float* __restrict__ out) {
  const int64_t total = T * n_fft;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / n_fft;
    const int n = (int)(i - t * n_fft);
    int64_t j = t * hop + n - n_fft / 2;             // center=True: frame t is centred on sample t * hop
    if (j < 0) j = -j;                               // pad_mode="reflect"
    if (j >= L) j = 2 * (L - 1) - j;
    out[i] = wave[j] * win[n];
  }
}

__global__ void power_kernel(const float* __restrict__ z, int64_t ldz, int64_t T, int bins, int im_offset,
                             float* __restrict__ p, int64_t ldp) {
  const int64_t total = T * ldp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / ldp;
    const int k = (int)(i - t * ldp);
    float v = 0.0f;
    if (k < bins) {
      const float re = z[t * ldz + k], im = z[t * ldz + im_offset + k];
      v = fmaf(im, im, re * re);
    }
    p[i] = v;
  }
}

__global__ void amplitude_to_db_kernel(const float* __restrict__ x, int64_t count, float multiplier, float amin,
                                       float db_offset, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = multiplier * log10f(fmaxf(x[i], amin)) - db_offset;
}

__global__ void unfold_segments_kernel(const float* __restrict__ db, int64_t ldm, int n_mels, int n_frames, int step,
                                       int64_t S, float* __restrict__ out) {
  const int64_t total = S * n_mels * n_frames;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % n_frames);
    const int m = (int)((i / n_frames) % n_mels);
    const int64_t s = i / ((int64_t)n_frames * n_mels);
    out[i] = db[(s * step + f) * ldm + m];
  }
}

static inline unsigned grid_for(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace grafp

using namespace grafp;

extern "C" {

int grafp_frame_window_fwd(const float* wave, int64_t L, const float* win, int n_fft, int hop, int64_t T,
                           float* out, void* stream) {
  GRAFP_REQUIRE(T <= 0 || (wave && win && out), "frame_window: null pointer");
  GRAFP_REQUIRE(n_fft > 0 && hop > 0 && T >= 0, "frame_window: bad sizes");
  GRAFP_REQUIRE(T == 0 || L > n_fft / 2, "frame_window: reflect padding needs more than n_fft/2 samples (L=%lld)", (long long)L);
  GRAFP_REQUIRE(T == 0 || (T - 1) * hop < L + n_fft / 2, "frame_window: T=%lld frames do not fit L=%lld samples", (long long)T, (long long)L);
  if (T == 0) return 0;
  frame_window_kernel<<<grid_for(T * n_fft), 256, 0, as_stream(stream)>>>(wave, L, win, n_fft, hop, T, out);
  return check_launch("frame_window");
}

int grafp_power_spectrum_fwd(const float* z, int64_t ldz, int64_t T, int bins, int im_offset, float* p, int64_t ldp,
                             void* stream) {
  GRAFP_REQUIRE(T <= 0 || (z && p), "power_spectrum: null pointer");
  GRAFP_REQUIRE(bins > 0 && ldp >= bins && im_offset >= bins && ldz >= im_offset + bins && T >= 0, "power_spectrum: bad sizes");
  if (T == 0) return 0;
  power_kernel<<<grid_for(T * ldp), 256, 0, as_stream(stream)>>>(z, ldz, T, bins, im_offset, p, ldp);
  return check_launch("power_spectrum");
}

int grafp_amplitude_to_db_fwd(const float* x, int64_t count, float multiplier, float amin, float db_offset, float* out,
                              void* stream) {
  GRAFP_REQUIRE(count <= 0 || (x && out), "amplitude_to_db: null pointer");
  GRAFP_REQUIRE(amin > 0.0f, "amplitude_to_db: amin must be positive");
  if (count <= 0) return 0;
  amplitude_to_db_kernel<<<grid_for(count), 256, 0, as_stream(stream)>>>(x, count, multiplier, amin, db_offset, out);
  return check_launch("amplitude_to_db");
}

int grafp_unfold_segments_fwd(const float* db, int64_t ldm, int64_t T, int n_mels, int n_frames, int step, int64_t S,
                              float* out, void* stream) {
  GRAFP_REQUIRE(S <= 0 || (db && out), "unfold_segments: null pointer");
  GRAFP_REQUIRE(n_mels > 0 && n_frames > 0 && step > 0 && ldm >= n_mels, "unfold_segments: bad sizes");
  GRAFP_REQUIRE(S <= 0 || (S - 1) * step + n_frames <= T, "unfold_segments: S=%lld segments do not fit T=%lld frames",
                (long long)S, (long long)T);
  if (S <= 0) return 0;
  unfold_segments_kernel<<<grid_for(S * n_mels * n_frames), 256, 0, as_stream(stream)>>>(db, ldm, n_mels, n_frames, step,
                                                                                         S, out);
  return check_launch("unfold_segments");
}

}  // extern "C"
