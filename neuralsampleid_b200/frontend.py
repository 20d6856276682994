"""Log-mel front end on the B200: waveform -> (S, n_mels, n_frames) spectrogram
segments, i.e. what ``GPUTransformSampleID`` produces in its eval branch (reference modules/transformations.py:27-34
``MelSpectrogram`` + ``AmplitudeToDB``, :96-104 ``unfold``) and what ``SimCLR`` / the peak extractor consume.

    frames  = window * reflect-padded, centred frames            grafp_frame_window_fwd   (T, n_fft)
    Z       = frames @ [cos | -sin]^T                            grafp_gemm_fwd, fp32 engine (T, 2 * bins padded)
    P       = re^2 + im^2                                        grafp_power_spectrum_fwd
    mel     = P @ filterbank                                     grafp_gemm_fwd, fp32 engine
    dB      = 10 log10(max(mel, 1e-10))                          grafp_amplitude_to_db_fwd
    segments[s, m, f] = dB[s * step + f, m]                      grafp_unfold_segments_fwd

The DFT is a dense contraction (2.1 MFLOP per frame): on this part that is cheaper than a multi-pass FFT and it
reuses the GEMM engine.  The fp32 FFMA engine is used on purpose: a quiet mel band sits 100 dB under the frame's
energy and a 3 x bf16 operand split (1e-5 of the frame norm) would swamp it."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib, ops
from ._lib import GrafpError, check
from ._prep import Linear


def _mel_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') (host-side constant): (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class LogMelSpectrogram(torch.nn.Module):
    """``nn.Sequential(MelSpectrogram(sample_rate, win_length, hop_length, n_fft, n_mels), AmplitudeToDB())`` of the
    reference (cfg keys fs / win_len / hop_len / n_fft / n_mels), plus the eval-branch segmentation."""

    def __init__(self, cfg):
        super().__init__()
        self.sample_rate, self.n_fft = int(cfg["fs"]), int(cfg["n_fft"])
        self.win_length, self.hop_length, self.n_mels = int(cfg["win_len"]), int(cfg["hop_len"]), int(cfg["n_mels"])
        self.n_frames = int(cfg.get("n_frames", 128))
        self.overlap = float(cfg.get("overlap", 0.875))
        if self.win_length > self.n_fft:
            raise ValueError("win_len must not exceed n_fft")
        bins = self.n_fft // 2 + 1
        self.bins = bins
        self.bins_pad = (bins + 31) // 32 * 32
        win = torch.zeros(self.n_fft)
        off = (self.n_fft - self.win_length) // 2                   # torch.stft centres a short window
        win[off:off + self.win_length] = torch.hann_window(self.win_length, periodic=True)
        self.register_buffer("window", win, persistent=False)
        # DFT basis rows: cos(2 pi k n / N) for k < bins, then -sin(2 pi k n / N); float64 -> fp32
        k = torch.arange(bins, dtype=torch.float64).unsqueeze(1)
        n = torch.arange(self.n_fft, dtype=torch.float64).unsqueeze(0)
        ang = 2.0 * math.pi * ((k * n) % self.n_fft) / self.n_fft
        basis = torch.zeros((2 * self.bins_pad, self.n_fft), dtype=torch.float64)
        basis[:bins] = torch.cos(ang)
        basis[self.bins_pad:self.bins_pad + bins] = -torch.sin(ang)
        self.register_buffer("dft_basis", basis.float(), persistent=False)
        fb = torch.zeros((self.n_mels, self.bins_pad))
        fb[:, :bins] = _mel_fbanks_htk(bins, 0.0, float(self.sample_rate // 2), self.n_mels, self.sample_rate).t()
        self.register_buffer("mel_fb", fb.contiguous(), persistent=False)

    def _db_frames(self, wave: torch.Tensor):
        """wave (L,) fp32 CUDA -> dB mel spectrogram in (T, n_mels) layout."""
        if wave.dim() != 1:
            raise ValueError("expected a mono waveform (L,)")
        wave = ops._chk(wave, name="wave")
        L = wave.shape[0]
        T = 1 + L // self.hop_length
        lib = _lib.load()
        P = C.c_void_p
        dev = wave.device
        frames = torch.empty((T, self.n_fft), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            st = ops._stream(wave)
            check(lib.grafp_frame_window_fwd(P(wave.data_ptr()), L, P(self.window.data_ptr()), self.n_fft,
                                             self.hop_length, T, P(frames.data_ptr()), st), "frame_window")
            z = ops.linear(frames, Linear(self.dft_basis, None, None), engine=_lib.ENGINE_SIMT)      # (T, 2*bins_pad)
            power = torch.empty((T, self.bins_pad), device=dev, dtype=torch.float32)
            check(lib.grafp_power_spectrum_fwd(P(z.data_ptr()), z.stride(0), T, self.bins, self.bins_pad,
                                               P(power.data_ptr()), power.stride(0), st), "power_spectrum")
            mel = ops.linear(power, Linear(self.mel_fb, None, None), engine=_lib.ENGINE_SIMT)        # (T, n_mels)
            db = torch.empty_like(mel)
            check(lib.grafp_amplitude_to_db_fwd(P(mel.data_ptr()), mel.numel(), 10.0, 1e-10, 0.0, P(db.data_ptr()), st),
                  "amplitude_to_db")
        return db, T

    def forward(self, wave: torch.Tensor) -> torch.Tensor:
        """(L,) -> (n_mels, T) dB, like the reference's ``logmelspec`` on a mono waveform."""
        db, T = self._db_frames(wave)
        return ops.nodes_to_nchw(db, 1, T)[0]

    def segments(self, wave: torch.Tensor) -> torch.Tensor:
        """(L,) -> (S, n_mels, n_frames): the eval branch of GPUTransformSampleID.forward (transformations.py:96-104).
        A waveform shorter than one segment returns the (T, n_mels) spectrogram, as the reference's ``except``."""
        db, T = self._db_frames(wave)
        step = int(self.n_frames * (1 - self.overlap))
        if T < self.n_frames:
            return db
        S = (T - self.n_frames) // step + 1
        out = torch.empty((S, self.n_mels, self.n_frames), device=db.device, dtype=torch.float32)
        with torch.cuda.device(db.device):
            check(_lib.load().grafp_unfold_segments_fwd(C.c_void_p(db.data_ptr()), db.stride(0), T, self.n_mels,
                                                        self.n_frames, step, S, C.c_void_p(out.data_ptr()),
                                                        ops._stream(db)), "unfold_segments")
        return out
