"""TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of the reference's log-mel front end:
``modules/transformations.py:27-34`` builds ``nn.Sequential(torchaudio.transforms.MelSpectrogram(sample_rate,
win_length, hop_length, n_fft, n_mels), AmplitudeToDB())`` and ``:96-104`` (eval branch) transposes the (n_mels, T)
spectrogram to (T, n_mels) and ``unfold``s it into n_frames-long segments every ``int(n_frames * (1 - overlap))``
frames.  The arithmetic lives in torchaudio (third-party, torchaudio==2.3.0 pinned by requirements.txt, not part
of /root/reference); its published algorithm is restated with plain torch ops: power spectrogram of a centred,
reflect-padded STFT with a periodic Hann window, HTK mel filterbank without normalisation
(``torchaudio.functional.melscale_fbanks``), ``10 * log10(clamp(x, 1e-10))``.  Pinned against the torchaudio of
this image and a committed vector (tests/test_oracle_golden.py)."""
import math

import torch


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(..., norm=None, mel_scale='htk') -> (n_freqs, n_mels), fp32."""
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


def log_mel_spectrogram(wave: torch.Tensor, sample_rate: int, n_fft: int, win_length: int, hop_length: int,
                        n_mels: int) -> torch.Tensor:
    """wave (L,) -> (n_mels, T) dB, T = 1 + L // hop_length."""
    window = torch.hann_window(win_length, periodic=True)
    spec = torch.stft(wave, n_fft, hop_length, win_length, window, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    power = spec.abs().pow(2.0)                                                  # (n_fft/2+1, T)
    fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, float(sample_rate // 2), n_mels, sample_rate)
    mel = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)            # (n_mels, T)
    return 10.0 * torch.log10(torch.clamp(mel, min=1e-10))                       # AmplitudeToDB(power), ref 1.0


def segment_spectrogram(X: torch.Tensor, n_frames: int, overlap: float) -> torch.Tensor:
    """Eval branch of GPUTransformSampleID.forward (transformations.py:96-104): (n_mels, T) -> (S, n_mels, n_frames)."""
    Xt = X.transpose(1, 0)
    try:
        return Xt.unfold(0, size=n_frames, step=int(n_frames * (1 - overlap)))
    except RuntimeError:
        return Xt
