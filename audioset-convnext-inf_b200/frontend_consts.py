"""Frozen front-end constants: windowed DFT kernels and the Slaney mel filterbank.

These are the tensors the reference obtains from torchlibrosa 0.0.9 / librosa 0.8.1 when it
constructs `Spectrogram(n_fft=1024, hop=320, window='hann', center=True, pad_mode='reflect')`
and `LogmelFilterBank(sr=32000, n_fft=1024, n_mels=224, fmin=50, fmax=14000, ...)`
(reference convnext.py:179-200) and that a checkpoint stores under
`spectrogram_extractor.stft.conv_{real,imag}.weight` and `logmel_extractor.melW`.
A loaded checkpoint always overrides them; they only matter for a randomly initialised model.
"""
import numpy as np
import torch


def windowed_dft(n_fft):
    """(conv_real, conv_imag): float32 (n_fft//2+1, 1, n_fft); periodic Hann x exp(-2 pi i n k / N)."""
    n = np.arange(n_fft)
    hann = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)            # get_window('hann', fftbins=True)
    omega = np.exp(-2j * np.pi / n_fft)
    basis = np.power(omega, np.outer(n, n))[:, : n_fft // 2 + 1]   # torchlibrosa DFTBase.dft_matrix
    kern = basis * hann[:, None]
    real = np.ascontiguousarray(kern.real.T).astype(np.float32)
    imag = np.ascontiguousarray(kern.imag.T).astype(np.float32)
    return torch.from_numpy(real)[:, None, :], torch.from_numpy(imag)[:, None, :]


def _mel_of_hz(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f * 3.0 / 200.0
    log_part = 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * (27.0 / np.log(6.4))
    return np.where(f >= 1000.0, log_part, lin)


def _hz_of_mel(m):
    m = np.asarray(m, dtype=np.float64)
    lin = m * 200.0 / 3.0
    log_part = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0))
    return np.where(m >= 15.0, log_part, lin)


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """float32 (n_fft//2+1, n_mels): librosa.filters.mel(htk=False, norm='slaney').T"""
    n_bins = n_fft // 2 + 1
    bin_hz = np.linspace(0.0, sr / 2.0, n_bins)
    edges = _hz_of_mel(np.linspace(_mel_of_hz(fmin), _mel_of_hz(fmax), n_mels + 2))
    width = np.diff(edges)
    offset = edges[:, None] - bin_hz[None, :]
    rising = -offset[:-2] / width[:-1, None]
    falling = offset[2:] / width[1:, None]
    tri = np.maximum(0.0, np.minimum(rising, falling)).astype(np.float32)
    tri *= (2.0 / (edges[2:] - edges[:-2]))[:, None]               # Slaney area normalisation
    return torch.from_numpy(np.ascontiguousarray(tri.T))
