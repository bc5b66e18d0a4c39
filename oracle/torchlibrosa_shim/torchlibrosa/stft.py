"""TEST INFRASTRUCTURE ONLY. Restatement of torchlibrosa.stft.{Spectrogram,LogmelFilterBank}.

Third-party dependency absent from /root/reference: torchlibrosa==0.0.9 (which builds its
constants with librosa==0.8.1).  The reference's call sites: convnext.py:179-187
(Spectrogram), convnext.py:190-200 (LogmelFilterBank), use at convnext.py:298-299.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ---- librosa 0.8.1 constants (filters.mel / filters.get_window / util.pad_center) -------
def _hz_to_mel(freqs):
    """Slaney mel scale (librosa htk=False)."""
    freqs = np.asanyarray(freqs, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = freqs / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if freqs.ndim:
        log_t = freqs >= min_log_hz
        mels[log_t] = min_log_mel + np.log(freqs[log_t] / min_log_hz) / logstep
    elif freqs >= min_log_hz:
        mels = min_log_mel + np.log(freqs / min_log_hz) / logstep
    return mels


def _mel_to_hz(mels):
    mels = np.asanyarray(mels, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def librosa_mel(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with defaults htk=False,
    norm='slaney', dtype=float32.  Returns (n_mels, 1 + n_fft//2) float32."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    weights = np.zeros((n_mels, n_bins), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    min_mel = _hz_to_mel(fmin)
    max_mel = _hz_to_mel(fmax)
    mel_f = _mel_to_hz(np.linspace(min_mel, max_mel, n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True) == periodic Hann, float64."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def dft_matrix(n):
    """torchlibrosa DFTBase.dft_matrix: W[x, y] = exp(-2*pi*1j/n) ** (x*y), complex128."""
    (x, y) = np.meshgrid(np.arange(n), np.arange(n))
    omega = np.exp(-2 * np.pi * 1j / n)
    return np.power(omega, x * y)


class STFT(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann",
                 center=True, pad_mode="reflect", freeze_parameters=True):
        super().__init__()
        assert pad_mode in ["constant", "reflect"]
        assert window == "hann", "shim restates the only window the reference uses (convnext.py:161)"
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length if win_length is not None else n_fft
        self.window = window
        self.center = center
        self.pad_mode = pad_mode
        if self.hop_length is None:
            self.hop_length = int(self.win_length // 4)
        fft_window = hann_periodic(self.win_length)
        if self.win_length < n_fft:  # librosa.util.pad_center
            lpad = (n_fft - self.win_length) // 2
            fft_window = np.pad(fft_window, (lpad, n_fft - self.win_length - lpad))
        W = dft_matrix(n_fft)
        out_channels = n_fft // 2 + 1
        self.conv_real = nn.Conv1d(1, out_channels, kernel_size=n_fft, stride=self.hop_length,
                                   padding=0, dilation=1, groups=1, bias=False)
        self.conv_imag = nn.Conv1d(1, out_channels, kernel_size=n_fft, stride=self.hop_length,
                                   padding=0, dilation=1, groups=1, bias=False)
        self.conv_real.weight.data = torch.Tensor(
            np.real(W[:, 0:out_channels] * fft_window[:, None]).T)[:, None, :]
        self.conv_imag.weight.data = torch.Tensor(
            np.imag(W[:, 0:out_channels] * fft_window[:, None]).T)[:, None, :]
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, input):
        x = input[:, None, :]
        if self.center:
            x = F.pad(x, pad=(self.n_fft // 2, self.n_fft // 2), mode=self.pad_mode)
        real = self.conv_real(x)
        imag = self.conv_imag(x)
        real = real[:, None, :, :].transpose(2, 3)
        imag = imag[:, None, :, :].transpose(2, 3)
        return real, imag


class Spectrogram(nn.Module):
    def __init__(self, n_fft=2048, hop_length=None, win_length=None, window="hann",
                 center=True, pad_mode="reflect", power=2.0, freeze_parameters=True):
        super().__init__()
        self.power = power
        self.stft = STFT(n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                         window=window, center=center, pad_mode=pad_mode, freeze_parameters=True)

    def forward(self, input):
        (real, imag) = self.stft.forward(input)
        spectrogram = real ** 2 + imag ** 2
        if self.power == 2.0:
            pass
        else:
            spectrogram = spectrogram ** (self.power / 2.0)
        return spectrogram


class LogmelFilterBank(nn.Module):
    def __init__(self, sr=22050, n_fft=2048, n_mels=64, fmin=0.0, fmax=None, is_log=True,
                 ref=1.0, amin=1e-10, top_db=80.0, freeze_parameters=True):
        super().__init__()
        self.is_log = is_log
        self.ref = ref
        self.amin = amin
        self.top_db = top_db
        if fmax is None:
            fmax = sr // 2
        self.melW = nn.Parameter(torch.Tensor(librosa_mel(sr, n_fft, n_mels, fmin, fmax).T))
        if freeze_parameters:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, input):
        mel_spectrogram = torch.matmul(input, self.melW)
        if self.is_log:
            return self.power_to_db(mel_spectrogram)
        return mel_spectrogram

    def power_to_db(self, input):
        ref_value = self.ref
        log_spec = 10.0 * torch.log10(torch.clamp(input, min=self.amin, max=np.inf))
        log_spec -= 10.0 * np.log10(np.maximum(self.amin, ref_value))
        if self.top_db is not None:
            if self.top_db < 0:
                raise ValueError("top_db must be non-negative")
            log_spec = torch.clamp(log_spec, min=log_spec.max().item() - self.top_db, max=np.inf)
        return log_spec
