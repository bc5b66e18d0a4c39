"""Clip preparation in front of the hot path: what reference demo_convnext.py:52-67 does with
torchaudio on the host -- resample to 32 kHz, then constant-pad or crop to 10 s -- as ONE libacx launch on the GPU.

The resampler is torchaudio.functional.resample's algorithm (sinc interpolation, Hann-windowed, lowpass_filter_width 6,
rolloff 0.99), restated here: torchaudio is not a dependency of the product; tests/ compare against it.
"""
import math

import torch

from . import _native

SAMPLE_RATE = 32000
CLIP_SAMPLES = 10 * SAMPLE_RATE        # demo_convnext.py:47-48


def sinc_resample_taps(orig_freq, new_freq, lowpass_filter_width=6, rolloff=0.99, dtype=torch.float32):
    """(taps (K, new) transposed polyphase kernel, width, orig, new) with orig/new reduced by their gcd.

    Same arithmetic, in the same order and dtype, as torchaudio's `_get_sinc_resample_kernel` when called from
    `resample(waveform_fp32, ...)` (it builds the kernel in the waveform's dtype): kernel[p, k] =
    sinc(pi t) * cos^2(pi t / (2 lw)) * scale with t = clamp((k - width)/orig - p/new) * base_freq, +-lw)."""
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq or orig_freq <= 0 or new_freq <= 0:
        raise ValueError("resample: frequencies must be positive integers")
    g = math.gcd(int(orig_freq), int(new_freq))
    orig, new = int(orig_freq) // g, int(new_freq) // g
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=dtype)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=dtype)[:, None, None] / new + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t *= math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0).to(t), t.sin() / t)
    kernels *= window * scale
    taps = kernels.to(torch.float32)[:, 0, :].t().contiguous()            # (K, new): lanes read consecutive phases
    return taps, width, orig, new


_TAPS_CACHE = {}


def resample_fit(waveform, orig_freq, new_freq=SAMPLE_RATE, n_out=CLIP_SAMPLES):
    """waveform (B, L) or (L,) float32 CUDA tensor at orig_freq -> (B, n_out) at new_freq: resampled, then zero-padded
    or cropped to n_out samples (n_out=None keeps torchaudio's length ceil(new * L / orig)).  No CPU fallback."""
    if not waveform.is_cuda:
        raise RuntimeError("resample_fit needs a CUDA tensor: the B200 path has no CPU fallback")
    x = waveform.reshape(-1, waveform.shape[-1]).to(torch.float32).contiguous()
    B, L = x.shape
    key = (int(orig_freq), int(new_freq), x.device)
    if key not in _TAPS_CACHE:
        if int(orig_freq) == int(new_freq):
            # torchaudio's resample() returns the input unchanged: a 3-tap delta keeps the single launch (pad / crop)
            taps, width, orig, new = torch.tensor([[0.0], [1.0], [0.0]]), 1, 1, 1
        else:
            taps, width, orig, new = sinc_resample_taps(orig_freq, new_freq)
        _TAPS_CACHE[key] = (taps.to(x.device), width, orig, new)
    taps, width, orig, new = _TAPS_CACHE[key]
    target = -(-new * L // orig)
    n = target if n_out is None else int(n_out)
    out = torch.empty(B, n, device=x.device, dtype=torch.float32)
    _native.call("acx_resample_fit", x.data_ptr(), L, taps.data_ptr(), out.data_ptr(), n, B, L, orig, new, width, n,
                 torch.cuda.current_stream(x.device).cuda_stream)
    return out if waveform.dim() > 1 else out[0]


def tags_above(probs, threshold=0.25):
    """Indices of the classes whose probability exceeds `threshold`, per clip (demo_convnext.py:87-88)."""
    p = probs.detach()
    return [torch.nonzero(row > threshold).flatten().cpu().numpy() for row in p.reshape(-1, p.shape[-1])]
