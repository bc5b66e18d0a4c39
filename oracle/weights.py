"""TEST INFRASTRUCTURE ONLY. Deterministic (numpy PCG64) weights and inputs.

The Zenodo/HF checkpoint is not available offline, so parity and throughput use seeded
random weights with the reference's own key names / shapes (module
structure at convnext.py:145-261, factory convnext.py:641-708).  numpy's Generator is used
instead of torch's RNG so that the very same numbers are regenerated on any box.

 * kind="init"   : the reference's random init -- Conv/Linear weights ~ N(0, 0.02^2)
                   (trunc_normal_(std=.02, a=-2, b=2), timm_weight_init.py:49-73 via
                   convnext.py:263-267 and :705-706), biases 0, LayerNorm (1, 0), bn0
                   identity, layer-scale gamma 1e-6 (convnext.py:56,67-71).
 * kind="parity" : same, then gamma ~ U(0.1, 0.6), LN / bn0 affine perturbed, non-trivial
                   bn0 running stats (mean -20, var 400 -- typical log-mel dB statistics),
                   biases ~ N(0, 0.02^2); with gamma = 1e-6 every Block is numerically the
                   identity and the MLP kernels would go untested.
 * kind="parity_stock_head" : "parity" with the classifier head left at the reference's own
                   init width (std 0.02 instead of the 3x wider, label-discriminative head of
                   "parity"; same random stream, so every other tensor is identical).  This
                   is the state dict the north_star tolerances (logits 2e-2 bf16 / 1e-4 fp32)
                   are asserted on UNSCALED.
"""
import numpy as np
import torch

from oracle.torchlibrosa_shim.torchlibrosa.stft import (dft_matrix, hann_periodic,
                                                        librosa_mel)

DEPTHS = [3, 3, 9, 3]
DIMS = [96, 192, 384, 768]
N_CLASSES = 527
N_MELS = 224
N_FFT = 1024
HOP = 320
SR = 32000
FMIN, FMAX = 50, 14000


def frontend_constants():
    """conv_real / conv_imag (513,1,1024) and melW (513,224), float32 -- SURVEY Appendix A."""
    win = hann_periodic(N_FFT)
    W = dft_matrix(N_FFT)[:, : N_FFT // 2 + 1] * win[:, None]
    conv_real = torch.from_numpy(np.real(W).T.astype(np.float32).copy())[:, None, :]
    conv_imag = torch.from_numpy(np.imag(W).T.astype(np.float32).copy())[:, None, :]
    melW = torch.from_numpy(librosa_mel(SR, N_FFT, N_MELS, FMIN, FMAX).T.copy())
    return conv_real, conv_imag, melW


def _key_shapes():
    """Ordered (key, shape) list of the 188 non-front-end... all learnable/buffer keys."""
    ks = []
    ks += [("bn0.weight", (N_MELS,)), ("bn0.bias", (N_MELS,)),
           ("bn0.running_mean", (N_MELS,)), ("bn0.running_var", (N_MELS,))]
    ks += [("downsample_layers.0.0.weight", (DIMS[0], 1, 4, 4)), ("downsample_layers.0.0.bias", (DIMS[0],)),
           ("downsample_layers.0.1.weight", (DIMS[0],)), ("downsample_layers.0.1.bias", (DIMS[0],))]
    for i in range(1, 4):
        ks += [(f"downsample_layers.{i}.0.weight", (DIMS[i - 1],)),
               (f"downsample_layers.{i}.0.bias", (DIMS[i - 1],)),
               (f"downsample_layers.{i}.1.weight", (DIMS[i], DIMS[i - 1], 2, 2)),
               (f"downsample_layers.{i}.1.bias", (DIMS[i],))]
    for s in range(4):
        C = DIMS[s]
        for j in range(DEPTHS[s]):
            p = f"stages.{s}.{j}."
            ks += [(p + "gamma", (C,)),
                   (p + "dwconv.weight", (C, 1, 7, 7)), (p + "dwconv.bias", (C,)),
                   (p + "norm.weight", (C,)), (p + "norm.bias", (C,)),
                   (p + "pwconv1.weight", (4 * C, C)), (p + "pwconv1.bias", (4 * C,)),
                   (p + "pwconv2.weight", (C, 4 * C)), (p + "pwconv2.bias", (C,))]
    ks += [("norm.weight", (DIMS[3],)), ("norm.bias", (DIMS[3],)),
           ("head_audioset.weight", (N_CLASSES, DIMS[3])), ("head_audioset.bias", (N_CLASSES,))]
    return ks


def make_state_dict(kind="parity", seed=1):
    """Full 190-key state dict (float32, + int64 num_batches_tracked)."""
    assert kind in ("init", "parity", "parity_stock_head")
    stock_head = kind == "parity_stock_head"
    if stock_head:
        kind = "parity"
    rng = np.random.default_rng(seed)
    conv_real, conv_imag, melW = frontend_constants()
    sd = {
        "spectrogram_extractor.stft.conv_real.weight": conv_real,
        "spectrogram_extractor.stft.conv_imag.weight": conv_imag,
        "logmel_extractor.melW": melW,
        "bn0.num_batches_tracked": torch.zeros((), dtype=torch.int64),
    }

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32))

    def uniform(shape, lo, hi):
        return torch.from_numpy(rng.uniform(lo, hi, size=shape).astype(np.float32))

    for key, shape in _key_shapes():
        leaf = key.rsplit(".", 1)[-1]
        is_matrix = len(shape) >= 2
        if is_matrix:
            # parity head: wider weights + negative bias so the 0.25-thresholded label set is
            # sparse and discriminative like a trained tagger's (demo_convnext.py:87-92)
            t = normal(shape, 0.06 if (kind == "parity" and not stock_head and key == "head_audioset.weight") else 0.02)
        elif key == "bn0.running_mean":
            t = torch.zeros(shape) if kind == "init" else normal(shape, 3.0) - 20.0
        elif key == "bn0.running_var":
            t = torch.ones(shape) if kind == "init" else uniform(shape, 300.0, 500.0)
        elif leaf == "gamma":
            t = torch.full(shape, 1e-6) if kind == "init" else uniform(shape, 0.1, 0.6)
        elif leaf == "weight":      # LayerNorm / bn0 scale
            t = torch.ones(shape) if kind == "init" else uniform(shape, 0.8, 1.2)
        elif leaf == "bias":
            is_norm = (key.startswith("bn0") or ".norm." in key or key.startswith("norm.")
                       or key in ("downsample_layers.0.1.bias",)
                       or (key.startswith("downsample_layers.") and key.split(".")[1] != "0"
                           and key.split(".")[2] == "0"))
            if kind == "init":
                t = torch.zeros(shape)
            elif key == "head_audioset.bias":
                t = normal(shape, 0.5) - 2.0
            else:
                t = normal(shape, 0.05 if is_norm else 0.02)
        else:
            raise AssertionError(key)
        sd[key] = t
    assert len(sd) == 190, len(sd)
    return sd


def make_waveforms(batch, n_samples=320000, kind="noise", seed=0):
    """Synthetic clips: 'noise' = 0.1*N(0,1) clamped to [-1,1];
    'tones' = 5 sinusoids below 4 kHz + 1e-4 noise (band-limited: the hard case for the
    log-mel numerics)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        w = np.clip(rng.standard_normal((batch, n_samples)) * 0.1, -1.0, 1.0)
    elif kind == "tones":
        t = np.arange(n_samples, dtype=np.float64) / SR
        w = np.zeros((batch, n_samples))
        for b in range(batch):
            for _ in range(5):
                f = rng.uniform(80.0, 4000.0)
                a = rng.uniform(0.02, 0.2)
                ph = rng.uniform(0, 2 * np.pi)
                w[b] += a * np.sin(2 * np.pi * f * t + ph)
            w[b] += 1e-4 * rng.standard_normal(n_samples)
    else:
        raise ValueError(kind)
    return torch.from_numpy(w.astype(np.float32))
