"""TEST INFRASTRUCTURE ONLY -- the parity oracle.  Never imported by the product package.

Functional CPU restatement (torch CPU ops, fp32 or fp64) of the reference hot path.
Every function cites the reference lines it follows; citations are relative to
/root/reference/src/audioset_convnext_inf/pytorch/convnext.py ("CX") or to the
torchlibrosa 0.0.9 algorithm restated in SURVEY.md Appendix A ("TL").

Pinned by tests/test_oracle.py against (a) the unmodified reference imported in the build
container (when /root/reference is mounted) and (b) the committed fixtures in
tests/golden/ that oracle/make_golden.py produced from that reference.
"""
import torch
import torch.nn.functional as F

import contextlib

DEPTHS = [3, 3, 9, 3]
DIMS = [96, 192, 384, 768]
_DEVICE = "cpu"          # the oracle is a CPU checker; bench.py's `gpu_eager_baseline` MEASUREMENT leg alone moves it


@contextlib.contextmanager
def on_device(device):
    """Run the same stock-PyTorch ops on another device (bench.py: eager cuDNN/cuBLAS comparator on the same B200)."""
    global _DEVICE
    old, _DEVICE = _DEVICE, device
    try:
        yield
    finally:
        _DEVICE = old


def _c(sd, key, dtype):
    return sd[key].detach().to(_DEVICE, dtype)


def spectrogram(wave, sd, dtype=torch.float32, n_fft=1024, hop=320):
    """TL STFT.forward + Spectrogram.forward (called at CX:298; ctor CX:179-187):
    reflect-pad n_fft//2, two strided conv1d (windowed DFT real / imag), re^2 + im^2.
    wave (B, L) -> power (B, T, 513), T = L // hop + 1."""
    x = wave.to(_DEVICE, dtype)[:, None, :]
    x = F.pad(x, (n_fft // 2, n_fft // 2), mode="reflect")
    real = F.conv1d(x, _c(sd, "spectrogram_extractor.stft.conv_real.weight", dtype), stride=hop)
    imag = F.conv1d(x, _c(sd, "spectrogram_extractor.stft.conv_imag.weight", dtype), stride=hop)
    return (real ** 2 + imag ** 2).transpose(1, 2)


def logmel(power, sd, dtype=torch.float32, amin=1e-10, ref=1.0):
    """TL LogmelFilterBank.forward / power_to_db (called at CX:299; ctor CX:190-200):
    matmul with melW (513, 224), 10*log10(clamp(., amin)) - 10*log10(max(amin, ref));
    top_db=None -> no dynamic-range clamp (CX:166)."""
    mel = torch.matmul(power, _c(sd, "logmel_extractor.melW", dtype))
    out = 10.0 * torch.log10(torch.clamp(mel, min=amin))
    out = out - 10.0 * torch.log10(torch.tensor(max(amin, ref), dtype=dtype))
    return out


def bn0(x, sd, dtype=torch.float32, eps=1e-5):
    """CX:304-306 -- eval-mode BatchNorm2d(224) over the mel axis (per-mel-bin affine)."""
    mean = _c(sd, "bn0.running_mean", dtype)
    var = _c(sd, "bn0.running_var", dtype)
    w = _c(sd, "bn0.weight", dtype)
    b = _c(sd, "bn0.bias", dtype)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def frontend(wave, sd, dtype=torch.float32):
    """wave (B, L) -> normalised log-mel (B, T, 224)   [CX:298-306]."""
    return bn0(logmel(spectrogram(wave, sd, dtype), sd, dtype), sd, dtype)


def layernorm_cf(x, w, b, eps=1e-6):
    """CX:536-541 -- channels_first LayerNorm on NCHW (biased variance over C)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return w[:, None, None] * x + b[:, None, None]


def block_dwconv_ln(x, sd, prefix, dtype):
    """CX:76-78 -- depthwise 7x7 (pad 3, bias) -> NHWC -> LayerNorm(C, eps 1e-6).  NCHW in, NHWC out."""
    C = x.shape[1]
    x = F.conv2d(x, _c(sd, prefix + "dwconv.weight", dtype), _c(sd, prefix + "dwconv.bias", dtype),
                 padding=3, groups=C)
    x = x.permute(0, 2, 3, 1)
    return F.layer_norm(x, (C,), _c(sd, prefix + "norm.weight", dtype), _c(sd, prefix + "norm.bias", dtype), 1e-6)


def block_mlp(y, sd, prefix, dtype):
    """CX:79-83 -- Linear(C,4C) -> exact-erf GELU -> Linear(4C,C) -> gamma * x.  NHWC in / out."""
    y = F.linear(y, _c(sd, prefix + "pwconv1.weight", dtype), _c(sd, prefix + "pwconv1.bias", dtype))
    y = F.gelu(y)
    y = F.linear(y, _c(sd, prefix + "pwconv2.weight", dtype), _c(sd, prefix + "pwconv2.bias", dtype))
    return _c(sd, prefix + "gamma", dtype) * y


def block(x, sd, prefix, dtype):
    """CX:74-87 -- dwconv7x7 -> NHWC -> LayerNorm -> Linear -> GELU(erf) -> Linear ->
    gamma * x -> NCHW -> residual (DropPath is Identity at rate 0, CX:72)."""
    y = block_mlp(block_dwconv_ln(x, sd, prefix, dtype), sd, prefix, dtype)
    return x + y.permute(0, 3, 1, 2)


def pool_head(x, sd, dtype):
    """CX:279-285 + CX:321-325 on a (B, 768, H, 7) NCHW tensor -> (scene, logits, probs)."""
    x = torch.mean(x, dim=3)
    (x1, _) = torch.max(x, dim=2)
    x2 = torch.mean(x, dim=2)
    emb = F.layer_norm(x1 + x2, (DIMS[3],), _c(sd, "norm.weight", dtype), _c(sd, "norm.bias", dtype), 1e-6)
    logits = F.linear(emb, _c(sd, "head_audioset.weight", dtype), _c(sd, "head_audioset.bias", dtype))
    return emb, logits, torch.sigmoid(logits)


def stem(x, sd, dtype):
    """CX:688-691 (Conv2d(1,96,4x4,s4,pad=(4,0))) + CX:227 channels_first LayerNorm."""
    x = F.conv2d(x, _c(sd, "downsample_layers.0.0.weight", dtype),
                 _c(sd, "downsample_layers.0.0.bias", dtype), stride=(4, 4), padding=(4, 0))
    return layernorm_cf(x, _c(sd, "downsample_layers.0.1.weight", dtype),
                        _c(sd, "downsample_layers.0.1.bias", dtype))


def downsample(x, sd, i, dtype):
    """CX:230-235 -- channels_first LayerNorm then Conv2d(C, 2C, k=2, s=2)."""
    x = layernorm_cf(x, _c(sd, f"downsample_layers.{i}.0.weight", dtype),
                     _c(sd, f"downsample_layers.{i}.0.bias", dtype))
    return F.conv2d(x, _c(sd, f"downsample_layers.{i}.1.weight", dtype),
                    _c(sd, f"downsample_layers.{i}.1.bias", dtype), stride=2)


def forward_features(x, sd, dtype=torch.float32, return_frame_embeddings=False, taps=None):
    """CX:269-285.  x: normalised log-mel as (B, 1, T, 224) NCHW."""
    for i in range(4):
        x = stem(x, sd, dtype) if i == 0 else downsample(x, sd, i, dtype)
        if taps is not None:
            taps[f"ds{i}"] = x
        for j in range(DEPTHS[i]):
            x = block(x, sd, f"stages.{i}.{j}.", dtype)
        if taps is not None:
            taps[f"stage{i}"] = x
    if return_frame_embeddings:
        return x                                           # CX:276-277
    x = torch.mean(x, dim=3)                               # CX:279
    (x1, _) = torch.max(x, dim=2)                          # CX:280
    x2 = torch.mean(x, dim=2)                              # CX:281
    x = x1 + x2                                            # CX:282
    return F.layer_norm(x, (DIMS[3],), _c(sd, "norm.weight", dtype), _c(sd, "norm.bias", dtype), 1e-6)  # CX:285


def forward(wave, sd, dtype=torch.float32, taps=None):
    """CX:287-331 (eval mode): returns dict(clipwise_output, clipwise_logits)."""
    lm = frontend(wave, sd, dtype)
    if taps is not None:
        taps["logmel_bn"] = lm
    emb = forward_features(lm[:, None], sd, dtype, taps=taps)
    if taps is not None:
        taps["scene"] = emb
    logits = F.linear(emb, _c(sd, "head_audioset.weight", dtype), _c(sd, "head_audioset.bias", dtype))  # CX:321
    return {"clipwise_output": torch.sigmoid(logits), "clipwise_logits": logits}  # CX:325-329


def forward_scene_embeddings(wave, sd, dtype=torch.float32):
    """CX:333-366 -> (B, 768) post-LayerNorm pooled embedding."""
    return forward_features(frontend(wave, sd, dtype)[:, None], sd, dtype)


def forward_frame_embeddings(wave, sd, dtype=torch.float32):
    """CX:369-402 -> (B, 768, T', 7) NCHW, raw stage-3 output."""
    return forward_features(frontend(wave, sd, dtype)[:, None], sd, dtype, return_frame_embeddings=True)
