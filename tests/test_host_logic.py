"""CPU: host-side logic of the product package and the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import audioset_convnext_inf_b200 as acx
from audioset_convnext_inf_b200 import _native
from audioset_convnext_inf_b200.engine import PackedWeights, out_time_dims
from oracle import convnext_oracle as O
from oracle import weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny():
    return acx.convnext_tiny(pretrained=False, strict=False, drop_path_rate=0.0, after_stem_dim=[252, 56])


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "acx.h")).read()
    declared = set(re.findall(r"ACX_API\s+[\w\s\*]+?\b(acx_\w+)\s*\(", header))
    assert len(declared) >= 14
    lib = ctypes.CDLL(_native.load()._name)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"libacx.so does not export: {missing}"
    assert set(_native.SIGNATURES) == declared, set(_native.SIGNATURES) ^ declared
    assert lib.acx_version() == 100


def test_module_surface_matches_reference_contract():
    m = _tiny()
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 28222767      # README.md:49
    sd = m.state_dict()
    ref = weights.make_state_dict("init", 0)
    assert len(sd) == 190 and set(sd) == set(ref)
    for k in ref:
        assert sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype, k
    # frozen front-end tensors are bit-identical to torchlibrosa's constants
    for k in ("spectrogram_extractor.stft.conv_real.weight", "spectrogram_extractor.stft.conv_imag.weight",
              "logmel_extractor.melW"):
        assert torch.equal(sd[k], ref[k]) and not dict(m.named_parameters())[k].requires_grad
    for name in ("forward", "forward_scene_embeddings", "forward_frame_embeddings", "from_pretrained"):
        assert callable(getattr(acx.ConvNeXt, name))


def test_no_cpu_fallback_and_training_mode_raise():
    m = _tiny()
    with pytest.raises(RuntimeError, match="eval"):
        m(torch.zeros(1, 32000))
    m.eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 32000))


def test_unsupported_configs_raise_explicitly():
    with pytest.raises(ValueError):
        acx.convnext_tiny(after_stem_dim=[7, 7], drop_path_rate=0.0)
    with pytest.raises(NotImplementedError):
        acx.convnext_tiny(after_stem_dim=[56], drop_path_rate=0.0)
    with pytest.raises(NotImplementedError):
        acx.convnext_tiny(after_stem_dim=[252, 56])          # default drop_path_rate=0.1 is training-only
    with pytest.raises(NotImplementedError):
        acx.LayerNorm(8, data_format="nope")


def test_checkpoint_round_trips(tmp_path):
    from safetensors.torch import save_model
    sd = weights.make_state_dict("parity", 3)
    m = _tiny()
    m.load_state_dict(sd, strict=True)
    st_path = str(tmp_path / "model.safetensors")
    save_model(m, st_path)                                   # convert_pytorch_ckpt_to_safetensors.py:18
    m2 = acx.ConvNeXt.from_pretrained(st_path)
    assert m2.training                                       # like the reference: caller must .eval()
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k]), k
    pth = str(tmp_path / "ckpt.pth")
    torch.save({"model": sd}, pth)                           # evaluate_convnext_on_audioset.py:36-38
    m3 = acx.ConvNeXt.from_pretrained(pth)
    for k, v in m3.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_packed_weights_layouts():
    sd = weights.make_state_dict("parity", 8)
    pw = PackedWeights(sd, torch.device("cpu"), "bf16")
    # mel band table reproduces the dense filterbank
    melW = sd["logmel_extractor.melW"]
    dense = torch.zeros_like(melW)
    for m in range(224):
        lo, hi = int(pw.mel_lo[m]), int(pw.mel_hi[m])
        dense[lo:hi, m] = pw.melT[m, lo:hi]
    assert torch.equal(dense, melW)
    assert pw.n_chunks == 7                                  # bins 2..447 carry all 884 non-zeros
    # split-fp16 DFT rows (scaled by 2^8): hi + lo reproduces fp32 to ~2^-22 relative
    re = sd["spectrogram_extractor.stft.conv_real.weight"][:448, 0]
    assert pw.dft_hi.dtype == torch.float16
    chunks = ((pw.dft_hi.float() + pw.dft_lo.float()) / 256.0).view(7, 2, 64, 1024)
    assert (chunks[:, 0].reshape(448, 1024) - re).abs().max() < 5e-7
    # mel chunks: K-major (mel, bin) tiles
    mel = (pw.melc_hi.float() + pw.melc_lo.float()).view(7, 256, 64)
    assert (mel[:, :224].permute(0, 2, 1).reshape(448, 224) - melW[:448]).abs().max() < 1e-6
    assert mel[:, 224:].abs().max() == 0
    # downsample conv as GEMM over (dy, dx, cin) patches
    w = sd["downsample_layers.1.1.weight"]
    x = torch.randn(1, 96, 4, 6)
    ref = torch.nn.functional.conv2d(x, w, stride=2)
    patches = x.permute(0, 2, 3, 1).reshape(1, 2, 2, 3, 2, 96).permute(0, 1, 3, 2, 4, 5).reshape(6, 384)
    got = (patches @ pw.ds[0]["w"].float().t()).reshape(1, 2, 3, 192).permute(0, 3, 1, 2)
    assert (got - ref).abs().max() < 0.05                    # bf16-rounded weights
    # dwconv taps transposed to (49, C)
    assert torch.equal(pw.blocks[0][0]["dw_w"].float().t().reshape(96, 1, 7, 7),
                       sd["stages.0.0.dwconv.weight"].to(torch.bfloat16).float())


@pytest.mark.parametrize("L", [320000, 96123, 960000, 32000])
def test_out_time_dims_match_oracle_shapes(L):
    sd = weights.make_state_dict("init", 0)
    T, hs = out_time_dims(L)
    lm = torch.zeros(1, 1, T, 224)
    x = O.stem(lm, sd, torch.float32)
    assert x.shape[2:] == (hs[0], 56)
    assert T == L // 320 + 1
    assert hs == [hs[0], hs[0] // 2, hs[0] // 4, hs[0] // 8] or hs[3] == ((hs[0] // 2) // 2) // 2


def test_gelu_fit_against_exact_erf():
    """The tensor-core epilogues' tanh-form GELU (csrc/common.cuh: gelu_stage_ta4 coefficients) vs nn.GELU()."""
    x = torch.linspace(-60, 60, 1200001)
    x2 = torch.clamp(x * x, max=50.0)
    inner = x * (7.97507884e-01 + x2 * (3.70056460e-02 + x2 * -3.51516788e-04))
    fast = 0.5 * x * (1 + torch.tanh(inner))
    exact = torch.nn.functional.gelu(x.double())
    assert (fast.double() - exact).abs().max() < 4e-5


# ---- SURVEY §8 "next" rows: demo preprocessing (f4), length bucketing (f3), checkpoint converter (f2) -------------
@pytest.mark.parametrize("orig,new", [(44100, 32000), (48000, 32000), (16000, 32000), (22050, 32000), (8000, 32000)])
def test_resample_taps_equal_torchaudio_kernel(orig, new):
    """preprocess.sinc_resample_taps restates torchaudio's polyphase kernel (what demo_convnext.py:53-59 applies)."""
    import math
    F = pytest.importorskip("torchaudio.functional.functional")
    k, w = F._get_sinc_resample_kernel(orig, new, math.gcd(orig, new), dtype=torch.float32)
    taps, width, o, n = acx.preprocess.sinc_resample_taps(orig, new)
    assert (width, o, n) == (w, orig // math.gcd(orig, new), new // math.gcd(orig, new))
    assert taps.shape == (2 * width + o, n)
    assert torch.equal(taps.t().contiguous(), k[:, 0, :])


def test_resample_fit_refuses_cpu_tensors():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        acx.preprocess.resample_fit(torch.zeros(1, 1000), 44100)


def test_length_buckets_group_exact_lengths_in_order():
    from audioset_convnext_inf_b200.extract import length_buckets
    b = length_buckets([5, 7, 5, 5, 7, 9], max_batch=2)
    assert b == [(5, [0, 2]), (5, [3]), (7, [1, 4]), (9, [5])]
    assert length_buckets([], 4) == []


def test_checkpoint_converter_round_trip(tmp_path):
    """tools/convert_checkpoint.py (reference convert_pytorch_ckpt_to_safetensors.py): .pth {"model": sd} -> strict
    safetensors with the 190 reference keys, loadable again by from_pretrained."""
    import importlib.util
    from safetensors.torch import load_file
    spec = importlib.util.spec_from_file_location(
        "convert_checkpoint", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "convert_checkpoint.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sd = weights.make_state_dict("parity", 3)
    src, dst = tmp_path / "ckpt.pth", tmp_path / "model.safetensors"
    torch.save({"model": sd}, src)
    mod.convert(str(src), str(dst))
    back = load_file(str(dst))
    assert set(back) == set(sd) and len(back) == 190
    assert all(torch.equal(back[k], sd[k]) for k in sd)
    m = acx.ConvNeXt.from_pretrained(str(dst))
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())


def test_reference_import_path_alias():
    """The reference's scripts import `audioset_convnext_inf.pytorch.convnext` (demo_convnext.py:13,
    evaluate_convnext_on_audioset.py:14, pytorch/extract_embeddings.py:12) and `pytorch_utils` (PU:9-15, 63-137): both
    resolve to the B200 package, so those scripts run unmodified against libacx."""
    from audioset_convnext_inf.pytorch.convnext import ConvNeXt, convnext_tiny
    from audioset_convnext_inf.pytorch import pytorch_utils
    assert ConvNeXt is acx.ConvNeXt and convnext_tiny is acx.convnext_tiny
    assert pytorch_utils.forward is acx.evalloop.forward
    import numpy as np
    assert pytorch_utils.move_data_to_device(np.zeros(3, np.float32), "cpu").dtype == torch.float32
    assert pytorch_utils.move_data_to_device(np.zeros(3, np.int64), "cpu").dtype == torch.int64
    obj = np.empty(2, dtype=object)
    assert pytorch_utils.move_data_to_device(obj, "cpu") is obj          # PU:13-14: returned unchanged


def test_checkpoint_format_is_not_sniffed_from_leading_bytes(tmp_path):
    """A safetensors header length is an arbitrary u64: low byte 0x80 (pickle opcode) or 'PK' (zip magic) must not send
    the file to torch.load.  Metadata padding is used to force both header lengths."""
    import json
    import struct
    from safetensors.torch import save_file
    sd = {k: v.contiguous() for k, v in weights.make_state_dict("parity", 3).items()}
    for want_low, name in ((b"\x80", "a.safetensors"), (b"PK", "b.safetensors"), (b"PK", "noext")):
        p = str(tmp_path / name)
        save_file(sd, p)
        raw = open(p, "rb").read()
        (n,) = struct.unpack("<Q", raw[:8])
        header, body = json.loads(raw[8:8 + n]), raw[8 + n:]
        target = n + 64
        while struct.pack("<Q", target)[:len(want_low)] != want_low:
            target += 1
        for pad in range(0, 1 << 17):               # pad the metadata until the serialised header has that length
            header["__metadata__"] = {"pad": "x" * pad}
            hb = json.dumps(header, separators=(",", ":")).encode()
            if len(hb) == target:
                break
            if len(hb) > target:
                target += 1 << (8 * len(want_low))
        assert len(hb) == target
        with open(p, "wb") as fh:
            fh.write(struct.pack("<Q", len(hb)) + hb + body)
        assert open(p, "rb").read(len(want_low)) == want_low
        m = acx.ConvNeXt.from_pretrained(p)
        assert torch.equal(m.state_dict()["stages.2.4.dwconv.weight"], sd["stages.2.4.dwconv.weight"])


def test_layernorm_fold_into_pwconv1_is_an_identity():
    """engine.fold_layernorm_into_pwconv1: pwconv1(LN(v)) == rstd * (v W1'^T - mean s) + b1' exactly (fp64), with the
    bf16-rounded W1' used consistently for the GEMM operand and for s (CX:78-79)."""
    from audioset_convnext_inf_b200.engine import fold_layernorm_into_pwconv1
    g = torch.Generator().manual_seed(5)
    C = 96
    v = (torch.randn(50, C, generator=g) * 2 + torch.randn(50, 1, generator=g) * 3).to(torch.bfloat16).double()
    w1, b1 = torch.randn(4 * C, C, generator=g) / C ** 0.5, torch.randn(4 * C, generator=g) * 0.1
    lw, lb = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    f = fold_layernorm_into_pwconv1(w1, b1, lw, lb)
    assert f["w1f"].dtype == torch.bfloat16 and f["s1"].shape == (4 * C,) and f["b1f"].shape == (4 * C,)
    mean = v.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(v.var(1, unbiased=False, keepdim=True) + 1e-6)
    folded = rstd * (v @ f["w1f"].double().t() - mean * f["s1"].double()[None]) + f["b1f"].double()[None]
    # reference with the SAME rounded operand W1' = bf16(W1 ln_w): LN without affine, then W1', then the folded bias
    direct = ((v - mean) * rstd) @ f["w1f"].double().t() + f["b1f"].double()[None]
    assert (folded - direct).abs().max().item() < 1e-5
    # and against the unfolded layer: differs only by the bf16 rounding of W1 ln_w
    ln = torch.nn.functional.layer_norm(v, (C,), lw.double(), lb.double(), 1e-6)
    exact = ln @ w1.double().t() + b1.double()[None]
    assert (folded - exact).abs().max().item() < 0.05
