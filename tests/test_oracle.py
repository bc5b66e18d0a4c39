"""CPU: pins the oracle (oracle/convnext_oracle.py) against the reference.

 * against the committed golden fixtures (outputs of the UNMODIFIED reference, oracle/make_golden.py);
 * against the reference itself when /root/reference is mounted (build container only).
The reference ships no tests / golden vectors of its own for this path.
"""
import os

import numpy as np
import pytest
import torch

from oracle import convnext_oracle as O
from oracle import weights
from oracle.ref_import import reference_available


def _demo_wave(g):
    return torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0)[None]


def test_oracle_matches_golden_demo_clip(golden_dir, parity_sd):
    g = np.load(os.path.join(golden_dir, "demo_clip.npz"))
    wave = _demo_wave(g)
    taps = {}
    out = O.forward(wave, parity_sd, taps=taps)
    assert np.abs(out["clipwise_logits"].numpy() - g["logits"]).max() < 2e-5
    assert np.abs(out["clipwise_output"].numpy() - g["probs"]).max() < 1e-5
    assert np.abs(taps["scene"].numpy() - g["scene"]).max() < 2e-5
    frame = O.forward_frame_embeddings(wave, parity_sd)
    assert frame.shape == (1, 768, 31, 7)                      # README.md:61
    assert np.abs(frame.numpy() - g["frame"]).max() < 5e-5
    lm = taps["logmel_bn"].numpy()[:, :: int(g["logmel_stride"])]
    assert np.abs(lm - g["logmel_bn"]).max() < 1e-4
    labels = np.where(out["clipwise_output"][0].numpy() > 0.25)[0]
    assert np.array_equal(labels, g["labels"])                 # thresholded label set, demo_convnext.py:87-88


@pytest.mark.parametrize("kind", ["noise", "tones"])
def test_oracle_matches_golden_synthetic(golden_dir, parity_sd, kind):
    g = np.load(os.path.join(golden_dir, f"synth_{kind}.npz"))
    wave = weights.make_waveforms(2, kind=kind, seed=0)
    cks = np.array([wave.double().sum().item(), wave.double().abs().sum().item()])
    assert np.allclose(cks, g["wave_cks"], rtol=1e-12)
    out = O.forward(wave, parity_sd)
    assert np.abs(out["clipwise_logits"].numpy() - g["logits"]).max() < 2e-5
    frame = O.forward_frame_embeddings(wave, parity_sd).numpy()
    assert np.abs(frame[:, :, :: int(g["frame_stride"])] - g["frame_t0"]).max() < 5e-5


def test_oracle_variable_length(golden_dir, parity_sd):
    g = np.load(os.path.join(golden_dir, "synth_short.npz"))
    wave = weights.make_waveforms(1, n_samples=int(g["n_samples"]), kind="noise", seed=3)
    frame = O.forward_frame_embeddings(wave, parity_sd).numpy()
    assert frame.shape == g["frame"].shape
    assert np.abs(frame - g["frame"]).max() < 5e-5
    assert np.abs(O.forward(wave, parity_sd)["clipwise_logits"].numpy() - g["logits"]).max() < 2e-5


def test_oracle_fp64_close_to_fp32(parity_sd):
    """The fp64 evaluation of the same formulas is the 'truth' used for error budgets."""
    wave = weights.make_waveforms(1, n_samples=32000, kind="tones", seed=5)
    a = O.forward(wave, parity_sd, torch.float32)["clipwise_logits"]
    b = O.forward(wave, parity_sd, torch.float64)["clipwise_logits"]
    assert (a.double() - b).abs().max() < 5e-4


@pytest.mark.skipif(not reference_available(), reason="/root/reference not mounted (GPU box)")
def test_oracle_matches_live_reference(parity_sd):
    from oracle.ref_import import build_reference_tiny
    model = build_reference_tiny()
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 28222767   # README.md:49
    ref_sd = model.state_dict()
    assert set(ref_sd) == set(parity_sd) and len(ref_sd) == 190
    for k in ("spectrogram_extractor.stft.conv_real.weight", "spectrogram_extractor.stft.conv_imag.weight",
              "logmel_extractor.melW"):
        assert torch.equal(ref_sd[k], parity_sd[k]), k       # constants == the reference's own
    model.load_state_dict(parity_sd, strict=True)
    model.eval()
    wave = weights.make_waveforms(1, n_samples=48000, kind="noise", seed=11)
    with torch.no_grad():
        ref = model(wave)
        ref_scene = model.forward_scene_embeddings(wave)
        ref_frame = model.forward_frame_embeddings(wave)
    out = O.forward(wave, parity_sd)
    assert (ref["clipwise_logits"] - out["clipwise_logits"]).abs().max() < 2e-5
    assert (ref_scene - O.forward_scene_embeddings(wave, parity_sd)).abs().max() < 2e-5
    assert (ref_frame - O.forward_frame_embeddings(wave, parity_sd)).abs().max() < 5e-5
