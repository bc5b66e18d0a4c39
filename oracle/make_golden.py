"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
The reference ships no golden vectors for this path, so these fixtures are
outputs of the reference's own PyTorch code (`ConvNeXt.forward`, `forward_scene_embeddings`,
`forward_frame_embeddings`, convnext.py:287-402) on
  * the bundled demo clip audio_samples/f62-S-v2swA_200000_210000.wav (copied as int16
    samples -- it is the reference's own test input, demo_convnext.py:28), and
  * seeded synthetic clips (oracle/weights.py::make_waveforms),
with the seeded weights of oracle/weights.py::make_state_dict (numpy PCG64, reproducible
anywhere).  The fixtures also record weight checksums so a test can prove it regenerated
the same weights.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import weights  # noqa: E402
from oracle.ref_import import REFERENCE_ROOT, build_reference_tiny  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
PARITY_SEED = 8
THRESH = 0.25


def read_wav_int16(path):
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    assert sr == 32000 and data.dtype == np.int16 and data.ndim == 1, (sr, data.dtype, data.shape)
    return data


def checksums(sd):
    keys = ["stages.0.0.pwconv1.weight", "stages.2.4.dwconv.weight", "stages.3.2.gamma",
            "head_audioset.weight", "bn0.running_mean", "logmel_extractor.melW",
            "spectrogram_extractor.stft.conv_real.weight"]
    return {k: np.array([sd[k].double().sum().item(), sd[k].double().abs().sum().item()]) for k in keys}


def run_reference(model, wave):
    with torch.no_grad():
        out = model(wave)
        scene = model.forward_scene_embeddings(wave)
        frame = model.forward_frame_embeddings(wave)
        # normalised log-mel exactly as the reference computes it (convnext.py:298-306)
        x = model.logmel_extractor(model.spectrogram_extractor(wave))
        x = model.bn0(x.transpose(1, 3)).transpose(1, 3)
    return out["clipwise_logits"], out["clipwise_output"], scene, frame, x[:, 0]


def stock_head_fixture(model):
    """parity_stock_head.npz: gamma ~ U(0.1, 0.6) trunk + the reference's own head width (std 0.02) -- the state
    dict the north_star tolerances are asserted on as written (no scaling): demo clip, noise and tones."""
    sd = weights.make_state_dict("parity_stock_head", PARITY_SEED)
    model.load_state_dict(sd, strict=True)
    model.eval()
    pcm = read_wav_int16(os.path.join(REFERENCE_ROOT, "audio_samples", "f62-S-v2swA_200000_210000.wav"))
    waves = {"demo": torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None],
             "noise": weights.make_waveforms(2, kind="noise", seed=0),
             "tones": weights.make_waveforms(2, kind="tones", seed=0)}
    out = {"parity_seed": np.array(PARITY_SEED),
           "cks/head_audioset.weight": checksums(sd)["head_audioset.weight"],
           "cks/stages.3.2.gamma": checksums(sd)["stages.3.2.gamma"]}
    for name, wave in waves.items():
        logits, probs, scene, frame, lm = run_reference(model, wave)
        out[f"{name}/logits"] = logits.numpy()
        out[f"{name}/probs"] = probs.numpy()
        out[f"{name}/scene"] = scene.numpy()
        print("stock head", name, "logits std", float(logits.std()), "range", float(logits.min()), float(logits.max()))
    np.savez_compressed(os.path.join(GOLDEN, "parity_stock_head.npz"), **out)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if "--stock-head-only" in sys.argv:      # adds the round-2 fixture without rewriting the round-1 ones
        stock_head_fixture(build_reference_tiny())
        return
    sd = weights.make_state_dict("parity", PARITY_SEED)
    model = build_reference_tiny()
    model.load_state_dict(sd, strict=True)
    model.eval()
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert n_params == 28222767, n_params  # README.md:49

    # ---- demo clip ------------------------------------------------------------------
    pcm = read_wav_int16(os.path.join(REFERENCE_ROOT, "audio_samples", "f62-S-v2swA_200000_210000.wav"))
    assert pcm.shape[0] == 320000
    wave = torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None]   # torchaudio.load normalisation
    logits, probs, scene, frame, lm = run_reference(model, wave)
    labels = np.where(probs[0].numpy() > THRESH)[0]
    margin = float((logits[0] - np.log(THRESH / (1 - THRESH))).abs().min())
    print("demo: labels", labels.shape, "min |logit - thr| =", margin)
    np.savez_compressed(
        os.path.join(GOLDEN, "demo_clip.npz"),
        pcm=pcm, logits=logits.numpy(), probs=probs.numpy(), scene=scene.numpy(),
        frame=frame.numpy(), logmel_bn=lm.numpy()[:, ::7].copy(), logmel_stride=np.array(7),
        labels=labels, label_margin=np.array(margin), n_params=np.array(n_params),
        parity_seed=np.array(PARITY_SEED),
        **{"cks/" + k: v for k, v in checksums(sd).items()})

    # ---- synthetic clips ------------------------------------------------------------
    for kind in ("noise", "tones"):
        wave = weights.make_waveforms(2, kind=kind, seed=0)
        logits, probs, scene, frame, lm = run_reference(model, wave)
        np.savez_compressed(
            os.path.join(GOLDEN, f"synth_{kind}.npz"),
            logits=logits.numpy(), scene=scene.numpy(),
            frame_t0=frame.numpy()[:, :, ::6, :].copy(), frame_stride=np.array(6),
            frame_mean=np.array(frame.double().mean().item()), frame_std=np.array(frame.double().std().item()),
            logmel_bn=lm.numpy()[:, ::25].copy(), logmel_stride=np.array(25),
            wave_cks=np.array([wave.double().sum().item(), wave.double().abs().sum().item()]))

    # ---- variable length (extract_embeddings.py:72-83 feeds arbitrary L) --------------
    wave = weights.make_waveforms(1, n_samples=3 * 32000 + 123, kind="noise", seed=3)
    logits, probs, scene, frame, lm = run_reference(model, wave)
    np.savez_compressed(os.path.join(GOLDEN, "synth_short.npz"), logits=logits.numpy(),
                        scene=scene.numpy(), frame=frame.numpy(), n_samples=np.array(wave.shape[1]))

    # ---- stock init (gamma=1e-6): what bench.py runs ------------------------------------
    sd0 = weights.make_state_dict("init", 0)
    model.load_state_dict(sd0, strict=True)
    wave = weights.make_waveforms(1, kind="noise", seed=0)
    logits, probs, scene, frame, lm = run_reference(model, wave)
    np.savez_compressed(os.path.join(GOLDEN, "synth_init.npz"), logits=logits.numpy(), scene=scene.numpy())
    stock_head_fixture(model)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
