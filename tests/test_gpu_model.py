"""GPU (B200): end-to-end parity of the drop-in ConvNeXt (libacx kernels) against the committed golden
outputs of the unmodified reference and against the CPU oracle, in both arithmetic modes.

Tolerances (BASELINE.json north_star): logits 2e-2 abs (bf16 mode) / 1e-4 abs (fp32-accurate mode);
thresholded (0.25) label set identical on the bundled demo clip (up to logits that the reference itself
places within the tolerance of the threshold)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import audioset_convnext_inf_b200 as acx                      # noqa: E402
from oracle import convnext_oracle as O                       # noqa: E402
from oracle import weights                                    # noqa: E402

DEV = "cuda:0"
TOL = {"fp32": dict(logits=1e-4 * 5, scene=5e-4, frame=2e-3), "bf16": dict(logits=2e-2 * 3, scene=0.12, frame=0.6)}
# NOTE: the "parity" weights use a 3x wider head (std 0.06) than the reference init so the label set is
# discriminative; logit error scales with the head width, hence the 3x / 5x factors above.  The north_star tolerances
# are applied AS WRITTEN (2e-2 bf16 / 1e-4 fp32, no factor) by test_stock_head_north_star_tolerances_unscaled below, on
# the "parity_stock_head" state dict: the same gamma ~ U(0.1, 0.6) trunk with the reference's own head width (0.02).
NORTH_STAR = {"fp32": 1e-4, "bf16": 2e-2}
# log-mel vs the reference's golden, in dB (SURVEY Appendix C metrics).  north_star asks 1e-3 dB (bf16 mode) / 1e-5 (fp32
# mode) max-abs; the reference's OWN fp32 evaluation is 2.6e-3 dB (max) away from an fp64 evaluation of the same formula
# on this clip (bins 60 dB below the peak carry ~1e-4 relative error in fp32), so max-abs is asserted at the level of
# that intrinsic noise and the north_star figure on the p99 / mean, where it is meaningful.
LOGMEL_DB = {"bf16": dict(max=1e-2, p99=1e-3, mean=2e-4, inband_max=1e-2),
             "fp32": dict(max=5e-3, p99=2e-4, mean=2e-5, inband_max=5e-3)}


@pytest.fixture(scope="module")
def models(parity_sd):
    out = {}
    for prec in ("fp32", "bf16"):
        m = acx.convnext_tiny(pretrained=False, strict=False, drop_path_rate=0.0, after_stem_dim=[252, 56])
        m.load_state_dict(parity_sd, strict=True)
        out[prec] = m.to(DEV).eval().set_precision(prec)
    return out


def _report(tag, got, ref):
    d = (got - ref).abs()
    print(f"  {tag}: max {d.max():.3e} mean {d.mean():.3e} (ref std {ref.std():.3f})")
    return d.max().item()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_demo_clip_against_reference_golden(models, golden_dir, prec):
    g = np.load(os.path.join(golden_dir, "demo_clip.npz"))
    wave = torch.from_numpy(g["pcm"].astype(np.float32) / 32768.0)[None].to(DEV)
    m = models[prec]
    out = m(wave)
    assert set(out) == {"clipwise_output", "clipwise_logits"}
    logits, probs = out["clipwise_logits"].cpu(), out["clipwise_output"].cpu()
    assert logits.shape == (1, 527) and logits.dtype == torch.float32       # README.md:53-55
    scene = m.forward_scene_embeddings(wave).cpu()
    frame = m.forward_frame_embeddings(wave).cpu()
    assert scene.shape == (1, 768) and frame.shape == (1, 768, 31, 7)       # README.md:59-61
    print(f"[{prec}] demo clip vs reference golden")
    t = TOL[prec]
    assert _report("logits", logits, torch.from_numpy(g["logits"])) < t["logits"]
    assert _report("probs", probs, torch.from_numpy(g["probs"])) < t["logits"]
    assert _report("scene", scene, torch.from_numpy(g["scene"])) < t["scene"]
    assert _report("frame", frame, torch.from_numpy(g["frame"])) < t["frame"]
    # log-mel against the reference's own output, ASSERTED, in dB: undo bn0 (per-mel affine) on both sides
    lm = m.forward_logmel(wave).cpu()[:, :: int(g["logmel_stride"])]
    sd = m.state_dict()
    scale = (sd["bn0.weight"] / torch.sqrt(sd["bn0.running_var"] + 1e-5)).cpu()
    shift = (sd["bn0.bias"].cpu() - sd["bn0.running_mean"].cpu() * scale)
    ref_db = (torch.from_numpy(g["logmel_bn"]) - shift) / scale
    d = ((lm - shift) / scale - ref_db).abs()
    inband = ref_db > ref_db.max() - 80.0
    stats = dict(max=d.max().item(), p99=d.flatten().kthvalue(int(0.99 * d.numel())).values.item(),
                 mean=d.mean().item(), inband_max=d[inband].max().item())
    print("  logmel [dB]: " + " ".join(f"{k} {v:.2e}" for k, v in stats.items()))
    for k, lim in LOGMEL_DB[prec].items():
        assert stats[k] < lim, (k, stats[k], lim)
    # thresholded label set (demo_convnext.py:87-88)
    thr = float(np.log(0.25 / 0.75))
    ref_logits = g["logits"][0]
    ours = set(np.where(probs[0].numpy() > 0.25)[0].tolist())
    sure_on = set(np.where(ref_logits > thr + t["logits"])[0].tolist())
    sure_off = set(np.where(ref_logits < thr - t["logits"])[0].tolist())
    assert sure_on <= ours and not (ours & sure_off)
    if prec == "fp32":
        assert ours == set(g["labels"].tolist())


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["noise", "tones"])
def test_synthetic_clips_against_reference_golden(models, golden_dir, prec, kind):
    g = np.load(os.path.join(golden_dir, f"synth_{kind}.npz"))
    wave = weights.make_waveforms(2, kind=kind, seed=0).to(DEV)
    m = models[prec]
    res = m.forward_all(wave)
    print(f"[{prec}] synthetic {kind} vs reference golden")
    t = TOL[prec]
    assert _report("logits", res["clipwise_logits"].cpu(), torch.from_numpy(g["logits"])) < t["logits"]
    assert _report("scene", res["scene_embeddings"].cpu(), torch.from_numpy(g["scene"])) < t["scene"]
    fr = res["frame_embeddings"].cpu()[:, :, :: int(g["frame_stride"])]
    assert _report("frame", fr, torch.from_numpy(g["frame_t0"])) < t["frame"]


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_variable_length_and_batch_independence(models, golden_dir, prec):
    g = np.load(os.path.join(golden_dir, "synth_short.npz"))
    L = int(g["n_samples"])
    wave = weights.make_waveforms(1, n_samples=L, kind="noise", seed=3).to(DEV)
    m = models[prec]
    frame = m.forward_frame_embeddings(wave).cpu()
    assert frame.shape == g["frame"].shape
    t = TOL[prec]
    assert _report("frame(short)", frame, torch.from_numpy(g["frame"])) < t["frame"]
    assert _report("logits(short)", m(wave)["clipwise_logits"].cpu(), torch.from_numpy(g["logits"])) < t["logits"]
    # clips are independent: a clip's result must not depend on its batch neighbours or on chunking
    w5 = weights.make_waveforms(5, n_samples=L, kind="noise", seed=9).to(DEV)
    w5[2] = wave[0]
    a = m(w5)["clipwise_logits"]
    b = m(wave)["clipwise_logits"]
    assert torch.equal(a[2], b[0])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_stock_head_north_star_tolerances_unscaled(golden_dir, prec):
    """north_star as written: |logits - reference| < 2e-2 (bf16 mode) / 1e-4 (fp32 mode), no scaling, on a trunk whose
    blocks all contribute (gamma ~ U(0.1, 0.6)) and the reference's own classifier width (std 0.02): demo clip, white
    noise and band-limited tones, against outputs of the UNMODIFIED reference (oracle/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "parity_stock_head.npz"))
    sd = weights.make_state_dict("parity_stock_head", int(g["parity_seed"]))
    m = acx.convnext_tiny(pretrained=False, strict=False, drop_path_rate=0.0, after_stem_dim=[252, 56])
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval().set_precision(prec)
    demo = np.load(os.path.join(golden_dir, "demo_clip.npz"))["pcm"]
    waves = {"demo": torch.from_numpy(demo.astype(np.float32) / 32768.0)[None],
             "noise": weights.make_waveforms(2, kind="noise", seed=0),
             "tones": weights.make_waveforms(2, kind="tones", seed=0)}
    print(f"[{prec}] stock-head parity weights vs reference golden (north_star tolerance {NORTH_STAR[prec]:g})")
    for name, w in waves.items():
        out = m(w.to(DEV))
        e_l = _report(f"logits[{name}]", out["clipwise_logits"].cpu(), torch.from_numpy(g[f"{name}/logits"]))
        e_p = _report(f"probs[{name}]", out["clipwise_output"].cpu(), torch.from_numpy(g[f"{name}/probs"]))
        assert e_l < NORTH_STAR[prec] and e_p < NORTH_STAR[prec]
        scene = m.forward_scene_embeddings(w.to(DEV)).cpu()
        _report(f"scene[{name}]", scene, torch.from_numpy(g[f"{name}/scene"]))


def test_stock_init_weights_unscaled_tolerances(golden_dir):
    """Reference random init (what bench.py runs): north_star tolerances applied as written."""
    g = np.load(os.path.join(golden_dir, "synth_init.npz"))
    sd = weights.make_state_dict("init", 0)
    wave = weights.make_waveforms(1, kind="noise", seed=0).to(DEV)
    for prec, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56])
        m.load_state_dict(sd)
        m = m.to(DEV).eval().set_precision(prec)
        err = _report(f"logits[{prec}, init]", m(wave)["clipwise_logits"].cpu(), torch.from_numpy(g["logits"]))
        assert err < tol


def test_bf16_mode_against_oracle_per_stage(models, parity_sd):
    """Localises error: compares the bf16 engine's stage outputs with the oracle's taps."""
    wave = weights.make_waveforms(1, n_samples=64000, kind="tones", seed=7)
    taps = {}
    O.forward(wave, parity_sd, taps=taps)
    m = models["bf16"]
    lm = m.forward_logmel(wave.to(DEV)).cpu()
    err = (lm - taps["logmel_bn"]).abs()
    print(f"  logmel_bn: max {err.max():.3e} mean {err.mean():.3e}")
    fr = m.forward_frame_embeddings(wave.to(DEV)).cpu()
    ref = taps["stage3"]
    rel = (fr - ref).abs().max() / ref.std()
    print(f"  stage3: max|d|/std {rel:.3e}")
    assert rel < 0.5


def test_host_pipeline_matches_direct_calls(models):
    """HostPipeline (double-buffered H2D / compute / D2H; fp32 and int16 PCM inputs) == direct forward calls."""
    m = models["bf16"]
    L = 32000 * 2
    batches = [weights.make_waveforms(3, n_samples=L, kind="noise", seed=20 + i).pin_memory() for i in range(5)]
    pipe = acx.HostPipeline(m, want=("logits", "frame"))
    res = pipe.run(batches)
    assert len(res) == 5
    for hb, r in zip(batches, res):
        d = m.forward_all(hb.to(DEV))
        assert torch.equal(r["logits"], d["clipwise_logits"].cpu())
        assert torch.equal(r["probs"], d["clipwise_output"].cpu())
        assert torch.equal(r["frame"], d["frame_embeddings"].cpu())
    # int16 PCM as stored in the AudioSet HDF5 files (data_generator.py:70-74, utilities.py:226-227)
    pcm = [(b * 32767.0).round().clamp(-32767, 32767).to(torch.int16).pin_memory() for b in batches[:2]]
    res16 = acx.HostPipeline(m).run(pcm)
    for p, r in zip(pcm, res16):
        ref_wave = (p.numpy() / 32767.0).astype("float32")        # utilities.py:226-227, on the host
        d = m(torch.from_numpy(ref_wave).to(DEV))
        assert torch.equal(r["logits"], d["clipwise_logits"].cpu())


def test_eval_loop_drop_in_matches_reference_loop_semantics(models, parity_sd):
    """evalloop.forward == the reference's pytorch_utils.forward (PU:63-137): per batch model(x)['clipwise_output'],
    concatenated; accepts the collate_fn's dtype=object arrays and int16 PCM."""
    from audioset_convnext_inf_b200 import evalloop
    m = models["fp32"]
    L = 32000
    waves = weights.make_waveforms(7, n_samples=L, kind="noise", seed=33)
    obj = np.empty(3, dtype=object)
    for i in range(3):
        obj[i] = waves[4 + i].numpy()
    gen = [{"waveform": waves[:4].numpy(), "target": np.zeros((4, 527), np.float32)},
           {"waveform": obj, "target": np.ones((3, 527), np.float32)}]
    out = evalloop.forward(m, gen, return_input=True, return_target=True)
    assert out["clipwise_output"].shape == (7, 527) and out["target"].shape == (7, 527) and out["waveform"].shape == (7, L)
    ref = O.forward(waves, parity_sd)["clipwise_output"].numpy()       # what the reference loop would return
    assert np.abs(out["clipwise_output"] - ref).max() < 1e-4
    direct = torch.cat([m(waves[:4].to(DEV))["clipwise_output"], m(waves[4:].to(DEV))["clipwise_output"]]).cpu().numpy()
    assert np.array_equal(out["clipwise_output"], direct)


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_edge_shapes(models, prec):
    """B=1, batch larger than the chunk, the shortest clip the 4-stage trunk accepts, non-contiguous input."""
    m = models[prec]
    eng = m._get_engine()
    L = 16000
    w = weights.make_waveforms(eng.chunk + 3 if prec == "fp32" else 5, n_samples=L, kind="noise", seed=50).to(DEV)
    full = m(w)["clipwise_logits"]
    one = m(w[1:2])["clipwise_logits"]
    assert torch.equal(full[1], one[0])
    nc = torch.stack([w, w], 2)[:, :, 0]                              # non-contiguous view
    assert not nc.is_contiguous()
    assert torch.equal(m(nc)["clipwise_logits"], full)
    short = weights.make_waveforms(2, n_samples=7 * 320 * 4 + 1, kind="noise", seed=51).to(DEV)   # T'=1 at stage 3
    fr = m.forward_frame_embeddings(short)
    ref = O.forward_frame_embeddings(short.cpu(), weights.make_state_dict("parity", 8))
    assert fr.shape == ref.shape and fr.shape[2] >= 1
    empty = m(torch.zeros(0, L, device=DEV))                         # empty batch: empty outputs, no launch
    assert empty["clipwise_output"].shape == (0, 527) and empty["clipwise_logits"].shape == (0, 527)
    assert m.forward_scene_embeddings(torch.zeros(0, L, device=DEV)).shape == (0, 768)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 2000, device=DEV))                          # too short for the trunk
    with pytest.raises(ValueError):
        m(torch.zeros(320000, device=DEV))                           # not (batch, samples)


def test_baseline_config_shapes_front_end_512_and_ragged_200(models, parity_sd):
    """BASELINE.json configs #3 (front end only, batch 512) and #5 (batch sweep, sizes that are not multiples of the
    64-clip chunk) at full size: per-clip results must equal the same clip processed alone / in a small batch
    (clips are independent), and a sample is checked against the oracle."""
    m = models["bf16"]
    g = torch.Generator(device=DEV).manual_seed(7)
    w = (torch.randn(512, 320000, device=DEV, generator=g) * 0.1).clamp_(-1, 1)
    lm = m.forward_logmel(w)
    assert lm.shape == (512, 1001, 224) and torch.isfinite(lm).all()
    idx = [0, 63, 64, 300, 511]
    small = m.forward_logmel(w[idx].contiguous())
    assert torch.equal(lm[idx], small)
    ref = O.frontend(w[idx[:2]].cpu(), parity_sd, torch.float64)
    assert (lm[idx[:2]].cpu().double() - ref).abs().mean() < 2e-5           # white noise: near-fp32 accuracy
    out = m(w[:200])                                                         # 64 + 64 + 64 + 8 clips
    ref_small = m(w[[0, 70, 199]].contiguous())
    assert torch.equal(out["clipwise_logits"][[0, 70, 199]], ref_small["clipwise_logits"])
    assert out["clipwise_output"].shape == (200, 527)


# ---- SURVEY §8 "next" rows f3 / f4 --------------------------------------------------------------------------------
@pytest.mark.parametrize("orig", [44100, 48000, 16000, 22050, 32000])
def test_resample_fit_matches_torchaudio_then_pad_crop(orig):
    """acx_resample_fit vs torchaudio.functional.resample on the CPU (what demo_convnext.py:52-67 does), including
    the constant pad (short clip) and the crop (long clip) to a fixed length, several channels as the batch."""
    TAF = pytest.importorskip("torchaudio.functional")
    g = torch.Generator().manual_seed(orig)
    for secs, n_out in ((0.73, 32000), (1.9, 32000), (1.0, None)):
        L = int(orig * secs) + 3
        wave = torch.randn(2, L, generator=g) * 0.3
        ref = TAF.resample(wave, orig, 32000)
        if n_out is not None:
            ref = torch.nn.functional.pad(ref, (0, max(n_out - ref.shape[-1], 0)))[:, :n_out]
        got = acx.preprocess.resample_fit(wave.to(DEV), orig, 32000, n_out).cpu()
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() < 2e-6 * max(1.0, ref.abs().max().item()) * 5      # fp32 summation order


def test_resample_rejects_bad_arguments():
    from audioset_convnext_inf_b200 import _native as N
    t = torch.zeros(16, device=DEV)
    rc = N.load().acx_resample_fit(t.data_ptr(), 16, t.data_ptr(), t.data_ptr(), 8, 1, 16, 0, 2, 3, 16, 0)
    assert rc != 0 and "rates" in N.last_error()
    rc = N.load().acx_resample_fit(t.data_ptr(), 8, t.data_ptr(), t.data_ptr(), 16, 1, 16, 3, 2, 3, 16, 0)
    assert rc != 0 and "pitch" in N.last_error()


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_extract_clipwise_buckets_equal_per_clip_loop(models, prec):
    """extract.extract_clipwise (exact-length buckets, batched) returns, in input order, exactly what the reference's
    per-file loop (pytorch/extract_embeddings.py:64-92: one forward per clip) would: per-clip results do not depend
    on the batch they ran in."""
    m = models[prec]
    g = torch.Generator().manual_seed(11)
    lengths = [40000, 64000, 40000, 33333, 64000, 40000, 33333]
    waves = [(torch.randn(n, generator=g) * 0.1).numpy() for n in lengths]
    res = acx.extract.extract_clipwise(m, waves, max_batch=2, want=("clipwise_logits", "scene_embeddings"))
    for i, w in enumerate(waves):
        with torch.no_grad():
            one = m.forward_all(torch.from_numpy(w)[None].to(DEV))
        assert torch.equal(res["clipwise_logits"][i], one["clipwise_logits"][0].float().cpu())
        assert torch.equal(res["scene_embeddings"][i], one["scene_embeddings"][0].float().cpu())


def test_demo_script_runs_resample_pad_and_three_outputs(tmp_path):
    """demo.py (counterpart of demo_convnext.py) end to end on a 16 kHz, 3 s wav: resample + pad on the GPU, tags,
    scene and frame embeddings with the reference's shapes."""
    import subprocess
    import sys
    from scipy.io import wavfile
    rng = np.random.default_rng(0)
    wav = tmp_path / "clip16k.wav"
    wavfile.write(wav, 16000, (rng.standard_normal(48000) * 3000).astype(np.int16))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "demo.py"), "--wav", str(wav)], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Resampling from 16000 to 32000 Hz" in r.stdout
    assert "logits size: (1, 527)" in r.stdout
    assert "Scene embedding, shape: (1, 768)" in r.stdout
    assert "Frame-level embeddings, shape: (1, 768, 31, 7)" in r.stdout


def test_workspace_lru_keeps_shapes_and_graphs_across_alternating_lengths(models):
    """Engine._workspace is an LRU over (clips, samples) shapes and each workspace owns its CUDA graphs: alternating
    lengths (variable-length extraction, ragged last batches) must neither re-allocate nor re-capture per call, a shape
    is captured on its second use only, and replayed results equal eager ones bit for bit."""
    m = models["bf16"]
    eng = m._get_engine()
    if not eng.use_graph:
        pytest.skip("graphs disabled (ACX_GRAPH=0)")
    eng._ws.clear()
    La, Lb = 32000, 48000
    wa = weights.make_waveforms(3, n_samples=La, kind="noise", seed=70).to(DEV)
    wb = weights.make_waveforms(2, n_samples=Lb, kind="noise", seed=71).to(DEV)
    first_a = m(wa)["clipwise_logits"].clone()                      # eager (first use of the shape)
    first_b = m(wb)["clipwise_logits"].clone()
    assert set(eng._ws) == {(3, La), (2, Lb)}
    assert not eng._ws[(3, La)]["graphs"] and not eng._ws[(2, Lb)]["graphs"]
    ids = {k: id(v) for k, v in eng._ws.items()}
    for _ in range(3):                                              # alternate: second use captures, later uses replay
        assert torch.equal(m(wa)["clipwise_logits"], first_a)
        assert torch.equal(m(wb)["clipwise_logits"], first_b)
    assert {k: id(v) for k, v in eng._ws.items()} == ids            # no re-allocation
    assert len(eng._ws[(3, La)]["graphs"]) == 1 and len(eng._ws[(2, Lb)]["graphs"]) == 1
    ga = next(iter(eng._ws[(3, La)]["graphs"].values()))[0]
    m(wb), m(wa)
    assert next(iter(eng._ws[(3, La)]["graphs"].values()))[0] is ga  # no re-capture
    # eviction: least recently used shape goes first, together with its graphs
    for i in range(eng.max_workspaces):
        m(weights.make_waveforms(1, n_samples=20000 + 320 * i, kind="noise", seed=80 + i).to(DEV))
    assert (3, La) not in eng._ws and len(eng._ws) == eng.max_workspaces


def test_eval_loop_streams_a_generator_with_bounded_pinned_memory(models):
    """evalloop.forward pulls batches from a generator one at a time (the loader overlaps the kernels) and stages
    pageable batches through the pipeline's `depth` pinned buffers: nothing else is page-locked."""
    from audioset_convnext_inf_b200 import evalloop
    m = models["bf16"]
    L = 32000
    pulled = []

    def gen():
        for i in range(6):
            pulled.append(i)
            yield {"waveform": weights.make_waveforms(2, n_samples=L, kind="noise", seed=90 + i).numpy()}

    out = evalloop.forward(m, gen())
    assert pulled == list(range(6)) and out["clipwise_output"].shape == (12, 527)
    ref = torch.cat([m(weights.make_waveforms(2, n_samples=L, kind="noise", seed=90 + i).to(DEV))["clipwise_output"]
                     for i in range(6)]).cpu().numpy()
    assert np.array_equal(out["clipwise_output"], ref)
    pipe = acx.HostPipeline(m)
    pipe.run([weights.make_waveforms(2, n_samples=L, kind="noise", seed=1) for _ in range(5)])   # pageable inputs
    assert sum(s["stage"] is not None for s in pipe._slots) == pipe.depth                        # depth staging buffers


def test_amplitude_contract_check_is_opt_in(models, monkeypatch):
    """INTEGRATION.md input contract: bf16 mode carries samples as 2^8-scaled fp16 pairs (|x| < 255).  ACX_CHECK_AMPLITUDE=1
    turns a violation into an explicit error; the fp32-accurate mode accepts any float amplitude like the reference."""
    m = models["bf16"]
    w = weights.make_waveforms(1, n_samples=32000, kind="noise", seed=5).to(DEV)
    monkeypatch.setenv("ACX_CHECK_AMPLITUDE", "1")
    assert torch.isfinite(m(w)["clipwise_logits"]).all()
    with pytest.raises(ValueError, match="amplitude"):
        m(w * 32767.0)
    ref = O.forward((w * 300.0).cpu(), weights.make_state_dict("parity", 8))["clipwise_logits"]
    got = models["fp32"](w * 300.0)["clipwise_logits"].cpu()
    assert (got - ref).abs().max() < 5e-4


@pytest.mark.parametrize("env", [{"ACX_DWCONV": "simt"}, {"ACX_GP": "0"}, {"ACX_LN": "smem"}, {"ACX_DWCONV_TC_STAGES": "0"},
                                 {"ACX_DS_GP": "0"}, {"ACX_DS_FUSED": "0"}, {"ACX_DS_FUSED_C": "96,192,384"}, {"ACX_PDL": "0"},
                                 {"ACX_PDL": "31"}, {"ACX_FRONTEND": "folded"}])
def test_alternative_kernel_routes_agree(parity_sd, monkeypatch, env):
    """Every selectable route of the block -- CUDA-core depthwise conv + fused LayerNorm (round 1), tensor-core conv on
    row-major tensors, LayerNorm applied in shared memory instead of folded, planar layout in stage 0 only, transpose pass
    instead of a planar downsample GEMM, ln_patchify + GEMM instead of the implicit-GEMM downsample kernel (and that kernel at
    all three widths), programmatic dependent launch off / on for every kernel family, the front end on folded frames -- gives the same logits as the default route within the bf16-mode tolerance:
    the routes differ only in where roundings to bf16 happen."""
    wave = weights.make_waveforms(2, n_samples=64000, kind="tones", seed=3).to(DEV)

    def run():
        m = acx.convnext_tiny(pretrained=False, strict=False, drop_path_rate=0.0, after_stem_dim=[252, 56])
        m.load_state_dict(parity_sd, strict=True)
        m = m.to(DEV).eval().set_precision("bf16")
        return m(wave)["clipwise_logits"].cpu()

    base = run()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    alt = run()
    ref = O.forward(wave.cpu(), parity_sd)["clipwise_logits"]
    assert (alt - ref).abs().max().item() < TOL["bf16"]["logits"]
    assert (alt - base).abs().max().item() < TOL["bf16"]["logits"]


def test_graph_replays_are_bit_identical_under_dependent_launch(models):
    """Programmatic dependent launch lets every kernel's prologue start under the previous kernel's tail; a read or write
    placed above `griddepcontrol.wait` by mistake would show up as run-to-run differences.  200 replays of the captured
    graph at full batch width (64 clips: every persistent kernel has a tail), eager launches and a different batch in
    between: each result equals the first one bit for bit."""
    m = models["bf16"]
    a = weights.make_waveforms(64, n_samples=64000, kind="tones", seed=11).to(DEV)
    b = weights.make_waveforms(64, n_samples=64000, kind="noise", seed=12).to(DEV)
    first_a = m(a)["clipwise_logits"].clone()
    first_b = m(b)["clipwise_logits"].clone()
    for i in range(100):
        assert torch.equal(m(a)["clipwise_logits"], first_a), i
        assert torch.equal(m(b)["clipwise_logits"], first_b), i
    eng = m._get_engine()
    eng.start_timing(None)                      # eager path: an event pair around every launch
    eager = eng.run(a)["logits"].clone()
    eng.stop_timing()
    assert torch.equal(eager, first_a)
