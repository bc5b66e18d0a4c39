"""Thermally fair A/B of programmatic dependent launch: one engine per ACX_PDL mask in ONE process (the mask is read
at capture time), graph replays interleaved mask by mask so every variant sees the same clocks.
usage: python tools/ab_pdl.py [masks, comma separated] [batch] [rounds]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audioset_convnext_inf_b200 as acx  # noqa: E402

masks = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,31").split(",")]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 12
torch.manual_seed(0)
waves = [(torch.randn(B, 320000, device="cuda") * 0.1).clamp(-1, 1) for _ in range(3)]
engines = {}
ref = None
for mk in masks:
    os.environ["ACX_PDL"] = str(mk)
    torch.manual_seed(1)
    m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56]).cuda().eval()
    eng = m._get_engine()
    for i in range(4):
        out = eng.run(waves[i % 3])       # 2nd use of the shape captures the graph under this mask
    torch.cuda.synchronize()
    lg = eng.run(waves[0])["logits"].clone()
    if ref is None:
        ref = lg
    print(f"mask {mk}: max |logit - logit(mask {masks[0]})| = {(lg - ref).abs().max().item():.3e}")
    engines[mk] = (m, eng)
tot = {mk: [] for mk in masks}
for r in range(rounds):
    for mk in masks:
        eng = engines[mk][1]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            eng.run(waves[i % 3])
        e1.record()
        torch.cuda.synchronize()
        tot[mk].append(e0.elapsed_time(e1) / 10)
for mk in masks:
    v = sorted(tot[mk])
    print(f"ACX_PDL={mk:2d}: median {v[len(v) // 2]:.4f}  min {v[0]:.4f}  max {v[-1]:.4f} ms/step")
