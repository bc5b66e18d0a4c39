"""Step time of the batch-64 forward (CUDA-graph replay, device-resident input): the quick A/B number.
usage: [ACX_PDL=mask] python tools/time_step.py [batch] [replays] [repeats]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audioset_convnext_inf_b200 as acx  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56]).cuda().eval()
waves = [(torch.randn(B, 320000, device="cuda") * 0.1).clamp(-1, 1) for _ in range(3)]
eng = m._get_engine()
for i in range(6):
    eng.run(waves[i % 3])
torch.cuda.synchronize()
res = []
for _ in range(rounds):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        eng.run(waves[i % 3])
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / reps)
print(f"ACX_PDL={os.environ.get('ACX_PDL', 'default')} B={B}: " + " ".join(f"{r:.4f}" for r in res) + f" ms/step (best {min(res):.4f})")
