"""Two forwards of one 32-clip chunk (bf16 tensor-core mode): the short command ncu wraps.
usage: python tools/run_once.py [clips]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audioset_convnext_inf_b200 as acx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56]).cuda().eval()
wave = (torch.randn(n, 320000, device="cuda") * 0.1).clamp(-1, 1)
for _ in range(2):
    out = m(wave)
torch.cuda.synchronize()
print("ok", out["clipwise_logits"].shape, m._get_engine().launches)
