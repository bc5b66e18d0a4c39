"""One forward of a batch (bf16 tensor-core mode) inside a cudaProfilerStart/Stop range: the short command ncu wraps
(`ncu --profile-from-start off ...`).  CUDA-graph replay is disabled so every kernel is an individual launch.
usage: ACX_GRAPH=0 python tools/run_once.py [clips]"""
import os
import sys

os.environ.setdefault("ACX_GRAPH", "0")
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audioset_convnext_inf_b200 as acx  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56]).cuda().eval()
wave = (torch.randn(n, 320000, device="cuda") * 0.1).clamp(-1, 1)
out = m(wave)                      # warm-up (weight repack, attribute setup)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = m(wave)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok", out["clipwise_logits"].shape, m._get_engine().launches)
