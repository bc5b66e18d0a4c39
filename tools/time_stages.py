"""Per-kernel device-time breakdown of one forward (CUDA events around every libacx launch).
usage: python tools/time_stages.py [batch] [precision]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audioset_convnext_inf_b200 as acx  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
torch.manual_seed(0)
m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56])
m = m.cuda().eval().set_precision(prec)
wave = (torch.randn(B, 320000, device="cuda") * 0.1).clamp(-1, 1)
eng = m._get_engine()
for _ in range(2):
    eng.run(wave)
torch.cuda.synchronize()
tot = collections.OrderedDict()
cnt = collections.Counter()
runs = 3
for _ in range(runs):
    for tag, ms in eng.profile(wave):
        tot[tag] = tot.get(tag, 0.0) + ms / runs
        cnt[tag] += 1
total = sum(tot.values())
print(f"B={B} precision={prec} chunk={eng.chunk} frontend={eng.frontend} mlp={eng.mlp}: {total:.3f} ms/forward "
      f"-> {B / total * 1e3:.0f} clips/s (serialised per-kernel timing)")
for tag, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {tag:24s} {ms:9.3f} ms  {100 * ms / total:5.1f}%  x{cnt[tag] // runs}")
# un-instrumented wall time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    eng.run(wave)
e1.record()
torch.cuda.synchronize()
print(f"un-instrumented: {e0.elapsed_time(e1) / 5:.3f} ms/forward -> {B / (e0.elapsed_time(e1) / 5) * 1e3:.0f} clips/s")
