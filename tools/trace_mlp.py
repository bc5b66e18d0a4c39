"""Bring-up aid: SM-clock timeline of CTA 0 of the weight-resident fused MLP kernel (first 16 tiles).
role 0 = MMA thread, role 1 = epilogue-1 warp 4 (group 0: even chunks) and epilogue-2 warp 12.  usage: python tools/trace_mlp.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioset_convnext_inf_b200 import _native as N  # noqa: E402

DEV = "cuda:0"
C, M = 96, 14112 * 32
g = torch.Generator().manual_seed(0)
y = torch.randn(M, C, generator=g).to(torch.bfloat16).to(DEV)
x = torch.randn(M, C, generator=g).to(torch.bfloat16).to(DEV)
w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(torch.bfloat16).to(DEV)
w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16).to(DEV)
b1 = torch.zeros(4 * C, device=DEV)
b2 = torch.zeros(C, device=DEV)
gamma = torch.ones(C, device=DEV)
trace = torch.zeros(16 * 2 * 32, dtype=torch.int64, device=DEV)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    N.call("acx_mlp_fused", y.data_ptr(), x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(), M, C, st)
torch.cuda.synchronize()
os.environ["ACX_TRACE_PTR"] = str(trace.data_ptr())
N.call("acx_mlp_fused", y.data_ptr(), x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(), M, C, st)
torch.cuda.synchronize()
t = trace.cpu().view(16, 2, 32)
t0 = int(t[t > 0].min())
names0 = [f"g2({h}) {w}" for h in range(6) for w in ("h_full seen", "issued")]
names1 = [f"E1({2 * hc}) {w}" for hc in range(3) for w in ("start", "d1_full", "d1 released", "h_full")]
names1 += ["E2 resid requested", "E2 D2 released", "-", "tile end"] + ["-"] * 9
for it in range(4, 6):
    print(f"--- tile iteration {it} (cycles since first stamp)")
    ev = [(int(t[it, 0, i]) - t0, "MMA " + names0[i]) for i in range(12) if t[it, 0, i] > 0]
    ev += [(int(t[it, 1, i]) - t0, "EPI " + names1[i]) for i in range(25) if t[it, 1, i] > 0]
    prev = None
    for c, n in sorted(ev):
        print(f"  {c:9d}  (+{0 if prev is None else c - prev:6d})  {n}")
        prev = c
