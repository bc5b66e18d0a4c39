"""Times the fused-MLP variants on 64 clips of 10 s (stage 0: C=96, stage 1: C=192): row-major with / without the
in-kernel LayerNorm, group-planar with / without.  python tools/time_mlp.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioset_convnext_inf_b200 import _native as N  # noqa: E402

DEV = "cuda:0"
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
for C, M in ((96, 64 * 252 * 56), (192, 64 * 126 * 28)):
    g = torch.Generator(device=DEV).manual_seed(1)
    vs = [torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16) for _ in range(3)]
    x = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, device=DEV, generator=g) / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, device=DEV, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    b1 = torch.zeros(4 * C, device=DEV)
    b2 = torch.zeros(C, device=DEV)
    gamma = torch.full((C,), 1e-3, device=DEV)
    lw, lb = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)

    def t(fn, iters=12):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    a = (w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(), M, C)
    rm = t(lambda i: N.call("acx_mlp_fused", vs[i % 3].data_ptr(), x.data_ptr(), *a, st()))
    rm_ln = t(lambda i: N.call("acx_mlp_fused_ln", vs[i % 3].data_ptr(), x.data_ptr(), lw.data_ptr(), lb.data_ptr(), *a, st()))
    gp = t(lambda i: N.call("acx_mlp_fused_gp", vs[i % 3].data_ptr(), x.data_ptr(), 0, 0, 0, *a, st()))
    gp_ln = t(lambda i: N.call("acx_mlp_fused_gp", vs[i % 3].data_ptr(), x.data_ptr(), lw.data_ptr(), lb.data_ptr(), 0, *a, st()))
    s1 = torch.zeros(4 * C, device=DEV)
    gp_fold = t(lambda i: N.call("acx_mlp_fused_gp", vs[i % 3].data_ptr(), x.data_ptr(), 0, 0, s1.data_ptr(), *a, st()))
    print(f"C={C}: row-major {rm:6.1f} us, +LN in smem {rm_ln:6.1f} us | group-planar {gp:6.1f} us, +LN in smem {gp_ln:6.1f} us, "
          f"+LN folded {gp_fold:6.1f} us")
