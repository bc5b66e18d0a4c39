import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioset_convnext_inf_b200 import _native as N
DEV = "cuda:0"
def to_gp(t):
    M, C = t.shape
    Mp = (M + 127) // 128 * 128                      # plane stride: rows rounded up to 128
    out = torch.zeros(C // 8, Mp, 8, device=t.device, dtype=t.dtype)
    out[:, :M] = t.view(M, C // 8, 8).permute(1, 0, 2)
    return out
def from_gp(t, M, C):
    return t.view(C // 8, -1, 8)[:, :M].permute(1, 0, 2).reshape(M, C).contiguous()
for C, M, ln in [(96, 19277, False), (96, 100, False), (96, 128 * 3, False), (192, 19077, False), (96, 19277, True)]:
    g = torch.Generator().manual_seed(11 * C + M)
    v = (torch.randn(M, C, generator=g) * 2 + torch.randn(M, 1, generator=g) * 5).to(torch.bfloat16)
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(4 * C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    gamma = torch.rand(C, generator=g) * 0.5 + 0.1
    lw = torch.rand(C, generator=g) * 0.4 + 0.8
    lb = torch.randn(C, generator=g) * 0.05
    vd, xd, w1d, w2d, b1d, b2d, gd, lwd, lbd = (t.to(DEV) for t in (v, x, w1, w2, b1, b2, gamma, lw, lb))
    st = torch.cuda.current_stream().cuda_stream
    x_rm = xd.clone()
    if ln:
        N.call("acx_mlp_fused_ln", vd.data_ptr(), x_rm.data_ptr(), lwd.data_ptr(), lbd.data_ptr(), w1d.data_ptr(), b1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    else:
        N.call("acx_mlp_fused", vd.data_ptr(), x_rm.data_ptr(), w1d.data_ptr(), b1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    v_gp, x_gp = to_gp(vd), to_gp(xd)
    N.call("acx_mlp_fused_gp", v_gp.data_ptr(), x_gp.data_ptr(), lwd.data_ptr() if ln else 0, lbd.data_ptr() if ln else 0, 0, w1d.data_ptr(), b1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    torch.cuda.synchronize()
    got = from_gp(x_gp, M, C)
    d = (got.float() - x_rm.float()).abs()
    bad = (d > 0)
    rows = bad.any(1).nonzero().flatten()
    cols = bad.any(0).nonzero().flatten()
    print(f"C={C} M={M} ln={ln}: max diff {d.max().item():.4g}, mismatching elements {int(bad.sum())} of {bad.numel()}, rows {rows.numel()} (first {rows[:8].tolist()}, tiles {sorted(set((rows // 128).tolist()))[:12]}), cols {cols[:16].tolist()}")
    if rows.numel():
        r = int(rows[0])
        print("   row", r, "got", got[r, :8].tolist(), "ref", x_rm[r, :8].tolist(), "x", xd[r, :8].tolist())
