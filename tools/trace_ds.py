"""In-kernel clock trace of the fused downsample kernel (CTA 0): needs a -DACX_DS_TRACE build of libacx
(ACX_NVCC_EXTRA=-DACX_DS_TRACE python audioset-convnext-inf_b200/build.py --force; cp libacx.so libacx_trace.so) loaded
through ACX_LIBACX.  usage: ACX_LIBACX=.../libacx_trace.so python tools/trace_ds.py [stage] [clips]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioset_convnext_inf_b200 import _native as N  # noqa: E402
from audioset_convnext_inf_b200.engine import pack_downsample_weight  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
C, Wd, H = (96, 192, 384)[stage], 56 >> stage, (252, 126, 63)[stage]
dev = "cuda:0"
trace = torch.zeros(32, dtype=torch.int64, device=dev)
os.environ["ACX_DS_TRACE_PTR"] = str(trace.data_ptr())
M = B * H * Wd
Mp = (M + 127) // 128 * 128
Mo = B * (H // 2) * (Wd // 2)
xg = torch.randn(C // 8, Mp, 8, device=dev).to(torch.bfloat16)
lw, lb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
w = pack_downsample_weight(torch.randn(2 * C, C, 2, 2, device=dev) * 0.05).to(torch.bfloat16).contiguous()
bias = torch.zeros(2 * C, device=dev)
out = torch.empty(2 * C // 8, (Mo + 127) // 128 * 128, 8, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N.call("acx_downsample_fused_gp", xg.data_ptr(), lw.data_ptr(), lb.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(),
           B, H, Wd, C, 1, st)
    e1.record()
    torch.cuda.synchronize()
t = trace.cpu().tolist()
tiles = -(-Mo // 128) * (2 if C == 384 else 1)
print(f"C={C}: {e0.elapsed_time(e1) * 1e3:.1f} us, {tiles} tiles, ~{tiles / 148:.1f} per CTA; SM clocks of CTA 0:")
print(f"  weight producer : total {t[0]:>8d}  wait empty {t[1]:>8d}")
print(f"  MMA issuer      : total {t[3]:>8d}  wait full  {t[4]:>8d}  wait tmem-empty {t[5]:>8d}")
print(f"  gather thread 0 : total {t[6]:>8d}  wait empty {t[7]:>8d}  bar.sync {t[8]:>8d}")
print(f"  gather thread255: total {t[9]:>8d}  wait empty {t[10]:>8d}  bar.sync {t[11]:>8d}")
print(f"  gather thread 0 : stats part {t[16]:>8d}  load + normalise {t[17]:>8d}  store + fence + arrive {t[18]:>8d}")
print(f"  epilogue warp 0 : total {t[12]:>8d}  wait tmem-full {t[13]:>8d}")
