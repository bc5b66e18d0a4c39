"""GPU bring-up aid for the tcgen05 GEMM: exact-integer operands and one-hot probes that reveal
descriptor / swizzle / lane-mapping mistakes.  Writes gpurun_out/diag_umma.npz.  Not a test."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audioset_convnext_inf_b200 import _native as N  # noqa: E402

DEV = "cuda:0"
os.makedirs("gpurun_out", exist_ok=True)
dump = {}


def gemm(A, W, epi=0, bias=None):
    M, K = A.shape
    Nn = W.shape[0]
    out = torch.full((M, Nn), float("nan"), device=DEV, dtype=torch.bfloat16)
    b = torch.zeros(Nn, device=DEV) if bias is None else bias
    N.call("acx_gemm_bf16", A.data_ptr(), W.data_ptr(), out.data_ptr(), M, Nn, K, epi, b.data_ptr(), 0, 0,
           torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.float()


def main():
    print("device ok:", N.load().acx_device_ok(), torch.cuda.get_device_name(0))
    g = torch.Generator().manual_seed(0)
    for (M, Nn, K) in [(128, 96, 64), (128, 128, 64), (128, 256, 64), (128, 192, 64), (128, 128, 16), (128, 128, 32),
                       (128, 128, 128), (256, 128, 256), (200, 96, 96), (128, 128, 512)]:
        A = torch.randint(-3, 4, (M, K), generator=g).float().to(torch.bfloat16).to(DEV)
        W = torch.randint(-3, 4, (Nn, K), generator=g).float().to(torch.bfloat16).to(DEV)
        ref = A.float() @ W.float().t()
        try:
            out = gemm(A, W)
        except Exception as e:  # noqa: BLE001
            print(f"M{M} N{Nn} K{K}: EXCEPTION {e}")
            return 1
        bad = (out != ref)
        print(f"M{M} N{Nn} K{K}: mismatches {int(bad.sum())}/{bad.numel()}  nan {int(torch.isnan(out).sum())}"
              f"  max|d| {float((out - ref).abs().nan_to_num(1e9).max()):.1f}")
        dump[f"out_{M}_{Nn}_{K}"] = out.cpu().numpy()
        dump[f"ref_{M}_{Nn}_{K}"] = ref.cpu().numpy()
        if bad.any() and (M, Nn, K) == (128, 128, 64):
            # one-hot K probe: which k does each instruction slice really read?
            for k0 in range(0, 64, 8):
                A1 = torch.zeros(M, K, device=DEV, dtype=torch.bfloat16)
                A1[:, k0] = 1
                Wk = (torch.arange(Nn, device=DEV)[:, None] * 0 + torch.arange(K, device=DEV)[None, :] + 1).float().to(torch.bfloat16)
                o = gemm(A1, Wk)
                print(f"  one-hot A[:, {k0}] -> out[0,:4] = {o[0, :4].tolist()} (expect {k0 + 1})")
            A2 = torch.zeros(M, K, device=DEV, dtype=torch.bfloat16)
            A2[:, 0] = (torch.arange(M, device=DEV) + 1).float().to(torch.bfloat16)
            W2 = torch.zeros(Nn, K, device=DEV, dtype=torch.bfloat16)
            W2[:, 0] = 1
            o = gemm(A2, W2)
            print("  row probe out[:8,0] =", o[:8, 0].tolist(), " out[32:36,0] =", o[32:36, 0].tolist())
            dump["row_probe"] = o.cpu().numpy()
    np.savez_compressed("gpurun_out/diag_umma.npz", **dump)
    return 0


if __name__ == "__main__":
    sys.exit(main())
