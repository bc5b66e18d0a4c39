#!/usr/bin/env python
"""Counterpart of the reference's convert_pytorch_ckpt_to_safetensors.py: load a `.pth` checkpoint (`{"model": ...}`,
as evaluate_convnext_on_audioset.py:36-38 expects) or a `.safetensors` file through ConvNeXt.from_pretrained and write
`model.safetensors` (strict 190-key state dict, loadable by the reference's `safetensors.torch.load_model`).
CPU only: no forward is run.

  python tools/convert_checkpoint.py convnext_tiny_471mAP.pth model.safetensors
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import audioset_convnext_inf_b200 as acx  # noqa: E402


def convert(src, dst):
    from safetensors.torch import save_model
    model = acx.ConvNeXt.from_pretrained(src, map_location="cpu")
    print("# params:", sum(p.numel() for p in model.parameters() if p.requires_grad))
    save_model(model, dst)
    return model


if __name__ == "__main__":
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    convert(sys.argv[1], sys.argv[2])
