#!/usr/bin/env python
"""Counterpart of the reference's demo_convnext.py (one clip -> tags @0.25, scene and frame embeddings), on the
B200 path.  Reads the wav with scipy (torchaudio.load needs torchcodec, absent here); resampling to 32 kHz and the
pad / crop to 10 s of demo_convnext.py:52-67 are one GPU launch (preprocess.resample_fit), then ONE forward gives all
three outputs.

  python demo.py --checkpoint model.safetensors --wav clip.wav [--labels class_labels_indices.csv] [--precision fp32]
"""
import argparse
import csv
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import audioset_convnext_inf_b200 as acx  # noqa: E402

SR, TARGET = 32000, 10 * 32000


def read_wav(path):
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.ndim > 1:
        data = data[:, 0]                                   # demo clip is mono; take the first channel otherwise
    if data.dtype == np.int16:
        wave = data.astype(np.float32) / 32768.0            # torchaudio.load normalisation
    else:
        wave = data.astype(np.float32)
    return wave, sr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", help=".safetensors or .pth; random init if omitted")
    ap.add_argument("--wav", required=True)
    ap.add_argument("--labels", help="class_labels_indices.csv (index,mid,display_name)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--threshold", type=float, default=0.25)
    args = ap.parse_args()
    model = (acx.ConvNeXt.from_pretrained(args.checkpoint) if args.checkpoint else
             acx.convnext_tiny(pretrained=False, strict=False, drop_path_rate=0.0, after_stem_dim=[252, 56]))
    print("# params:", sum(p.numel() for p in model.parameters() if p.requires_grad))
    model = model.to("cuda").eval().set_precision(args.precision)
    wave, sr = read_wav(args.wav)
    if sr != SR:
        print("Resampling from %d to 32000 Hz" % sr)
    # resample (identity at 32 kHz) + constant pad / crop to 10 s, on the GPU
    clip = acx.preprocess.resample_fit(torch.from_numpy(wave)[None].to("cuda"), sr, SR, TARGET)
    out = model.forward_all(clip)
    probs = out["clipwise_output"][0].cpu().numpy()
    print("logits size:", tuple(out["clipwise_logits"].shape))
    labels = np.where(probs > args.threshold)[0]
    print(f"Predicted labels using activity threshold {args.threshold}:\n\n{labels}")
    if args.labels:
        with open(args.labels) as fh:
            names = [row[2] for row in list(csv.reader(fh))[1:]]
        for ix in labels:
            print("%s: %.3f" % (names[ix], probs[ix]))
    print("\nScene embedding, shape:", tuple(out["scene_embeddings"].shape))
    print("\nFrame-level embeddings, shape:", tuple(out["frame_embeddings"].shape))


if __name__ == "__main__":
    main()
