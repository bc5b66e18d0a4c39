"""Counterpart of the reference's per-file extraction loop (pytorch/extract_embeddings.py:64-92: one forward per wav,
whatever its length) for a LIST of variable-length clips.

The model's result for a clip depends on its exact length (reflect padding, frame count), so clips are never padded
to a common length: they are bucketed by EXACT length, every bucket runs as batched forwards, and the outputs are
returned in the caller's order -- identical to looping clip by clip, at batch throughput."""
import collections

import numpy as np
import torch


def length_buckets(lengths, max_batch):
    """[(length, [indices...])...]: indices grouped by equal length, split into runs of at most max_batch."""
    by_len = collections.OrderedDict()
    for i, n in enumerate(lengths):
        by_len.setdefault(int(n), []).append(i)
    out = []
    for n, idx in by_len.items():
        for s in range(0, len(idx), max_batch):
            out.append((n, idx[s:s + max_batch]))
    return out


def extract_clipwise(model, waveforms, max_batch=64, want=("clipwise_logits",)):
    """waveforms: sequence of 1-D float arrays / tensors at 32 kHz (any lengths).  Returns {name: list of per-clip
    tensors on the host, in input order} for the names in `want` (keys of ConvNeXt.forward_all)."""
    dev = next(model.parameters()).device
    res = {k: [None] * len(waveforms) for k in want}
    for n, idx in length_buckets([len(w) for w in waveforms], max_batch):
        batch = torch.stack([torch.as_tensor(np.asarray(waveforms[i]), dtype=torch.float32) for i in idx]).to(dev)
        with torch.no_grad():
            out = model.forward_all(batch)
        for k in want:
            host = out[k].float().cpu()
            for j, i in enumerate(idx):
                res[k][i] = host[j]
    return res
