"""Clip-sharded multi-GPU inference: one process per GPU, replicas of the 28 M-parameter model, the
batch split by clip, and ONE collective -- the final all-gather of the outputs.

The reference has no inference-time multi-GPU path (one `cuda` device, evaluate_convnext_on_audioset.py:45-47);
clips are independent in eval mode (BatchNorm uses running stats, pooling is per clip), so no other exchange
is needed and per-clip results are bit-identical to a single-GPU run.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_clips, world_size, rank):
    """Contiguous block split: rank r owns [r*ceil(N/W), min(N, (r+1)*ceil(N/W)))."""
    per = -(-n_clips // world_size)
    lo = min(n_clips, rank * per)
    return lo, min(n_clips, lo + per)


def gather_rows(local, n_total, group=None):
    """All-gather row blocks of possibly unequal length (last ranks may own fewer / zero clips):
    pad to ceil(N/W) rows, one all_gather_into_tensor, trim."""
    world = dist.get_world_size(group)
    per = -(-n_total // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return out[:n_total]


def sharded_forward(run_local, waveforms, group=None):
    """run_local(wave_block) -> dict of tensors with leading clip dimension, evaluated on this rank's block
    of `waveforms` (N, L) (every rank passes the same N; only its block needs to hold real data).
    Returns the dict for all N clips on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = waveforms.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    if hi > lo:
        local = run_local(waveforms[lo:hi])
    else:
        probe = run_local(waveforms[:1])            # shapes only; contributes zero rows
        local = {k: v[:0] for k, v in probe.items()}
    return {k: gather_rows(v, n, group) for k, v in local.items()}
