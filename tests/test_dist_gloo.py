"""CPU, world_size 2 over gloo: the clip-sharding + single all-gather plumbing of the N>1 path."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audioset_convnext_inf_b200.dist import shard_bounds, sharded_forward
    g = torch.Generator().manual_seed(0)
    waves = torch.randn(n_clips, 64, generator=g)

    def fake_model(w):      # stands in for the per-rank engine: any per-clip function
        return {"logits": torch.stack([w.sum(1), w.abs().sum(1)], 1), "frame": w[:, :6].reshape(-1, 2, 3) * 2}

    lo, hi = shard_bounds(n_clips, world, rank)
    mine = waves.clone()
    mine[:lo] = float("nan")        # a rank must never need data outside its shard
    mine[hi:] = float("nan")
    if hi == lo:
        mine[:1] = 0.0
    out = sharded_forward(fake_model, mine)
    ref = fake_model(waves)
    ok = all(torch.equal(out[k], ref[k]) for k in ref)
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


def _run(n_clips, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    return sorted(res)


def test_shard_bounds_cover_all_clips():
    from audioset_convnext_inf_b200.dist import shard_bounds
    for n in (0, 1, 5, 64, 1024, 20481):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_sharded_forward_even_split():
    res = _run(8)
    assert all(ok for _, ok, _ in res), res


def test_sharded_forward_ragged_and_empty_rank():
    assert all(ok for _, ok, _ in _run(5)), "ragged split"
    res = _run(1)
    assert all(ok for _, ok, _ in res) and res[1][2] == (1, 1), res
