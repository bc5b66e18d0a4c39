"""GPU (>= 2 B200 on one box): clip-sharded replicas + NCCL all-gather == single-GPU result, bit for bit
(clips are independent).  Skipped on a single-GPU box."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, L, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import audioset_convnext_inf_b200 as acx
    from audioset_convnext_inf_b200.dist import sharded_forward
    from oracle import weights
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56])
    m.load_state_dict(weights.make_state_dict("parity", 8))
    m = m.to(f"cuda:{rank}").eval()
    waves = weights.make_waveforms(n_clips, n_samples=L, kind="noise", seed=42).to(f"cuda:{rank}")
    out = sharded_forward(lambda w: m.forward_all(w), waves)
    ok = True
    if rank == 0:
        ref = m.forward_all(waves)
        ok = all(torch.equal(out[k], ref[k]) for k in ref)
    q.put((rank, ok, tuple(out["frame_embeddings"].shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("n_clips", [6, 5])
def test_sharded_forward_bit_identical_to_single_gpu(n_clips):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_clips, 48000, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert res[0][2][0] == n_clips


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_devices_in_one_process_give_identical_results():
    """One process driving cuda:0 and then cuda:1 (not the one-process-per-GPU layout of bench.py): per-device kernel
    attributes (dynamic shared memory opt-in), tensor maps and workspaces must all follow the model's device."""
    sys.path.insert(0, ROOT)
    import audioset_convnext_inf_b200 as acx
    from oracle import weights
    sd = weights.make_state_dict("parity", 8)
    waves = weights.make_waveforms(3, n_samples=64000, kind="noise", seed=9)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = acx.convnext_tiny(drop_path_rate=0.0, after_stem_dim=[252, 56])
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        o = m.forward_all(waves.to(dev))
        assert all(v.device == torch.device(dev) for v in o.values())
        outs.append({k: v.cpu() for k, v in o.items()})
    assert all(torch.equal(outs[0][k], outs[1][k]) for k in outs[0])
