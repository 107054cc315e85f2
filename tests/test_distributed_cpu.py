"""world_size-2 gloo tests of the multi-GPU host logic: sample sharding, one-time weight broadcast, gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from sketch2img_b200 import distributed as D
    r, w, _ = D.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)                      # ranks start with DIFFERENT weights
    sd = {"a.weight": torch.randn(7, 5), "b.bias": torch.randn(3).half(), "steps": torch.tensor(3)}
    D.broadcast_state_dict(sd, src=0, bucket_bytes=64)
    torch.manual_seed(100)
    want = {"a.weight": torch.randn(7, 5), "b.bias": torch.randn(3).half()}
    ok = torch.equal(sd["a.weight"], want["a.weight"]) and torch.equal(sd["b.bias"], want["b.bias"]) and sd["b.bias"].dtype == torch.float16
    # matrices as fp16 on the wire (the engine packs them to fp16 operands anyway), vectors exact
    torch.manual_seed(200 + rank)
    sd2 = {"w": torch.randn(6, 4), "conv.weight": torch.randn(2, 3, 3, 3), "w.bias": torch.randn(6)}
    D.broadcast_state_dict(sd2, src=0, half_matrices=True)
    torch.manual_seed(200)
    w2 = {"w": torch.randn(6, 4), "conv.weight": torch.randn(2, 3, 3, 3), "w.bias": torch.randn(6)}
    ok = ok and torch.equal(sd2["w"].half(), w2["w"].half()) and torch.equal(sd2["conv.weight"].half(), w2["conv.weight"].half())
    ok = ok and torch.equal(sd2["w.bias"], w2["w.bias"]) and sd2["w"].dtype == torch.float32
    n = 5
    mine = D.shard_samples(n, rank, world)
    local = torch.stack([torch.full((4, 2, 2), float(k)) for k in mine])
    full = D.gather_latents(local, n, rank, world)
    ok = ok and full.shape == (n, 4, 2, 2) and all(float(full[k, 0, 0, 0]) == k for k in range(n))
    ok = ok and D.max_over_ranks(1.0 + rank) == float(world)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_sharding_is_a_partition():
    from sketch2img_b200.distributed import shard_samples
    for n in (0, 1, 7, 32, 64):
        for w in (1, 2, 4, 8):
            seen = sorted(k for r in range(w) for k in shard_samples(n, r, w))
            assert seen == list(range(n))
            assert max(len(shard_samples(n, r, w)) for r in range(w)) - min(len(shard_samples(n, r, w)) for r in range(w)) <= 1
