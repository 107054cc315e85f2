"""Multi-GPU driver for the sampling path: one process per GPU, samples sharded, no per-step collective.

Each image (its CFG pair) is a closed computation, so rank r takes samples r, r+W, r+2W, ... and runs
them through its own engine replica.  The only collective is the one-time weight broadcast from rank 0 (NCCL over
NVLink on GPUs, gloo in the CPU tests); final latents are gathered for convenience.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK/WORLD_SIZE/MASTER_* variables (no-op for world size 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    if torch.cuda.is_available():
        torch.cuda.set_device(local % torch.cuda.device_count())
    return rank, world, local


def shard_samples(n, rank, world):
    """Indices of the samples rank `rank` owns (round-robin: sample k -> GPU k mod world)."""
    return list(range(rank, n, world))


def broadcast_state_dict(state_dict, src=0, device=None, bucket_bytes=1 << 30, half_matrices=False):
    """One-time weight broadcast: rank `src`'s tensors overwrite everyone's, in flat buckets sized for launch latency
    (NVSwitch gives every peer full bandwidth; the bucket count, not link count, is what matters).

    half_matrices: tensors with two or more dimensions (conv / linear weights) travel as fp16 -- the engine only ever uses
    them as fp16 tensor-core operands (packed with round-to-nearest at load), so every rank still packs bit-identical
    operands from half the bytes (SD1.5: 1.7 GB instead of 3.4 GB); vectors (biases, norm affine) stay fp32."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return state_dict
    keys = sorted(k for k, v in state_dict.items() if torch.is_tensor(v) and v.dtype.is_floating_point)
    dev = torch.device(device) if device is not None else (
        torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))

    def wire(k):
        return torch.float16 if half_matrices and state_dict[k].dim() >= 2 else torch.float32

    for wdt in (torch.float16, torch.float32):
        bucket, size = [], 0

        def flush():
            nonlocal bucket, size
            if not bucket:
                return
            flat = torch.cat([state_dict[k].detach().reshape(-1).to(dev, wdt) for k in bucket])
            dist.broadcast(flat, src=src)
            off = 0
            for k in bucket:
                n = state_dict[k].numel()
                state_dict[k] = flat[off:off + n].reshape(state_dict[k].shape).to(state_dict[k].device, state_dict[k].dtype)
                off += n
            bucket, size = [], 0

        for k in keys:
            if wire(k) != wdt:
                continue
            bucket.append(k)
            size += state_dict[k].numel() * (2 if wdt == torch.float16 else 4)
            if size >= bucket_bytes:
                flush()
        flush()
    return state_dict


def gather_latents(local_latents, n_total, rank, world):
    """All ranks' final latents reassembled in sample order on every rank (optional convenience; off the step)."""
    if world == 1:
        return local_latents
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local_latents.shape[1:]), device=local_latents.device, dtype=local_latents.dtype)
    pad[:local_latents.shape[0]] = local_latents
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    full = torch.empty((n_total,) + tuple(local_latents.shape[1:]), device=local_latents.device, dtype=local_latents.dtype)
    for r in range(world):
        idx = shard_samples(n_total, r, world)
        if idx:
            full[idx] = outs[r][:len(idx)]
    return full


def max_over_ranks(value, device=None):
    """Device-time aggregation for benchmarks: MAX over ranks of a python float."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
