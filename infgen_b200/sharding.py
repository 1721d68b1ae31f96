"""Scene sharding across ranks and the final metric gather.

The decode path never exchanges data between scenes, so multi-GPU is: scene i -> rank i mod world (what Lightning's
DistributedSampler does for the reference's `trainer.validate`, run.py:130-139), one process per GPU, and ONE
collective after all rollouts - the gather of per-scenario metric states (`LongMetric` sync,
infgen/metrics/compute_metrics.py:1200-1205).  NCCL on GPUs, gloo in the CPU tests.
"""
from typing import List, Sequence
import torch
import torch.distributed as dist


def shard_scenes(n_scenes: int, rank: int, world: int) -> List[int]:
    """Indices of the scenes rank `rank` rolls out (round-robin, every scene exactly once)."""
    return list(range(rank, n_scenes, world))


def gather_scene_metrics(local_ids: Sequence[int], local_values: torch.Tensor, n_scenes: int) -> torch.Tensor:
    """All-gather ragged per-scene metric rows into one [n_scenes, D] tensor on every rank.

    local_values: [len(local_ids), D] on the rank's device. Works for world == 1 without a process group."""
    d = local_values.shape[1] if local_values.dim() == 2 else 1
    local_values = local_values.reshape(len(local_ids), d)
    out = torch.zeros(n_scenes, d, dtype=local_values.dtype, device=local_values.device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out[list(local_ids)] = local_values
        return out
    world = dist.get_world_size()
    per_rank = (n_scenes + world - 1) // world
    pad = torch.zeros(per_rank, d + 1, dtype=local_values.dtype, device=local_values.device)
    pad[:, 0] = -1
    if len(local_ids):
        pad[:len(local_ids), 0] = torch.as_tensor(list(local_ids), dtype=local_values.dtype, device=local_values.device)
        pad[:len(local_ids), 1:] = local_values
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    for b in bufs:
        ok = b[:, 0] >= 0
        out[b[ok, 0].long()] = b[ok, 1:]
    return out
