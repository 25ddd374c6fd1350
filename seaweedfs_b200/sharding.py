"""Multi-GPU driver logic for batches of independent volumes (BASELINE configs[3]).

RS encode has no cross-volume (or cross-column) data dependency, so the path shards with NO data-path
collective: volume v goes to rank v mod N (the shell's ec.encode runs volumes concurrently the same
way, weed/shell/command_ec_encode.go:302-315), each rank drives its own GPU on its own stream, and
torch.distributed is used only for the timing bracket (barrier, MAX of per-rank device time) and for
gathering per-volume digests so rank 0 can print one checksum-of-checksums for the whole batch.
Backend: nccl on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Callable

SEED0 = 0x5EA3EED5F00DCAFE
_MASK = (1 << 64) - 1


def volumes_for_rank(n_volumes: int, world_size: int, rank: int) -> list[int]:
    """Round-robin placement: volume v → rank v mod N."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world size")
    return list(range(rank, n_volumes, world_size))


def volume_seed(v: int, seed0: int = SEED0) -> int:
    return (seed0 + v) & _MASK


def combine_digests(digests: dict[int, int]) -> int:
    """Order-independent checksum of per-volume digests (checksum of checksums)."""
    acc = 0
    for v, d in digests.items():
        z = (d + (v + 1) * 0x9E3779B97F4A7C15) & _MASK
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
        acc = (acc + (z ^ (z >> 31))) & _MASK
    return acc


def run_batch(n_volumes: int, encode_volume: Callable[[int, int], tuple[int, float]], dist=None,
              device=None) -> dict:
    """Encode this rank's share of the batch.  encode_volume(v, seed) → (digest, device_ms).
    Returns on every rank {"ms_max", "volumes", "digest", "per_rank_ms"}; collective calls are
    control-plane only (a few bytes)."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    mine = volumes_for_rank(n_volumes, world, rank)
    if dist is not None:
        dist.barrier()
    local_ms = 0.0
    digests: dict[int, int] = {}
    for v in mine:
        d, ms = encode_volume(v, volume_seed(v))
        digests[v] = d & _MASK
        local_ms += ms
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    per_rank = [t.clone() for _ in range(world)]
    if dist is not None:
        dist.all_gather(per_rank, t)
        gathered: list = [None] * world
        dist.all_gather_object(gathered, digests)
        digests = {k: v for part in gathered for k, v in part.items()}
    per_rank_ms = [float(x.item()) for x in per_rank]
    if sorted(digests) != list(range(n_volumes)):
        raise RuntimeError("volume placement lost or duplicated a volume")
    return {"ms_max": max(per_rank_ms), "volumes": n_volumes, "digest": combine_digests(digests),
            "per_rank_ms": per_rank_ms}
