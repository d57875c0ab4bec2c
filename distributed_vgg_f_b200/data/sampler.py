"""Index arithmetic of the sharded training sampler.

Reference behaviour (distributedVggf.py:114-116 -> torch.utils.data.DistributedSampler with its
defaults): ``randperm(N)`` seeded with ``seed + epoch``, padded by wrap-around to
``ceil(N / ws) * ws`` entries, rank ``r`` takes ``indices[r::ws]``.  The reference never calls
``set_epoch`` so every epoch replays the epoch-0 order; ``ShardedSampler`` reshuffles
per epoch by default and keeps ``reference_order=True`` for bit-parity with that quirk.
"""
from __future__ import annotations

import math
from typing import List

import torch


def shard_indices(n: int, world_size: int, rank: int, epoch: int = 0, seed: int = 0,
                  shuffle: bool = True, drop_last: bool = False, pad: bool = True) -> List[int]:
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        indices = torch.randperm(n, generator=g).tolist()
    else:
        indices = list(range(n))
    if not pad:          # exact partition (sharded evaluation: every sample counted once)
        return indices[rank::world_size]
    if drop_last and n % world_size != 0:
        num = math.ceil((n - world_size) / world_size)
    else:
        num = math.ceil(n / world_size)
    total = num * world_size
    if not drop_last:
        pad = total - len(indices)
        if pad <= len(indices):
            indices += indices[:pad]
        else:
            indices += (indices * math.ceil(pad / len(indices)))[:pad]
    else:
        indices = indices[:total]
    return indices[rank:total:world_size]


class ShardedSampler:
    def __init__(self, n: int, world_size: int = 1, rank: int = 0, shuffle: bool = True,
                 seed: int = 0, reference_order: bool = False, pad: bool = True) -> None:
        if not 0 <= rank < world_size:
            raise ValueError("invalid rank %d for world size %d" % (rank, world_size))
        self.n, self.world_size, self.rank = n, world_size, rank
        self.shuffle, self.seed = shuffle, seed
        self.reference_order = reference_order
        self.pad = pad
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = 0 if self.reference_order else epoch

    def __len__(self) -> int:
        if not self.pad:
            return len(range(self.rank, self.n, self.world_size))
        return math.ceil(self.n / self.world_size)

    def __iter__(self):
        return iter(shard_indices(self.n, self.world_size, self.rank, self.epoch, self.seed,
                                  self.shuffle, pad=self.pad))
