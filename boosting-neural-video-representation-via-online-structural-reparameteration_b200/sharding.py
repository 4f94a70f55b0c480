"""Frame sharding for data-parallel fitting.

The reference only gestures at this (`DistributedSampler`, main_train.py:206, never initialised); the
semantics kept here are the ones that make K ranks x batch 1 equal the reference run with `-b K`:
a per-epoch seeded permutation of the N frames, padded (by wrapping) to a multiple of K, rank r taking
positions r, r+K, ...  Every rank therefore runs ceil(N/K) steps per epoch.
"""
import math

import torch


def epoch_permutation(n_frames, epoch, seed=1, shuffle=True):
    if not shuffle:
        return list(range(n_frames))
    g = torch.Generator().manual_seed(seed + epoch)
    return torch.randperm(n_frames, generator=g).tolist()


def shard_indices(n_frames, world_size, rank, epoch, seed=1, shuffle=True):
    """Frame indices rank `rank` fits in epoch `epoch` (length ceil(n_frames / world_size))."""
    perm = epoch_permutation(n_frames, epoch, seed, shuffle)
    steps = math.ceil(n_frames / world_size)
    total = steps * world_size
    perm = perm + perm[: total - n_frames]
    return perm[rank:total:world_size]


def steps_per_epoch(n_frames, world_size):
    return math.ceil(n_frames / world_size)
