"""Frame sharding for data-parallel fitting.

The reference only gestures at this (`DistributedSampler`, main_train.py:206, never initialised); the semantics kept
here are the ones that make K ranks x batch 1 equal the reference run with `-b K`: a per-epoch seeded permutation of
the N frames, cut into floor(N / K) global batches of K frames — the reference DataLoader has `drop_last=True`
(main_train.py:207-209), so the N % K frames at the tail of an epoch's permutation are skipped, a different set every
epoch — rank r taking position r of every batch.  Every rank therefore runs floor(N / K) steps per epoch.
"""
import torch


def epoch_permutation(n_frames, epoch, seed=1, shuffle=True):
    if not shuffle:
        return list(range(n_frames))
    g = torch.Generator().manual_seed(seed + epoch)
    return torch.randperm(n_frames, generator=g).tolist()


def steps_per_epoch(n_frames, world_size):
    return max(1, n_frames // world_size)


def shard_indices(n_frames, world_size, rank, epoch, seed=1, shuffle=True):
    """Frame indices rank `rank` fits in epoch `epoch` (length floor(n_frames / world_size), at least 1)."""
    perm = epoch_permutation(n_frames, epoch, seed, shuffle)
    steps = steps_per_epoch(n_frames, world_size)
    total = steps * world_size
    if total > n_frames:                       # fewer frames than ranks: wrap so that every rank has work
        perm = (perm * (total // n_frames + 1))[:total]
    return perm[rank:total:world_size]
