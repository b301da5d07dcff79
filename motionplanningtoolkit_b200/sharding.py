"""Pose-set sharding over ranks / devices: contiguous, verdict-word aligned slices, no exchange step.

SURVEY section 8e: pose-sets are independent units; every rank holds the whole scene and checks its own
slice; the slices' verdict words concatenate without bit shifts because every boundary is a multiple
of 32 pose-sets.  The C library uses the same rule for the devices of one context
(csrc/mptcc_api.cpp, mptcc_check_poses).
"""
import numpy as np


def shard_range(n_pose_sets, world_size, rank):
    """[begin, end) of rank's slice; begin is a multiple of 32."""
    n_words = (n_pose_sets + 31) // 32
    per = (n_words + world_size - 1) // world_size
    b = min(n_words, rank * per) * 32
    e = min(n_words, (rank + 1) * per) * 32
    return min(b, n_pose_sets), min(e, n_pose_sets)


def concat_verdict_words(slices):
    """Verdict words of consecutive shards -> words of the whole batch."""
    return np.concatenate([np.asarray(s, dtype=np.uint32) for s in slices]) if slices else np.zeros(0, np.uint32)


def check_sharded(check_fn, poses, world_size, rank):
    """Run `check_fn(pose_slice) -> uint32 words` on this rank's slice."""
    b, e = shard_range(len(poses), world_size, rank)
    return check_fn(poses[b:e]) if e > b else np.zeros(0, np.uint32)


def all_gather_verdicts(local_words, n_pose_sets, world_size, rank, dist):
    """Gather every rank's verdict words on all ranks (torch.distributed, any backend).  Only a
    convenience for callers that want the whole bitmask everywhere; the hot path has no collective."""
    import torch
    n_words = (n_pose_sets + 31) // 32
    per = (n_words + world_size - 1) // world_size
    buf = torch.zeros(per, dtype=torch.int32)
    lw = torch.from_numpy(np.asarray(local_words, dtype=np.uint32).view(np.int32).copy())
    buf[:lw.numel()] = lw
    out = [torch.zeros(per, dtype=torch.int32) for _ in range(world_size)]
    dist.all_gather(out, buf)
    words = torch.cat(out).numpy().view(np.uint32)[:n_words]
    return words.copy()
