"""Multi-GPU plumbing: envs are independent (no inter-env term anywhere in Drone.step, components.py:220-248), so
the batch is cut into contiguous per-rank slices, each rank steps its slice with no data-path collective, and the only
communication is the all-reduce of the 8-double episode-statistics vector (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

import torch

STAT_KEYS = ("env_steps", "crashes", "episodes", "episode_len_sum", "reward_sum", "reward_sq_sum", "nonfinite")


def env_shard(total_envs: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous slice [start, start+count) of `total_envs` owned by `rank`.  When total_envs divides evenly the slices are
    equal; otherwise every start is a multiple of 64 (one warp chunk of the kernel) and the counts differ by at most 64
    (by at most one when total_envs < 64 * world_size)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    if total_envs < 0:
        raise ValueError("total_envs must be >= 0")
    base, rem = divmod(total_envs, world_size)
    if rem == 0 or total_envs < 64 * world_size:
        start = rank * base + min(rank, rem)
        return start, base + (1 if rank < rem else 0)
    chunks = -(-total_envs // 64)
    cb, cr = divmod(chunks, world_size)
    c0 = rank * cb + min(rank, cr)
    c1 = c0 + cb + (1 if rank < cr else 0)
    return min(c0 * 64, total_envs), min(c1 * 64, total_envs) - min(c0 * 64, total_envs)


def reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the fpv_stats_t vector over the process group (no-op without an initialised group)."""
    s = stats.clone()
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(s, op=torch.distributed.ReduceOp.SUM, group=group)
    return s


def stats_dict(stats: torch.Tensor) -> dict:
    v = stats.tolist()
    out = dict(zip(STAT_KEYS, v))
    out["mean_episode_len"] = out["episode_len_sum"] / out["episodes"] if out["episodes"] else float("nan")
    return out
