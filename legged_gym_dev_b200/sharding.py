"""Env sharding across GPUs (one process per GPU).  Environments are independent: each rank owns a
contiguous range of GLOBAL env ids, keys its random streams with them (so per-env results do not depend on the number
of ranks) and only ever exchanges reduced statistics."""
import torch


def env_shard(rank, world, total_envs):
    """[lo, hi) of rank `rank`: contiguous, near-equal, boundaries on multiples of 4 (16-byte rows for the TMA tiles)."""
    per = (total_envs // world) // 4 * 4
    if per == 0:
        per = total_envs // world
    lo = rank * per
    hi = total_envs if rank == world - 1 else lo + per
    return lo, hi


def combine_advantage_stats(stats):
    """(sum, sum^2, n) summed over ranks -> (mean, unbiased std), as RolloutStorage.compute_returns needs them."""
    s, s2, n = stats[0], stats[1], stats[2]
    mean = s / n
    var = torch.clamp((s2 - n * mean * mean) / (n - 1.0), min=0.0)
    return mean, torch.sqrt(var)


def combine_episode_stats(raw, term_names, max_episode_length_s, total_envs, terrain_level=False):
    """extras["episode"] (legged_robot.py:175-182) from RAW per-step rows `raw` [..., K + 2] = (per-term sums over the envs that reset, sum of
    terrain levels over all envs, number of resets) — the rows the step kernel leaves in `extras_raw`, summed over the env shards by ONE
    all-reduce: the means a single process over all envs would log.  Rows without a reset give NaN (nothing to average)."""
    K = len(term_names)
    cnt = raw[..., K + 1]
    out = {"rew_" + n: raw[..., i] / cnt / max_episode_length_s for i, n in enumerate(term_names)}
    if terrain_level:
        out["terrain_level"] = raw[..., K] / float(total_envs)
    return out


def init_distributed(device=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE (torchrun) and sets up NCCL over NVLink; returns (rank, local_rank, world)."""
    import os
    import torch.distributed as dist
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        dev = torch.device("cuda", local_rank) if device is None else device
        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=dev)
    return rank, local_rank, world
