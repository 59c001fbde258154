"""Environment sharding over GPUs: env i -> rank floor(i * G / N), contiguous ranges, no data-path collective."""
from __future__ import annotations

import os
from typing import Tuple


def shard_range(total_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the global env ids owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_from_torchrun():
    """(rank, local_rank, world_size) from the torchrun environment (defaults for a single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_sharded_env(cfg: dict, total_envs: int, resource_dir: str = ""):
    """One FlexibleGymEnv per process holding this rank's contiguous shard; env_offset keys the RNG with global ids."""
    from ._flexible_robot import FlexibleGymEnv
    from .cfg import dump_yaml
    rank, local, world = env_from_torchrun()
    if total_envs % world != 0:
        # PPO2 averages per-rank mean gradients with equal weights and keys the act model with rank * n_envs: shards must be equal
        raise ValueError(f"total_envs = {total_envs} is not a multiple of the world size {world}: environment shards must be equal")
    lo, hi = shard_range(total_envs, world, rank)
    c = dict(cfg, num_envs=hi - lo)
    return FlexibleGymEnv(resource_dir, dump_yaml(c), device=local, env_offset=lo)
