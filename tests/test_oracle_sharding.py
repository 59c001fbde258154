"""Environments shard as independent units: the counter-based RNG is keyed by the GLOBAL env id, so a job
gives the same per-env results however it is split.  Checked on the oracle here (CPU) and on the CUDA path in the gpu suite."""
import numpy as np

from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import trot_cfg
from oracle_lib import Oracle


def test_two_shards_equal_one_job():
    cfg = trot_cfg(num_envs=8, num_threads=2, StochasticDynamics=True, ObsNoise=2.0)
    whole = Oracle(cfg)
    half = dict(cfg, num_envs=4)
    a, b = Oracle(half, env_offset=0), Oracle(half, env_offset=4)
    ow = whole.reset(); oa = a.reset(); ob = b.reset()
    assert np.array_equal(ow[:4], oa) and np.array_equal(ow[4:], ob)
    rng = np.random.default_rng(0)
    for t in range(30):
        act = np.clip(rng.normal(0, 0.2, size=(8, 12)), -1, 1).astype(np.float32)
        w = whole.step(act); x = a.step(act[:4]); y = b.step(act[4:])
        for i in range(4):
            assert np.array_equal(w[i][:4], x[i]) and np.array_equal(w[i][4:], y[i])


def test_shard_partition_helper():
    from high_speed_quadrupedal_locomotion_by_irrl_b200.sharding import shard_range
    total = 65536
    for world in (1, 2, 4, 8):
        spans = [shard_range(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    assert [shard_range(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
