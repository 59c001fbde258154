"""flex_gym.env.RaisimGymVecEnv (reference: flex_gym/env/RaisimGymVecEnv.py:6-189)."""
from high_speed_quadrupedal_locomotion_by_irrl_b200.vec_env import RaisimGymVecEnv, LazyInfo, Box  # noqa: F401
