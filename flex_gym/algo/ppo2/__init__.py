"""flex_gym.algo.ppo2 (reference: flex_gym/algo/ppo2/ppo2.py) -> the PyTorch / CUDA re-host."""
from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import PPO2, LstmActorCritic  # noqa: F401
