"""Drop-in for the reference's pybind11 module `_flexible_robot` (raisim_gym.cpp:14-47): `from _flexible_robot import FlexibleGymEnv`."""
from high_speed_quadrupedal_locomotion_by_irrl_b200._flexible_robot import FlexibleGymEnv  # noqa: F401
