"""flex_gym.env.env.BlackPanther_V55 (reference: .../BlackPanther_V55/__init__.py:1-6): the native class and the resource directory."""
import os

from _flexible_robot import *  # noqa: F401,F403
from _flexible_robot import FlexibleGymEnv  # noqa: F401

__BLACKPANTHER_V55_RESOURCE_DIRECTORY__ = os.path.dirname(os.path.abspath(__file__)) + '/urdf'
