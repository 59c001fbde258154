"""flex_gym.archi.policies: only the pieces run_bp_v5.py's CustomLSTMPolicy uses survive."""
from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import LstmActorCritic as ActorCriticPolicy  # noqa: F401
from high_speed_quadrupedal_locomotion_by_irrl_b200.ppo2 import LstmActorCritic as LstmPolicy  # noqa: F401
from high_speed_quadrupedal_locomotion_by_irrl_b200.policy import FusedLstmPolicy  # noqa: F401
