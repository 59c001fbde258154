"""`flex_gym.env.env.BlackPanther_V55`: exports the native vec-env class and the resource directory constant that
run_bp_v5.py imports (run_bp_v5.py:9,13).  The resource directory's `default_cfg.yaml` is materialised on first import from
`cfg.train_cfg()` (the values of the reference's shipped training file) so that `--cfg` defaults keep working."""
import pathlib

import yaml

from _flexible_robot import FlexibleGymEnv  # noqa: F401
from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg import train_cfg

_RSC = pathlib.Path(__file__).resolve().parent / "urdf"
__BLACKPANTHER_V55_RESOURCE_DIRECTORY__ = str(_RSC)


def _ensure_default_cfg() -> None:
    target = _RSC / "default_cfg.yaml"
    if target.exists():
        return
    try:
        _RSC.mkdir(parents=True, exist_ok=True)
        target.write_text("# generated from high_speed_quadrupedal_locomotion_by_irrl_b200.cfg.train_cfg()\n"
                          + yaml.safe_dump({"seed": 1, "record_video": False, "environment": train_cfg()}, sort_keys=False))
    except OSError:
        pass


_ensure_default_cfg()
