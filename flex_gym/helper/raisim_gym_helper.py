"""Run-directory bookkeeping used by the training script (role of flex_gym/helper/raisim_gym_helper.py in the reference:
snapshot the config next to the logs; the TensorBoard launcher has nothing to launch on a headless GPU box)."""
import pathlib
import shutil
import time


class ConfigurationSaver:
    """Creates `<log_dir>/<timestamp>/` and copies the given files into it; `.data_dir` is that directory."""

    def __init__(self, log_dir, save_items=()):
        stamp = time.strftime("%Y-%m-%d-%H-%M-%S")
        self._dir = pathlib.Path(log_dir) / stamp
        self._dir.mkdir(parents=True, exist_ok=True)
        for item in save_items or ():
            src = pathlib.Path(item) if item else None
            if src is not None and src.is_file():
                shutil.copy2(src, self._dir / src.name)

    @property
    def data_dir(self):
        return str(self._dir)


def TensorboardLauncher(directory_path):
    """Accepted for call compatibility; returns None (no browser / display here)."""
    return None
