"""ConfigurationSaver / TensorboardLauncher equivalents (reference: flex_gym/helper/raisim_gym_helper.py:6-28)."""
import datetime
import os
import shutil


class ConfigurationSaver:
    def __init__(self, log_dir, save_items):
        self._data_dir = log_dir + '/' + datetime.datetime.now().strftime('%Y-%m-%d-%H-%M-%S')
        os.makedirs(self._data_dir, exist_ok=True)
        for save_item in save_items or []:
            if save_item and os.path.exists(save_item):
                shutil.copyfile(save_item, self._data_dir + '/' + os.path.basename(save_item))

    @property
    def data_dir(self):
        return self._data_dir


def TensorboardLauncher(directory_path):   # no display / browser on a GPU box: accepted no-op
    return None
