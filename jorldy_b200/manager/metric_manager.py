"""Running-mean metric accumulation + scalar logging: the minimal observer side of
jorldy/manager/metric_manager.py:9-21 and log_manager.py:21-35 (TensorBoard/GIF/eval are out of scope,
SURVEY.md §2 row 13; TensorBoard scalars are written when the package is importable)."""
import os
import time
from collections import defaultdict

import numpy as np


class MetricManager:
    def __init__(self):
        self.metrics = defaultdict(list)

    def append(self, result):
        for key, value in result.items():
            self.metrics[key].append(value)

    def get_statistics(self, mode="mean"):
        ret = {}
        for key, values in self.metrics.items():
            if values:
                ret[key] = getattr(np, mode)(values)
                if isinstance(ret[key], (float, np.floating)):
                    ret[key] = round(float(ret[key]), 4)
        self.metrics.clear()
        return ret


class LogManager:
    def __init__(self, env, id, experiment=None):
        self.id = id
        now = time.strftime("%Y%m%d%H%M%S")
        self.path = f"./logs/{experiment}/{env}/{id}/{now}/" if experiment else f"./logs/{env}/{id}/{now}/"
        os.makedirs(self.path, exist_ok=True)
        self.stamp = time.time()
        self.writer = None
        try:
            from torch.utils.tensorboard import SummaryWriter
            self.writer = SummaryWriter(self.path)
        except Exception:
            pass

    def write(self, scalar_dict, step):
        if self.writer is None:
            return
        for key, value in scalar_dict.items():
            self.writer.add_scalar(f"{self.id}/{key}", value, step)
            self.writer.add_scalar("all/" + key, value, step)
            if "score" in key:
                self.writer.add_scalar(f"{self.id}/{key}_per_time", value, time.time() - self.stamp)
