"""Buffer contract (jorldy/core/buffer/base.py:5-56): store(list[dict]), sample(...), size,
stack_transition.  Transitions are dicts of arrays whose leading dimension is the batch of
envs that produced them (1 in the reference's per-actor loop)."""
from abc import ABC, abstractmethod

import numpy as np


class BaseBuffer(ABC):
    def __init__(self):
        self.first_store = True

    def check_dim(self, transition):
        print("########################################")
        print("You should check dimension of transition")
        for key, val in transition.items():
            if isinstance(val, (list, tuple)):
                for i, v in enumerate(val):
                    print(f"{key}{i}: {np.shape(v)}")
            else:
                print(f"{key}: {np.shape(val)}")
        print("########################################")
        self.first_store = False

    @abstractmethod
    def store(self, transitions):
        ...

    @abstractmethod
    def sample(self, batch_size):
        ...

    def stack_transition(self, batch):
        """list of transitions -> dict of arrays with the batch on axis 0.  A value that is a list /
        tuple is a multimodal observation and is stacked per modality (base.py:45-52)."""
        out = {}
        first = batch[0]
        for key, val in first.items():
            if isinstance(val, (list, tuple)):
                out[key] = [np.concatenate([np.asarray(b[key][i]) for b in batch], axis=0) for i in range(len(val))]
            else:
                out[key] = np.concatenate([np.asarray(b[key]) for b in batch], axis=0)
        return out
