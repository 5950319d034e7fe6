"""Env contract (jorldy/core/env/base.py:4-38): reset() -> state, step(action) -> (next_state,
reward, done), close(), recordable(); attrs state_size / action_size / action_type / score."""
from abc import ABC, abstractmethod


class BaseEnv(ABC):
    @abstractmethod
    def reset(self):
        ...

    @abstractmethod
    def step(self, action):
        ...

    @abstractmethod
    def close(self):
        ...

    def recordable(self):
        return False
