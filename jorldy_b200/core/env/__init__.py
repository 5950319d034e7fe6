"""Env factory with the reference's registry keys (jorldy/core/env/__init__.py:41-64)."""
from collections import OrderedDict

from .classic import Cartpole, Pendulum, MountainCar

env_dict = OrderedDict(cartpole=Cartpole, mountain_car=MountainCar, pendulum=Pendulum)


def register(name, cls):
    env_dict[name] = cls


class Env:
    def __new__(cls, name, *args, **kwargs):
        if type(name) != str:
            print("### name variable must be string! ###")
            raise Exception
        name = name.lower()
        if name not in env_dict.keys():
            print(f"### can use only follows {[opt for opt in env_dict.keys()]}")
            raise Exception
        return env_dict[name](*args, **kwargs)
