"""Agent factory with the reference's registry keys (jorldy/core/agent/__init__.py:32-42)."""
from collections import OrderedDict

from .ppo import PPO
from .dqn import DQN, Double, Dueling, Multistep, PER, Noisy, C51, Rainbow, ApeX
from .ddpg import DDPG, TD3, SAC

agent_dict = OrderedDict(sorted(dict(ape_x=ApeX, c51=C51, ddpg=DDPG, double=Double, dqn=DQN, dueling=Dueling,
                                     multistep=Multistep, noisy=Noisy, per=PER, ppo=PPO, rainbow=Rainbow, sac=SAC,
                                     td3=TD3).items()))


def register(name, cls):
    agent_dict[name] = cls


class Agent:
    def __new__(cls, name, *args, **kwargs):
        if type(name) != str:
            print("### name variable must be string! ###")
            raise Exception
        name = name.lower()
        if name not in agent_dict.keys():
            print(f"### can use only follows {[opt for opt in agent_dict.keys()]}")
            raise Exception
        return agent_dict[name](*args, **kwargs)
