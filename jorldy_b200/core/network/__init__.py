"""Network factory with the reference's registry keys (jorldy/core/network/__init__.py:30-40:
snake_case(ClassName))."""
from collections import OrderedDict

from .policy_value import DiscretePolicyValue, ContinuousPolicyValue, DiscreteQ_Network
from .dueling import Dueling
from .noisy import Noisy, Rainbow
from .policy import ContinuousPolicy, DeterministicPolicy
from .q_network import ContinuousQ_Network

network_dict = OrderedDict(
    continuous_policy=ContinuousPolicy,
    continuous_policy_value=ContinuousPolicyValue,
    continuous_q_network=ContinuousQ_Network,
    deterministic_policy=DeterministicPolicy,
    discrete_policy_value=DiscretePolicyValue,
    discrete_q_network=DiscreteQ_Network,
    dueling=Dueling,
    noisy=Noisy,
    rainbow=Rainbow,
)


def register(name, cls):
    network_dict[name] = cls


class Network:
    def __new__(cls, name, *args, **kwargs):
        if type(name) != str:
            print("### name variable must be string! ###")
            raise Exception
        name = name.lower()
        if name not in network_dict.keys():
            print(f"### can use only follows {[opt for opt in network_dict.keys()]}")
            raise Exception
        return network_dict[name](*args, **kwargs)
