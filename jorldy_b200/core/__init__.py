from .agent import Agent  # noqa: F401
from .env import Env  # noqa: F401
from .network import Network  # noqa: F401
from .optimizer import Optimizer  # noqa: F401
from . import buffer  # noqa: F401
