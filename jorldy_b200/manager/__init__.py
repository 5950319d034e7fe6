from .config_manager import ConfigManager, CustomDict, type_cast  # noqa: F401
from .metric_manager import MetricManager, LogManager  # noqa: F401
