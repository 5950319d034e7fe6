from .rollout_buffer import RolloutBuffer, DeviceRollout  # noqa: F401
