from . import episode_wrapper, reward_scaling_wrapper  # noqa: F401
from .episode_wrapper import EpisodeWrapper  # noqa: F401
from .reward_scaling_wrapper import RewardScalingWrapper  # noqa: F401
