from . import config, ppo, rollout, types  # noqa: F401
