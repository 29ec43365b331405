from . import episode_wrapper  # noqa: F401
