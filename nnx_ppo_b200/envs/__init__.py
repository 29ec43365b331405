from .synthetic import SyntheticEnv, SyntheticEnvState  # noqa: F401
