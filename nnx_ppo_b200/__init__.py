"""nnx_ppo_b200 — B200-native (sm_100a) implementation of the PPO hot path of emiwar/nnx-ppo
behind the reference's own API (train_ppo / TrainConfig / PPOConfig / make_mlp_actor_critic).

Host code is Python; all arithmetic of the path runs in hand-written CUDA kernels reached through
the C ABI of libb200ppo.so (include/b200ppo.h).  There is no CPU or eager fallback.
"""
from . import prng  # noqa: F401
from .prng import Rngs  # noqa: F401

__all__ = ["Rngs", "prng", "algorithms", "networks", "envs", "wrappers"]
__version__ = "0.1.0"
