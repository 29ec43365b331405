"""Timing of eval_rollout (EvalConfig defaults: 64 envs x 1000 steps) on the configs[1] network:
one launch of the persistent eval kernel vs the per-step path (profiling script, not a test)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nnx_ppo_b200 import Rngs, prng                                   # noqa: E402
from nnx_ppo_b200.algorithms import rollout                            # noqa: E402
from nnx_ppo_b200.envs import SyntheticEnv                             # noqa: E402
from nnx_ppo_b200.networks.factories import make_mlp_actor_critic      # noqa: E402


class PerStep:
    fused_rollout = False

    def __init__(self, env):
        self.reset, self.step = env.reset, env.step


def main():
    env = SyntheticEnv(64, 8, max_len=64)
    nets = make_mlp_actor_critic(64, 8, [64] * 4, [256] * 2, Rngs(0))
    nets.eval()
    for name, e, n_envs in (("fused", env, 64), ("fused", env, 4096), ("per-step", PerStep(env), 64)):
        rollout.eval_rollout(e, nets, n_envs, 1000, prng.key(1), (0, 50, 100))   # warm-up (incl. torch.quantile's first call)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = rollout.eval_rollout(e, nets, n_envs, 1000, prng.key(1), (0, 50, 100))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{name:9s} n_envs={n_envs:5d} L=1000: {dt * 1e3:8.2f} ms  lifespan p50 {m['lifespan/p50']:.1f} "
              f"reward p50 {m['episode_reward/p50']:.4f}", flush=True)


if __name__ == "__main__":
    main()
