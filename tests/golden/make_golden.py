"""Regenerates tests/golden/*.npz.

The reference (JAX) cannot be imported in this image, so golden vectors are limited to what the
reference's own tests define in plain NumPy: the ``test_gae`` recipe (ppo_test.py:229-264), i.e.
seeded inputs plus the inline float64 loop that the reference uses as ITS oracle.  Threefry /
JAX-split known answers are literal constants in tests/test_oracle_prng.py.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def gae_kat():
    np.random.seed(23)
    T, B, gamma, lambda_ = 100, 512, 0.8, 0.95
    rewards = np.random.normal(size=(T, B))
    values = np.random.normal(size=(T + 1, B))
    done = np.random.choice([True, False], size=(T, B), p=[0.01, 0.99])
    truncation = np.random.choice([True, False], size=(T, B))
    truncation = np.logical_and(done, truncation)
    advantages = np.full((T, B), np.nan)
    for t in reversed(range(T)):
        next_values = values[t + 1, :].copy()
        next_values[done[t, :]] = 0.0
        advantages[t, :] = rewards[t, :] + gamma * next_values - values[t, :]
        advantages[t, truncation[t, :]] = 0.0
        if t < T - 1:
            advantages[t, :] += gamma * lambda_ * advantages[t + 1, :] * (1 - done[t, :])
    np.savez_compressed(os.path.join(HERE, "gae_kat.npz"), adv_f64=advantages,
                        done_bits=np.packbits(done), trunc_bits=np.packbits(truncation),
                        rewards_f32=rewards.astype(np.float32), values_f32=values.astype(np.float32))


if __name__ == "__main__":
    gae_kat()
