"""Synthetic benchmark environment (ORACLE side — test infrastructure).

The reference ships no env for BASELINE.json configs 2/5 ("synthetic dummy env obs=64 act=8"), so
this env is builder-defined.  It is modelled on the reference's ``MockEnv``
(``test_dummies/mock_env.py:25-63``: normal-noise obs on reset) wrapped in ``EpisodeWrapper``
(``wrappers/episode_wrapper.py:12-32``: ``step_counter`` initialised with
``randint(rng, 0, max_len // 2)``, ``truncated |= counter >= max_len``, ``done |= truncated``).

Single-env semantics (the library vmaps it, ``rollout.py:21,39``):

reset(key):  k_base, k_cnt = split(key);  obs = normal(k_base, (O,));
             step_counter = randint(k_cnt, (), 0, max_len // 2);  term_state = k_cnt[0] ^ k_cnt[1]
step(s, a):  obs' = tanh(obs @ Wo + a @ Wa);  reward = -mean(obs'^2);  counter' = counter + 1;
             term_state' = term_state * 1664525 + 1013904223 (mod 2^32);
             terminated = (term_state' >> 16) < term_thresh16;  truncated = counter' >= max_len;
             done = terminated | truncated

Termination is driven by an integer stream on purpose: reset masks and episode bookkeeping are
then bit-exact between the oracle and the CUDA path regardless of float rounding.
"""

from __future__ import annotations

import dataclasses

import numpy as np

from . import prng

F = np.float32


def make_env_weights(obs_dim: int, act_dim: int, seed: int = 0):
    g = np.random.default_rng(seed)
    Wo = (g.standard_normal((obs_dim, obs_dim)) * (0.5 / np.sqrt(obs_dim))).astype(F)
    Wa = (g.standard_normal((act_dim, obs_dim)) * (0.5 / np.sqrt(act_dim))).astype(F)
    return Wo, Wa


@dataclasses.dataclass
class EnvState:
    obs: np.ndarray          # [B, O] f32
    step_counter: np.ndarray  # [B] i32
    term_state: np.ndarray   # [B] u32
    reward: np.ndarray       # [B] f32
    done: np.ndarray         # [B] bool
    truncated: np.ndarray    # [B] bool


class SyntheticEnv:
    def __init__(self, obs_dim=64, act_dim=8, max_len=64, term_thresh16=512, seed=0):
        self.obs_dim, self.act_dim = obs_dim, act_dim
        self.max_len, self.term_thresh16 = int(max_len), int(term_thresh16)
        self.Wo, self.Wa = make_env_weights(obs_dim, act_dim, seed)
        self.observation_size, self.action_size = obs_dim, act_dim

    def reset(self, keys: np.ndarray) -> EnvState:
        """Batched reset: ``keys`` is [B, 2] uint32 (one key per env)."""
        B = keys.shape[0]
        obs = np.empty((B, self.obs_dim), F)
        cnt = np.empty(B, np.int32)
        term = np.empty(B, np.uint32)
        # vectorised over envs: split(key) = threefry(key, (0, j)), j = 0, 1
        b0, b1 = prng.threefry2x32(keys[:, 0], keys[:, 1], np.uint32(0), np.uint32(0))
        c0, c1 = prng.threefry2x32(keys[:, 0], keys[:, 1], np.uint32(0), np.uint32(1))
        idx = np.arange(self.obs_dim, dtype=np.uint32)[None, :]
        o0, o1 = prng.threefry2x32(b0[:, None], b1[:, None], np.uint32(0), idx)
        obs[:] = prng.bits_to_normal(o0 ^ o1)
        for i in range(B):  # randint is scalar-shaped per env; loop keeps it a literal restatement
            cnt[i] = prng.randint(np.array([c0[i], c1[i]], np.uint32), (), 0, self.max_len // 2)
        term[:] = c0 ^ c1
        return EnvState(obs, cnt, term, np.zeros(B, F), np.zeros(B, bool), np.zeros(B, bool))

    def reset_fast(self, keys: np.ndarray) -> EnvState:
        """Same as ``reset`` with the per-env randint vectorised (used for large B)."""
        B = keys.shape[0]
        b0, b1 = prng.threefry2x32(keys[:, 0], keys[:, 1], np.uint32(0), np.uint32(0))
        c0, c1 = prng.threefry2x32(keys[:, 0], keys[:, 1], np.uint32(0), np.uint32(1))
        idx = np.arange(self.obs_dim, dtype=np.uint32)[None, :]
        o0, o1 = prng.threefry2x32(b0[:, None], b1[:, None], np.uint32(0), idx)
        obs = prng.bits_to_normal(o0 ^ o1)
        # randint(k_cnt, (), 0, span): k1,k2 = split(k_cnt); bits at element 0 of each
        k10, k11 = prng.threefry2x32(c0, c1, np.uint32(0), np.uint32(0))
        k20, k21 = prng.threefry2x32(c0, c1, np.uint32(0), np.uint32(1))
        h0, h1 = prng.threefry2x32(k10, k11, np.uint32(0), np.uint32(0))
        l0, l1 = prng.threefry2x32(k20, k21, np.uint32(0), np.uint32(0))
        span = max(self.max_len // 2, 1)
        mult = (2 ** 16) % span
        mult = (mult * mult) % span
        hb, lb = (h0 ^ h1), (l0 ^ l1)
        with np.errstate(over="ignore"):
            off = (hb % np.uint32(span)) * np.uint32(mult) + (lb % np.uint32(span))
        cnt = (off % np.uint32(span)).astype(np.int32)
        term = (c0 ^ c1).astype(np.uint32)
        return EnvState(obs.astype(F), cnt, term, np.zeros(B, F), np.zeros(B, bool),
                        np.zeros(B, bool))

    def step(self, s: EnvState, action: np.ndarray) -> EnvState:
        pre = (s.obs @ self.Wo + action @ self.Wa).astype(F)
        obs = np.tanh(pre).astype(F)
        reward = (-np.mean(obs * obs, axis=1, dtype=F)).astype(F)
        cnt = (s.step_counter + 1).astype(np.int32)
        with np.errstate(over="ignore"):
            term = (s.term_state * np.uint32(1664525) + np.uint32(1013904223)).astype(np.uint32)
        terminated = (term >> np.uint32(16)) < np.uint32(self.term_thresh16)
        truncated = cnt >= self.max_len
        done = terminated | truncated
        return EnvState(obs, cnt, term, reward, done, truncated)
