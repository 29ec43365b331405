"""Device-resident synthetic benchmark env (BASELINE.json configs 2 / 5: obs=64, act=8).

Builder-defined (the reference ships no such env), modelled on the reference's
``MockEnv`` (test_dummies/mock_env.py:25-63) behind ``EpisodeWrapper``
(wrappers/episode_wrapper.py:12-32).  Definition (single env; see oracle/env.py for the CPU
restatement the parity tests check against):

  reset(key):  k_base, k_cnt = split(key); obs = normal(k_base, (O,));
               step_counter = randint(k_cnt, (), 0, max_len // 2); term_state = k_cnt[0] ^ k_cnt[1]
  step(s, a):  obs' = tanh(obs @ Wo + a @ Wa); reward = -mean(obs'^2); counter += 1;
               term_state = term_state * 1664525 + 1013904223; terminated = (term_state >> 16) < thr;
               truncated = counter >= max_len; done = terminated | truncated

The env step is fused into the persistent rollout kernel (csrc/rollout.cu), so a rollout is one
launch.  Termination uses an integer stream so reset masks are bit-exact vs the oracle.
"""
from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np

from .. import _lib


def make_env_weights(obs_dim: int, act_dim: int, seed: int = 0):
    g = np.random.default_rng(seed)
    Wo = (g.standard_normal((obs_dim, obs_dim)) * (0.5 / np.sqrt(obs_dim))).astype(np.float32)
    Wa = (g.standard_normal((act_dim, obs_dim)) * (0.5 / np.sqrt(act_dim))).astype(np.float32)
    return Wo, Wa


@dataclasses.dataclass
class SyntheticEnvState:
    obs: Any             # [B, O] float32 (cuda)
    step_counter: Any    # [B] int32
    term_state: Any      # [B] int32 holding uint32 bits
    reward: Any = None
    done: Any = None
    info: dict = dataclasses.field(default_factory=dict)
    metrics: dict = dataclasses.field(default_factory=dict)


class SyntheticEnv:
    fused_rollout = True

    def __init__(self, obs_dim: int = 64, act_dim: int = 8, max_len: int = 64,
                 term_thresh16: int = 512, seed: int = 0):
        self.obs_dim, self.act_dim = int(obs_dim), int(act_dim)
        self.max_len, self.term_thresh16 = int(max_len), int(term_thresh16)
        self.Wo, self.Wa = make_env_weights(obs_dim, act_dim, seed)
        self._dev = None
        self._w = None

    @property
    def observation_size(self):
        return self.obs_dim

    @property
    def action_size(self):
        return self.act_dim

    def c_struct(self, device) -> _lib.SynthEnv:
        import torch
        if self._w is None or self._dev != device:
            w = np.concatenate([self.Wo, self.Wa], axis=0)     # [(O + A), O], Wa right after Wo
            self._w = torch.from_numpy(np.ascontiguousarray(w)).to(device)
            self._dev = device
        s = _lib.SynthEnv()
        s.obs_dim, s.act_dim, s.max_len, s.term_thresh16 = (self.obs_dim, self.act_dim, self.max_len,
                                                            self.term_thresh16)
        s.Wo = self._w.data_ptr()
        s.Wa = self._w.data_ptr() + 4 * self.obs_dim * self.obs_dim
        return s

    def reset_from_split(self, key, n_envs: int, device) -> SyntheticEnvState:
        """vmap(env.reset)(jax.random.split(key, n_envs)) — ppo.py:548-549, on the device."""
        import torch
        lib = _lib.load()
        s = _lib.current_stream()
        keys = torch.empty(n_envs, 2, dtype=torch.int32, device=device)
        _lib.check(lib.b200ppo_synth_init_keys(s, key[0], key[1], n_envs, _lib.ptr(keys)), "synth_init_keys")
        return self.reset(keys)

    def reset(self, keys) -> SyntheticEnvState:
        """Batched reset from per-env keys [B, 2] (uint32 bits in an int32 tensor)."""
        import torch
        lib = _lib.load()
        _lib.require_cuda(keys)
        B = keys.shape[0]
        dev = keys.device
        obs = torch.empty(B, self.obs_dim, dtype=torch.float32, device=dev)
        cnt = torch.empty(B, dtype=torch.int32, device=dev)
        term = torch.empty(B, dtype=torch.int32, device=dev)
        es = self.c_struct(dev)
        _lib.check(lib.b200ppo_synth_reset(_lib.current_stream(), es, _lib.ptr(keys), B, _lib.ptr(obs),
                                           _lib.ptr(cnt), _lib.ptr(term)), "synth_reset")
        return SyntheticEnvState(obs, cnt, term, torch.zeros(B, device=dev),
                                 torch.zeros(B, device=dev), {"truncated": torch.zeros(B, dtype=torch.bool, device=dev)}, {})

    def step(self, state: SyntheticEnvState, action) -> SyntheticEnvState:
        """Single batched env step in user-land torch ops (eval / generic path only; training
        rollouts use the fused kernel, which implements the same definition)."""
        import torch
        self.c_struct(action.device)
        O = self.obs_dim
        pre = state.obs @ self._w[:O] + action @ self._w[O:]
        obs = torch.tanh(pre)
        reward = -(obs * obs).mean(dim=1)
        cnt = state.step_counter + 1
        term = (state.term_state.to(torch.int64) & 0xFFFFFFFF) * 1664525 + 1013904223
        term = term & 0xFFFFFFFF
        terminated = (term >> 16) < self.term_thresh16
        truncated = cnt >= self.max_len
        done = terminated | truncated
        term32 = torch.where(term >= 2 ** 31, term - 2 ** 32, term).to(torch.int32)
        return SyntheticEnvState(obs, cnt.to(torch.int32), term32, reward, done.float(),
                                 {"truncated": truncated}, {})
