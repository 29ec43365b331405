"""EpisodeWrapper — mirrors nnx_ppo/wrappers/episode_wrapper.py:7-39 for batched torch envs.

(The synthetic benchmark env has these semantics fused into its rollout kernel.)"""
from __future__ import annotations

import dataclasses

from .. import _lib


class EpisodeWrapper:
    def __init__(self, env, max_len: int):
        self.env = env
        self.max_len = max_len

    def step(self, state, action):
        import torch
        nxt = self.env.step(state, action)
        info = dict(nxt.info)
        info["step_counter"] = state.info["step_counter"] + 1
        truncated = torch.logical_or(info.get("truncated", torch.zeros_like(info["step_counter"], dtype=torch.bool)),
                                     info["step_counter"] >= self.max_len)
        info["truncated"] = truncated
        return dataclasses.replace(nxt, info=info,
                                   done=torch.logical_or(nxt.done.bool(), truncated).float())

    def reset(self, keys):
        """keys: int32 [B, 2].  base_rng, step_counter_rng = split(rng) per env
        (episode_wrapper.py:25); the counter draw is jax.random.randint(rng, (), 0, max_len // 2)."""
        import torch
        from ..algorithms.rollout import split_keys_device  # noqa: F401
        lib = _lib.load()
        B = keys.shape[0]
        # per-env split and randint run on the device through the synthetic-env reset kernel,
        # which implements exactly this wrapper's key flow
        obs_dummy = torch.empty(B, 1, device=keys.device)
        cnt = torch.empty(B, dtype=torch.int32, device=keys.device)
        term = torch.empty(B, dtype=torch.int32, device=keys.device)
        es = _lib.SynthEnv()
        es.obs_dim, es.act_dim, es.max_len, es.term_thresh16 = 1, 1, self.max_len, 0
        _lib.check(lib.b200ppo_synth_reset(_lib.current_stream(), es, _lib.ptr(keys.contiguous()), B,
                                           _lib.ptr(obs_dummy), _lib.ptr(cnt), _lib.ptr(term)), "episode reset")
        base_keys = _split0(keys)
        st = self.env.reset(base_keys)
        info = dict(st.info)
        info["step_counter"] = cnt
        info["truncated"] = torch.zeros(B, dtype=torch.bool, device=keys.device)
        return dataclasses.replace(st, info=info)

    @property
    def observation_size(self):
        return self.env.observation_size

    @property
    def action_size(self):
        return self.env.action_size


def _split0(keys):
    """Element 0 of jax.random.split(key) for every row of an int32 [B, 2] key tensor, on the device
    (`b200ppo_split_rows`): the generic rollout calls `env.reset` for all envs on every step like the
    reference (rollout.py:39), so this must not touch the host."""
    import torch
    lib = _lib.load()
    keys = keys.contiguous()
    out = torch.empty_like(keys)
    _lib.check(lib.b200ppo_split_rows(_lib.current_stream(), _lib.ptr(keys), keys.shape[0], 0, _lib.ptr(out)), "split_rows")
    return out
