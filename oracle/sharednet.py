"""TEST ORACLE (not product code; only tests/, smoke() and bench.py's cpu legs may import this).

Shared-trunk actor-critic: the network of the reference's composition tutorial
(docs/tutorials/02_composition.rst, "shared body") -

    Sequential([Normalizer, trunk MLP, PPOAdapter(action=Sequential([actor head, sampler]), value=critic head)])

``Sequential.__call__`` (nnx_ppo/networks/containers.py:18-39) feeds the trunk's output to the adapter, whose
two ports (ppo_adapter.py:51-78) both read it; jax.grad therefore sums the two paths' contributions for every
trunk parameter.  Restated on top of oracle/nets.py: both chains START with the trunk's layers (same arrays),
the flat parameter order is [trunk, actor head, critic head], and ``fold_grads`` adds the critic chain's
trunk gradient onto the actor chain's.

Parity unpinned against a live JAX run (no JAX in this image); pinned by tests/test_shared_trunk.py against
float64 torch autograd of the same composite function.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import prng
from .nets import ACT_IDS, ActorCritic, Chain, F


@dataclasses.dataclass
class SharedTrunkActorCritic(ActorCritic):
    n_trunk: int = 0

    def _logical(self):
        """(chain, layer) pairs in logical order: trunk (actor's arrays), actor head, critic head."""
        nt = self.n_trunk
        return ([(self.actor, l) for l in range(self.actor.n_layers)] +
                [(self.critic, l) for l in range(nt, self.critic.n_layers)])

    def flat_params(self) -> np.ndarray:
        parts = []
        for ch, l in self._logical():
            parts += [ch.W[l].ravel(), ch.b[l].ravel()]
        return np.concatenate(parts).astype(F)

    def set_flat_params(self, p: np.ndarray) -> None:
        o = 0
        for ch, l in self._logical():
            n = ch.W[l].size
            ch.W[l] = p[o:o + n].reshape(ch.W[l].shape).astype(F).copy(); o += n
            n = ch.b[l].size
            ch.b[l] = p[o:o + n].astype(F).copy(); o += n
        assert o == p.size
        for l in range(self.n_trunk):                      # one trunk, seen by both chains
            self.critic.W[l], self.critic.b[l] = self.actor.W[l], self.actor.b[l]

    def fold_grads(self, grads_full: np.ndarray) -> np.ndarray:
        """[actor chain (trunk + head), critic chain (trunk + head)] -> logical order, the two trunk
        contributions added (float32, actor path + critic path)."""
        sizes = lambda ch: [(ch.W[l].size, ch.b[l].size) for l in range(ch.n_layers)]
        na = sum(w + b for w, b in sizes(self.actor))
        nt = sum(w + b for w, b in sizes(self.actor)[:self.n_trunk])
        ga, gc = grads_full[:na], grads_full[na:]
        trunk = (ga[:nt].astype(F) + gc[:nt].astype(F)).astype(F)
        return np.concatenate([trunk, ga[nt:], gc[nt:]]).astype(F)


def make_shared_trunk_actor_critic(obs_size, action_size, trunk_sizes, actor_head_sizes, critic_head_sizes,
                                   seed: int = 0, activation="relu", normalize_obs=True, entropy_weight=1e-2,
                                   min_std=1e-1, std_scale=1.0) -> SharedTrunkActorCritic:
    """Key-draw order: trunk layers, actor head layers, critic head layers (kernel then bias key each, as
    nnx.Linear does - factories.py:116-137), the sampler continuing on the same stream."""
    rngs = prng.Rngs(seed)
    act = ACT_IDS[activation] if isinstance(activation, str) else int(activation)

    def layers(sizes):
        Ws, bs = [], []
        for din, dout in zip(sizes[:-1], sizes[1:]):
            Ws.append(prng.variance_scaling_uniform(rngs.params(), din, dout, 1.0))
            rngs.params()
            bs.append(np.zeros(dout, F))
        return Ws, bs

    tw, tb = layers([obs_size] + list(trunk_sizes))
    aw, ab = layers([trunk_sizes[-1]] + list(actor_head_sizes) + [2 * action_size])
    cw, cb = layers([trunk_sizes[-1]] + list(critic_head_sizes) + [1])
    dims_t = [obs_size] + list(trunk_sizes)
    actor = Chain(dims_t + list(actor_head_sizes) + [2 * action_size], act, tw + aw, tb + ab)
    critic = Chain(dims_t + list(critic_head_sizes) + [1], act, list(tw) + cw, list(tb) + cb)
    return SharedTrunkActorCritic(obs_size, action_size, actor, critic, normalize_obs, entropy_weight, min_std,
                                  std_scale, np.zeros(obs_size, F), np.zeros(obs_size, F), F(0.0), rngs.key.copy(),
                                  rngs.count, n_trunk=len(trunk_sizes))
