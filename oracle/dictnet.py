"""CPU restatement of the dict-observation path (BASELINE configs[3]; reference containers.py:55-110
``Concat``): every key of the observation dict goes through its own Dense encoder stack (activation
after every layer), the encoder outputs are concatenated in key order and feed a trunk MLP whose last
layer is linear.  Actor and critic each own such a tower; the Normalizer keeps per-feature statistics,
which for a dict of vectors is the same arithmetic as one Normalizer over the concatenation.

TEST INFRASTRUCTURE ONLY (oracle/__init__.py).  Parity unpinned like the rest (the reference cannot be
run here); the analytic gradient is checked against float64 autograd in tests/test_oracle_dictnet.py.
``EncChain`` exposes the same interface as ``nets.Chain`` (``W``, ``b``, ``forward(x, keep)``) plus
``backward``, so ``oracle.ppo`` (rollout, loss, Adam, ppo_step) runs unchanged on it.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from . import prng
from .nets import ACT_IDS, ActorCritic, F, act_fwd, act_grad


@dataclasses.dataclass
class EncChain:
    """Parameter order (= flat order): for every key in order, its encoder layers; then the trunk."""
    act: int
    W: list
    b: list
    enc: list            # per key: (col0, col1, [layer indices])
    trunk: list          # layer indices; the last one is linear

    @property
    def n_layers(self) -> int:
        return len(self.W)

    def forward(self, x: np.ndarray, keep: bool = False):
        zs = [None] * len(self.W)
        outs = []
        for c0, c1, idx in self.enc:
            h = x[:, c0:c1]
            for li in idx:
                z = (h @ self.W[li] + self.b[li]).astype(F)
                zs[li] = z
                h = act_fwd(z, self.act)
            outs.append(h)
        h = np.concatenate(outs, axis=1)
        for n, li in enumerate(self.trunk):
            z = (h @ self.W[li] + self.b[li]).astype(F)
            zs[li] = z
            h = act_fwd(z, self.act) if n < len(self.trunk) - 1 else z
        return h, (zs if keep else [])

    def backward(self, x: np.ndarray, zs, d_out: np.ndarray):
        L = len(self.W)
        dWs, dbs = [None] * L, [None] * L
        enc_out = [act_fwd(zs[idx[-1]], self.act) for _, _, idx in self.enc]
        cat = np.concatenate(enc_out, axis=1)
        d = d_out
        for n in reversed(range(len(self.trunk))):
            li = self.trunk[n]
            hin = cat if n == 0 else act_fwd(zs[self.trunk[n - 1]], self.act)
            dWs[li] = (hin.T @ d).astype(F)
            dbs[li] = d.sum(axis=0, dtype=F)
            d = (d @ self.W[li].T).astype(F)
            if n > 0:
                d = (d * act_grad(zs[self.trunk[n - 1]], self.act)).astype(F)
        o = 0
        for (c0, c1, idx), eo in zip(self.enc, enc_out):
            w = eo.shape[1]
            dk = d[:, o:o + w]
            o += w
            for n in reversed(range(len(idx))):
                li = idx[n]
                dk = (dk * act_grad(zs[li], self.act)).astype(F)
                hin = x[:, c0:c1] if n == 0 else act_fwd(zs[idx[n - 1]], self.act)
                dWs[li] = (hin.T @ dk).astype(F)
                dbs[li] = dk.sum(axis=0, dtype=F)
                dk = (dk @ self.W[li].T).astype(F)
        return dWs, dbs


def make_dict_actor_critic(obs_sizes: dict, action_size: int, encoder_hidden: dict, actor_trunk_sizes,
                           critic_trunk_sizes, seed: int = 0, activation="relu",
                           normalize_obs=True) -> ActorCritic:
    """Key-draw order of the shared Rngs stream: actor tower (encoders in key order, then trunk), critic
    tower, then the sampler — two counts per Linear, as in make_mlp_actor_critic."""
    rngs = prng.Rngs(seed)
    act = ACT_IDS[activation]

    def tower(trunk, out):
        Ws, bs, enc, col = [], [], [], 0

        def lin(din, dout):
            Ws.append(prng.variance_scaling_uniform(rngs.params(), din, dout, 1.0))
            rngs.params()
            bs.append(np.zeros(dout, F))
            return len(Ws) - 1
        for k, n in obs_sizes.items():
            sizes = [n] + list(encoder_hidden[k])
            enc.append((col, col + n, [lin(a, b) for a, b in zip(sizes[:-1], sizes[1:])]))
            col += n
        width = sum(encoder_hidden[k][-1] for k in obs_sizes)
        sizes = [width] + list(trunk) + [out]
        tr = [lin(a, b) for a, b in zip(sizes[:-1], sizes[1:])]
        return EncChain(act, Ws, bs, enc, tr)

    actor = tower(actor_trunk_sizes, 2 * action_size)
    critic = tower(critic_trunk_sizes, 1)
    O = sum(obs_sizes.values())
    return ActorCritic(O, action_size, actor, critic, normalize_obs, 1e-2, 1e-1, 1.0, np.zeros(O, F),
                       np.zeros(O, F), F(0.0), rngs.key.copy(), rngs.count)
