"""NumPy float32 restatement of the reference network forward path (ORACLE — test infrastructure).

Follows: ``networks/factories.py:72-146`` (make_mlp_actor_critic topology and init),
``networks/feedforward.py:42-51`` (Dense), ``networks/normalizer.py:63-136`` (Normalizer forward
and Welford merge), ``networks/sampling_layers.py:82-147`` (NormalTanhSampler),
``networks/adapter.py:75-117`` (PPOAdapter: actor + critic on the same input, value squeeze).
flax.nnx ``Linear`` (kernel ``[in, out]``, ``x @ W + b``, two ``rngs.params()`` draws per layer)
and ``variance_scaling`` are third-party and restated from their published semantics (unpinned).
"""

from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from . import prng

F = np.float32
ACT_NONE, ACT_RELU, ACT_SWISH, ACT_TANH = 0, 1, 2, 3
ACT_IDS = {"none": ACT_NONE, "relu": ACT_RELU, "swish": ACT_SWISH, "tanh": ACT_TANH}
LOG2 = F(np.log(2.0))
HALF_LOG_2PI = F(0.5 * np.log(2.0 * np.pi))


def act_fwd(z: np.ndarray, act: int) -> np.ndarray:
    if act == ACT_RELU:
        return np.maximum(z, F(0))
    if act == ACT_TANH:
        return np.tanh(z).astype(F)
    if act == ACT_SWISH:
        return (z * sigmoid(z)).astype(F)
    return z


def act_grad(z: np.ndarray, act: int) -> np.ndarray:
    """d act(z) / dz evaluated from the pre-activation (jax.nn.relu: 0 at z == 0)."""
    if act == ACT_RELU:
        return (z > 0).astype(F)
    if act == ACT_TANH:
        h = np.tanh(z).astype(F)
        return (F(1) - h * h).astype(F)
    if act == ACT_SWISH:
        s = sigmoid(z)
        return (s * (F(1) + z * (F(1) - s))).astype(F)
    return np.ones_like(z)


def sigmoid(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, F)
    e = np.exp(-np.abs(x)).astype(F)
    return np.where(x >= 0, F(1) / (F(1) + e), e / (F(1) + e)).astype(F)


def softplus(x: np.ndarray) -> np.ndarray:
    """jax.nn.softplus = logaddexp(x, 0) = max(x, 0) + log1p(exp(-|x|))."""
    x = np.asarray(x, F)
    return (np.maximum(x, F(0)) + np.log1p(np.exp(-np.abs(x)).astype(F)).astype(F)).astype(F)


@dataclasses.dataclass
class Chain:
    """A stack of Dense layers: hidden layers use ``act``, the last layer is linear."""
    dims: list[int]
    act: int
    W: list[np.ndarray]
    b: list[np.ndarray]

    @property
    def n_layers(self) -> int:
        return len(self.W)

    def forward(self, x: np.ndarray, keep: bool = False):
        """Returns (output, [pre-activations z_1..z_L]) — ``feedforward.py:48-50``."""
        zs = []
        h = x
        for l in range(self.n_layers):
            z = (h @ self.W[l] + self.b[l]).astype(F)
            if keep:
                zs.append(z)
            h = act_fwd(z, self.act) if l < self.n_layers - 1 else z
        return h, zs


@dataclasses.dataclass
class ActorCritic:
    """``make_mlp_actor_critic`` (``factories.py:72-146``) as plain arrays."""
    obs_dim: int
    act_dim: int
    actor: Chain
    critic: Chain
    normalize: bool = True
    entropy_weight: float = 1e-2
    min_std: float = 1e-1
    std_scale: float = 1.0
    # Normalizer state (normalizer.py:52-61): float32 mean, M2, counter.
    mean: Optional[np.ndarray] = None
    M2: Optional[np.ndarray] = None
    counter: F = F(0.0)
    # nnx.Rngs default stream shared by the Linear inits and the sampler (factories.py:116-137)
    rng_key: Optional[np.ndarray] = None
    rng_count: int = 0

    # ---- parameter arena (same order as the CUDA plan: actor W0,b0,..., critic W0,b0,...) ----
    def flat_params(self) -> np.ndarray:
        parts = []
        for ch in (self.actor, self.critic):
            for W, b in zip(ch.W, ch.b):
                parts += [W.ravel(), b.ravel()]
        return np.concatenate(parts).astype(F)

    def set_flat_params(self, p: np.ndarray) -> None:
        o = 0
        for ch in (self.actor, self.critic):
            for l in range(ch.n_layers):
                n = ch.W[l].size
                ch.W[l] = p[o:o + n].reshape(ch.W[l].shape).astype(F).copy(); o += n
                n = ch.b[l].size
                ch.b[l] = p[o:o + n].astype(F).copy(); o += n
        assert o == p.size

    # ---- Normalizer forward (normalizer.py:63-96) ----
    def norm_std(self) -> np.ndarray:
        if not self.normalize:
            return np.ones(self.obs_dim, F)
        if self.counter > 0:
            return np.sqrt(np.maximum(self.M2 / self.counter, F(1e-6))).astype(F)
        return np.full(self.obs_dim, 10.0, F)

    def normalize_obs(self, x: np.ndarray) -> np.ndarray:
        if not self.normalize:
            return x.astype(F)
        return ((x - self.mean) / self.norm_std()).astype(F)

    # ---- Normalizer.update_statistics (normalizer.py:98-136) ----
    def update_statistics(self, obs_TBO: np.ndarray) -> None:
        if not self.normalize:
            return
        flat = obs_TBO.reshape(-1, obs_TBO.shape[-1]).astype(F)
        n = F(flat.shape[0])
        new_count = F(self.counter + n)
        frac = F(n / new_count)
        bm = flat.mean(axis=0, dtype=F)
        bM2 = np.square(flat - bm).sum(axis=0, dtype=F)
        delta = (bm - self.mean).astype(F)
        new_mean = (self.mean + delta * frac).astype(F)
        new_M2 = (self.M2 + bM2 + (delta * delta) * self.counter * n / new_count).astype(F)
        self.mean, self.M2, self.counter = new_mean, new_M2, new_count

    # ---- nnx.Rngs stream ----
    def next_key(self) -> np.ndarray:
        k = prng.fold_in(self.rng_key, self.rng_count)
        self.rng_count = (self.rng_count + 1) & 0xFFFFFFFF
        return k


def make_mlp_actor_critic(obs_size, action_size, actor_hidden_sizes, critic_hidden_sizes,
                          seed: int = 0, activation="relu", normalize_obs=True,
                          initializer_scale=1.0, entropy_weight=1e-2, min_std=1e-1,
                          std_scale=1.0) -> ActorCritic:
    """Restates ``factories.py:72-146`` with ``rngs = nnx.Rngs(seed)``: actor Linear layers are
    created first, then the critic's, each drawing kernel then bias keys from the shared stream;
    the sampler keeps using the same stream afterwards."""
    rngs = prng.Rngs(seed)
    act = ACT_IDS[activation] if isinstance(activation, str) else int(activation)

    def chain(sizes):
        Ws, bs = [], []
        for din, dout in zip(sizes[:-1], sizes[1:]):
            kk = rngs.params()
            Ws.append(prng.variance_scaling_uniform(kk, din, dout, initializer_scale))
            rngs.params()  # bias key: zeros initializer ignores it but the count advances
            bs.append(np.zeros(dout, F))
        return Chain(list(sizes), act, Ws, bs)

    actor = chain([obs_size] + list(actor_hidden_sizes) + [2 * action_size])
    critic = chain([obs_size] + list(critic_hidden_sizes) + [1])
    return ActorCritic(obs_size, action_size, actor, critic, normalize_obs, entropy_weight,
                       min_std, std_scale, np.zeros(obs_size, F), np.zeros(obs_size, F), F(0.0),
                       rngs.key.copy(), rngs.count)


# ------------------------------------------------------------------------------------------
# NormalTanhSampler (sampling_layers.py:82-147)
# ------------------------------------------------------------------------------------------

def sampler_std(rho: np.ndarray, min_std, std_scale) -> np.ndarray:
    return ((softplus(rho) + F(min_std)) * F(std_scale)).astype(F)


def log_det_jac(z: np.ndarray) -> np.ndarray:
    return (F(2.0) * (LOG2 - z - softplus(F(-2.0) * z))).astype(F)


def loglikelihood(z, mu, sigma) -> np.ndarray:
    """``_loglikelihood`` (sampling_layers.py:115-135)."""
    log_unnorm = F(-0.5) * np.square((z - mu) / sigma)
    log_norm = HALF_LOG_2PI + np.log(sigma).astype(F)
    lp = (log_unnorm - log_norm).astype(F)
    lp = lp - log_det_jac(z)
    return lp.sum(axis=-1, dtype=F)


def entropy(mu, sigma, eps2) -> np.ndarray:
    """``_entropy`` (sampling_layers.py:137-147) with the second normal draw ``eps2``."""
    normal_entropy = F(0.5) + HALF_LOG_2PI + np.log(sigma).astype(F)
    z = (mu + sigma * eps2).astype(F)
    return (normal_entropy + log_det_jac(z)).sum(axis=-1, dtype=F)


def policy_forward(net: ActorCritic, obs: np.ndarray, raw_action: Optional[np.ndarray] = None,
                   deterministic: bool = False, advance_rng: bool = True):
    """One network call ``networks(state, obs, rollout_extras)`` for the MLP actor-critic.

    Draw order follows ``sampling_layers.py:93-96,143-145``: key #1 for the action sample (skipped
    when deterministic), key #2 for the entropy estimate.  Returns a dict with raw_action, action,
    loglik, value, reg (per-sample regularisation loss), mu, sigma.
    """
    B = obs.shape[0]
    A = net.act_dim
    x = net.normalize_obs(obs)
    y, _ = net.actor.forward(x)
    v, _ = net.critic.forward(x)
    mu, rho = y[:, :A], y[:, A:]
    sigma = sampler_std(rho, net.min_std, net.std_scale)
    if deterministic:
        sampled = mu
    else:
        eps1 = prng.normal(net.next_key(), (B, A))
        sampled = (mu + sigma * eps1).astype(F)
    raw = sampled if raw_action is None else raw_action
    action = np.tanh(raw).astype(F)
    ll = loglikelihood(raw, mu, sigma)
    eps2 = prng.normal(net.next_key(), (B, A))
    reg = (F(-net.entropy_weight) * entropy(mu, sigma, eps2)).astype(F)
    return dict(raw_action=raw, action=action, loglik=ll, value=v[:, 0], reg=reg, mu=mu,
                sigma=sigma)
