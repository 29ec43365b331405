"""CPU restatement of the RECURRENT hot path (BASELINE configs[2]):
actor = Dense chain -> LSTM -> Dense chain, MLP critic, carry reset on done, BPTT through the replay.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Status: parity unpinned — the reference cannot be
imported here (no jax / flax), so this follows the reference source and flax's published LSTM cell:

* ``nnx_ppo/networks/recurrent.py:89-161``: the carry tuple goes straight to the flax cell, the
  layer output is the new hidden state, ``reset_state`` returns zeros - or, with
  ``trainable_initial_state`` (:85-87, 135-141, 154-157), the learned ``initial_h`` / ``initial_c``
  broadcast into carry slot 0 / slot 1 (slot 0 is what the flax cell reads as c); regularisation
  loss is zero.
* flax ``nnx.OptimizedLSTMCell`` (flax/nnx/nn/recurrent.py): four input kernels without bias and
  four hidden kernels with bias, gate order (i, f, g, o):
      i = sigmoid(x Wii + h Whi + bi)   f = sigmoid(x Wif + h Whf + bf)
      g = tanh   (x Wig + h Whg + bg)   o = sigmoid(x Wio + h Who + bo)
      c' = f * c + i * g                h' = o * tanh(c')
  Here the four kernels of a kind are stored concatenated: Wi [in, 4H], Wh [H, 4H], b [4H].
* ``rollout.py:11-45`` / ``ppo.py:409-437``: after every step the carry of envs that finished is
  replaced by ``reset_state`` (zeros); the replay starts from the carry the rollout started with
  and applies the same resets, so no gradient flows across an episode boundary.

Initialisation is builder-defined (uniform variance scaling for all kernels; flax's defaults are
lecun-normal / orthogonal, which cannot be matched bit-for-bit without jax): only the arithmetic of
the step, the reset semantics and the gradients are the contract.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from . import prng
from .nets import (ACT_IDS, ActorCritic, Chain, act_fwd, act_grad, entropy, loglikelihood,
                   sampler_std, sigmoid)
from .ppo import F, Rollout, _chain_backward, loss_head


@dataclasses.dataclass
class LSTMParams:
    Wi: np.ndarray   # [in, 4H]  (no bias on the input kernels)
    Wh: np.ndarray   # [H, 4H]
    b: np.ndarray    # [4H]      (bias of the hidden kernels)

    @property
    def hidden(self) -> int:
        return self.Wh.shape[0]


def lstm_step(p: LSTMParams, c: np.ndarray, h: np.ndarray, x: np.ndarray):
    """One cell step.  Returns (c', h', cache)."""
    H = p.hidden
    a = (x @ p.Wi + h @ p.Wh + p.b).astype(F)
    i, f, g, o = sigmoid(a[:, :H]), sigmoid(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]).astype(F), sigmoid(a[:, 3 * H:])
    c2 = (f * c + i * g).astype(F)
    tc = np.tanh(c2).astype(F)
    h2 = (o * tc).astype(F)
    return c2, h2, (i, f, g, o, tc)


@dataclasses.dataclass
class RecurrentActorCritic:
    """Normalizer -> {actor: pre Chain -> LSTM -> post Chain -> NormalTanhSampler, critic: Chain}."""
    obs_dim: int
    act_dim: int
    pre: Chain
    lstm: LSTMParams
    post: Chain
    critic: Chain
    rng_key: np.ndarray
    rng_count: int = 0
    entropy_weight: float = 1e-2
    min_std: float = 0.1
    std_scale: float = 1.0
    normalize: bool = True
    mean: Optional[np.ndarray] = None
    M2: Optional[np.ndarray] = None
    counter: F = F(0.0)
    # trainable_initial_state: learned carry (slot 0 = c, slot 1 = h of this restatement), [H] each
    init_c: Optional[np.ndarray] = None
    init_h: Optional[np.ndarray] = None

    norm_std = ActorCritic.norm_std
    normalize_obs = ActorCritic.normalize_obs
    update_statistics = ActorCritic.update_statistics
    next_key = ActorCritic.next_key

    def initialize_state(self, B: int):
        H = self.lstm.hidden
        if self.init_c is not None:
            return np.broadcast_to(self.init_c, (B, H)).astype(F).copy(), np.broadcast_to(self.init_h, (B, H)).astype(F).copy()
        return np.zeros((B, H), F), np.zeros((B, H), F)          # (c, h)

    def reset_carry(self, carry, done):
        """rollout.py:33-40 / ppo.py:419-425: finished envs get ``reset_state``."""
        d = done[:, None]
        r0 = F(0) if self.init_c is None else self.init_c[None, :]
        r1 = F(0) if self.init_h is None else self.init_h[None, :]
        return np.where(d, r0, carry[0]).astype(F), np.where(d, r1, carry[1]).astype(F)

    def param_list(self):
        out = []
        for ch in (self.pre,):
            for W, b in zip(ch.W, ch.b):
                out += [W, b]
        out += [self.lstm.Wi, self.lstm.Wh, self.lstm.b]
        if self.init_c is not None:
            out += [self.init_c, self.init_h]
        for ch in (self.post, self.critic):
            for W, b in zip(ch.W, ch.b):
                out += [W, b]
        return out

    def flat_params(self) -> np.ndarray:
        return np.concatenate([p.ravel() for p in self.param_list()]).astype(F)

    def set_flat_params(self, flat: np.ndarray) -> None:
        o = 0
        for p in self.param_list():
            p[...] = flat[o:o + p.size].reshape(p.shape)
            o += p.size
        assert o == flat.size


def make_recurrent_actor_critic(obs_size, action_size, pre_sizes, lstm_hidden, post_sizes, critic_hidden_sizes,
                                seed=0, activation="relu", normalize_obs=True,
                                trainable_initial_state=False) -> RecurrentActorCritic:
    """Key-draw order: pre Dense layers, LSTM (Wi, Wh), post Dense layers, critic layers, then the
    sampler's stream key — every Linear consumes two counts of the Rngs stream like the MLP factory."""
    rngs = prng.Rngs(seed)
    act = ACT_IDS[activation]

    def uniform(shape, fan_in):
        return prng.variance_scaling_uniform(rngs(), shape[0], shape[1], 1.0)

    def chain(sizes, last_linear=True):   # (a pre chain is evaluated with _chain_all_act instead)
        Ws, bs = [], []
        for din, dout in zip(sizes[:-1], sizes[1:]):
            Ws.append(uniform((din, dout), din))
            rngs()                                                 # bias init key (zeros)
            bs.append(np.zeros(dout, F))
        return Chain(list(sizes), act, Ws, bs)

    pre = chain([obs_size] + list(pre_sizes), last_linear=False)
    H = lstm_hidden
    lstm = LSTMParams(uniform((pre_sizes[-1], 4 * H), pre_sizes[-1]), uniform((H, 4 * H), H), np.zeros(4 * H, F))
    post = chain([H] + list(post_sizes) + [2 * action_size], last_linear=True)
    critic = chain([obs_size] + list(critic_hidden_sizes) + [1], last_linear=True)
    # the sampler keeps drawing from the same nnx.Rngs stream (factories.py:116-137)
    net = RecurrentActorCritic(obs_size, action_size, pre, lstm, post, critic, rng_key=rngs.key.copy(),
                               rng_count=rngs.count)
    net.normalize = normalize_obs
    if normalize_obs:
        net.mean, net.M2 = np.zeros(obs_size, F), np.zeros(obs_size, F)
    if trainable_initial_state:                                   # recurrent.py:85-87: zeros, no key drawn
        net.init_c, net.init_h = np.zeros(H, F), np.zeros(H, F)
    return net


def _chain_all_act(chain: Chain, x: np.ndarray):
    """A 'pre' chain applies the activation after EVERY layer (Dense(..., activation=act))."""
    zs, h = [], x
    for W, b in zip(chain.W, chain.b):
        z = (h @ W + b).astype(F)
        zs.append(z)
        h = act_fwd(z, chain.act)
    return h, zs


def actor_step(net: RecurrentActorCritic, c, h, x):
    """obs (already normalised) -> (c', h', y, caches) through pre -> LSTM -> post."""
    u, zs_pre = _chain_all_act(net.pre, x)
    c2, h2, cache = lstm_step(net.lstm, c, h, u)
    y, zs_post = net.post.forward(h2, keep=True)
    return c2, h2, y, (u, zs_pre, cache, zs_post)


def policy_forward(net: RecurrentActorCritic, carry, obs, raw_action=None):
    """One network call (sampling_layers.py:88-147 on top of the recurrent actor).  Consumes two
    counts of the sampler stream like the MLP path.  Returns (next_carry, outputs)."""
    A = net.act_dim
    x = net.normalize_obs(obs)
    c2, h2, y, _ = actor_step(net, carry[0], carry[1], x)
    v = net.critic.forward(x)[0][:, 0]
    mu, rho = y[:, :A], y[:, A:]
    sigma = sampler_std(rho, net.min_std, net.std_scale)
    k1, k2 = net.next_key(), net.next_key()
    if raw_action is None:
        z = (mu + sigma * prng.normal(k1, mu.shape)).astype(F)
    else:
        z = raw_action
    eps2 = prng.normal(k2, mu.shape)
    out = {"raw_action": z, "action": np.tanh(z).astype(F), "loglik": loglikelihood(z, mu, sigma), "value": v,
           "reg": (F(-net.entropy_weight) * entropy(mu, sigma, eps2)).astype(F)}
    return (c2, h2), out


def unroll_env(env, env_state, net: RecurrentActorCritic, carry, T: int, reset_key):
    """rollout.unroll_env with a recurrent policy.  Returns (env_state, carry, rollout, start_carry)."""
    from .env import EnvState
    B = env_state.obs.shape[0]
    keys = prng.split(reset_key, (T, B))
    O, A = net.obs_dim, net.act_dim
    ro = Rollout(np.zeros((T, B, O), F), np.zeros((T, B, A), F), np.zeros((T, B, A), F),
                 np.zeros((T, B), F), np.zeros((T, B), F), np.zeros((T, B), F),
                 np.zeros((T, B), bool), np.zeros((T, B), bool), np.zeros((B, O), F))
    start = (carry[0].copy(), carry[1].copy())
    s = env_state
    for t in range(T):
        carry, out = policy_forward(net, carry, s.obs)
        nxt = env.step(s, out["action"])
        ro.obs[t] = s.obs
        ro.raw_action[t], ro.action[t] = out["raw_action"], out["action"]
        ro.loglik[t], ro.value[t] = out["loglik"], out["value"]
        ro.reward[t], ro.done[t], ro.truncated[t] = nxt.reward, nxt.done, nxt.truncated
        if t == T - 1:
            ro.next_obs_last[:] = nxt.obs
        rs = env.reset_fast(keys[t])
        d = nxt.done
        carry = net.reset_carry(carry, d)
        s = EnvState(np.where(d[:, None], rs.obs, nxt.obs),
                     np.where(d, rs.step_counter, nxt.step_counter).astype(np.int32),
                     np.where(d, rs.term_state, nxt.term_state).astype(np.uint32),
                     np.where(d, rs.reward, nxt.reward), np.where(d, rs.done, nxt.done),
                     np.where(d, rs.truncated, nxt.truncated))
    return s, carry, ro, start


def ppo_loss_and_grads(net: RecurrentActorCritic, ro: Rollout, start_carry, inds: np.ndarray,
                       rng_count_base: int, want_grads=True, **kw):
    """``ppo_loss`` for the recurrent actor: replay T steps from the minibatch's start carry with the
    rollout's resets (ppo.py:409-431), bootstrap value (ppo.py:433-437), shared loss head, then
    back-propagation through time.  Gradient layout = ``RecurrentActorCritic.param_list()``."""
    T = ro.obs.shape[0]
    mb = inds.shape[0]
    N = T * mb
    done = ro.done[:, inds]
    xs = net.normalize_obs(ro.obs[:, inds].reshape(N, -1)).reshape(T, mb, -1)
    c, h = start_carry[0][inds].copy(), start_carry[1][inds].copy()
    ys, caches, c_in, h_in = [], [], [], []
    for t in range(T):
        c_in.append(c)
        h_in.append(h)
        c, h, y, cache = actor_step(net, c, h, xs[t])
        ys.append(y)
        caches.append(cache)
        if net.init_c is None:
            keep = (~done[t])[:, None]
            c, h = (c * keep).astype(F), (h * keep).astype(F)                  # reset_state -> zeros
        else:
            c, h = net.reset_carry((c, h), done[t])                            # ... or the learned carry
    y = np.concatenate(ys, axis=0)
    xf = xs.reshape(N, -1)
    v, zs_c = net.critic.forward(xf, keep=True)
    v = v[:, 0]
    v_last = net.critic.forward(net.normalize_obs(ro.next_obs_last[inds]))[0][:, 0]
    total, metrics, d_y, d_v = loss_head(net, ro, inds, rng_count_base, y, v, v_last, want_grads=want_grads, **kw)
    if not want_grads:
        return total, metrics, None

    H = net.lstm.hidden
    d_y = d_y.reshape(T, mb, -1)
    g_pre_W = [np.zeros_like(W) for W in net.pre.W]
    g_pre_b = [np.zeros_like(b) for b in net.pre.b]
    g_post_W = [np.zeros_like(W) for W in net.post.W]
    g_post_b = [np.zeros_like(b) for b in net.post.b]
    gWi, gWh, gb = np.zeros_like(net.lstm.Wi), np.zeros_like(net.lstm.Wh), np.zeros_like(net.lstm.b)
    dc_next = np.zeros((mb, H), F)
    dh_next = np.zeros((mb, H), F)
    g_init_c, g_init_h = np.zeros(H, F), np.zeros(H, F)
    for t in reversed(range(T)):
        u, zs_pre, (i, f, g, o, tc), zs_post = caches[t]
        h_t = (o * tc).astype(F)
        # post chain: y_t = post(h_t)
        L = net.post.n_layers
        d = d_y[t]
        for l in reversed(range(L)):
            hin = h_t if l == 0 else act_fwd(zs_post[l - 1], net.post.act)
            g_post_W[l] += (hin.T @ d).astype(F)
            g_post_b[l] += d.sum(axis=0, dtype=F)
            d = (d @ net.post.W[l].T).astype(F)
            if l > 0:
                d = (d * act_grad(zs_post[l - 1], net.post.act)).astype(F)
        keep = (~done[t])[:, None]
        if net.init_c is not None:                   # where done[t], step t+1 started from the learned carry
            g_init_c += (dc_next * ~keep).sum(axis=0, dtype=F)
            g_init_h += (dh_next * ~keep).sum(axis=0, dtype=F)
        dh = (d + dh_next * keep).astype(F)          # the carry handed to step t+1 was zeroed where done[t]
        dc = (dh * o * (F(1) - tc * tc) + dc_next * keep).astype(F)
        da = np.concatenate([dc * g * i * (F(1) - i), dc * c_in[t] * f * (F(1) - f), dc * i * (F(1) - g * g),
                             dh * tc * o * (F(1) - o)], axis=1).astype(F)
        gWi += (u.T @ da).astype(F)
        gWh += (h_in[t].T @ da).astype(F)
        gb += da.sum(axis=0, dtype=F)
        dh_next = (da @ net.lstm.Wh.T).astype(F)
        dc_next = (dc * f).astype(F)
        # pre chain (activation after every layer)
        d = (da @ net.lstm.Wi.T).astype(F)
        for l in reversed(range(net.pre.n_layers)):
            d = (d * act_grad(zs_pre[l], net.pre.act)).astype(F)
            hin = xs[t] if l == 0 else act_fwd(zs_pre[l - 1], net.pre.act)
            g_pre_W[l] += (hin.T @ d).astype(F)
            g_pre_b[l] += d.sum(axis=0, dtype=F)
            d = (d @ net.pre.W[l].T).astype(F)
    dWc, dbc = _chain_backward(net.critic, xf, zs_c, d_v)
    parts = []
    for W, b in zip(g_pre_W, g_pre_b):
        parts += [W.ravel(), b.ravel()]
    parts += [gWi.ravel(), gWh.ravel(), gb.ravel()]
    if net.init_c is not None:
        parts += [g_init_c, g_init_h]
    for W, b in zip(g_post_W, g_post_b):
        parts += [W.ravel(), b.ravel()]
    for W, b in zip(dWc, dbc):
        parts += [W.ravel(), b.ravel()]
    return total, metrics, np.concatenate(parts).astype(F)


# ------------------------------------------------------------------------------------------
# one full PPO iteration with the recurrent actor (ppo.py:254-348)
# ------------------------------------------------------------------------------------------
@dataclasses.dataclass
class RecurrentTrainingState:
    net: RecurrentActorCritic
    env_state: object
    carry: tuple            # network_states (types.py:50-56): the LSTM (c, h) of every env
    opt: object
    rng_key: np.ndarray
    steps_taken: F = F(0.0)


def new_training_state(env, net: RecurrentActorCritic, n_envs: int, seed: int) -> RecurrentTrainingState:
    from .ppo import AdamState
    k = prng.key(seed)
    k, training_key = prng.split(k)
    P = net.flat_params().size
    return RecurrentTrainingState(net, env.reset_fast(prng.split(k, n_envs)), net.initialize_state(n_envs),
                                  AdamState(np.zeros(P, F), np.zeros(P, F), 0), training_key, F(0.0))


def ppo_step(env, ts: RecurrentTrainingState, n_envs, rollout_length, n_epochs=4, n_minibatches=4,
             learning_rate=1e-4, trace: Optional[dict] = None, **loss_kw):
    """Same orchestration as oracle.ppo.ppo_step; every update replays from the carry the rollout
    STARTED with (ppo.py:296-300 gathers ``network_state[inds]`` of the pre-rollout state)."""
    from .ppo import adam_update, minibatch_indices
    net = ts.net
    reset_key, new_key = prng.split(ts.rng_key)
    env_state, carry, ro, start = unroll_env(env, ts.env_state, net, ts.carry, rollout_length, reset_key)
    all_inds = minibatch_indices(new_key, n_envs, n_epochs, n_minibatches)
    per_update = []
    for u in range(all_inds.shape[0]):
        base = net.rng_count
        total, m, grads = ppo_loss_and_grads(net, ro, start, all_inds[u], base, **loss_kw)
        net.rng_count = (base + 2 * (rollout_length + 1)) & 0xFFFFFFFF
        net.set_flat_params(adam_update(net.flat_params(), grads, ts.opt, learning_rate))
        per_update.append((m["losses/actor"], m["losses/critic"], m["losses/regularization"]))
        if trace is not None and u == 0:
            trace["first_update"] = dict(m, grads=grads, total=total)
    lm = np.array(per_update, F)
    metrics = {f"{n}/mean": lm[:, i].mean(dtype=F) for i, n in
               enumerate(("losses/actor", "losses/critic", "losses/regularization"))}
    steps = F(ts.steps_taken + F(rollout_length * n_envs))
    metrics["total_steps"] = steps
    net.update_statistics(ro.obs)
    if trace is not None:
        trace["rollout"], trace["indices"], trace["start_carry"] = ro, all_inds, start
    return RecurrentTrainingState(net, env_state, carry, ts.opt, new_key, steps), metrics
