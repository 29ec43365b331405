"""NumPy float32 restatement of the PPO iteration (ORACLE — test infrastructure).

Follows ``nnx_ppo/algorithms/ppo.py`` (``gae`` :351-394, ``ppo_loss`` :397-531, ``ppo_step``
:254-348, ``new_training_state`` :534-572) and ``nnx_ppo/algorithms/rollout.py``
(``single_transition`` :11-45, ``unroll_env`` :48-73).  The optimizer arithmetic is optax's
(third-party, unpinned): ``adam`` / ``adamw`` / ``clip_by_global_norm`` restated from its
published update rule.  Gradients are analytic and are validated against
torch.autograd in float64 by ``tests/test_oracle_ppo.py``.
"""

from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from . import prng
from .env import SyntheticEnv, EnvState
from .nets import (F, ActorCritic, act_fwd, act_grad, sampler_std, sigmoid, loglikelihood,
                   entropy, policy_forward)


# ------------------------------------------------------------------------------------------
# gae — ppo.py:351-394 (same op order; float32)
# ------------------------------------------------------------------------------------------
def gae(rewards, values_excl_last, last_value, done, truncation, lambda_, gamma, dtype=F):
    T, B = rewards.shape
    f = dtype
    values = np.concatenate([values_excl_last, last_value.reshape(1, B)], axis=0).astype(f)
    adv = np.zeros((T, B), f)
    nxt = np.zeros(B, f)
    g, lam = f(gamma), f(lambda_)
    for t in reversed(range(T)):
        nv = np.where(done[t], f(0.0), values[t + 1])
        new_value = rewards[t].astype(f) + g * nv
        a = new_value - values[t]
        a = np.where(truncation[t], f(0.0), a)
        nxt = (a + (f(1) - done[t].astype(f)) * g * lam * nxt).astype(f)
        adv[t] = nxt
    return adv


# ------------------------------------------------------------------------------------------
# rollout — rollout.py:11-73
# ------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Rollout:
    obs: np.ndarray           # [T, B, O] raw observations (also the Normalizer's rollout_extras)
    raw_action: np.ndarray    # [T, B, A]
    action: np.ndarray        # [T, B, A]
    loglik: np.ndarray        # [T, B]
    value: np.ndarray         # [T, B]
    reward: np.ndarray        # [T, B]
    done: np.ndarray          # [T, B] bool
    truncated: np.ndarray     # [T, B] bool
    next_obs_last: np.ndarray  # [B, O] = next_obs[-1] (pre-reset terminal obs)


def unroll_env(env: SyntheticEnv, env_state: EnvState, net: ActorCritic, T: int,
               reset_key: np.ndarray):
    B = env_state.obs.shape[0]
    keys = prng.split(reset_key, (T, B))                                # rollout.py:57-59
    O, A = net.obs_dim, net.act_dim
    ro = Rollout(np.zeros((T, B, O), F), np.zeros((T, B, A), F), np.zeros((T, B, A), F),
                 np.zeros((T, B), F), np.zeros((T, B), F), np.zeros((T, B), F),
                 np.zeros((T, B), bool), np.zeros((T, B), bool), np.zeros((B, O), F))
    s = env_state
    for t in range(T):
        out = policy_forward(net, s.obs)                                 # rollout.py:18
        nxt = env.step(s, out["action"])                                 # rollout.py:21
        ro.obs[t] = s.obs
        ro.raw_action[t], ro.action[t] = out["raw_action"], out["action"]
        ro.loglik[t], ro.value[t] = out["loglik"], out["value"]
        ro.reward[t], ro.done[t], ro.truncated[t] = nxt.reward, nxt.done, nxt.truncated
        if t == T - 1:
            ro.next_obs_last[:] = nxt.obs
        rs = env.reset_fast(keys[t])                                     # rollout.py:39-40
        d = nxt.done
        s = EnvState(np.where(d[:, None], rs.obs, nxt.obs),
                     np.where(d, rs.step_counter, nxt.step_counter).astype(np.int32),
                     np.where(d, rs.term_state, nxt.term_state).astype(np.uint32),
                     np.where(d, rs.reward, nxt.reward), np.where(d, rs.done, nxt.done),
                     np.where(d, rs.truncated, nxt.truncated))
    return s, ro


def eval_rollout(env: SyntheticEnv, net: ActorCritic, n_envs: int, max_episode_length: int,
                 key: np.ndarray, deterministic: bool = True):
    """rollout.py:97-148 -> (cumulative reward [B], lifespan [B]).  ``done`` is sticky (:115-117),
    a step's reward counts while the env was not done BEFORE it (:119-123), lifespan counts the
    steps that did not end in done (:124).  train_ppo calls it under ``networks.eval()``
    (ppo.py:122), i.e. with the deterministic sampler."""
    s = env.reset_fast(prng.split(key, n_envs))                          # rollout.py:105-106
    prev_done = np.zeros(n_envs, bool)
    cuml = np.zeros(n_envs, F)
    lifespan = np.zeros(n_envs, F)
    for _ in range(max_episode_length):
        out = policy_forward(net, s.obs, deterministic=deterministic)
        nxt = env.step(s, out["action"])
        done = nxt.done | prev_done
        cuml = (cuml + np.where(prev_done, F(0), nxt.reward)).astype(F)
        lifespan = (lifespan + np.where(done, F(0), F(1))).astype(F)
        prev_done = done
        s = nxt
    return cuml, lifespan


# ------------------------------------------------------------------------------------------
# ppo_loss forward + analytic backward — ppo.py:397-531, SURVEY App. A
# ------------------------------------------------------------------------------------------
def _chain_backward(chain, x0, zs, d_out):
    """Backprop ``d_out`` (grad w.r.t. the last layer's output) through a Chain.
    Returns ([dW], [db])."""
    if hasattr(chain, "backward"):          # dictnet.EncChain: per-key encoders + trunk
        return chain.backward(x0, zs, d_out)
    L = chain.n_layers
    dWs, dbs = [None] * L, [None] * L
    d = d_out
    for l in reversed(range(L)):
        h_in = x0 if l == 0 else act_fwd(zs[l - 1], chain.act)
        dWs[l] = (h_in.T @ d).astype(F)
        dbs[l] = d.sum(axis=0, dtype=F)
        if l > 0:
            d = ((d @ chain.W[l].T) * act_grad(zs[l - 1], chain.act)).astype(F)
    return dWs, dbs


def loss_head(net, ro: Rollout, inds: np.ndarray, rng_count_base: int, y, v, v_last,
              clip_range=0.2, normalize_advantages=True, discounting_factor=0.99, gae_lambda=0.95,
              critic_loss_weight=1.0, want_grads=True, n_global: Optional[int] = None, adv_stats=None):
    """Everything of ``ppo_loss`` downstream of the network outputs (ppo.py:447-531): sampler
    log-lik / entropy, GAE, advantage normalisation, clipped surrogate, value loss — and the
    analytic gradients w.r.t. the actor output ``y`` [T*mb, 2A] and the values ``v`` [T*mb].
    Shared by the MLP path below and the recurrent path (oracle/recurrent.py)."""
    T = ro.obs.shape[0]
    mb = inds.shape[0]
    A = net.act_dim
    N = T * mb
    Ng = F(N if n_global is None else n_global)
    z = ro.raw_action[:, inds].reshape(N, A)
    ll_old = ro.loglik[:, inds].reshape(N)
    mu, rho = y[:, :A], y[:, A:]
    sigma = sampler_std(rho, net.min_std, net.std_scale)
    ll = loglikelihood(z, mu, sigma)
    eps2 = np.empty((T, mb, A), F)
    for t in range(T):
        k2 = prng.fold_in(net.rng_key, (rng_count_base + 2 * t + 1) & 0xFFFFFFFF)
        eps2[t] = prng.normal(k2, (mb, A))
    eps2 = eps2.reshape(N, A)
    ent = entropy(mu, sigma, eps2)
    reg = (F(-net.entropy_weight) * ent).astype(F)

    adv = gae(ro.reward[:, inds], v.reshape(T, mb), v_last, ro.done[:, inds],
              ro.truncated[:, inds], gae_lambda, discounting_factor).reshape(N)
    target = (v + adv).astype(F)                                        # ppo.py:456-458
    if normalize_advantages:                                            # ppo.py:477-480
        if adv_stats is None:
            a_mean, a_std = adv.mean(dtype=F), adv.std(dtype=F)
        else:
            a_mean, a_std = F(adv_stats[0]), F(adv_stats[1])
        adv_n = ((adv - a_mean) / (a_std + F(1e-8))).astype(F)
    else:
        adv_n = adv
    ratio = np.exp(ll - ll_old).astype(F)                               # ppo.py:482-488
    lo, hi = F(1 - clip_range), F(1 + clip_range)
    c1 = ratio * adv_n
    c2 = np.clip(ratio, lo, hi) * adv_n
    actor_loss = F(-(np.minimum(c1, c2).sum(dtype=F)) / Ng)
    diff = (v - target).astype(F)
    critic_loss = F(0.5) * F((diff * diff).sum(dtype=F) / Ng)           # ppo.py:496-500
    reg_loss = F(reg.sum(dtype=F) / Ng)                                 # ppo.py:503
    total = F(actor_loss + F(critic_loss_weight) * critic_loss + reg_loss)
    metrics = {"losses/actor": actor_loss, "losses/critic": critic_loss,
               "losses/regularization": reg_loss, "adv_mean": adv.mean(dtype=F),
               "adv_std": adv.std(dtype=F), "adv": adv.reshape(T, mb), "values": v.reshape(T, mb),
               "v_last": v_last, "loglik": ll.reshape(T, mb), "eps2": eps2.reshape(T, mb, A),
               # ppo.py:514-527 (ACTOR_EXTRA / CRITIC_EXTRA)
               "losses/clipping_fraction": F(np.mean(np.abs(ratio - F(1)) > F(clip_range))),
               "losses/critic_R^2": F(1.0 - 2.0 * float(critic_loss) / (float(np.var(target.astype(np.float64))) + 1e-8))}
    if not want_grads:
        return total, metrics, None, None

    # ---- analytic backward (JAX tie rules: minimum / clip split the gradient 0.5/0.5) ----
    w1 = np.where(c1 < c2, F(1), np.where(c1 == c2, F(0.5), F(0)))
    w2 = F(1) - w1
    dclip = np.where((ratio > lo) & (ratio < hi), F(1),
                     np.where((ratio == lo) | (ratio == hi), F(0.5), F(0)))
    g_ratio = (-(w1 * adv_n + w2 * adv_n * dclip) / Ng).astype(F)
    g_ll = (g_ratio * ratio).astype(F)
    dmu_ll = ((z - mu) / (sigma * sigma)).astype(F)
    dsig_ll = (np.square(z - mu) / (sigma * sigma * sigma) - F(1) / sigma).astype(F)
    zp = (mu + sigma * eps2).astype(F)
    th = np.tanh(zp).astype(F)
    we = F(net.entropy_weight) / Ng
    d_mu = (g_ll[:, None] * dmu_ll + we * F(2) * th).astype(F)
    d_sig = (g_ll[:, None] * dsig_ll - we * (F(1) / sigma - F(2) * th * eps2)).astype(F)
    d_rho = (d_sig * sigmoid(rho) * F(net.std_scale)).astype(F)
    d_y = np.concatenate([d_mu, d_rho], axis=1).astype(F)
    d_v = (F(critic_loss_weight) * diff / Ng).astype(F)[:, None]
    metrics["d_y"], metrics["d_v"] = d_y, d_v[:, 0]
    return total, metrics, d_y, d_v


def ppo_loss_and_grads(net: ActorCritic, ro: Rollout, inds: np.ndarray, rng_count_base: int,
                       clip_range=0.2, normalize_advantages=True, discounting_factor=0.99,
                       gae_lambda=0.95, critic_loss_weight=1.0, want_grads=True,
                       n_global: Optional[int] = None, adv_stats=None):
    """Loss (and flat analytic gradient) of one minibatch, restating ``ppo_loss``.

    ``rng_count_base`` is the sampler stream count at the start of this loss call; replay step t
    consumes counts base+2t (unused sample draw) and base+2t+1 (entropy noise), the bootstrap
    call consumes two more (``sampling_layers.py:96,143-145``; ``ppo.py:425-436``).
    ``n_global`` / ``adv_stats`` let a data-parallel caller use global means (defaults: local).
    """
    T = ro.obs.shape[0]
    mb = inds.shape[0]
    N = T * mb
    obs = ro.obs[:, inds].reshape(N, -1)
    x = net.normalize_obs(obs)
    y, zs_a = net.actor.forward(x, keep=True)
    v, zs_c = net.critic.forward(x, keep=True)
    v = v[:, 0]
    x_last = net.normalize_obs(ro.next_obs_last[inds])
    v_last = net.critic.forward(x_last)[0][:, 0]                       # ppo.py:433-437
    total, metrics, d_y, d_v = loss_head(net, ro, inds, rng_count_base, y, v, v_last, clip_range,
                                         normalize_advantages, discounting_factor, gae_lambda,
                                         critic_loss_weight, want_grads, n_global, adv_stats)
    if not want_grads:
        return total, metrics, None
    dWa, dba = _chain_backward(net.actor, x, zs_a, d_y)
    dWc, dbc = _chain_backward(net.critic, x, zs_c, d_v)
    parts = []
    for dWs, dbs in ((dWa, dba), (dWc, dbc)):
        for dW, db in zip(dWs, dbs):
            parts += [dW.ravel(), db.ravel()]
    grads = np.concatenate(parts).astype(F)
    if hasattr(net, "fold_grads"):            # shared trunk (oracle/sharednet.py): one parameter, two paths
        grads = net.fold_grads(grads)
    return total, metrics, grads


# ------------------------------------------------------------------------------------------
# optax adam / adamw / clip_by_global_norm — ppo.py:555-569
# ------------------------------------------------------------------------------------------
@dataclasses.dataclass
class AdamState:
    mu: np.ndarray
    nu: np.ndarray
    count: int = 0


def adam_update(params, grads, st: AdamState, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8,
                gradient_clipping=None, weight_decay=None):
    g = grads.astype(F)
    if gradient_clipping is not None:
        gn = np.sqrt((g * g).sum(dtype=F)).astype(F)
        c = F(gradient_clipping)
        g = g if gn < c else ((g / gn) * c).astype(F)
    st.mu = (F(1 - b1) * g + F(b1) * st.mu).astype(F)
    st.nu = (F(1 - b2) * (g * g) + F(b2) * st.nu).astype(F)
    st.count += 1
    bc1 = F(1) - F(b1) ** F(st.count)
    bc2 = F(1) - F(b2) ** F(st.count)
    mu_hat = (st.mu / bc1).astype(F)
    nu_hat = (st.nu / bc2).astype(F)
    upd = (mu_hat / (np.sqrt(nu_hat).astype(F) + F(eps))).astype(F)
    if weight_decay is not None:
        wd = 1e-4 if (isinstance(weight_decay, bool) and weight_decay) else weight_decay
        upd = (upd + F(wd) * params).astype(F)
    return (params + F(-lr) * upd).astype(F)


# ------------------------------------------------------------------------------------------
# ppo_step / new_training_state — ppo.py:254-348, 534-572
# ------------------------------------------------------------------------------------------
@dataclasses.dataclass
class TrainingState:
    net: ActorCritic
    env_state: EnvState
    opt: AdamState
    rng_key: np.ndarray
    steps_taken: F = F(0.0)


def new_training_state(env: SyntheticEnv, net: ActorCritic, n_envs: int, seed: int):
    k = prng.key(seed)
    k, training_key = prng.split(k)                                     # ppo.py:544-545
    env_keys = prng.split(k, n_envs)                                    # ppo.py:548
    P = net.flat_params().size
    return TrainingState(net, env.reset_fast(env_keys),
                         AdamState(np.zeros(P, F), np.zeros(P, F), 0), training_key, F(0.0))


def minibatch_indices(new_key, n_envs, n_epochs, n_minibatches):
    """ppo.py:284-294 → [E*M, n_envs // M] int32."""
    mb = n_envs // n_minibatches
    rows = []
    for e in range(n_epochs):
        perm = prng.permutation(prng.fold_in(new_key, e), n_envs)
        rows.append(perm[: n_minibatches * mb].reshape(n_minibatches, mb))
    return np.concatenate(rows, axis=0).astype(np.int32)


def ppo_step(env: SyntheticEnv, ts: TrainingState, n_envs, rollout_length, gae_lambda=0.95,
             discounting_factor=0.99, clip_range=0.2, normalize_advantages=True, n_epochs=4,
             n_minibatches=4, critic_loss_weight=1.0, learning_rate=1e-4,
             gradient_clipping=None, weight_decay=None, trace: Optional[dict] = None):
    net = ts.net
    reset_key, new_key = prng.split(ts.rng_key)                          # ppo.py:271
    next_env_state, ro = unroll_env(env, ts.env_state, net, rollout_length, reset_key)
    all_inds = minibatch_indices(new_key, n_envs, n_epochs, n_minibatches)
    per_update = []
    for u in range(all_inds.shape[0]):
        base = net.rng_count
        total, m, grads = ppo_loss_and_grads(net, ro, all_inds[u], base, clip_range,
                                             normalize_advantages, discounting_factor,
                                             gae_lambda, critic_loss_weight)
        net.rng_count = (base + 2 * (rollout_length + 1)) & 0xFFFFFFFF
        p = adam_update(net.flat_params(), grads, ts.opt, learning_rate,
                        gradient_clipping=gradient_clipping, weight_decay=weight_decay)
        net.set_flat_params(p)
        per_update.append((m["losses/actor"], m["losses/critic"], m["losses/regularization"]))
        if trace is not None and u == 0:
            trace["first_update"] = dict(m, grads=grads, total=total)
    lm = np.array(per_update, F)
    metrics = {}
    for i, name in enumerate(("losses/actor", "losses/critic", "losses/regularization")):
        metrics[f"{name}/mean"] = lm[:, i].mean(dtype=F)
        metrics[f"{name}/std"] = lm[:, i].std(dtype=F)
    total_steps = F(ts.steps_taken + F(rollout_length * n_envs))        # ppo.py:329 (float32)
    metrics["total_steps"] = total_steps
    net.update_statistics(ro.obs)                                       # ppo.py:336
    if trace is not None:
        trace["rollout"], trace["indices"], trace["loss_per_update"] = ro, all_inds, lm
    return TrainingState(net, next_env_state, ts.opt, new_key, total_steps), metrics
