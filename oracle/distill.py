"""TEST ORACLE (not product code; only tests/, smoke() and bench.py's cpu legs may import this).

Policy distillation (nnx_ppo/algorithms/distillation.py) restated on top of oracle/nets.py + oracle/ppo.py:

* rollout (:67-157): the STUDENT's sampled actions drive the env; the teacher, in eval mode, is run on the same
  observations and its rollout_extras holds its action mean in raw (pre-tanh) space;
* loss (:160-232): the student is replayed with the teacher's extras, so its sampler returns
  log p_student(mu_teacher | obs) (sampling_layers.py:96-108 with a given raw action) and its regulariser is the
  usual entropy term with fresh noise; total = -mean(loglik) + mean(regularisation); the value head is not in
  the loss (zero gradient);
* step (:235-360): split(rng_key) -> rollout -> E x M minibatch updates with the permutation of ppo.py:287-294
  -> student.update_statistics -> counters.  Sampler stream of the student: 2 counts per rollout step and
  2 per replayed step (no bootstrap call here), i.e. 2 T per update.

Parity unpinned against a live JAX run (no JAX in this image); the analytic gradient is pinned by float64
torch autograd in tests/test_oracle_distill.py.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from . import prng
from .nets import ActorCritic, entropy, loglikelihood, sampler_std, sigmoid
from .ppo import AdamState, F, _chain_backward, adam_update, minibatch_indices, unroll_env


def teacher_means(teacher: ActorCritic, obs_TBO: np.ndarray) -> np.ndarray:
    """The teacher's rollout_extras at the sampler position: raw_action = mu (deterministic mode)."""
    T, B, O = obs_TBO.shape
    x = teacher.normalize_obs(obs_TBO.reshape(T * B, O))
    y = teacher.actor.forward(x)[0]
    return y[:, :teacher.act_dim].reshape(T, B, teacher.act_dim).astype(F)


def distillation_loss_and_grads(student: ActorCritic, obs_TBO: np.ndarray, mu_teacher: np.ndarray, inds: np.ndarray,
                                rng_count_base: int, want_grads=True, n_global: Optional[int] = None):
    T = obs_TBO.shape[0]
    mb = inds.shape[0]
    A = student.act_dim
    N = T * mb
    Ng = F(N if n_global is None else n_global)
    x = student.normalize_obs(obs_TBO[:, inds].reshape(N, -1))
    y, zs = student.actor.forward(x, keep=True)
    mu, rho = y[:, :A], y[:, A:]
    sigma = sampler_std(rho, student.min_std, student.std_scale)
    z = mu_teacher[:, inds].reshape(N, A).astype(F)
    ll = loglikelihood(z, mu, sigma)
    eps2 = np.empty((T, mb, A), F)
    for t in range(T):
        k2 = prng.fold_in(student.rng_key, (rng_count_base + 2 * t + 1) & 0xFFFFFFFF)
        eps2[t] = prng.normal(k2, (mb, A))
    eps2 = eps2.reshape(N, A)
    ent = entropy(mu, sigma, eps2)
    reg = (F(-student.entropy_weight) * ent).astype(F)
    nll = F(-(ll.sum(dtype=F)) / Ng)
    reg_loss = F(reg.sum(dtype=F) / Ng)
    total = F(nll + reg_loss)
    metrics = {"losses/distillation_nll": nll, "losses/regularization": reg_loss, "loglik": ll.reshape(T, mb)}
    if not want_grads:
        return total, metrics, None
    g_ll = np.full(N, F(-1) / Ng, F)
    dmu_ll = ((z - mu) / (sigma * sigma)).astype(F)
    dsig_ll = (np.square(z - mu) / (sigma * sigma * sigma) - F(1) / sigma).astype(F)
    th = np.tanh((mu + sigma * eps2).astype(F)).astype(F)
    we = F(student.entropy_weight) / Ng
    d_mu = (g_ll[:, None] * dmu_ll + we * F(2) * th).astype(F)
    d_sig = (g_ll[:, None] * dsig_ll - we * (F(1) / sigma - F(2) * th * eps2)).astype(F)
    d_y = np.concatenate([d_mu, (d_sig * sigmoid(rho) * F(student.std_scale)).astype(F)], axis=1).astype(F)
    metrics["d_y"] = d_y
    dWa, dba = _chain_backward(student.actor, x, zs, d_y)
    parts = []
    for dW, db in zip(dWa, dba):
        parts += [dW.ravel(), db.ravel()]
    for W, b in zip(student.critic.W, student.critic.b):               # the value head is not in the loss
        parts += [np.zeros(W.size, F), np.zeros(b.size, F)]
    return total, metrics, np.concatenate(parts).astype(F)


@dataclasses.dataclass
class DistillationState:
    student: ActorCritic
    env_state: object
    opt: AdamState
    rng_key: np.ndarray
    steps_taken: F = F(0.0)


def new_distillation_state(env, student: ActorCritic, n_envs: int, seed: int) -> DistillationState:
    k = prng.key(seed)
    k, training_key = prng.split(k)                                     # distillation.py:387-388
    P = student.flat_params().size
    return DistillationState(student, env.reset_fast(prng.split(k, n_envs)),
                             AdamState(np.zeros(P, F), np.zeros(P, F), 0), training_key, F(0.0))


def distillation_step(env, teacher: ActorCritic, ds: DistillationState, n_envs, rollout_length, n_epochs=4, n_minibatches=4,
                      learning_rate=1e-4, gradient_clipping=None, weight_decay=None, trace: Optional[dict] = None):
    student = ds.student
    reset_key, new_key = prng.split(ds.rng_key)                          # distillation.py:264
    next_env_state, ro = unroll_env(env, ds.env_state, student, rollout_length, reset_key)
    mu_t = teacher_means(teacher, ro.obs)
    all_inds = minibatch_indices(new_key, n_envs, n_epochs, n_minibatches)
    per_update = []
    for u in range(all_inds.shape[0]):
        base = student.rng_count
        total, m, grads = distillation_loss_and_grads(student, ro.obs, mu_t, all_inds[u], base)
        student.rng_count = (base + 2 * rollout_length) & 0xFFFFFFFF
        student.set_flat_params(adam_update(student.flat_params(), grads, ds.opt, learning_rate,
                                            gradient_clipping=gradient_clipping, weight_decay=weight_decay))
        per_update.append((m["losses/distillation_nll"], m["losses/regularization"]))
        if trace is not None and u == 0:
            trace["first_update"] = dict(m, grads=grads, total=total)
    lm = np.array(per_update, F)
    metrics = {}
    for i, name in enumerate(("losses/distillation_nll", "losses/regularization")):
        metrics[f"{name}/mean"] = lm[:, i].mean(dtype=F)
        metrics[f"{name}/std"] = lm[:, i].std(dtype=F)
    total_steps = F(ds.steps_taken + F(rollout_length * n_envs))
    metrics["total_steps"] = total_steps
    student.update_statistics(ro.obs)                                   # distillation.py:327
    if trace is not None:
        trace["rollout"], trace["indices"], trace["teacher_mu"], trace["loss_per_update"] = ro, all_inds, mu_t, lm
    return DistillationState(student, next_env_state, ds.opt, new_key, total_steps), metrics
