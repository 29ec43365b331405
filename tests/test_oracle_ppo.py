"""Oracle self-checks: the reference's own KATs plus an independent float64 autograd check of
the analytic gradients."""
import os

import numpy as np
import pytest
import torch

from oracle import env as oenv
from oracle import nets, ppo, prng

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _gae_kat_inputs():
    """Verbatim recipe of the reference's test_gae (ppo_test.py:229-246)."""
    np.random.seed(23)
    T, B = 100, 512
    rewards = np.random.normal(size=(T, B))
    values = np.random.normal(size=(T + 1, B))
    done = np.random.choice([True, False], size=(T, B), p=[0.01, 0.99])
    trunc = np.random.choice([True, False], size=(T, B))
    trunc = np.logical_and(done, trunc)
    return rewards, values, done, trunc


def test_gae_known_answer_matches_reference_test():
    """ppo_test.py:229-264: float32 gae vs the test's inline float64 loop, max|diff| < 1e-6."""
    rewards, values, done, trunc = _gae_kat_inputs()
    g = np.load(os.path.join(GOLDEN, "gae_kat.npz"))
    assert done.sum() == 515 and trunc.sum() == 246          # SURVEY App. F checksums
    assert np.array_equal(np.packbits(done), g["done_bits"])
    adv = ppo.gae(rewards.astype(np.float32), values[:-1].astype(np.float32),
                  values[-1].astype(np.float32), done, trunc, 0.95, 0.8)
    assert np.abs(adv - g["adv_f64"]).max() < 1e-6
    assert abs(g["adv_f64"].sum() - (-1453.3261142478)) < 1e-6


def test_normalizer_moments_and_default_std():
    """normalizer_test.py:33-65: std=10 before any update; Welford merge == mean/std (1e-5)."""
    net = nets.make_mlp_actor_critic(8, 2, [16], [16], seed=1)
    x = np.array([[1, 2, 3, 4, 5, 6, 7, 8]], np.float32)
    assert np.allclose(net.normalize_obs(x), x / 10.0)
    g = np.random.default_rng(42)
    data = (g.standard_normal(8) + g.standard_normal((5, 16, 8))).astype(np.float32)
    net.update_statistics(data)
    assert float(net.counter) == 5 * 16
    assert np.abs(net.mean - data.mean(axis=(0, 1))).max() < 1e-5
    assert np.abs(np.sqrt(net.M2 / net.counter) - data.std(axis=(0, 1))).max() < 1e-5
    # a second merge equals the moments of the concatenation (Chan merge)
    data2 = (3 + 2 * g.standard_normal((7, 16, 8))).astype(np.float32)
    net.update_statistics(data2)
    both = np.concatenate([data.reshape(-1, 8), data2.reshape(-1, 8)])
    assert np.abs(net.mean - both.mean(0)).max() < 1e-5
    assert np.abs(np.sqrt(net.M2 / net.counter) - both.std(0)).max() < 2e-5


def test_replay_reproduces_actions_and_loglik():
    """adapter_test.py:61-75: feeding the emitted raw action back reproduces action and loglik."""
    net = nets.make_mlp_actor_critic(5, 3, [16, 16], [16], seed=0)
    obs = np.ones((4, 5), np.float32)
    o1 = nets.policy_forward(net, obs)
    o2 = nets.policy_forward(net, obs, raw_action=o1["raw_action"])
    assert np.allclose(o1["action"], o2["action"]) and np.allclose(o1["loglik"], o2["loglik"])


def _torch_loss(net, ro, inds, eps2, clip=0.2, gamma=0.99, lam=0.95, cw=1.0, act="relu"):
    """Independent float64 autograd restatement of ppo_loss (ppo.py:397-531)."""
    dt = torch.float64
    T = ro.obs.shape[0]
    mb = len(inds)
    A = net.act_dim
    N = T * mb
    params = [torch.tensor(a, dtype=dt, requires_grad=True) for ch in (net.actor, net.critic)
              for pair in zip(ch.W, ch.b) for a in pair]
    it = iter(params)
    actf = {"relu": torch.relu, "tanh": torch.tanh, "swish": lambda t: t * torch.sigmoid(t)}[act]

    def chain(x, n):
        for l in range(n):
            W, b = next(it), next(it)
            x = x @ W + b
            if l < n - 1:
                x = actf(x)
        return x
    mean = torch.tensor(net.mean, dtype=dt)
    std = torch.tensor(net.norm_std(), dtype=dt)
    x = (torch.tensor(ro.obs[:, inds].reshape(N, -1), dtype=dt) - mean) / std
    xl = (torch.tensor(ro.next_obs_last[inds], dtype=dt) - mean) / std
    y = chain(x, net.actor.n_layers)
    itc = list(params[2 * net.actor.n_layers:])

    def critic(xx):
        h = xx
        n = net.critic.n_layers
        for l in range(n):
            h = h @ itc[2 * l] + itc[2 * l + 1]
            if l < n - 1:
                h = actf(h)
        return h[:, 0]
    v, v_last = critic(x), critic(xl)
    mu, rho = y[:, :A], y[:, A:]
    sigma = (torch.nn.functional.softplus(rho) + net.min_std) * net.std_scale
    z = torch.tensor(ro.raw_action[:, inds].reshape(N, A), dtype=dt)
    ldj = lambda q: 2.0 * (np.log(2.0) - q - torch.nn.functional.softplus(-2.0 * q))
    ll = (-0.5 * ((z - mu) / sigma) ** 2 - (0.5 * np.log(2 * np.pi) + torch.log(sigma)) - ldj(z)).sum(-1)
    zp = mu + sigma * torch.tensor(eps2.reshape(N, A), dtype=dt)
    ent = (0.5 + 0.5 * np.log(2 * np.pi) + torch.log(sigma) + ldj(zp)).sum(-1)
    reg = -net.entropy_weight * ent
    with torch.no_grad():
        vv = torch.cat([v.reshape(T, mb), v_last.reshape(1, mb)]).numpy()
        adv = ppo.gae(ro.reward[:, inds].astype(np.float64), vv[:-1], vv[-1], ro.done[:, inds],
                      ro.truncated[:, inds], lam, gamma, dtype=np.float64).reshape(N)
        adv_t = torch.tensor(adv, dtype=dt)
        target = v.detach() + adv_t
        adv_n = (adv_t - adv_t.mean()) / (adv_t.std(unbiased=False) + 1e-8)
    ratio = torch.exp(ll - torch.tensor(ro.loglik[:, inds].reshape(N), dtype=dt))
    la = -torch.minimum(ratio * adv_n, torch.clip(ratio, 1 - clip, 1 + clip) * adv_n).mean()
    lc = 0.5 * ((v - target) ** 2).mean()
    total = la + cw * lc + reg.mean()
    total.backward()
    return float(total), np.concatenate([p.grad.numpy().ravel() for p in params])


@pytest.mark.parametrize("act", ["relu", "tanh", "swish"])
def test_analytic_gradients_match_float64_autograd(act):
    env = oenv.SyntheticEnv(12, 3, max_len=16, term_thresh16=3000)
    net = nets.make_mlp_actor_critic(12, 3, [32, 32], [48, 48], seed=3, activation=act)
    ts = ppo.new_training_state(env, net, 64, 5)
    # run one iteration so the normalizer has statistics, params moved, and policies diverged
    ts, _ = ppo.ppo_step(env, ts, 64, 12, n_epochs=2, n_minibatches=2, learning_rate=3e-3)
    reset_key, new_key = prng.split(ts.rng_key)
    _, ro = ppo.unroll_env(env, ts.env_state, net, 12, reset_key)
    # perturb old log-probs so that a good fraction of samples is clipped
    g = np.random.default_rng(0)
    ro.loglik += (0.3 * g.standard_normal(ro.loglik.shape)).astype(np.float32)
    inds = np.arange(0, 64, 2, dtype=np.int32)
    total, m, grads = ppo.ppo_loss_and_grads(net, ro, inds, net.rng_count)
    ref_total, ref_grads = _torch_loss(net, ro, inds, m["eps2"], act=act)
    assert abs(total - ref_total) < 2e-5
    scale = np.abs(ref_grads).max()
    assert np.abs(grads - ref_grads).max() < 2e-5 * max(scale, 1.0)
    assert m["d_y"].shape == (12 * 32, 6)


def test_adam_matches_closed_form_first_step():
    p = np.array([1.0, -2.0, 0.5], np.float32)
    g = np.array([0.1, -0.2, 0.0], np.float32)
    st = ppo.AdamState(np.zeros(3, np.float32), np.zeros(3, np.float32))
    p2 = ppo.adam_update(p, g, st, lr=1e-2)
    # first step: m_hat = g, v_hat = g^2 -> update = g / (|g| + eps)
    assert np.allclose(p2, p - 1e-2 * g / (np.abs(g) + 1e-8), atol=1e-7)
    assert st.count == 1


def test_ppo_step_bookkeeping():
    """ppo_test.py:307-349 style: steps_taken == n*T*B, Normalizer counter == n*T*B, finite."""
    env = oenv.SyntheticEnv(6, 2, max_len=10, term_thresh16=2000)
    net = nets.make_mlp_actor_critic(6, 2, [16, 16], [16, 16], seed=22)
    ts = ppo.new_training_state(env, net, 32, 22)
    c0 = net.rng_count
    for n in range(1, 4):
        tr = {}
        ts, metrics = ppo.ppo_step(env, ts, 32, 8, n_epochs=2, n_minibatches=4, trace=tr)
        assert float(ts.steps_taken) == n * 8 * 32
        assert float(net.counter) == n * 8 * 32
        assert all(np.isfinite(v) for v in metrics.values())
        assert net.rng_count == c0 + n * (2 * 8 + 2 * 4 * 2 * 9)      # R-RNG row of SURVEY §8a
        ro = tr["rollout"]
        assert ro.truncated.sum() > 0 or n == 1
        assert np.all(ro.done[ro.truncated])                           # truncated ⇒ done
        assert sorted(tr["indices"][:4].ravel().tolist()) == list(range(32))


def test_eval_rollout_bookkeeping_against_per_env_simulation():
    """rollout.py:97-148: the batched scan against a literal one-env-at-a-time restatement
    (deterministic sampler, so an env's trajectory does not depend on the batch it is in): the reward
    of the terminal step counts, later ones do not; lifespan = steps before the first done."""
    O, A, B, L = 6, 2, 24, 20
    env = oenv.SyntheticEnv(O, A, max_len=12, term_thresh16=4000)
    net = nets.make_mlp_actor_critic(O, A, [16], [16], seed=3)
    key = prng.key(11)
    c0 = net.rng_count                                                   # the init draws of the shared stream
    cuml, life = ppo.eval_rollout(env, net, B, L, key, deterministic=True)
    assert net.rng_count == c0 + L                                       # one (entropy) draw per call
    keys = prng.split(key, B)
    s0 = env.reset_fast(keys)
    for b in range(B):
        s = oenv.EnvState(s0.obs[b:b + 1], s0.step_counter[b:b + 1], s0.term_state[b:b + 1],
                          s0.reward[b:b + 1], s0.done[b:b + 1], s0.truncated[b:b + 1])
        total, steps, dead = np.float32(0), 0, False
        for _ in range(L):
            out = nets.policy_forward(net, s.obs, deterministic=True)
            s = env.step(s, out["action"])
            if not dead:
                total = np.float32(total + s.reward[0])
            dead = dead or bool(s.done[0])
            if not dead:
                steps += 1
        assert life[b] == steps
        assert abs(cuml[b] - total) < 1e-5
    assert life.max() <= 11 and life.min() >= 0 and len(set(life.tolist())) > 1
