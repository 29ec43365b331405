"""CPU checks of the recurrent (LSTM) restatement, oracle/recurrent.py — SURVEY section 8 row a15.

The reference's own tests for this path only assert finiteness / shapes (recurrent_test.py:232-330),
so the restatement is pinned internally: the analytic BPTT gradient against torch.autograd in
float64, the reset semantics against a hand-rolled loop, and the reference's shape / zero-reg facts.
"""
import numpy as np
import pytest
import torch

from oracle import env as oenv, prng, recurrent as orec
from oracle.nets import loglikelihood, sampler_std

F = np.float32


def _setup(seed=3, O=10, A=3, B=12, T=9, H=8, trainable=False):
    net = orec.make_recurrent_actor_critic(O, A, [7], H, [], [6], seed=seed, activation="tanh",
                                           trainable_initial_state=trainable)
    if trainable:                   # the reference starts them at zero; non-zero values make the checks meaningful
        g = np.random.default_rng(9)
        net.init_c[:] = 0.5 * g.standard_normal(H)
        net.init_h[:] = 0.5 * g.standard_normal(H)
    e = oenv.SyntheticEnv(O, A, max_len=6, term_thresh16=9000)      # frequent resets inside T steps
    k = prng.key(5)
    es = e.reset(prng.split(k, B))
    carry = net.initialize_state(B)
    es, carry, ro, start = orec.unroll_env(e, es, net, carry, T, prng.key(11))
    return net, e, es, carry, ro, start


def test_lstm_cell_matches_flax_formula():
    """Gate order (i, f, g, o), c' = f c + i g, h' = o tanh(c') — flax OptimizedLSTMCell."""
    rng = np.random.default_rng(0)
    H, I, B = 5, 4, 3
    p = orec.LSTMParams(rng.normal(size=(I, 4 * H)).astype(F), rng.normal(size=(H, 4 * H)).astype(F),
                        rng.normal(size=4 * H).astype(F))
    c = rng.normal(size=(B, H)).astype(F); h = rng.normal(size=(B, H)).astype(F); x = rng.normal(size=(B, I)).astype(F)
    c2, h2, _ = orec.lstm_step(p, c, h, x)
    sg = lambda v: 1 / (1 + np.exp(-v))
    a = x.astype(np.float64) @ p.Wi + h.astype(np.float64) @ p.Wh + p.b
    i, f, g, o = sg(a[:, :H]), sg(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]), sg(a[:, 3 * H:])
    cr = f * c + i * g
    assert np.allclose(c2, cr, atol=1e-6) and np.allclose(h2, o * np.tanh(cr), atol=1e-6)


def test_rollout_resets_carry_and_runs_finite():
    """recurrent_test.py:232-283: resets occur, outputs finite; plus: the carry after a done step is zero."""
    net, e, es, carry, ro, start = _setup()
    assert ro.done.sum() > 0 and np.all(np.isfinite(ro.loglik)) and np.all(np.isfinite(ro.value))
    assert np.all(start[0] == 0) and np.all(start[1] == 0)
    d_last = ro.done[-1]
    assert np.all(carry[0][d_last] == 0) and np.all(carry[1][d_last] == 0)
    assert np.any(carry[1][~d_last] != 0)
    # init draws: pre Dense 2, LSTM's two kernels 2, post Dense 2, critic 2 x 2; then 2 sampler draws per step
    assert net.rng_count == 2 + 2 + 2 + 4 + 2 * ro.obs.shape[0]


def _torch_loss(net, ro, start, inds, base, flat64):
    """float64 autograd replay of the same loss as a function of the flat parameter vector."""
    T = ro.obs.shape[0]
    mb = len(inds)
    A = net.act_dim
    t64 = lambda a: torch.tensor(np.asarray(a, np.float64))
    shapes = [p.shape for p in net.param_list()]
    ps, o = [], 0
    for sh in shapes:
        n = int(np.prod(sh))
        ps.append(flat64[o:o + n].reshape(sh))
        o += n
    npre, npost = net.pre.n_layers, net.post.n_layers
    pre = ps[:2 * npre]
    Wi, Wh, b = ps[2 * npre:2 * npre + 3]
    ni = 2 if net.init_c is not None else 0
    init = ps[2 * npre + 3:2 * npre + 3 + ni]
    post = ps[2 * npre + 3 + ni:2 * npre + 3 + ni + 2 * npost]
    crit = ps[2 * npre + 3 + ni + 2 * npost:]
    act = torch.tanh
    H = net.lstm.hidden
    x_all = t64(net.normalize_obs(ro.obs[:, inds].reshape(T * mb, -1))).reshape(T, mb, -1)
    done = torch.tensor(ro.done[:, inds])
    c, h = t64(start[0][inds]), t64(start[1][inds])
    ys = []
    for t in range(T):
        u = x_all[t]
        for l in range(npre):
            u = act(u @ pre[2 * l] + pre[2 * l + 1])
        a = u @ Wi + h @ Wh + b
        i, f, g, o_ = torch.sigmoid(a[:, :H]), torch.sigmoid(a[:, H:2 * H]), torch.tanh(a[:, 2 * H:3 * H]), torch.sigmoid(a[:, 3 * H:])
        c = f * c + i * g
        h = o_ * torch.tanh(c)
        y = h
        for l in range(npost):
            y = y @ post[2 * l] + post[2 * l + 1]
            if l < npost - 1:
                y = act(y)
        ys.append(y)
        keep = (~done[t]).double()[:, None]
        if ni:
            c, h = c * keep + (1 - keep) * init[0][None, :], h * keep + (1 - keep) * init[1][None, :]
        else:
            c, h = c * keep, h * keep
    y = torch.cat(ys, 0)

    def critic(x):
        hh = x
        nl = len(crit) // 2
        for l in range(nl):
            hh = hh @ crit[2 * l] + crit[2 * l + 1]
            if l < nl - 1:
                hh = act(hh)
        return hh[:, 0]
    v = critic(x_all.reshape(T * mb, -1))
    v_last = critic(t64(net.normalize_obs(ro.next_obs_last[inds])))
    mu, rho = y[:, :A], y[:, A:]
    sigma = (torch.nn.functional.softplus(rho) + net.min_std) * net.std_scale
    z = t64(ro.raw_action[:, inds].reshape(T * mb, A))
    LOG2, HL = np.log(2.0), 0.5 * np.log(2 * np.pi)
    ldj = lambda q: 2.0 * (LOG2 - q - torch.nn.functional.softplus(-2.0 * q))
    ll = (-0.5 * ((z - mu) / sigma) ** 2 - HL - torch.log(sigma) - ldj(z)).sum(1)
    eps2 = np.stack([prng.normal(prng.fold_in(net.rng_key, base + 2 * t + 1), (mb, A)) for t in range(T)]).reshape(T * mb, A)
    zp = mu + sigma * t64(eps2)
    ent = (0.5 + HL + torch.log(sigma) + ldj(zp)).sum(1)
    # GAE on detached values (stop_gradient), float64
    vd, vl = v.detach().reshape(T, mb), v_last.detach()
    rew, dn, tr = t64(ro.reward[:, inds]), torch.tensor(ro.done[:, inds]), torch.tensor(ro.truncated[:, inds])
    adv = torch.zeros(T, mb, dtype=torch.float64)
    nxt_adv, nxt_v = torch.zeros(mb, dtype=torch.float64), vl
    for t in reversed(range(T)):
        nv = torch.where(dn[t], torch.zeros_like(nxt_v), nxt_v)
        delta = rew[t] + 0.99 * nv - vd[t]
        delta = torch.where(tr[t], torch.zeros_like(delta), delta)
        nxt_adv = delta + (~dn[t]).double() * 0.99 * 0.95 * nxt_adv
        adv[t] = nxt_adv
        nxt_v = vd[t]
    adv = adv.reshape(-1)
    target = (v.detach() + adv)
    an = (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-8)
    ratio = torch.exp(ll - t64(ro.loglik[:, inds].reshape(-1)))
    actor = -torch.minimum(ratio * an, torch.clamp(ratio, 0.8, 1.2) * an).mean()
    crit_l = 0.5 * ((v - target) ** 2).mean()
    reg = (-net.entropy_weight * ent).mean()
    return actor + crit_l + reg


def test_trainable_initial_state_resets_and_gradient():
    """recurrent.py:85-87, 135-141, 154-157 (recurrent_test.py:95-150: the state is the learned vector broadcast
    over the batch, at initialisation and after a reset) and the gradient those parameters get through the
    resets inside the replayed window, against float64 autograd."""
    net, e, es, carry, ro, start = _setup(trainable=True)
    assert np.all(start[0] == net.init_c[None, :]) and np.all(start[1] == net.init_h[None, :])
    d_last = ro.done[-1]
    assert d_last.any() and np.all(carry[0][d_last] == net.init_c) and np.all(carry[1][d_last] == net.init_h)
    inds = np.array([0, 3, 4, 7, 9, 11])
    assert ro.done[:-1, inds].any()
    base = net.rng_count
    total, m, g = orec.ppo_loss_and_grads(net, ro, start, inds, base)
    flat = torch.tensor(net.flat_params().astype(np.float64), requires_grad=True)
    loss = _torch_loss(net, ro, start, inds, base, flat)
    loss.backward()
    gt = flat.grad.numpy()
    assert abs(float(loss) - float(total)) < 2e-5 * max(1.0, abs(float(loss)))
    scale = np.abs(gt).max()
    assert np.abs(g - gt).max() < 2e-4 * scale, (np.abs(g - gt).max(), scale)
    H = net.lstm.hidden
    o = sum(p.size for p in net.param_list()[:2 * net.pre.n_layers + 3])
    assert np.abs(gt[o:o + 2 * H]).max() > 1e-3 * scale               # the learned carry does get gradient
    # the start carry is data (types.py:50-56 network_states), not a function of the parameters
    net2, *_ = _setup(trainable=True)
    ro.done[:] = False
    _t, _m, g3 = orec.ppo_loss_and_grads(net, ro, start, inds, base)
    assert np.all(g3[o:o + 2 * H] == 0)


def test_bptt_gradient_matches_autograd():
    net, e, es, carry, ro, start = _setup()
    inds = np.array([0, 3, 4, 7, 9, 11])
    base = net.rng_count
    total, m, g = orec.ppo_loss_and_grads(net, ro, start, inds, base)
    flat = torch.tensor(net.flat_params().astype(np.float64), requires_grad=True)
    loss = _torch_loss(net, ro, start, inds, base, flat)
    loss.backward()
    gt = flat.grad.numpy()
    assert abs(float(loss) - float(total)) < 2e-5 * max(1.0, abs(float(loss)))
    assert g.shape == gt.shape
    scale = np.abs(gt).max()
    assert np.abs(g - gt).max() < 2e-4 * scale, (np.abs(g - gt).max(), scale)
    # no gradient crosses an episode boundary: with every step done, the recurrent kernel gets none
    ro2 = ro
    ro2.done[:] = True
    _t, _m, g2 = orec.ppo_loss_and_grads(net, ro2, start, inds, base)
    o = sum(p.size for p in net.param_list()[:2 * net.pre.n_layers]) + net.lstm.Wi.size
    assert np.all(g2[o:o + net.lstm.Wh.size] == 0)


def test_recurrent_ppo_step_bookkeeping():
    """recurrent_test.py:285-330 (ppo_step with an LSTM actor): parameters move, losses finite,
    counters follow the same rules as the MLP path (ppo_test.py:307-349)."""
    O, A, B, T, E, M, H = 10, 3, 16, 12, 2, 2, 8
    net = orec.make_recurrent_actor_critic(O, A, [7], H, [], [6], seed=1)
    e = oenv.SyntheticEnv(O, A, max_len=8, term_thresh16=4000)
    ts = orec.new_training_state(e, net, B, 17)
    p0 = net.flat_params().copy()
    c0 = net.rng_count
    for it in range(2):
        ts, m = orec.ppo_step(e, ts, B, T, n_epochs=E, n_minibatches=M)
    assert all(np.isfinite(float(v)) for v in m.values())
    assert float(ts.steps_taken) == 2 * T * B == float(net.counter)
    assert ts.opt.count == 2 * E * M
    assert net.rng_count - c0 == 2 * (2 * T + E * M * 2 * (T + 1))
    assert np.abs(net.flat_params() - p0).max() > 0
    assert ts.carry[0].shape == (B, H)
