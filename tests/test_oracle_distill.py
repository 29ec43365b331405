"""CPU checks of the distillation restatement (oracle/distill.py; reference nnx_ppo/algorithms/distillation.py):
the analytic gradient against float64 torch autograd, the facts the reference's distillation_test.py:41-198
asserts (finite losses, the teacher untouched, the student changes), and the stream bookkeeping."""
import numpy as np
import torch

from oracle import distill, env as oenv, nets as onets, prng

F = np.float32


def _pair(act="tanh"):
    student = onets.make_mlp_actor_critic(10, 3, [16, 12], [8], seed=1, activation=act)
    teacher = onets.make_mlp_actor_critic(10, 3, [20], [8], seed=2, activation=act)
    return student, teacher


def test_distillation_gradient_matches_autograd():
    student, teacher = _pair()
    g = np.random.default_rng(0)
    T, B = 5, 12
    obs = g.standard_normal((T, B, 10)).astype(F)
    student.update_statistics(obs)
    teacher.update_statistics(2.0 * obs)
    mu_t = distill.teacher_means(teacher, obs)
    inds = np.array([0, 2, 3, 5, 8, 11], np.int32)
    base = student.rng_count
    total, m, grads = distill.distillation_loss_and_grads(student, obs, mu_t, inds, base)
    # float64 autograd of the same loss over the actor parameters
    tdt = torch.float64
    Ws = [torch.tensor(W.astype(np.float64), requires_grad=True) for W in student.actor.W]
    bs = [torch.tensor(b.astype(np.float64), requires_grad=True) for b in student.actor.b]
    N, A = T * len(inds), 3
    x = torch.tensor(student.normalize_obs(obs[:, inds].reshape(N, -1)), dtype=tdt)
    h = x
    for l in range(len(Ws)):
        h = h @ Ws[l] + bs[l]
        if l < len(Ws) - 1:
            h = torch.tanh(h)
    mu, rho = h[:, :A], h[:, A:]
    sigma = (torch.nn.functional.softplus(rho) + student.min_std) * student.std_scale
    z = torch.tensor(mu_t[:, inds].reshape(N, A), dtype=tdt)
    LOG2, HL = np.log(2.0), 0.5 * np.log(2 * np.pi)
    ldj = lambda q: 2.0 * (LOG2 - q - torch.nn.functional.softplus(-2.0 * q))
    ll = (-0.5 * ((z - mu) / sigma) ** 2 - HL - torch.log(sigma) - ldj(z)).sum(1)
    eps2 = np.stack([prng.normal(prng.fold_in(student.rng_key, base + 2 * t + 1), (len(inds), A)) for t in range(T)]).reshape(N, A)
    ent = (0.5 + HL + torch.log(sigma) + ldj(mu + sigma * torch.tensor(eps2, dtype=tdt))).sum(1)
    loss = -ll.mean() + (-student.entropy_weight * ent).mean()
    loss.backward()
    ref = np.concatenate([np.concatenate([W.grad.numpy().ravel(), b.grad.numpy().ravel()]) for W, b in zip(Ws, bs)])
    assert abs(float(loss) - float(total)) < 2e-5 * max(1.0, abs(float(loss)))
    n_actor = ref.size
    scale = np.abs(ref).max()
    assert np.abs(grads[:n_actor] - ref).max() < 2e-5 * scale
    assert np.all(grads[n_actor:] == 0)                       # the value head is not in the loss
    assert m["loglik"].shape == (T, len(inds))


def test_distillation_step_bookkeeping_and_teacher_frozen():
    student, teacher = _pair("relu")
    e = oenv.SyntheticEnv(10, 3, max_len=6, term_thresh16=5000)
    ds = distill.new_distillation_state(e, student, 16, 17)
    t_before = teacher.flat_params().copy()
    s_before = student.flat_params().copy()
    c0 = student.rng_count
    tr = {}
    nll = []
    for _ in range(3):
        ds, m = distill.distillation_step(e, teacher, ds, 16, 5, n_epochs=2, n_minibatches=2, learning_rate=3e-3, trace=tr)
        assert np.isfinite(m["losses/distillation_nll/mean"]) and np.isfinite(m["losses/regularization/mean"])
        nll.append(float(m["losses/distillation_nll/mean"]))
    assert np.array_equal(teacher.flat_params(), t_before) and teacher.counter == 0       # frozen, statistics too
    assert not np.array_equal(student.flat_params(), s_before)
    assert student.rng_count == c0 + 3 * (2 * 5 + 4 * 2 * 5) and ds.opt.count == 12
    assert float(ds.steps_taken) == 3 * 16 * 5 and float(student.counter) == 3 * 16 * 5
    assert tr["teacher_mu"].shape == (5, 16, 3) and tr["indices"].shape == (4, 8)
    # the critic's parameters never move (zero gradient, Adam's update of a zero moment is zero)
    na = sum(W.size + b.size for W, b in zip(student.actor.W, student.actor.b))
    assert np.array_equal(student.flat_params()[na:], s_before[na:])
    # on a fixed batch the updates do what distillation is for: the NLL of the teacher's means goes down
    from oracle.ppo import AdamState, adam_update
    ro, mu_t = tr["rollout"], tr["teacher_mu"]
    inds = np.arange(16, dtype=np.int32)
    opt = AdamState(np.zeros_like(s_before), np.zeros_like(s_before), 0)
    first = last = None
    for it in range(30):
        total, mm, grads = distill.distillation_loss_and_grads(student, ro.obs, mu_t, inds, 1000)
        student.set_flat_params(adam_update(student.flat_params(), grads, opt, 1e-2))
        first = float(mm["losses/distillation_nll"]) if first is None else first
        last = float(mm["losses/distillation_nll"])
    assert last < first - 0.05
