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


def test_distillation_api_mirrors_the_reference():
    """Names, positional order and defaults of the drop-in surface (nnx_ppo/algorithms/distillation.py:63,235,363,422;
    config.py:72-95,120-127; types.py:85-125) - the product module is importable without a GPU."""
    import dataclasses
    import inspect
    from nnx_ppo_b200.algorithms import config as cfg, distillation as d, types as t

    def params(f):
        return [(p.name, p.default if p.default is not inspect.Parameter.empty else None, p.kind.name)
                for p in inspect.signature(f).parameters.values()]

    assert [p[0] for p in params(d.distillation_step)] == [
        "env", "teacher", "distillation_state", "n_envs", "rollout_length", "n_epochs", "n_minibatches",
        "logging_level", "logging_percentiles"]
    assert [p[0] for p in params(d.distillation_loss)][:4] == ["student", "student_state", "rollout_data", "logging_level"]
    assert [p[0] for p in params(d.new_distillation_state)] == [
        "env", "teacher", "student", "n_envs", "seed", "learning_rate", "gradient_clipping", "weight_decay"]
    assert dict((p[0], p[1]) for p in params(d.new_distillation_state))["learning_rate"] == 1e-4
    td = params(d.train_distillation)
    assert [p[0] for p in td] == ["env", "teacher", "student", "config", "total_steps", "seed", "log_fn", "video_fn",
                                  "checkpoint_fn", "eval_env", "initial_state"]
    assert all(p[2] == "KEYWORD_ONLY" for p in td[4:])
    dc = cfg.DistillationConfig()
    assert [f.name for f in dataclasses.fields(dc)] == [
        "n_envs", "rollout_length", "total_steps", "learning_rate", "n_epochs", "n_minibatches", "gradient_clipping",
        "weight_decay", "logging_level", "logging_percentiles"]
    assert (dc.n_envs, dc.rollout_length, dc.total_steps, dc.n_epochs, dc.n_minibatches) == (256, 20, 512_000, 4, 4)
    tc = d.default_distillation_config()
    assert isinstance(tc, cfg.DistillationTrainConfig) and tc.seed == 17 and tc.checkpoint_every_steps == 500_000
    assert [f.name for f in dataclasses.fields(cfg.DistillationTrainResult)] == [
        "training_state", "final_metrics", "eval_history", "total_steps", "total_iterations"]
    assert [f.name for f in dataclasses.fields(t.DistillationState)] == [
        "student", "student_states", "teacher_states", "env_states", "optimizer", "rng_key", "steps_taken"]
    assert [f.name for f in dataclasses.fields(t.DistillationTransition)] == [
        "obs", "student_output", "rewards", "done", "truncated", "next_obs", "metrics", "student_rollout_extras",
        "teacher_rollout_extras"]
