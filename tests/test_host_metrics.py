"""Host-side metric assembly (algorithms/ppo.py) against a literal NumPy restatement of the reference's
``compute_metrics`` / ``_log_metric`` / ``log_weight_stats`` (nnx_ppo/algorithms/metrics.py:17-121) and of
the ``losses/*`` extras of ``ppo_loss`` (ppo.py:509-528).  CPU tensors stand in for the engine's device
buffers: the functions only reduce what the kernels wrote."""
import types

import numpy as np
import torch

from nnx_ppo_b200.algorithms import ppo
from nnx_ppo_b200.algorithms.types import LoggingLevel


def _ref_log(m, name, x, pct):
    """metrics.py:72-100."""
    if isinstance(x, dict):
        for k, v in x.items():
            _ref_log(m, f"{name}/{k}", v, pct)
    elif x.dtype == bool:
        m[name] = x.mean()
    elif not pct:
        m[f"{name}/mean"], m[f"{name}/std"] = x.mean(), x.std()
    else:
        for p, v in zip(pct, np.percentile(x, pct)):
            m[f"{name}/p{int(p)}"] = v


def _fake_engine(g, T=6, B=10, A=3, P=40):
    f = lambda *s: torch.from_numpy(g.standard_normal(s).astype(np.float32))
    net = types.SimpleNamespace(arena=f(P), param_mask=torch.from_numpy((g.random(P) > 0.2).astype(np.uint8)))
    done = torch.from_numpy((g.random((T, B)) < 0.3).astype(np.uint8))
    trunc = done * torch.from_numpy((g.random((T, B)) < 0.5).astype(np.uint8))
    return types.SimpleNamespace(net=net, reward=f(T, B), action=f(T, B, A), done=done, trunc=trunc,
                                 loglik=f(T, B), value=f(T, B),
                                 env_metrics={"env": {"speed": f(T, B), "sub": {"alive": 1.0 - done.float()}}},
                                 hp=types.SimpleNamespace(grad_clip=-1.0))


def _per_update(g, n=8):
    pu = np.zeros((n, 12), np.float32)
    pu[:, 0] = g.standard_normal(n) * 0.01            # actor
    pu[:, 1] = g.random(n) + 0.1                      # critic
    pu[:, 2] = -g.random(n) * 0.05                    # regularization
    pu[:, 3] = g.random(n) + 0.5                      # grad norm
    pu[:, 4] = g.random(n) * 0.3                      # clipping fraction
    t = g.standard_normal((n, 50))
    adv = g.standard_normal((n, 50)) * 2 + 0.3
    pu[:, 5], pu[:, 6] = t.mean(1), (t * t).mean(1)
    pu[:, 7], pu[:, 8] = adv.mean(1), (adv * adv).mean(1)
    return pu, t, adv


def test_iteration_metrics_mean_std_mode():
    g = np.random.default_rng(0)
    eng = _fake_engine(g)
    pu, t, adv = _per_update(g)
    lvl = LoggingLevel.ALL
    m = ppo._iteration_metrics(pu, eng, lvl, None)
    want = {}
    for i, k in enumerate(("losses/actor", "losses/critic", "losses/regularization")):
        _ref_log(want, k, pu[:, i], None)
    _ref_log(want, "losses/clipping_fraction", pu[:, 4], None)                         # ppo.py:514-520
    _ref_log(want, "losses/critic_R^2", 1.0 - 2.0 * pu[:, 1] / (t.var(1) + 1e-8), None)  # ppo.py:524-527
    _ref_log(want, "losses/advantages", adv, None)       # ppo.py:523: the array whose E[a], E[a^2] columns 7 / 8 hold
    _ref_log(want, "env", {"speed": eng.env_metrics["env"]["speed"].numpy(),
                           "sub": {"alive": eng.env_metrics["env"]["sub"]["alive"].numpy()}}, None)
    _ref_log(want, "rollout_batch/reward", eng.reward.numpy(), None)
    _ref_log(want, "rollout_batch/action", eng.action.numpy(), None)
    want["rollout_batch/done_rate"] = eng.done.numpy().astype(bool).mean()
    want["rollout_batch/truncation_rate"] = eng.trunc.numpy().astype(bool).mean()
    _ref_log(want, "loglikelihood", eng.loglik.numpy(), None)
    _ref_log(want, "losses/predicted_value", eng.value.numpy(), None)
    _ref_log(want, "weights", eng.net.arena.numpy()[eng.net.param_mask.numpy() != 0], None)
    assert set(m) == set(want), set(m) ^ set(want)      # no grad_norm: clipping is off and the kernel did not run
    for k, v in want.items():
        assert np.allclose(np.float64(m[k]), v, rtol=2e-5, atol=2e-6), (k, m[k], v)


def test_iteration_metrics_percentile_mode_and_grad_norm():
    g = np.random.default_rng(1)
    eng = _fake_engine(g)
    eng.hp.grad_clip = 0.5
    pu, _, _ = _per_update(g)
    pct = (0, 25, 50, 100)
    lvl = LoggingLevel.LOSSES | LoggingLevel.TRAIN_ROLLOUT_STATS | LoggingLevel.GRAD_NORM | LoggingLevel.WEIGHTS
    m = ppo._iteration_metrics(pu, eng, lvl, pct)
    want = {}
    for i, k in enumerate(("losses/actor", "losses/critic", "losses/regularization")):
        _ref_log(want, k, pu[:, i], pct)
    _ref_log(want, "rollout_batch/reward", eng.reward.numpy(), pct)
    _ref_log(want, "rollout_batch/action", eng.action.numpy(), pct)
    want["rollout_batch/done_rate"] = eng.done.numpy().astype(bool).mean()
    want["rollout_batch/truncation_rate"] = eng.trunc.numpy().astype(bool).mean()
    _ref_log(want, "weights", eng.net.arena.numpy()[eng.net.param_mask.numpy() != 0], pct)
    _ref_log(want, "grad_norm", pu[:, 3], pct)         # ppo.py:313-315: one scalar per update, logged via _log_metric
    assert set(m) == set(want), set(m) ^ set(want)
    for k, v in want.items():
        assert np.allclose(np.float64(m[k]), v, rtol=2e-5, atol=2e-6), (k, m[k], v)


def test_logging_level_none_and_basic():
    g = np.random.default_rng(2)
    eng = _fake_engine(g)
    pu, _, _ = _per_update(g)
    assert ppo._iteration_metrics(pu, eng, LoggingLevel.NONE, None) == {}
    assert set(ppo._iteration_metrics(pu, eng, LoggingLevel.BASIC, None)) == {
        f"losses/{k}/{s}" for k in ("actor", "critic", "regularization") for s in ("mean", "std")}


def test_train_ppo_cadence_matches_reference_rules(monkeypatch):
    """Host loop of train_ppo (ppo.py:169-251) with the device work stubbed out: eval / checkpoint / log
    cadence (`_should_run`, incl. the step-0 calls), stop condition, TrainResult bookkeeping."""
    from nnx_ppo_b200.algorithms.config import EvalConfig, PPOConfig, TrainConfig
    from nnx_ppo_b200.algorithms.types import TrainingState

    class Nets:
        mode = []

        def eval(self):
            self.mode.append("eval")

        def train(self):
            self.mode.append("train")

    def fake_step(env, ts, n_envs, rollout_length, *rest):
        return ts.replace(steps_taken=np.float32(ts.steps_taken + n_envs * rollout_length)), {"losses/x": 1.0}

    evals, ckpts, logs = [], [], []
    monkeypatch.setattr(ppo, "ppo_step", fake_step)
    monkeypatch.setattr(ppo.rollout, "eval_rollout",
                        lambda env, nets, n, L, key, pct: evals.append((n, L, key, pct)) or {"episode_reward/mean": 0.5})
    nets = Nets()
    cfg = TrainConfig(ppo=PPOConfig(n_envs=10, rollout_length=10, total_steps=1000,
                                    logging_level=LoggingLevel.LOSSES | LoggingLevel.THROUGHPUT),
                      eval=EvalConfig(every_steps=250, n_envs=7, max_episode_length=33), seed=5,
                      checkpoint_every_steps=400)
    ts = TrainingState(nets, None, None, None, (0, 1), np.float32(0.0))
    res = ppo.train_ppo(None, nets, cfg, log_fn=lambda m, s: logs.append((s, dict(m))),
                        checkpoint_fn=lambda st, s: ckpts.append(s), initial_state=ts)
    assert res.total_iterations == 10 and res.total_steps == 1000
    assert [e["step"] for e in res.eval_history] == [0, 300, 500, 800, 1000]
    assert ckpts == [0, 400, 800]
    assert [s for s, _ in logs] == [0] + list(range(100, 1001, 100))
    assert evals[0] == (7, 33, (0, 5), (0, 25, 50, 75, 100))          # key(config.seed), eval config
    assert nets.mode == ["eval", "train"] * 5                         # ppo.py:122,139
    assert "throughput/train_sps" in logs[1][1] and "throughput/eval_sps" in logs[0][1]
    assert "episode_reward/mean" in logs[3][1] and "episode_reward/mean" not in logs[2][1]
    # total_steps / seed overrides, eval disabled: nothing at step 0, so no step-0 log either
    logs.clear(); evals.clear()
    cfg2 = TrainConfig(ppo=PPOConfig(n_envs=10, rollout_length=10), eval=EvalConfig(enabled=False))
    res = ppo.train_ppo(None, nets, cfg2, total_steps=250, seed=9, log_fn=lambda m, s: logs.append(s), initial_state=ts)
    assert res.total_steps == 300 and res.total_iterations == 3 and evals == [] and logs == [100, 200, 300]


def test_lazy_metrics_dict_materialises_on_any_read():
    """ppo_step returns a dict that is filled in on first access (the device work is still running when it
    returns): every read path of a plain dict must force it exactly once."""
    import json
    import pickle
    from nnx_ppo_b200.algorithms.ppo import LazyMetrics
    calls = []

    def build():
        calls.append(1)
        return {"losses/actor/mean": np.float32(1.5), "total_steps": np.float32(64.0)}

    for read in (lambda m: m["total_steps"], lambda m: list(m), lambda m: len(m), lambda m: "total_steps" in m,
                 lambda m: dict(m), lambda m: {**m}, lambda m: m.items(), lambda m: m.get("x"), lambda m: repr(m),
                 lambda m: json.dumps({k: float(v) for k, v in m.items()}), lambda m: pickle.dumps(m),
                 lambda m: m.update({"a": 1}), lambda m: m.__setitem__("b", 2), lambda m: m == {}, lambda m: m | {"c": 3}):
        calls.clear()
        m = LazyMetrics(build)
        assert not calls
        read(m)
        assert calls == [1]
        assert float(m["losses/actor/mean"]) == 1.5 and calls == [1]
    m = LazyMetrics(build)
    assert dict(m) == {"losses/actor/mean": np.float32(1.5), "total_steps": np.float32(64.0)}
    assert type(pickle.loads(pickle.dumps(m))) is dict
