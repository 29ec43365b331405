"""Python control flow of the per-step path (networks/plan.py::call_network,
algorithms/rollout.py::_unroll_generic) on CPU tensors with the kernel launches stubbed out: the
reference-shaped state / rollout_extras pytrees (containers.py:18-39, adapter.py:100-117), observation
adapters, RNG-count bookkeeping, the Transition record (rollout.py:11-45).  The arithmetic itself is
the `-m gpu` tests' business."""
import dataclasses
import types

import numpy as np
import pytest
import torch

from nnx_ppo_b200 import _lib, prng
from nnx_ppo_b200.algorithms import rollout
from nnx_ppo_b200.networks import factories, plan as plan_mod
from nnx_ppo_b200.networks.containers import Sequential
from nnx_ppo_b200.networks.normalizer import Normalizer
from nnx_ppo_b200.networks.plan import CompiledNet
from nnx_ppo_b200.networks.utils import Filter, Flattener


@pytest.fixture
def stubbed(monkeypatch):
    """No device: the policy-step launch records its arguments and returns 0."""
    calls = []

    def policy_step(stream, plan, params, mean, std, obs, B, mode, counters, off, raw_in, *outs):
        calls.append(dict(B=B, mode=mode, replay=bool(raw_in)))
        return 0

    fake = types.SimpleNamespace(b200ppo_policy_step=policy_step)
    monkeypatch.setattr(plan_mod._lib, "load", lambda: fake)
    monkeypatch.setattr(plan_mod._lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(plan_mod._lib, "current_stream", lambda: 0)
    monkeypatch.setattr(Normalizer, "prepare", lambda self, s=None: None)
    monkeypatch.setattr(CompiledNet, "norm_ptrs", lambda self: (0, 0))
    monkeypatch.setattr(rollout, "split_keys_device",
                        lambda key, n, dev: torch.zeros(n, 2, dtype=torch.int32))
    return calls


def _compile_cpu(nets):
    nets._b200_compiled = CompiledNet(nets, torch.device("cpu"))
    return nets._b200_compiled


def test_call_network_pytrees_plain_and_with_adapters(stubbed):
    plain = factories.make_mlp_actor_critic(9, 2, [8, 8], [8], prng.Rngs(0))
    bare = factories.make_mlp_actor_critic(9, 2, [8, 8], [8], prng.Rngs(0), normalize_obs=False)
    base = factories.make_mlp_actor_critic(9, 2, [8, 8], [8], prng.Rngs(0))
    adapted = Sequential([Filter({"p": ("arm", "proprio"), "h": "head"}), Flattener(), *base.layers])
    flat = torch.arange(18.0).reshape(2, 9)
    tree = {"arm": {"proprio": flat[:, 5:], "target": torch.zeros(2, 3)}, "head": flat[:, :5]}
    for nets, obs, n_lead in ((plain, flat, 1), (bare, flat, 0), (adapted, tree, 3)):
        net = _compile_cpu(nets)
        sampler = net.sampler
        c0 = sampler.rng.count
        out = nets(nets.initialize_state(2), obs)
        assert stubbed[-1] == dict(B=2, mode=0, replay=False) and sampler.rng.count == c0 + 2
        ad_state = {"action": [(), (), (), ()], "value": [(), ()]}
        if n_lead:
            assert out.next_state == [()] * n_lead + [ad_state]
            assert len(out.rollout_extras) == n_lead + 1
            assert out.rollout_extras[:n_lead - 1] == [None] * (n_lead - 1)
            assert torch.equal(out.rollout_extras[n_lead - 1], flat)         # the Normalizer's raw observation
        else:
            assert out.next_state == ad_state
        extras = net.adapter_extras(out.rollout_extras)
        assert extras["action"][:3] == [None] * 3 and extras["action"][3].shape == (2, 2) and extras["value"] == [None] * 2
        assert out.output.actions.shape == (2, 2) and out.output.value_estimates.shape == (2,)
        # replay with the extras of the rollout call (adapter_test.py:61-75); eval mode draws one count
        nets(nets.initialize_state(2), obs, out.rollout_extras)
        assert stubbed[-1] == dict(B=2, mode=1, replay=True)
        nets.eval()
        c1 = sampler.rng.count
        nets(nets.initialize_state(2), obs)
        nets.train()
        assert stubbed[-1]["mode"] == 2 and sampler.rng.count == c1 + 1
    with pytest.raises(TypeError):
        plain(plain.initialize_state(2), np.zeros((2, 9), np.float32))
    with pytest.raises(TypeError):
        plain(plain.initialize_state(2), {"a": flat})                         # dict obs need a Concat plan or a Flattener


@dataclasses.dataclass
class _S:
    obs: object
    reward: torch.Tensor
    done: torch.Tensor
    info: dict
    metrics: dict
    t: torch.Tensor


class _TreeEnv:
    """Batched torch env with a nested-dict observation; done after `life` steps."""

    def __init__(self, life):
        self.life = life

    def _obs(self, t):
        B = t.shape[0]
        return {"arm": {"proprio": t[:, None].expand(B, 4).clone(), "target": torch.zeros(B, 3)},
                "head": torch.ones(B, 5)}

    def reset(self, keys):
        B = keys.shape[0]
        t = torch.zeros(B)
        return _S(self._obs(t), torch.zeros(B), torch.zeros(B), {"truncated": torch.zeros(B, dtype=torch.bool)}, {}, t)

    def step(self, s, a):
        t = s.t + 1
        done = (t >= self.life).float()
        return _S(self._obs(t), t.clone(), done, {"truncated": done.bool()}, {"age": t.clone()}, t)


def test_unroll_generic_records_what_the_kernels_saw(stubbed):
    base = factories.make_mlp_actor_critic(9, 2, [8], [8], prng.Rngs(0))
    nets = Sequential([Filter({"p": ("arm", "proprio"), "h": "head"}), Flattener(), *base.layers])
    net = _compile_cpu(nets)
    env = _TreeEnv(life=3)
    B, T = 4, 7
    s0 = env.reset(torch.zeros(B, 2, dtype=torch.int32))
    _, s_end, tr = rollout._unroll_generic(env, s0, nets, nets.initialize_state(B), T, prng.key(1))
    assert len(stubbed) == T and net.sampler.rng.count == 2 * 4 + 2 * T      # 4 Linear inits (2 counts each) + 2 per step
    assert tr.obs.shape == (T, B, 9)                                          # flattened: head (5) then proprio (4)
    ages = [0, 1, 2, 0, 1, 2, 0]                                              # reset on done: tree_where over the dict obs
    assert torch.equal(tr.obs[:, 0, 5], torch.tensor(ages, dtype=torch.float32))
    assert torch.equal(tr.obs[:, :, :5], torch.ones(T, B, 5))
    assert tr.done[:, 0].tolist() == [False, False, True, False, False, True, False]
    assert torch.equal(tr.truncated, tr.done) and torch.equal(tr.rewards[:, 0], torch.tensor([1., 2, 3, 1, 2, 3, 1]))
    assert tr.next_obs.shape == (B, 9) and float(tr.next_obs[0, 5]) == 1.0     # pre-reset observation of the last step
    assert torch.equal(tr.metrics["env"]["age"], tr.rewards)
    assert tr.rollout_extras[:2] == [None, None] and torch.equal(tr.rollout_extras[2], tr.obs)
    assert net.adapter_extras(tr.rollout_extras)["action"][-1].shape == (T, B, 2)
    assert float(s_end.t[0]) == 1.0


def test_unroll_generic_plain_network_flat_observations(stubbed):
    """The topology the GPU tests run (Sequential([Normalizer, PPOAdapter]), tensor observations)."""
    class FlatEnv(_TreeEnv):
        def _obs(self, t):
            return t[:, None].expand(t.shape[0], 9).clone()

    for normalize in (True, False):
        nets = factories.make_mlp_actor_critic(9, 2, [8], [8], prng.Rngs(0), normalize_obs=normalize)
        net = _compile_cpu(nets)
        env = FlatEnv(life=2)
        s0 = env.reset(torch.zeros(3, 2, dtype=torch.int32))
        _, _, tr = rollout._unroll_generic(env, s0, nets, nets.initialize_state(3), 5, prng.key(1))
        assert tr.obs.shape == (5, 3, 9) and tr.obs[:, 0, 0].tolist() == [0, 1, 0, 1, 0]
        assert tr.done[:, 0].tolist() == [False, True, False, True, False]
        if normalize:
            assert len(tr.rollout_extras) == 2 and torch.equal(tr.rollout_extras[0], tr.obs)
        else:
            assert set(tr.rollout_extras) == {"action", "value"}
        assert net.adapter_extras(tr.rollout_extras)["action"][-1].shape == (5, 3, 2)
        assert tr.next_obs.shape == (3, 9) and float(tr.next_obs[0, 0]) == 1.0
