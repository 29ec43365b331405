"""Plan compiler: StatefulModule tree -> flat POD plan + device parameter arena.

The reference evaluates the module tree op by op under jit (containers.py:18-39,
adapter.py:75-117).  Here the tree of ``make_mlp_actor_critic`` (factories.py:72-146) is
recognised structurally and lowered to a ``b200ppo_plan``: layer sizes, activation, offsets into
ONE flat float32 parameter arena (so the optimizer, the gradient all-reduce and the kernels all
see a single contiguous buffer), plus the sampler's hyper-parameters.  Unsupported topologies
raise — there is no eager / PyTorch fallback.
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from .. import _lib
from .adapter import PPOAdapter
from .containers import Sequential
from .feedforward import Dense
from .normalizer import Normalizer
from .sampling_layers import NormalTanhSampler
from .types import PPONetworkOutput, StatefulModule, StatefulModuleOutput


def _align4(n: int) -> int:
    return (n + 3) & ~3


class CompiledNet:
    """Device-side view of one actor-critic network."""
    recurrent = False

    def __init__(self, network: StatefulModule, device):
        import torch
        self.network = network
        self.device = device
        normalizer: Optional[Normalizer] = None
        adapter = network
        if isinstance(network, Sequential):
            layers = list(network.layers)
            if len(layers) == 2 and isinstance(layers[0], Normalizer) and isinstance(layers[1], PPOAdapter):
                normalizer, adapter = layers
            elif len(layers) == 1 and isinstance(layers[0], PPOAdapter):
                adapter = layers[0]
            else:
                raise NotImplementedError(
                    "unsupported network topology: expected Sequential([Normalizer, PPOAdapter]) "
                    "or PPOAdapter (the MLP plan of make_mlp_actor_critic)")
        if not isinstance(adapter, PPOAdapter):
            raise NotImplementedError("unsupported network topology: no PPOAdapter found")
        action = adapter.action
        if not isinstance(action, Sequential) or not isinstance(action.layers[-1], NormalTanhSampler):
            raise NotImplementedError("action port must be Sequential([Dense..., NormalTanhSampler])")
        actor_layers = list(action.layers[:-1])
        self.sampler: NormalTanhSampler = action.layers[-1]
        value = adapter.value
        critic_layers = list(value.layers) if isinstance(value, Sequential) else [value]
        for l in actor_layers + critic_layers:
            if not isinstance(l, Dense):
                raise NotImplementedError(f"unsupported layer {type(l).__name__} in the MLP plan")
        self.normalizer = normalizer
        self.adapter = adapter
        self.actor_layers, self.critic_layers = actor_layers, critic_layers

        plan = _lib.Plan()
        off = 0

        def fill(chain, layers, out_dim_expected=None):
            nonlocal off
            if not 1 <= len(layers) <= _lib.MAX_LAYERS:
                raise NotImplementedError(f"MLP depth must be 1..{_lib.MAX_LAYERS}")
            chain.n_layers = len(layers)
            acts = {l.activation_name for l in layers[:-1]}
            if len(acts) > 1:
                raise NotImplementedError("all hidden layers of one MLP must share an activation")
            if layers[-1].activation_name != "none":
                raise NotImplementedError("the last layer of an actor / critic MLP must be linear")
            chain.act = _lib.ACT_IDS[acts.pop() if acts else "none"]
            chain.dims[0] = layers[0].in_features
            for i, l in enumerate(layers):
                if l.in_features != chain.dims[i]:
                    raise ValueError("layer sizes do not chain")
                chain.dims[i + 1] = l.out_features
                chain.w_off[i] = off
                off = _align4(off + l.in_features * l.out_features)
                chain.b_off[i] = off
                off = _align4(off + l.out_features)

        fill(plan.actor, actor_layers)
        fill(plan.critic, critic_layers)
        plan.obs_dim = actor_layers[0].in_features
        if plan.actor.dims[plan.actor.n_layers] % 2:
            raise ValueError("actor output must be 2 * action_size")
        plan.act_dim = plan.actor.dims[plan.actor.n_layers] // 2
        if critic_layers[0].in_features != plan.obs_dim or plan.critic.dims[plan.critic.n_layers] != 1:
            raise NotImplementedError("critic must map obs -> 1")
        plan.normalize = 1 if normalizer is not None else 0
        if normalizer is not None and normalizer.size != plan.obs_dim:
            raise ValueError("Normalizer size != obs size")
        plan.entropy_weight = float(self.sampler.entropy_weight)
        plan.min_std = float(self.sampler.min_std)
        plan.std_scale = float(self.sampler.std_scale)
        plan.n_params = off
        self.plan = plan
        self.n_params = off

        # one flat arena; every Dense parameter becomes a view into it
        host = np.zeros(off, np.float32)
        for chain, layers in ((plan.actor, actor_layers), (plan.critic, critic_layers)):
            for i, l in enumerate(layers):
                w, b = l.linear.kernel.numpy(), l.linear.bias.numpy()
                host[chain.w_off[i]:chain.w_off[i] + w.size] = w.ravel()
                host[chain.b_off[i]:chain.b_off[i] + b.size] = b.ravel()
        self.arena = torch.from_numpy(host).to(device)
        for chain, layers in ((plan.actor, actor_layers), (plan.critic, critic_layers)):
            for i, l in enumerate(layers):
                kin, kout = l.in_features, l.out_features
                l.linear.kernel._dev = self.arena[chain.w_off[i]:chain.w_off[i] + kin * kout].view(kin, kout)
                l.linear.bias._dev = self.arena[chain.b_off[i]:chain.b_off[i] + kout]
        if normalizer is not None:
            normalizer._bind(device)
        # counters: [0..1] sampler stream key, [2] sampler count, [3] adam count
        rng = self.sampler.rng
        self._counters_host = np.array([rng.key[0], rng.key[1], rng.count, 0], np.uint32)
        self.counters = torch.from_numpy(self._counters_host.view(np.int32).copy()).to(device)
        self.engines: dict = {}
        self.adam_step = 0   # host mirror of counters[3] (optimizer step count)

    # ---- flat <-> logical parameter order (actor W0,b0,..., critic W0,b0,...; no padding) ----
    def logical_slices(self):
        out = []
        for chain, layers in ((self.plan.actor, self.actor_layers), (self.plan.critic, self.critic_layers)):
            for i, l in enumerate(layers):
                out.append((int(chain.w_off[i]), l.in_features * l.out_features))
                out.append((int(chain.b_off[i]), l.out_features))
        return out

    def params_logical(self, arena=None) -> np.ndarray:
        a = (self.arena if arena is None else arena).detach().cpu().numpy()
        return np.concatenate([a[o:o + n] for o, n in self.logical_slices()])

    def load_params_logical(self, flat: np.ndarray) -> None:
        import torch
        host = np.zeros(self.n_params, np.float32)
        p = 0
        for o, n in self.logical_slices():
            host[o:o + n] = flat[p:p + n]
            p += n
        self.arena.copy_(torch.from_numpy(host))

    # ---- sampler stream bookkeeping (host mirror of counters[2]) ----
    @property
    def rng_count(self) -> int:
        return int(self.sampler.rng.count)

    def advance_rng(self, n: int) -> None:
        self.sampler.rng.count = (self.sampler.rng.count + n) & 0xFFFFFFFF

    def sync_counters_to_device(self) -> None:
        import torch
        self._counters_host[2] = self.sampler.rng.count
        self._counters_host[3] = self.adam_step
        self.counters.copy_(torch.from_numpy(self._counters_host.view(np.int32).copy()))

    def norm_ptrs(self):
        if self.normalizer is None:
            return 0, 0
        return _lib.ptr(self.normalizer.mean._dev), _lib.ptr(self.normalizer._std)


def compile_network(network: StatefulModule, device=None) -> CompiledNet:
    import torch
    c = getattr(network, "_b200_compiled", None)
    if c is not None:
        return c
    if device is None:
        if not torch.cuda.is_available():
            raise _lib.B200PPOError("no CUDA device: the B200 PPO path has no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    _lib.load()
    from . import rplan
    c = rplan.RecurrentCompiledNet(network, device) if rplan.is_recurrent(network) else CompiledNet(network, device)
    network._b200_compiled = c
    return c


def call_network(network: StatefulModule, state: Any, obs: Any, rollout_extras: Any = None):
    """``networks(state, obs, rollout_extras)`` for a plan-compilable actor-critic: one launch of
    the fused policy-step kernel (K1).  Mirrors the output structure of containers.py:18-39 /
    adapter.py:100-117: rollout_extras = [raw_obs, {"action": [None..., raw_action], "value": [...]}]."""
    import torch
    if not isinstance(obs, torch.Tensor):
        raise TypeError("observations must be a CUDA float32 torch tensor [B, obs_size]")
    net = compile_network(network, obs.device)
    lib = _lib.load()
    obs = obs.contiguous().float()
    _lib.require_cuda(obs)
    if net.recurrent:
        return _call_recurrent(net, state, obs, rollout_extras)
    B = obs.shape[0]
    A = net.plan.act_dim
    raw_in = None
    if rollout_extras is not None:
        extras = rollout_extras
        if net.normalizer is not None:
            extras = extras[1]
        raw_in = extras["action"][-1].contiguous().float()
    mode = (1 if raw_in is not None else 0) | (2 if net.sampler.deterministic else 0)
    dev = obs.device
    raw = torch.empty(B, A, device=dev)
    action = torch.empty(B, A, device=dev)
    ll = torch.empty(B, device=dev)
    value = torch.empty(B, device=dev)
    reg = torch.empty(B, device=dev)
    s = _lib.current_stream()
    if net.normalizer is not None:
        net.normalizer.prepare(s)
    net.sync_counters_to_device()
    mean_p, std_p = net.norm_ptrs()
    _lib.check(lib.b200ppo_policy_step(s, net.plan, _lib.ptr(net.arena), mean_p, std_p, _lib.ptr(obs), B,
                                       mode, _lib.ptr(net.counters), 0, _lib.ptr(raw_in), _lib.ptr(raw),
                                       _lib.ptr(action), _lib.ptr(ll), _lib.ptr(value), _lib.ptr(reg), 0),
               "policy_step")
    net.advance_rng(1 if net.sampler.deterministic else 2)
    na, nc = len(net.actor_layers), len(net.critic_layers)
    adapter_state = {"action": [()] * (na + 1), "value": [()] * nc}
    adapter_extras = {"action": [None] * na + [raw], "value": [None] * nc}
    out = PPONetworkOutput(actions=action, loglikelihoods=ll, value_estimates=value)
    if net.normalizer is not None:
        return StatefulModuleOutput([(), adapter_state], out, reg, {}, [obs, adapter_extras])
    return StatefulModuleOutput(adapter_state, out, reg, {}, adapter_extras)


def _call_recurrent(net, state: Any, obs, rollout_extras: Any):
    """One call of a recurrent actor-critic (networks/rplan.py): recurrent step kernel + sampler +
    critic.  Functional like the reference: the input state is not modified, the new carry is in
    ``next_state`` at the LSTM's position of the reference-shaped state pytree."""
    import copy
    import torch
    lib = _lib.load()
    B, A, Y = obs.shape[0], net.plan.act_dim, net.lplan.out_dim
    dev = obs.device
    c, h = net.get_carry(state)
    c, h = c.clone().contiguous(), h.clone().contiguous()
    raw_in = None
    if rollout_extras is not None:
        extras = rollout_extras[1] if net.normalizer is not None else rollout_extras
        raw_in = extras["action"][-1].contiguous().float()
    mode = (1 if raw_in is not None else 0) | (2 if net.sampler.deterministic else 0)
    y = torch.empty(B, Y, device=dev)
    raw, action = torch.empty(B, A, device=dev), torch.empty(B, A, device=dev)
    ll, reg = torch.empty(B, device=dev), torch.empty(B, device=dev)
    s = _lib.current_stream()
    if net.normalizer is not None:
        net.normalizer.prepare(s)
    net.sync_counters_to_device()
    mean_p, std_p = net.norm_ptrs()
    _lib.check(lib.b200ppo_lstm_step_fwd(s, net.lplan, _lib.ptr(net.arena), mean_p, std_p, _lib.ptr(obs), 0, 0, B,
                                         _lib.ptr(c), _lib.ptr(h), _lib.ptr(y), 0), "lstm_step_fwd")
    _lib.check(lib.b200ppo_sampler_step(s, _lib.ptr(y), B, A, mode, net.plan.min_std, net.plan.std_scale,
                                        net.plan.entropy_weight, _lib.ptr(net.counters), 0, _lib.ptr(raw_in),
                                        _lib.ptr(raw), _lib.ptr(action), _lib.ptr(ll), _lib.ptr(reg)), "sampler_step")
    from ..algorithms.rollout import policy_values
    value = policy_values(net, obs)
    net.advance_rng(1 if net.sampler.deterministic else 2)
    nc = len(net.critic_layers)
    adapter_state = {"action": [(), (c, h), (), ()], "value": [()] * nc}
    adapter_extras = {"action": [None, None, None, raw], "value": [None] * nc}
    out = PPONetworkOutput(actions=action, loglikelihoods=ll, value_estimates=value)
    del copy
    if net.normalizer is not None:
        return StatefulModuleOutput([(), adapter_state], out, reg, {}, [obs, adapter_extras])
    return StatefulModuleOutput(adapter_state, out, reg, {}, adapter_extras)
