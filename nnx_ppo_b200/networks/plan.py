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
from .containers import Concat, Sequential
from .feedforward import Dense
from .normalizer import Normalizer
from .sampling_layers import NormalTanhSampler
from .types import PPONetworkOutput, StatefulModule, StatefulModuleOutput
from .utils import Filter, Flattener, tree_leaves


def _align4(n: int) -> int:
    return (n + 3) & ~3


def _flatten_trunk(modules) -> list:
    """The layer modules of a shared trunk in order (nested Sequentials opened; a Concat may lead)."""
    out = []
    for m in modules:
        if isinstance(m, Sequential):
            out += _flatten_trunk(m.layers)
        elif isinstance(m, (Dense, Concat)):
            out.append(m)
        else:
            raise NotImplementedError(f"unsupported module {type(m).__name__} in a shared trunk (Dense stacks only)")
    return out


def _module_tree(m, leaf):
    """`leaf` arranged like the state / rollout_extras pytree of module `m` (containers.py:18-39)."""
    if isinstance(m, Sequential):
        return [_module_tree(c, leaf) for c in m.layers]
    if isinstance(m, Concat):
        return {k: _module_tree(c, leaf) for k, c in m.components.items()}
    return leaf


class _VLayer:
    """One layer of a lowered chain: a plain Dense (one block) or the block-diagonal union of the
    same-depth Dense layers of a Concat's per-key encoders (blocks = [(row0, col0, Dense)])."""

    def __init__(self, blocks):
        self.blocks = blocks
        self.in_features = sum(d.in_features for _, _, d in blocks)
        self.out_features = sum(d.out_features for _, _, d in blocks)
        names = {d.activation_name for _, _, d in blocks}
        if len(names) != 1:
            raise NotImplementedError("per-key encoders must use the same activation at the same depth")
        self.activation_name = names.pop()


def _lower_chain(layers):
    """[Concat(k=Sequential([Dense...])...)?, Dense...] -> ([_VLayer...], key order, per-key input sizes)."""
    out, keys, sizes = [], None, None
    layers = list(layers)
    if layers and isinstance(layers[0], Concat):
        comps = layers[0].components
        stacks = []
        for k, c in comps.items():
            ls = list(c.layers) if isinstance(c, Sequential) else [c]
            if not ls or not all(isinstance(l, Dense) for l in ls):
                raise NotImplementedError("Concat components must be Dense stacks")
            stacks.append(ls)
        depth = {len(ls) for ls in stacks}
        if len(depth) != 1:
            raise NotImplementedError("per-key encoders must have the same depth")
        keys, sizes = list(comps.keys()), [ls[0].in_features for ls in stacks]
        for j in range(depth.pop()):
            blocks, r0, c0 = [], 0, 0
            for ls in stacks:
                blocks.append((r0, c0, ls[j]))
                r0 += ls[j].in_features
                c0 += ls[j].out_features
            out.append(_VLayer(blocks))
        layers = layers[1:]
    for l in layers:
        if not isinstance(l, Dense):
            raise NotImplementedError(f"unsupported layer {type(l).__name__} in the MLP plan")
        out.append(_VLayer([(0, 0, l)]))
    return out, keys, sizes


class CompiledNet:
    """Device-side view of one actor-critic network."""
    recurrent = False
    obs_adapters: list = []        # leading Flattener / Filter layers (none for recurrent plans)

    def __init__(self, network: StatefulModule, device):
        import torch
        self.network = network
        self.device = device
        normalizer: Optional[Normalizer] = None
        adapter = network
        # leading observation adapters (networks/utils.py: Flattener / Filter) are host plumbing
        # applied to the observation before the kernels see it; the plan starts after them
        self.obs_adapters = []
        # shared trunk (docs/tutorials/02_composition.rst "shared body"): Dense stacks between the Normalizer and the
        # PPOAdapter feed BOTH ports.  Lowered as the head of both chains: the trunk's parameters exist twice in
        # the arena (actor copy, critic copy), tied - the update kernels add the two copies' gradients and give
        # both the same optimizer step, so they stay bit-identical.
        self.trunk_modules: list = []
        if isinstance(network, Sequential):
            layers = list(network.layers)
            while len(layers) > 1 and isinstance(layers[0], (Flattener, Filter)):
                self.obs_adapters.append(layers.pop(0))
            if layers and isinstance(layers[0], Normalizer):
                normalizer = layers.pop(0)
            if not layers or not isinstance(layers[-1], PPOAdapter):
                raise NotImplementedError(
                    "unsupported network topology: expected Sequential([Flattener / Filter..., Normalizer?, "
                    "shared Dense trunk..., PPOAdapter]) or PPOAdapter (the MLP plans)")
            adapter = layers[-1]
            self.trunk_modules = layers[:-1]
        if not isinstance(adapter, PPOAdapter):
            raise NotImplementedError("unsupported network topology: no PPOAdapter found")
        action = adapter.action
        if not isinstance(action, Sequential) or not isinstance(action.layers[-1], NormalTanhSampler):
            raise NotImplementedError("action port must be Sequential([Dense..., NormalTanhSampler])")
        self.sampler: NormalTanhSampler = action.layers[-1]
        value = adapter.value
        trunk_flat = _flatten_trunk(self.trunk_modules)
        self.n_trunk_layers = 0
        actor_layers, self.obs_keys, self.obs_sizes = _lower_chain(trunk_flat + list(action.layers[:-1]))
        critic_layers, ck, cs = _lower_chain(trunk_flat + (list(value.layers) if isinstance(value, Sequential) else [value]))
        if trunk_flat:
            self.n_trunk_layers = len(_lower_chain(trunk_flat)[0])
            for vl in actor_layers[:self.n_trunk_layers]:
                if vl.activation_name == "none":
                    raise NotImplementedError("a shared trunk must apply its activation after every layer")
        if ck is not None and self.obs_keys is not None and (ck != self.obs_keys or cs != self.obs_sizes):
            raise NotImplementedError("actor and critic must split the observation dict the same way")
        if self.obs_keys is None:
            self.obs_keys, self.obs_sizes = ck, cs
        self.normalizer = normalizer
        self.adapter = adapter
        # module lists (for the reference-shaped state / extras pytrees) vs lowered virtual layers
        self.actor_layers = list(action.layers[:-1])
        self.critic_layers = list(value.layers) if isinstance(value, Sequential) else [value]

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

        # one flat arena; every Dense parameter becomes a view into it (a sub-block view for the
        # per-key encoders of a Concat, whose layers are stored block-diagonally)
        host = np.zeros(off, np.float32)
        mask = np.ones(off, np.uint8)
        tie = np.full(off, -1, np.int32)
        masked = False
        nt = self.n_trunk_layers
        for chain, layers in ((plan.actor, actor_layers), (plan.critic, critic_layers)):
            for i, vl in enumerate(layers):
                K, N = vl.in_features, vl.out_features
                Wd = np.zeros((K, N), np.float32)
                Md = np.zeros((K, N), np.uint8)
                bd = np.zeros(N, np.float32)
                for r0, c0, d in vl.blocks:
                    Wd[r0:r0 + d.in_features, c0:c0 + d.out_features] = d.linear.kernel.numpy()
                    Md[r0:r0 + d.in_features, c0:c0 + d.out_features] = 1
                    bd[c0:c0 + d.out_features] = d.linear.bias.numpy()
                host[chain.w_off[i]:chain.w_off[i] + K * N] = Wd.ravel()
                mask[chain.w_off[i]:chain.w_off[i] + K * N] = Md.ravel()
                host[chain.b_off[i]:chain.b_off[i] + N] = bd
                masked = masked or len(vl.blocks) > 1
        for i in range(nt):                           # tie the two copies of every trunk layer
            K, N = actor_layers[i].in_features, actor_layers[i].out_features
            for oa, oc, n in ((plan.actor.w_off[i], plan.critic.w_off[i], K * N), (plan.actor.b_off[i], plan.critic.b_off[i], N)):
                tie[oa:oa + n] = np.arange(oc, oc + n, dtype=np.int32)
                tie[oc:oc + n] = np.arange(oa, oa + n, dtype=np.int32)
                mask[oc:oc + n] = np.where(mask[oc:oc + n] != 0, 2, 0)      # 2: duplicate copy (updated, not counted in the norm)
            masked = True
        self.arena = torch.from_numpy(host).to(device)
        self.param_mask = torch.from_numpy(mask).to(device) if masked else None
        self.param_tie = torch.from_numpy(tie).to(device) if nt else None
        self._tie_pairs = None
        if nt:
            dup = np.nonzero(mask == 2)[0]
            self._tie_pairs = (torch.from_numpy(dup.astype(np.int64)).to(device),
                               torch.from_numpy(tie[dup].astype(np.int64)).to(device))
        self._logical_params = []
        index = []
        for ci, (chain, layers) in enumerate(((plan.critic, critic_layers), (plan.actor, actor_layers))):
            # (critic first so that a shared trunk's Params end up as views of the ACTOR copy)
            for i, vl in enumerate(layers):
                K, N = vl.in_features, vl.out_features
                Wv = self.arena[chain.w_off[i]:chain.w_off[i] + K * N].view(K, N)
                bv = self.arena[chain.b_off[i]:chain.b_off[i] + N]
                for r0, c0, d in vl.blocks:
                    d.linear.kernel._dev = Wv[r0:r0 + d.in_features, c0:c0 + d.out_features]
                    d.linear.bias._dev = bv[c0:c0 + d.out_features]
                    if i < nt:
                        d.linear.kernel._after_set = d.linear.bias._after_set = self.sync_tied
        for ci, (chain, layers) in enumerate(((plan.actor, actor_layers), (plan.critic, critic_layers))):
            # logical (oracle) order: per key, the encoder's layers; then the shared trunk layers; the critic's
            # copy of a shared trunk is not a parameter of its own
            nkeys = max(len(vl.blocks) for vl in layers)
            enc_depth = sum(1 for vl in layers if len(vl.blocks) > 1)
            order = [(j, kb) for kb in range(nkeys) for j in range(enc_depth)] if nkeys > 1 else []
            order += [(j, 0) for j in range(enc_depth if nkeys > 1 else 0, len(layers))]
            for j, kb in order:
                if ci == 1 and j < nt:
                    continue
                r0, c0, d = layers[j].blocks[kb]
                N = layers[j].out_features
                self._logical_params += [d.linear.kernel, d.linear.bias]
                rr = np.arange(r0, r0 + d.in_features, dtype=np.int64)[:, None]
                cc = np.arange(c0, c0 + d.out_features, dtype=np.int64)[None, :]
                index.append((int(chain.w_off[j]) + rr * N + cc).ravel())
                index.append(int(chain.b_off[j]) + cc.ravel())
        self._logical_index = np.concatenate(index)
        if normalizer is not None:
            normalizer._bind(device)
        # counters: [0..1] sampler stream key, [2] sampler count, [3] adam count
        rng = self.sampler.rng
        self._counters_host = np.array([rng.key[0], rng.key[1], rng.count, 0], np.uint32)
        self.counters = torch.from_numpy(self._counters_host.view(np.int32).copy()).to(device)
        self.engines: dict = {}
        self.adam_step = 0   # host mirror of counters[3] (optimizer step count)

    # ---- observation plumbing and the reference-shaped state / extras pytrees (containers.py:18-39) ----
    def flat_obs(self, obs):
        """The [B, obs_size] float32 tensor the kernels consume: leading adapters applied, then a dict
        observation of a Concat plan concatenated in the Concat's key order."""
        import torch
        for m in self.obs_adapters:
            obs = m((), obs).output
        if isinstance(obs, dict):
            if self.obs_keys is None:
                raise TypeError("dict observations need a network whose actor / critic start with a Concat "
                                "(or a Flattener in front of the network)")
            obs = torch.cat([obs[k].float() for k in self.obs_keys], dim=-1)
        return obs

    def wrap(self, obs_flat, adapter_part, leaf):
        """Per-layer list of the enclosing Sequential: `leaf` for every adapter, then the Normalizer's
        entry (`obs_flat` for extras, `()` for state), then the PPOAdapter's part."""
        n = len(self.obs_adapters)
        trunk = [_module_tree(m, leaf) for m in self.trunk_modules]
        if self.normalizer is not None:
            return [leaf] * n + [obs_flat] + trunk + [adapter_part]
        if n or trunk:
            return [leaf] * n + trunk + [adapter_part]
        return adapter_part

    def adapter_extras(self, rollout_extras):
        """The PPOAdapter's part of a rollout_extras pytree produced by ``wrap``."""
        if self.normalizer is not None or self.obs_adapters or self.trunk_modules:
            return rollout_extras[-1]
        return rollout_extras

    # ---- the sampler's `metrics` dict (sampling_layers.py:111) inside the reference-shaped metrics tree ----
    def sampler_metric_keys(self) -> list:
        """Key path of the sampler's metrics under Transition.metrics["net"]: Sequential layers are keyed
        by position (containers.py:36), the PPOAdapter by port name (adapter.py:112)."""
        path = []
        if self.network is not self.adapter:
            path.append(len(self.obs_adapters) + (1 if self.normalizer is not None else 0) + len(self.trunk_modules))
        path += ["action", len(self.adapter.action.layers) - 1]
        return path

    def sampler_metric_path(self) -> str:
        return "/".join(["net"] + [str(k) for k in self.sampler_metric_keys()])

    def wrap_sampler_metrics(self, mu_sigma):
        """{"mu", "sigma"} of one network call nested like the reference's `out.metrics`."""
        A = mu_sigma.shape[-1] // 2
        tree = {"mu": mu_sigma[..., :A], "sigma": mu_sigma[..., A:]}
        for k in reversed(self.sampler_metric_keys()):
            tree = {k: tree}
        return tree

    def logical_index_dev(self):
        """Device int64 index of every real parameter in the arena (no padding, no structural zeros)."""
        import torch
        idx = getattr(self, "_logical_index_dev", None)
        if idx is None:
            idx = self._logical_index_dev = torch.from_numpy(np.asarray(self._logical_index, np.int64)).to(self.device)
        return idx

    # ---- flat <-> logical parameter order (actor W0,b0,..., critic W0,b0,...; no padding) ----
    def params_logical(self, arena=None) -> np.ndarray:
        """An arena-shaped tensor (parameters by default; also gradients, Adam moments) in the
        oracle's flat order: no padding, no structural zeros."""
        a = (self.arena if arena is None else arena).detach().cpu().numpy()
        return a[self._logical_index]

    def load_params_logical(self, flat: np.ndarray) -> None:
        import torch
        host = self.arena.detach().cpu().numpy().copy()
        host[self._logical_index] = np.asarray(flat, np.float32)
        self.arena.copy_(torch.from_numpy(host))
        self.sync_tied()

    def sync_tied(self) -> None:
        """Copy the actor copy of a shared trunk over the critic copy (after parameters were written from
        outside: checkpoint load, ``Param.set``); training keeps the two bit-identical by itself."""
        if getattr(self, "_tie_pairs", None) is not None:
            dup, src = self._tie_pairs
            self.arena[dup] = self.arena[src]

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

    def rollout_workspace(self, B: int):
        """(pointer, bytes) of the scratch buffer `b200ppo_rollout_synth_ws` wants for B envs: (0, 0) while the
        fused kernel keeps this network's weights resident in shared memory; otherwise a cached device buffer for
        the batched per-step path (include/b200ppo.h)."""
        import torch
        n = int(_lib.load().b200ppo_rollout_synth_workspace_bytes(self.plan, B))
        if n <= 0:
            return 0, 0
        # one buffer per env count, never freed while the network lives: a captured iteration graph keeps the pointer
        # it was recorded with (~100 MB at configs[3]; only plans wider than shared memory get here)
        cache = self.__dict__.setdefault("_rollout_ws", {})
        ws = cache.get(B)
        if ws is None or ws.numel() * 4 < n:
            ws = cache[B] = torch.zeros((n + 3) // 4, dtype=torch.float32, device=self.device)
        return ws.data_ptr(), n


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
    leaves = tree_leaves(obs)
    if not leaves or not all(isinstance(x, torch.Tensor) for x in leaves):
        raise TypeError("observations must be a CUDA float32 torch tensor [B, obs_size] (or a pytree of them)")
    # adapters / dict observations (Concat plan): plumbing only
    obs = compile_network(network, leaves[0].device).flat_obs(obs)
    if not isinstance(obs, torch.Tensor):
        raise TypeError("observations must be a CUDA float32 torch tensor [B, obs_size] (or a dict of them)")
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
        raw_in = net.adapter_extras(rollout_extras)["action"][-1].contiguous().float()
    mode = (1 if raw_in is not None else 0) | (2 if net.sampler.deterministic else 0)
    dev = obs.device
    raw = torch.empty(B, A, device=dev)
    action = torch.empty(B, A, device=dev)
    ll = torch.empty(B, device=dev)
    value = torch.empty(B, device=dev)
    reg = torch.empty(B, device=dev)
    musig = torch.empty(B, 2 * A, device=dev)            # the sampler's metrics {"mu", "sigma"} (sampling_layers.py:111)
    s = _lib.current_stream()
    if net.normalizer is not None:
        net.normalizer.prepare(s)
    net.sync_counters_to_device()
    mean_p, std_p = net.norm_ptrs()
    _lib.check(lib.b200ppo_policy_step(s, net.plan, _lib.ptr(net.arena), mean_p, std_p, _lib.ptr(obs), B,
                                       mode, _lib.ptr(net.counters), 0, _lib.ptr(raw_in), _lib.ptr(raw),
                                       _lib.ptr(action), _lib.ptr(ll), _lib.ptr(value), _lib.ptr(reg),
                                       _lib.ptr(musig)), "policy_step")
    net.advance_rng(1 if net.sampler.deterministic else 2)
    na, nc = len(net.actor_layers), len(net.critic_layers)
    adapter_state = {"action": [()] * (na + 1), "value": [()] * nc}
    adapter_extras = {"action": [None] * na + [raw], "value": [None] * nc}
    out = PPONetworkOutput(actions=action, loglikelihoods=ll, value_estimates=value)
    return StatefulModuleOutput(net.wrap((), adapter_state, ()), out, reg, net.wrap_sampler_metrics(musig),
                                net.wrap(obs, adapter_extras, None))


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
    _lib.check(lib.b200ppo_lstm_step_fwd(s, net.lplan_step, _lib.ptr(net.arena), mean_p, std_p, _lib.ptr(obs), 0, 0, B,
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
