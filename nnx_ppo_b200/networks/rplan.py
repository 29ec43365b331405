"""Plan compiler for the recurrent actor-critic:
Sequential([Normalizer?, PPOAdapter(action=Sequential([Dense, LSTM, Dense, NormalTanhSampler]),
value=Sequential([Dense...]))]) — the architecture of the reference's recurrent_test.py:245-258.

One flat parameter arena again:  [W1 b1 | Wi Wh bl | W2 b2 | stand-in actor | critic layers].
The critic, GAE, the loss head, the gradient reduction and Adam are the MLP path's kernels, driven
through an MLP plan whose "actor" is a 1-layer stand-in (obs -> 2A, zero weights, never used: the
recurrent replay overwrites its output slot in the workspace and its gradient slots belong to no
real parameter).  The recurrent actor itself runs in csrc/recurrent.cu (``b200ppo_lstm_plan``).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .. import _lib
from .adapter import PPOAdapter
from .containers import Sequential
from .feedforward import Dense
from .normalizer import Normalizer
from .plan import CompiledNet, _align4
from .recurrent import LSTM
from .sampling_layers import NormalTanhSampler


def is_recurrent(network) -> bool:
    return any(isinstance(m, LSTM) for m in network.iter_modules()) if hasattr(network, "iter_modules") else False


class RecurrentCompiledNet(CompiledNet):
    recurrent = True

    def __init__(self, network, device):   # noqa: super().__init__ is deliberately not called
        import torch
        self.network, self.device = network, device
        normalizer: Optional[Normalizer] = None
        adapter = network
        if isinstance(network, Sequential):
            layers = list(network.layers)
            if len(layers) == 2 and isinstance(layers[0], Normalizer) and isinstance(layers[1], PPOAdapter):
                normalizer, adapter = layers
            elif len(layers) == 1 and isinstance(layers[0], PPOAdapter):
                adapter = layers[0]
        if not isinstance(adapter, PPOAdapter):
            raise NotImplementedError("unsupported network topology: no PPOAdapter found")
        act_layers = list(adapter.action.layers)
        if not (len(act_layers) == 4 and isinstance(act_layers[0], Dense) and isinstance(act_layers[1], LSTM)
                and isinstance(act_layers[2], Dense) and isinstance(act_layers[3], NormalTanhSampler)):
            raise NotImplementedError("recurrent plan: the actor must be Sequential([Dense, LSTM, Dense, NormalTanhSampler])")
        pre, lstm, post, self.sampler = act_layers
        if post.activation_name != "none" or pre.activation_name == "none":
            raise NotImplementedError("recurrent plan: Dense(act) -> LSTM -> Dense(linear)")
        if pre.out_features != lstm.in_features or post.in_features != lstm.hidden_features or post.out_features % 2:
            raise ValueError("layer sizes do not chain")
        critic_layers = list(adapter.value.layers) if isinstance(adapter.value, Sequential) else [adapter.value]
        if not all(isinstance(l, Dense) for l in critic_layers):
            raise NotImplementedError("recurrent plan: the critic must be an MLP")
        self.normalizer, self.adapter = normalizer, adapter
        self.pre, self.lstm, self.post = pre, lstm, post
        self.actor_layers, self.critic_layers = [pre, lstm, post], critic_layers
        O, P, H, Y = pre.in_features, pre.out_features, lstm.hidden_features, post.out_features
        self.carry_path = (1 if normalizer is not None else None, "action", 1)

        lp = _lib.LstmPlan()
        lp.obs_dim, lp.pre_dim, lp.hidden, lp.out_dim = O, P, H, Y
        lp.act, lp.normalize = _lib.ACT_IDS[pre.activation_name], 1 if normalizer is not None else 0
        off = 0
        lp.w1_off = off; off += O * P
        lp.b1_off = off; off += P
        lp.wcat_off = off; off += (P + H) * 4 * H
        lp.bl_off = off; off += 4 * H
        lp.w2_off = off; off += H * Y
        lp.b2_off = off; off += Y
        if lstm.trainable_initial_state:                 # learned reset carry, [H] each, 16-byte aligned
            off = _align4(off)
            lp.init_c_off = off; off += H                # carry slot 0 <- `initial_h` (see networks/recurrent.py)
            lp.init_h_off = off; off += H
        self.n_recurrent = off
        off = _align4(off)

        plan = _lib.Plan()
        plan.actor.n_layers, plan.actor.act = 1, 0                       # stand-in actor (see module docstring)
        plan.actor.dims[0], plan.actor.dims[1] = O, Y
        plan.actor.w_off[0] = off; off = _align4(off + O * Y)
        plan.actor.b_off[0] = off; off = _align4(off + Y)
        c = plan.critic
        c.n_layers = len(critic_layers)
        acts = {l.activation_name for l in critic_layers[:-1]}
        if len(acts) > 1 or critic_layers[-1].activation_name != "none":
            raise NotImplementedError("critic: one shared hidden activation, linear last layer")
        c.act = _lib.ACT_IDS[acts.pop() if acts else "none"]
        c.dims[0] = critic_layers[0].in_features
        for i, l in enumerate(critic_layers):
            c.dims[i + 1] = l.out_features
            c.w_off[i] = off; off = _align4(off + l.in_features * l.out_features)
            c.b_off[i] = off; off = _align4(off + l.out_features)
        if c.dims[0] != O or c.dims[c.n_layers] != 1:
            raise NotImplementedError("critic must map obs -> 1")
        plan.obs_dim, plan.act_dim = O, Y // 2
        plan.normalize = lp.normalize
        plan.entropy_weight = float(self.sampler.entropy_weight)
        plan.min_std, plan.std_scale = float(self.sampler.min_std), float(self.sampler.std_scale)
        plan.n_params = off
        lp.n_params = off
        self.plan, self.lplan, self.n_params = plan, lp, off
        # a single step applies no reset: the per-step kernels get the plan without the learned carry
        self.lplan_step = _lib.LstmPlan.from_buffer_copy(lp)
        self.lplan_step.init_c_off = self.lplan_step.init_h_off = 0

        host = np.zeros(off, np.float32)
        params = [(lp.w1_off, pre.linear.kernel), (lp.b1_off, pre.linear.bias),
                  (lp.wcat_off, lstm.kernel_i), (lp.wcat_off + P * 4 * H, lstm.kernel_h), (lp.bl_off, lstm.bias),
                  (lp.w2_off, post.linear.kernel), (lp.b2_off, post.linear.bias)]
        self.init_views = None
        if lstm.trainable_initial_state:
            if not _lib.load().b200ppo_lstm_seq_supported(lp):
                raise NotImplementedError("trainable_initial_state needs the sequence kernels: hidden % 16 == 0 and "
                                          "pre Dense width % 4 == 0")
            params[5:5] = [(lp.init_c_off, lstm.initial_h), (lp.init_h_off, lstm.initial_c)]   # oracle order: after bl
        for i, l in enumerate(critic_layers):
            params += [(int(c.w_off[i]), l.linear.kernel), (int(c.b_off[i]), l.linear.bias)]
        self._params = params
        for o, p in params:
            v = p.numpy()
            host[o:o + v.size] = v.ravel()
        self.arena = torch.from_numpy(host).to(device)
        for o, p in params:
            p._dev = self.arena[o:o + int(np.prod(p.shape))].view(*p.shape)
        if lstm.trainable_initial_state:
            self.init_views = (lstm.initial_h._dev, lstm.initial_c._dev)
        if normalizer is not None:
            normalizer._bind(device)
        rng = self.sampler.rng
        self._counters_host = np.array([rng.key[0], rng.key[1], rng.count, 0], np.uint32)
        self.counters = torch.from_numpy(self._counters_host.view(np.int32).copy()).to(device)
        self.engines = {}
        self.adam_step = 0

        # oracle order (oracle/recurrent.py param_list): W1 b1 Wi Wh bl W2 b2, critic W0 b0 ...
        self._logical_index = np.concatenate([np.arange(o, o + int(np.prod(p.shape)), dtype=np.int64)
                                              for o, p in params])
        self.param_mask = None
        self.obs_keys = self.obs_sizes = None

    def reset_carry(self, carry, done):
        """rollout.py:33-40: envs that finished get ``reset_state`` - zeros, or the learned initial carry."""
        import torch
        c, h = carry
        d = done.bool()[:, None]
        if self.init_views is None:
            keep = (~d).to(torch.float32)
            return (c * keep, h * keep)
        return (torch.where(d, self.init_views[0], c), torch.where(d, self.init_views[1], h))

    # ---- carry inside the reference-shaped network_states pytree ----
    def get_carry(self, network_states):
        top, key, idx = self.carry_path
        s = network_states[top] if top is not None else network_states
        return s[key][idx]

    def set_carry(self, network_states, carry):
        top, key, idx = self.carry_path
        s = network_states[top] if top is not None else network_states
        s[key][idx] = carry
        return network_states
