"""Device engine for one PPO iteration (the body of the reference's jitted ``ppo_step``,
ppo.py:254-348) as a fixed launch sequence over preallocated HBM buffers:

    norm_prepare -> fused T-step rollout -> permutation indices -> E*M x update
    (FWD, GAE, LOSS, BWD, RED, ADAM) -> Normalizer batch statistics + merge -> counters

The sequence is captured once into a CUDA graph and replayed per iteration (the reference relies
on XLA's jit for the same purpose).  Per-iteration inputs — the two PRNG keys of ppo.py:271 and the
numeric hyper-parameters (gamma, lambda, clip range, critic weight, learning rate ..., which the
reference keeps as traced scalars / optax state, ppo.py:105) — travel in ONE 80-byte block copied
from a ring of pinned host slots into device memory that the kernels read, so the graph never needs
re-capturing and a schedule costs nothing.  With world_size > 1 every rank runs the same sequence on
its own env shard; the per-update exchanges (advantage moment sums, flat gradient) run inside the
update kernels over peer memory (stores into every rank's comm buffer + per-block epoch flags), or
as two NCCL all-reduces per update when peer access is unavailable (B200PPO_P2P=0); one all-gather
per iteration merges the Normalizer batch statistics.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from .. import _lib, parallel
from ..networks.plan import CompiledNet


class AdamOptimizer:
    """State of ``nnx.Optimizer(networks, optax.chain([clip?], adam|adamw))`` — ppo.py:555-569.
    ``mu`` / ``nu`` are flat device arenas parallel to the parameter arena."""

    def __init__(self, net: CompiledNet, learning_rate: float = 1e-4,
                 gradient_clipping: Optional[float] = None, weight_decay=None,
                 b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        import torch
        self.net = net
        self.learning_rate = float(learning_rate)
        self.gradient_clipping = gradient_clipping
        self.weight_decay = weight_decay
        self.b1, self.b2, self.eps = b1, b2, eps
        self.mu = torch.zeros(net.n_params, dtype=torch.float32, device=net.device)
        self.nu = torch.zeros(net.n_params, dtype=torch.float32, device=net.device)
        self.step = 0   # host mirror of counters[3]

    @property
    def wd_value(self) -> float:
        if self.weight_decay is None:
            return -1.0
        if isinstance(self.weight_decay, bool):
            return 1e-4 if self.weight_decay else -1.0   # optax.adamw default decay
        return float(self.weight_decay)


_SIDE_STREAMS: dict = {}

MAX_CACHED_ENGINES = 4


def cached_engine(net: CompiledNet, kind: str, env, opt, shape_key: tuple, factory):
    """Engine cache of one compiled network, keyed on SHAPES and launch structure only (hyper-parameter
    values are run-time data, `PPOEngine.set_hparams`).  An entry keeps strong references to the env
    and optimizer objects it was built for, so their ids cannot be recycled while it is cached, and
    identity is re-checked on every hit.  Least-recently-used engines beyond MAX_CACHED_ENGINES are
    closed (peer buffers unmapped) and dropped together with their ~100 MB workspaces."""
    key = (kind, id(env), id(opt)) + tuple(shape_key)
    eng = net.engines.pop(key, None)
    if eng is not None and not (eng._key_refs[0] is env and eng._key_refs[1] is opt):
        eng.close()
        eng = None
    if eng is None:
        eng = factory()
        eng._key_refs = (env, opt)
    net.engines[key] = eng                       # (re)inserted last = most recently used
    while len(net.engines) > MAX_CACHED_ENGINES:
        old_key = next(iter(net.engines))
        net.engines.pop(old_key).close()
    return eng


def _side_stream(dev):
    """One side stream per device, shared by every engine.  torch hands out streams round-robin from
    a pool of 32 per priority, so a stream per engine eventually aliases another stream object —
    including the one `torch.cuda.graph` captures on (seen as cudaErrorStreamCaptureInvalidated after
    ~40 engines in one process).  The high-priority pool is separate from the default-priority one
    torch's capture stream comes from."""
    import torch
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=dev, priority=-1)
        _SIDE_STREAMS[key] = st
    return st


class PendingMetrics:
    """Per-update metric rows of one iteration on their way to the host (pinned slot + event)."""

    def __init__(self, slot, event):
        self._slot, self._event, self._value = slot, event, None

    def wait(self) -> np.ndarray:
        if self._value is None:
            self._event.synchronize()
            self._value = self._slot.numpy().copy()
            self._slot = self._event = None
        return self._value


class PPOEngine:
    N_MH = 4          # slots of the metrics ring
    N_SLOTS = 8

    def __init__(self, net: CompiledNet, env, opt: AdamOptimizer, n_envs: int, rollout_length: int,
                 n_epochs: int, n_minibatches: int, gae_lambda: float, discounting_factor: float,
                 clip_range: float, normalize_advantages: bool, critic_loss_weight: float,
                 world_size: int = 1, group=None, use_graph: Optional[bool] = None):
        import torch
        if not getattr(env, "fused_rollout", False):
            raise NotImplementedError("PPOEngine needs a device env with a fused rollout kernel")
        if n_envs % n_minibatches:
            raise ValueError("n_envs must be divisible by n_minibatches (ppo.py:285-290 reshape)")
        self.lib = _lib.load()
        self.net, self.env, self.opt = net, env, opt
        self.B, self.T, self.E, self.M = n_envs, rollout_length, n_epochs, n_minibatches
        self.mb = n_envs // n_minibatches
        self.world, self.group = int(world_size), group
        dev = net.device
        self.dev = dev
        O, A, T, B = net.plan.obs_dim, net.plan.act_dim, self.T, self.B
        f32 = dict(dtype=torch.float32, device=dev)
        self.obs = torch.zeros(T, B, O, **f32)
        self.raw_action = torch.zeros(T, B, A, **f32)
        self.action = torch.zeros(T, B, A, **f32)
        self.loglik = torch.zeros(T, B, **f32)
        self.value = None            # [T, B] rollout-time value estimates: only with enable_values()
        self.mu_sigma = None         # [T, B, 2A] rollout-time sampler mu | sigma: only with enable_values(net_metrics=True)
        self.adv_log = None          # [updates, T * mb] raw advantages of every update: only with enable_adv_log()
        self.env_metrics: dict = {}  # Transition.metrics of the rollout (envs stepped from Python only)
        self.reward = torch.zeros(T, B, **f32)
        self.done = torch.zeros(T, B, dtype=torch.uint8, device=dev)
        self.trunc = torch.zeros(T, B, dtype=torch.uint8, device=dev)
        self.next_obs_last = torch.zeros(B, O, **f32)
        self.inds = torch.zeros(self.E, B, dtype=torch.int32, device=dev)
        self.perm_scratch = torch.zeros(int(self.lib.b200ppo_permutation_scratch_bytes(B, self.E)) // 4 + 64,
                                        dtype=torch.int32, device=dev)
        ws_bytes = int(self.lib.b200ppo_update_workspace_bytes(net.plan, T, self.mb))
        if ws_bytes <= 0:
            raise _lib.B200PPOError("update_workspace_bytes rejected the plan")
        # zero-initialised: tickets and the alignment padding of the gradient arena rely on it
        self.ws = torch.zeros(ws_bytes // 4, **f32)
        self.n_updates = self.E * self.M
        self.metrics = torch.zeros(self.n_updates, _lib.METRICS_STRIDE, **f32)
        # device -> host ring of the per-update metric rows: step(fetch_metrics="lazy") queues the copy and returns a
        # handle, so the host can enqueue the next iteration before it reads this one's losses (the reference's
        # jitted ppo_step is dispatched asynchronously in the same way)
        self._mh_ring = [torch.zeros(self.n_updates, _lib.METRICS_STRIDE, dtype=torch.float32).pin_memory()
                         for _ in range(self.N_MH)]
        self._mh_pending = [None] * self.N_MH
        self._mh_i = 0
        self.metrics_host = self._mh_ring[0]
        # per-iteration host -> device block: int32[4] keys | float32[HP_FLOATS] hyper-parameters.  A ring
        # of pinned slots with one event each: a slot is rewritten only after the copy that read it has
        # executed, so the host may run many iterations ahead of the GPU (fetch_metrics=False).
        self.iter_block = torch.zeros(4 + _lib.HP_FLOATS, dtype=torch.int32, device=dev)
        self.iter_keys = self.iter_block[:4]
        self.hp_dev = self.iter_block[4:].view(torch.float32)
        self._slots = [torch.zeros(4 + _lib.HP_FLOATS, dtype=torch.int32).pin_memory() for _ in range(self.N_SLOTS)]
        self._slot_events = [None] * self.N_SLOTS
        self._slot = 0
        self.comm_epoch = torch.zeros(1, dtype=torch.int32, device=dev)   # monotonic exchange epoch (never rolled back)
        self.norm_scratch = torch.zeros(int(self.lib.b200ppo_norm_scratch_bytes(O)) // 4 + 64, **f32)
        self.batch_stats = torch.zeros(2 * O, **f32)
        self.batch_stats_all = torch.zeros(self.world, 2 * O, **f32)
        wsp = self.ws.data_ptr()
        adv_p = self.lib.b200ppo_update_adv_sums_ptr(net.plan, T, self.mb, wsp)
        grad_p = self.lib.b200ppo_update_grad_ptr(net.plan, T, self.mb, wsp)
        ao, go = (adv_p - wsp) // 4, (grad_p - wsp) // 4
        self.adv_sums = self.ws[ao:ao + 4].view(torch.float64)
        self.grad = self.ws[go:go + net.n_params]
        self.hp = _lib.HParams()
        self.hp.gamma, self.hp.lambda_ = float(discounting_factor), float(gae_lambda)
        self.hp.clip_range, self.hp.critic_loss_weight = float(clip_range), float(critic_loss_weight)
        self.hp.learning_rate = opt.learning_rate
        self.hp.adam_b1, self.hp.adam_b2, self.hp.adam_eps = opt.b1, opt.b2, opt.eps
        self.hp.weight_decay = opt.wd_value
        self.hp.grad_clip = float(opt.gradient_clipping) if opt.gradient_clipping is not None else -1.0
        self.hp.normalize_advantages = 1 if normalize_advantages else 0
        self.hp.world_size = self.world
        self.bufs = []
        for u in range(self.n_updates):
            b = _lib.UpdateBufs()
            b.obs, b.raw_action, b.loglik_old = self.obs.data_ptr(), self.raw_action.data_ptr(), self.loglik.data_ptr()
            b.reward, b.done, b.truncated = self.reward.data_ptr(), self.done.data_ptr(), self.trunc.data_ptr()
            b.next_obs_last = self.next_obs_last.data_ptr()
            b.inds = self.inds.data_ptr() + 4 * u * self.mb        # [E*M][mb] view of [E][B]
            b.norm_mean, b.norm_std = net.norm_ptrs()
            b.params, b.adam_mu, b.adam_nu = net.arena.data_ptr(), opt.mu.data_ptr(), opt.nu.data_ptr()
            b.rng_state = net.counters.data_ptr()
            b.metrics_out = self.metrics.data_ptr() + 4 * _lib.METRICS_STRIDE * u
            b.ws = wsp
            pm = getattr(net, "param_mask", None)
            b.param_mask = pm.data_ptr() if pm is not None else 0
            b.hparams_dev = self.hp_dev.data_ptr()
            b.comm_epoch = self.comm_epoch.data_ptr()
            pt = getattr(net, "param_tie", None)
            b.param_tie = pt.data_ptr() if pt is not None else 0
            self.bufs.append(b)
        self._hp_host = np.zeros(_lib.HP_FLOATS, np.float32)
        self.set_hparams()
        self._upload_block((0, 0), (0, 0))
        # sampler counts one update's replay consumes (ppo.py:425-436: T replayed steps + the bootstrap call, two
        # draws each) and the loss head; the distillation engine overrides both (no bootstrap call, NLL head)
        self.rng_per_update = 2 * (T + 1)
        self.nll = False
        self.rng_per_iter = 2 * T + self.n_updates * self.rng_per_update
        # data-parallel exchange of the per-update advantage sums and gradients: peer memory
        # (flags + P2P loads inside the GAE / loss / Adam kernels) unless disabled or unsupported
        self.fuse_prep = os.environ.get("B200PPO_FUSE_PREP", "1") != "0"
        self._side = _side_stream(dev)
        self.p2p = False
        self._comm_local, self._comm_peers, self.comm_table = None, [], None
        if self.world > 1 and os.environ.get("B200PPO_P2P", "1") != "0":
            self._setup_p2p_collective()
        if use_graph is None:
            use_graph = os.environ.get("B200PPO_GRAPH", "1") != "0"
        self.use_graph = use_graph
        self.graph = None
        self.iters_run = 0
        self.kernel_launches_per_iter = 0
        self._env_state = None
        self._rng_mirror = None

    # ------------------------------------------------------------------------------------
    def set_hparams(self, gae_lambda=None, discounting_factor=None, clip_range=None, critic_loss_weight=None):
        """Refresh the host copy of the device hyper-parameter block from the arguments (None: keep)
        and from the optimizer (learning rate, Adam constants, decay / clip VALUES: a schedule just
        assigns ``opt.learning_rate``).  Whether decay / clipping exist at all is launch structure and
        is fixed when the engine is built.  The block reaches the device with the next step()."""
        hp, opt, h = self.hp, self.opt, self._hp_host
        if gae_lambda is not None:
            hp.lambda_ = float(gae_lambda)
        if discounting_factor is not None:
            hp.gamma = float(discounting_factor)
        if clip_range is not None:
            hp.clip_range = float(clip_range)
        if critic_loss_weight is not None:
            hp.critic_loss_weight = float(critic_loss_weight)
        hp.learning_rate = float(opt.learning_rate)
        hp.adam_b1, hp.adam_b2, hp.adam_eps = opt.b1, opt.b2, opt.eps
        if hp.weight_decay >= 0.0 and opt.wd_value >= 0.0:
            hp.weight_decay = opt.wd_value
        if hp.grad_clip > 0.0 and opt.gradient_clipping is not None:
            hp.grad_clip = float(opt.gradient_clipping)
        h[_lib.HP_GAMMA], h[_lib.HP_LAMBDA] = hp.gamma, hp.lambda_
        h[_lib.HP_CLIP_RANGE], h[_lib.HP_CRITIC_WEIGHT] = hp.clip_range, hp.critic_loss_weight
        h[_lib.HP_LEARNING_RATE] = hp.learning_rate
        h[_lib.HP_ADAM_B1], h[_lib.HP_ADAM_B2], h[_lib.HP_ADAM_EPS] = hp.adam_b1, hp.adam_b2, hp.adam_eps
        h[_lib.HP_WEIGHT_DECAY], h[_lib.HP_GRAD_CLIP] = hp.weight_decay, hp.grad_clip

    def _upload_block(self, reset_key, new_key) -> None:
        """Write (keys, hyper-parameters) into the next pinned slot and queue its copy to the device."""
        import torch
        i = self._slot
        self._slot = (i + 1) % self.N_SLOTS
        ev = self._slot_events[i]
        if ev is not None:
            ev.synchronize()                       # the copy that last read this slot has executed
        host = self._slots[i].numpy()
        host[:4] = np.array([reset_key[0], reset_key[1], new_key[0], new_key[1]], np.uint32).view(np.int32)
        host[4:] = self._hp_host.view(np.int32)
        self.iter_block.copy_(self._slots[i], non_blocking=True)
        if ev is None:
            ev = self._slot_events[i] = torch.cuda.Event()
        ev.record()

    def _setup_p2p_collective(self):
        """Allocate this rank's comm buffer, exchange CUDA IPC handles, map every peer's buffer
        (include/b200ppo.h, 'peer-memory exchange').  Every rank executes the same collectives whatever
        happens locally and the outcome is agreed with a MIN over ranks after each phase: if any rank
        cannot allocate / export / map (no P2P access between some GPUs, IPC disabled in the container)
        all ranks use the NCCL all-reduce path instead — still a GPU collective, reported in
        bench.py's `config.collectives`."""
        import ctypes as C
        import sys
        import torch
        import torch.distributed as dist
        lib, world = self.lib, self.world
        rank = dist.get_rank(self.group)

        def agree(ok: int) -> bool:
            flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            return int(flag.item()) == 1

        err = None
        handle = bytes(64)
        try:                                                     # phase 1: local allocation + export
            nbytes = int(lib.b200ppo_comm_bytes(self.net.plan, world))
            if nbytes <= 0 or world > 16:
                raise _lib.B200PPOError("peer exchange: unsupported plan or world size")
            p = C.c_void_p()
            _lib.check(lib.b200ppo_comm_alloc(nbytes, C.byref(p)), "comm_alloc")
            self._comm_local = p.value
            h = C.create_string_buffer(64)
            _lib.check(lib.b200ppo_comm_ipc_get(p, h), "comm_ipc_get")
            handle = h.raw
        except Exception as e:                                   # noqa: BLE001 - decided collectively
            err = e
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.dev)
        allh = torch.zeros(world * 64, dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        ok = agree(0 if err is not None else 1)
        ptrs = []
        if ok:
            allh = allh.cpu().numpy().reshape(world, 64)
            try:                                                 # phase 2: map the peers
                for r in range(world):
                    if r == rank:
                        ptrs.append(self._comm_local)
                        continue
                    q = C.c_void_p()
                    _lib.check(lib.b200ppo_comm_ipc_open(allh[r].tobytes(), C.byref(q)), f"comm_ipc_open(rank {r})")
                    self._comm_peers.append(q.value)
                    ptrs.append(q.value)
            except Exception as e:                               # noqa: BLE001
                err = e
            ok = agree(0 if err is not None else 1)
        if not ok:
            if err is not None:
                sys.stderr.write(f"[b200ppo] peer-memory exchange unavailable ({err}); using NCCL all-reduce\n")
            self.close()
            self.p2p = False
            dist.barrier(group=self.group)
            return
        self.comm_table = torch.tensor(ptrs, dtype=torch.int64, device=self.dev)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for bf in self.bufs:
            bf.comm = self.comm_table.data_ptr()
        self.hp.rank = rank
        self.p2p = True

    def close(self):
        """Unmap / free the peer-exchange buffers (safe to call twice)."""
        lib = self.lib
        for q in self._comm_peers:
            lib.b200ppo_comm_ipc_close(q)
        self._comm_peers = []
        if self._comm_local is not None:
            lib.b200ppo_comm_free(self._comm_local)
            self._comm_local = None

    def _allreduce(self, t):
        parallel.all_reduce_sum(t, self.group)

    def _enqueue(self, env_state) -> int:
        """Enqueue one whole iteration on the current stream.  Returns the number of kernel
        launches issued by this library (NCCL kernels not counted).

        Two small pieces do not depend on what runs next to them and go to a side stream (forked /
        joined with events, so the captured graph gets parallel branches): the minibatch permutation
        (depends only on the key) runs beside the rollout, and the Normalizer's batch statistics
        (depend only on the rollout's observations) run beside the updates."""
        import torch
        cur = torch.cuda.current_stream()
        side = self._side
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            n = self._enqueue_permutation()
        n += self._enqueue_rollout(env_state)
        if self.value is not None:
            n += self._enqueue_values()          # before the first Adam step touches the parameters
        cur.wait_stream(side)
        n += self._enqueue_updates(2 * self.T, self.rng_per_iter, permute=False, side_stats=True)
        return n

    def _enqueue_rollout(self, env_state) -> int:
        lib, net, T, B = self.lib, self.net, self.T, self.B
        s = _lib.current_stream()
        n = 0
        if net.normalizer is not None:
            net.normalizer.prepare(s); n += 1
        mean_p, std_p = net.norm_ptrs()
        es = self.env.c_struct(self.dev)
        ws_p, ws_n = net.rollout_workspace(B)
        _lib.check(lib.b200ppo_rollout_synth_ws(
            s, net.plan, es, net.arena.data_ptr(), mean_p, std_p, net.counters.data_ptr(),
            self.iter_keys.data_ptr(), T, B, env_state.obs.data_ptr(), env_state.step_counter.data_ptr(),
            env_state.term_state.data_ptr(), self.obs.data_ptr(), self.raw_action.data_ptr(),
            self.action.data_ptr(), self.loglik.data_ptr(), self.reward.data_ptr(), self.done.data_ptr(),
            self.trunc.data_ptr(), self.next_obs_last.data_ptr(), ws_p, ws_n), "rollout_synth")
        # one fused launch, or the batched per-step sequence for networks too wide for shared memory
        n += int(lib.b200ppo_rollout_synth_num_launches(net.plan, T, B, 1 if ws_n else 0))
        return n

    def enable_values(self, net_metrics: bool = False) -> None:
        """Also evaluate the network on the rollout's observations with the rollout-time parameters, for
        logging: ``Transition.network_output.value_estimates`` (losses/predicted_value at
        LoggingLevel.CRITIC_EXTRA, metrics.py:62-68) and, with ``net_metrics``, the sampler's mu / sigma
        (``Transition.metrics["net"]``, sampling_layers.py:111, logged at TRAINING_ENV_METRICS).  The
        training math never reads either (ppo.py:425-446 replays the network), so the pass is off unless
        asked for."""
        import torch
        T, B, A = self.T, self.B, self.net.plan.act_dim
        f32 = dict(dtype=torch.float32, device=self.dev)
        if self.value is None:
            self.value = torch.zeros(T, B, **f32)
            self._val_scratch = (torch.empty(T * B, A, **f32), torch.empty(T * B, A, **f32), torch.empty(T * B, **f32))
            self.graph = None        # the captured iteration does not contain the pass: capture again
        if net_metrics and self.mu_sigma is None:
            self.mu_sigma = torch.zeros(T, B, 2 * A, **f32)
            self.graph = None

    def enable_adv_log(self) -> None:
        """Keep every update's advantage tensor (``losses/advantages`` is the [updates, T, mb] array in the
        reference, ppo.py:523): needed only when percentiles of it are logged; mean / std come from the
        sums the loss kernel accumulates anyway."""
        import torch
        if self.adv_log is None:
            self.adv_log = torch.zeros(self.n_updates, self.T * self.mb, dtype=torch.float32, device=self.dev)
            wsp = self.ws.data_ptr()
            ap = ctypes.cast(self.lib.b200ppo_update_debug_ptr(self.net.plan, self.T, self.mb, wsp, 0), ctypes.c_void_p).value
            o = (ap - wsp) // 4
            self._adv_view = self.ws[o:o + self.T * self.mb]
            self.graph = None

    def enable_grad_norm(self) -> bool:
        """LoggingLevel.GRAD_NORM without gradient clipping (ppo.py:313-315): run the global-norm
        kernel with a clip threshold no finite norm reaches, so metrics[3] is written and the
        gradient is used unscaled (on the peer-memory path too: the norm is taken from the rank-ordered
        sum inside the exchange launch)."""
        if self.hp.grad_clip > 0.0:
            return True
        self.hp.grad_clip = float(np.finfo(np.float32).max)
        self.set_hparams()
        self.graph = None
        return True

    def _enqueue_values(self) -> int:
        net = self.net
        raw, act, ll = self._val_scratch
        mean_p, std_p = net.norm_ptrs()
        # replay mode on the stored raw actions: no sampler count is consumed
        _lib.check(self.lib.b200ppo_policy_step(
            _lib.current_stream(), net.plan, net.arena.data_ptr(), mean_p, std_p, self.obs.data_ptr(),
            self.T * self.B, 1, net.counters.data_ptr(), 0, self.raw_action.data_ptr(), raw.data_ptr(),
            act.data_ptr(), ll.data_ptr(), self.value.data_ptr(), 0,
            self.mu_sigma.data_ptr() if self.mu_sigma is not None else 0), "policy_step(values)")
        return 1

    def _enqueue_permutation(self) -> int:
        _lib.check(self.lib.b200ppo_permutation(_lib.current_stream(), self.iter_keys.data_ptr() + 8, self.B, self.E,
                                                self.inds.data_ptr(), self.perm_scratch.data_ptr()), "permutation")
        return 1

    def _enqueue_batch_stats(self) -> int:
        nz = self.net.normalizer
        _lib.check(self.lib.b200ppo_norm_batch_stats(_lib.current_stream(), self.obs.data_ptr(), self.T * self.B, nz.size,
                                                     self.batch_stats.data_ptr(), self.norm_scratch.data_ptr()),
                   "norm_batch_stats")
        return 2

    def _enqueue_updates(self, rng_offset0: int, rng_advance: int, permute: bool = True,
                         side_stats: bool = False) -> int:
        """Permutation indices, the E*M minibatch updates, Normalizer statistics, counters.
        ``rng_offset0`` = sampler counts already consumed since counters[2] was last advanced."""
        import torch
        lib, net, T, B = self.lib, self.net, self.T, self.B
        s = _lib.current_stream()
        n = 0
        if permute:
            n += self._enqueue_permutation()
        cur = torch.cuda.current_stream()
        if side_stats and net.normalizer is not None:
            self._side.wait_stream(cur)                      # the rollout's observations are complete
            with torch.cuda.stream(self._side):
                n += self._enqueue_batch_stats()
        # PPO: every stage; distillation: no GAE, the NLL head (include/b200ppo.h B200PPO_STAGE_NLL)
        gae_bit = 0 if self.nll else _lib.STAGE_GAE
        loss_bits = _lib.STAGE_LOSS | (_lib.STAGE_NLL if self.nll else 0)
        stage_all = _lib.STAGE_FWD | gae_bit | loss_bits | _lib.STAGE_BWD | _lib.STAGE_RED | _lib.STAGE_ADAM
        for u in range(self.n_updates):
            off = rng_offset0 + u * self.rng_per_update
            args = (s, net.plan, self.hp, self.bufs[u], T, B, self.mb, off, u)
            # the Adam kernel of update u-1 refreshed the split weight planes: only the first update
            # of an iteration (parameters may have been touched from outside) runs the prep launch
            noprep = _lib.STAGE_NO_PREP if (u > 0 and self.fuse_prep) else 0
            per_update = int(lib.b200ppo_update_num_launches(net.plan, self.hp, T, self.mb, stage_all | noprep))
            if (self.world == 1 or self.p2p) and self.adv_log is not None:
                head = _lib.STAGE_FWD | gae_bit | loss_bits
                _lib.check(lib.b200ppo_update(*args, head | noprep), "update/loss")
                self.adv_log[u].copy_(self._adv_view)
                _lib.check(lib.b200ppo_update(*args, stage_all & ~head), "update/bwd")
            elif self.world == 1 or self.p2p:
                _lib.check(lib.b200ppo_update(*args, stage_all | noprep), "update")
            else:
                _lib.check(lib.b200ppo_update(*args, _lib.STAGE_FWD | gae_bit | noprep), "update/fwd")
                if self.hp.normalize_advantages and not self.nll:
                    self._allreduce(self.adv_sums)
                _lib.check(lib.b200ppo_update(*args, loss_bits | _lib.STAGE_BWD | _lib.STAGE_RED), "update/bwd")
                if self.adv_log is not None:
                    self.adv_log[u].copy_(self._adv_view)
                self._allreduce(self.grad)
                _lib.check(lib.b200ppo_update(*args, _lib.STAGE_ADAM), "update/adam")
            n += per_update
        if self.world > 1:
            # every column is pre-divided by the GLOBAL sample count (SUM = global mean) except the gradient
            # norm [3], which is computed from the already summed gradient and identical on every rank
            if self.hp.grad_clip > 0.0 and parallel.dist_info()[1] != 0:
                self.metrics[:, 3].zero_()
            self._allreduce(self.metrics)
        if net.normalizer is not None:
            nz = net.normalizer
            if side_stats:
                cur.wait_stream(self._side)
            else:
                n += self._enqueue_batch_stats()
            src = self.batch_stats
            if self.world > 1:
                parallel.all_gather_into(self.batch_stats_all.view(-1), self.batch_stats, self.group)
                src = self.batch_stats_all
            _lib.check(lib.b200ppo_norm_merge(s, src.data_ptr(), self.world, float(T * B), nz.size,
                                              nz.mean._dev.data_ptr(), nz.M2._dev.data_ptr(),
                                              nz.counter._dev.data_ptr()), "norm_merge"); n += 1
        _lib.check(lib.b200ppo_iter_finalize(s, net.counters.data_ptr(), rng_advance, self.n_updates,
                                             self.comm_epoch.data_ptr()), "iter_finalize"); n += 1
        return n

    @property
    def env_state(self):
        """The env state the engine advances in place (the object the captured graph points at)."""
        return self._env_state

    def step(self, env_state, reset_key, new_key, fetch_metrics: bool = True):
        """Run one iteration.  ``reset_key, new_key = split(rng_key)`` (ppo.py:271) come from the
        host; everything else stays on the device.  Asynchronous unless ``fetch_metrics``.  The env
        state is advanced IN PLACE in engine-owned tensors (``self.env_state``): the first state passed
        in is adopted, a different object passed later (a loaded checkpoint, a rolled-back
        TrainingState) is copied over the live one."""
        import torch
        self._upload_block(reset_key, new_key)
        if self._env_state is None:
            self._env_state = env_state
        elif env_state is not self._env_state:
            self._env_state.obs.copy_(env_state.obs)
            self._env_state.step_counter.copy_(env_state.step_counter)
            self._env_state.term_state.copy_(env_state.term_state)
        env_state = self._env_state
        if self.iters_run == 0 or self.net.adam_step != self.opt.step or self._rng_mirror != self.net.rng_count:
            # first use, or the optimizer / sampler counters were changed from outside (checkpoint load,
            # rollback): bring the device counters in line.  The exchange epoch is NOT touched: it only
            # ever moves forward, so flags left by earlier iterations can never look current.
            self.net.adam_step = self.opt.step
            self.net.sync_counters_to_device()
        if not self.use_graph or self.iters_run == 0:
            self.kernel_launches_per_iter = self._enqueue(env_state)
        else:
            if self.graph is None:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(env_state)
                self.graph = g
            self.graph.replay()
        self.iters_run += 1
        self.net.advance_rng(self.rng_per_iter)
        self._rng_mirror = self.net.rng_count
        self.opt.step += self.n_updates
        self.net.adam_step = self.opt.step
        if fetch_metrics:
            i = self._mh_i
            self._mh_i = (i + 1) % self.N_MH
            old = self._mh_pending[i]
            if old is not None:
                old.wait()                        # an unread result is about to lose its slot: keep its values
            self.metrics_host = self._mh_ring[i]
            self.metrics_host.copy_(self.metrics, non_blocking=True)
            if fetch_metrics == "lazy":
                ev = torch.cuda.Event()
                ev.record()
                h = self._mh_pending[i] = PendingMetrics(self.metrics_host, ev)
                return h
            self._mh_pending[i] = None
            torch.cuda.current_stream().synchronize()
            return self.metrics_host.numpy().copy()
        return None

    def h2d_bytes_per_step(self) -> int:
        return 4 * (4 + _lib.HP_FLOATS)

    def d2h_bytes_per_step(self) -> int:
        return self.n_updates * 4 * _lib.METRICS_STRIDE
