"""Checkpoint callback / loader with the reference's API (algorithms/checkpointing.py:42-204):
``make_checkpoint_fn(directory, config)`` returns a ``checkpoint_fn(training_state, step)`` that
writes ``{directory}/step_{step:010d}/``; ``load_checkpoint(path, networks, optimizer)`` restores
into structurally identical templates IN PLACE and returns ``{"training_state", "step", "config"}``.

Format (the reference's orbax + pickle layout needs orbax / jax, absent here):
  state.npz      params, adam_mu, adam_nu  - flat float32 in the LOGICAL parameter order (actor layers
                 W0, b0, ... then critic; per-key encoders before the trunk; recurrent: W1 b1 Wi Wh bl
                 W2 b2 then critic) = the order of the reference's ``nnx.state(networks, nnx.Param)``
                 leaves for the same architecture; norm_mean, norm_M2, norm_counter; sampler_key (2 x
                 uint32), sampler_count, adam_count
  metadata.pkl   env_states / network_states (host copies of the tensors), rng_key, steps_taken, step,
                 config
A training run resumed from a checkpoint continues bit-identically (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import dataclasses
import os
import pickle
from typing import Any, Optional

import numpy as np

from ..networks.plan import compile_network
from .config import TrainConfig
from .types import TrainingState


def _to_host(tree: Any) -> Any:
    import torch
    if isinstance(tree, torch.Tensor):
        return ("__tensor__", tree.detach().cpu().numpy())
    if isinstance(tree, dict):
        return {k: _to_host(v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_to_host(v) for v in tree)
    if dataclasses.is_dataclass(tree) and not isinstance(tree, type):
        return ("__dataclass__", type(tree), {f.name: _to_host(getattr(tree, f.name)) for f in dataclasses.fields(tree)})
    return tree


def _to_device(tree: Any, device) -> Any:
    import torch
    if isinstance(tree, tuple) and len(tree) == 2 and tree[0] == "__tensor__":
        return torch.from_numpy(np.ascontiguousarray(tree[1])).to(device)
    if isinstance(tree, tuple) and len(tree) == 3 and tree[0] == "__dataclass__":
        return tree[1](**{k: _to_device(v, device) for k, v in tree[2].items()})
    if isinstance(tree, dict):
        return {k: _to_device(v, device) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_to_device(v, device) for v in tree)
    return tree


def make_checkpoint_fn(directory: str, config: Optional[TrainConfig] = None):
    directory = os.path.abspath(directory)

    def checkpoint_fn(training_state: TrainingState, step: int) -> None:
        import torch
        net = compile_network(training_state.networks)
        opt = training_state.optimizer
        step_dir = os.path.join(directory, f"step_{int(step):010d}")
        os.makedirs(step_dir, exist_ok=True)
        torch.cuda.synchronize()
        arrays = dict(params=net.params_logical(), adam_mu=net.params_logical(opt.mu),
                      adam_nu=net.params_logical(opt.nu), adam_count=np.int64(opt.step),
                      sampler_key=np.asarray(net.sampler.rng.key, np.uint32),
                      sampler_count=np.int64(net.sampler.rng.count))
        if net.normalizer is not None:
            nz = net.normalizer
            arrays.update(norm_mean=nz.mean.numpy(), norm_M2=nz.M2.numpy(), norm_counter=nz.counter.numpy())
        np.savez(os.path.join(step_dir, "state.npz"), **arrays)
        metadata = {"network_states": _to_host(training_state.network_states),
                    "env_states": _to_host(training_state.env_states),
                    "rng_key": tuple(int(x) for x in training_state.rng_key),
                    "steps_taken": np.float32(training_state.steps_taken), "step": int(step), "config": config}
        with open(os.path.join(step_dir, "metadata.pkl"), "wb") as f:
            pickle.dump(metadata, f)

    return checkpoint_fn


def load_checkpoint(path: str, networks: Any, optimizer: Any) -> dict[str, Any]:
    import torch
    path = os.path.abspath(path)
    net = compile_network(networks)
    z = np.load(os.path.join(path, "state.npz"))
    if z["params"].shape != net.params_logical().shape:
        raise ValueError("checkpoint does not match the network architecture")
    net.load_params_logical(z["params"])

    def load_arena(dst, flat):
        host = np.zeros(net.n_params, np.float32)
        host[net._logical_index] = flat
        dst.copy_(torch.from_numpy(host))
    load_arena(optimizer.mu, z["adam_mu"])
    load_arena(optimizer.nu, z["adam_nu"])
    optimizer.step = int(z["adam_count"])
    net.sampler.rng.key = tuple(int(x) for x in z["sampler_key"])
    net.sampler.rng.count = int(z["sampler_count"])
    if net.normalizer is not None:
        nz = net.normalizer
        nz.mean.set(z["norm_mean"]); nz.M2.set(z["norm_M2"]); nz.counter.set(z["norm_counter"])
    # device mirrors of the stream key / counts (the captured graph reads them from device memory)
    net._counters_host[0], net._counters_host[1] = net.sampler.rng.key[0], net.sampler.rng.key[1]
    net.adam_step = optimizer.step
    net.sync_counters_to_device()
    with open(os.path.join(path, "metadata.pkl"), "rb") as f:
        md = pickle.load(f)
    ts = TrainingState(networks=networks, network_states=_to_device(md["network_states"], net.device),
                       env_states=_to_device(md["env_states"], net.device), optimizer=optimizer,
                       rng_key=tuple(md["rng_key"]), steps_taken=np.float32(md["steps_taken"]))
    return {"training_state": ts, "step": md["step"], "config": md["config"]}
