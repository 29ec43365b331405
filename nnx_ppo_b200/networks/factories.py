"""Network factories — mirrors nnx_ppo/networks/factories.py:14-146 (same names, argument order,
defaults and layer / key-draw order)."""
from __future__ import annotations

from typing import Callable, Union

from .. import prng
from . import feedforward
from .adapter import PPOAdapter
from .containers import Concat, Sequential
from .feedforward import Dense
from .normalizer import Normalizer
from .sampling_layers import NormalTanhSampler
from .types import StatefulModule


def make_mlp_layers(sizes: list[int], rngs: prng.Rngs, activation: Callable = feedforward.relu,
                    activation_last_layer: bool = True, **linear_kwargs) -> list[Dense]:
    layers = []
    for i, (din, dout) in enumerate(zip(sizes[:-1], sizes[1:])):
        is_last = i == len(sizes) - 2
        act = activation if (not is_last or activation_last_layer) else None
        layers.append(Dense(din, dout, rngs, activation=act, **linear_kwargs))
    return layers


def make_mlp(sizes: list[int], rngs: prng.Rngs, activation: Callable = feedforward.relu,
             activation_last_layer: bool = True, **linear_kwargs) -> Sequential:
    return Sequential(make_mlp_layers(sizes, rngs, activation, activation_last_layer, **linear_kwargs))


def make_mlp_actor_critic(obs_size: int, action_size: int, actor_hidden_sizes: list[int],
                          critic_hidden_sizes: list[int], rngs: prng.Rngs,
                          activation: Union[Callable, str] = feedforward.relu,
                          normalize_obs: bool = True, initializer_scale: float = 1.0,
                          entropy_weight: float = 1e-2, min_std: float = 1e-1,
                          std_scale: float = 1.0) -> StatefulModule:
    """Sequential([Normalizer(obs_size)?, PPOAdapter(action=Sequential([actor, sampler]),
    value=critic)]) — factories.py:72-146."""
    if isinstance(activation, str):
        activation = {"swish": feedforward.swish, "tanh": feedforward.tanh,
                      "relu": feedforward.relu}[activation]

    def kernel_init(key, shape):
        return prng.variance_scaling_uniform(key, shape[0], shape[1], initializer_scale)

    actor_layers = make_mlp_layers([obs_size] + list(actor_hidden_sizes) + [action_size * 2], rngs,
                                   activation, activation_last_layer=False, kernel_init=kernel_init)
    critic = make_mlp([obs_size] + list(critic_hidden_sizes) + [1], rngs, activation,
                      activation_last_layer=False, kernel_init=kernel_init)
    sampler = NormalTanhSampler(rngs, entropy_weight=entropy_weight, min_std=min_std,
                                std_scale=std_scale)
    adapter = PPOAdapter(action=Sequential([*actor_layers, sampler]), value=critic)
    if normalize_obs:
        return Sequential([Normalizer(obs_size), adapter])
    return adapter


def make_recurrent_actor_critic(obs_size: int, action_size: int, pre_size: int, lstm_hidden: int,
                                critic_hidden_sizes: list[int], rngs: prng.Rngs,
                                activation: Union[Callable, str] = feedforward.relu,
                                normalize_obs: bool = True, entropy_weight: float = 1e-2,
                                min_std: float = 1e-1, std_scale: float = 1.0,
                                trainable_initial_state: bool = False) -> StatefulModule:
    """Actor Dense(act) -> LSTM -> Dense with an MLP critic, the network of the reference's
    recurrent_test.py:245-258 wrapped like make_mlp_actor_critic (Normalizer + PPOAdapter)."""
    from .recurrent import LSTM
    if isinstance(activation, str):
        activation = {"swish": feedforward.swish, "tanh": feedforward.tanh, "relu": feedforward.relu}[activation]

    def kernel_init(key, shape):
        return prng.variance_scaling_uniform(key, shape[0], shape[1], 1.0)

    actor = [feedforward.Dense(obs_size, pre_size, rngs, activation=activation, kernel_init=kernel_init),
             LSTM(pre_size, lstm_hidden, rngs, trainable_initial_state=trainable_initial_state),
             feedforward.Dense(lstm_hidden, action_size * 2, rngs, activation=None, kernel_init=kernel_init)]
    critic = make_mlp([obs_size] + list(critic_hidden_sizes) + [1], rngs, activation,
                      activation_last_layer=False, kernel_init=kernel_init)
    sampler = NormalTanhSampler(rngs, entropy_weight=entropy_weight, min_std=min_std, std_scale=std_scale)
    adapter = PPOAdapter(action=Sequential([*actor, sampler]), value=critic)
    if normalize_obs:
        return Sequential([Normalizer(obs_size), adapter])
    return adapter


def make_dict_actor_critic(obs_sizes: dict, action_size: int, encoder_hidden: dict,
                           actor_trunk_sizes: list[int], critic_trunk_sizes: list[int], rngs: prng.Rngs,
                           activation: Union[Callable, str] = feedforward.relu, normalize_obs: bool = True,
                           entropy_weight: float = 1e-2, min_std: float = 1e-1,
                           std_scale: float = 1.0) -> StatefulModule:
    """Dict observations routed to per-key encoders (BASELINE configs[3]): actor and critic each are
    Sequential([Concat(key=MLP encoder, ...), trunk Dense..., head]) over the observation dict
    (containers.py:55-110).  ``obs_sizes`` / ``encoder_hidden`` map key -> size / hidden sizes (all
    encoders must have the same depth; every encoder layer is followed by the activation)."""
    if isinstance(activation, str):
        activation = {"swish": feedforward.swish, "tanh": feedforward.tanh, "relu": feedforward.relu}[activation]

    def kernel_init(key, shape):
        return prng.variance_scaling_uniform(key, shape[0], shape[1], 1.0)

    def tower(trunk, out):
        encs = {k: make_mlp([obs_sizes[k]] + list(encoder_hidden[k]), rngs, activation,
                            activation_last_layer=True, kernel_init=kernel_init) for k in obs_sizes}
        width = sum(encoder_hidden[k][-1] for k in obs_sizes)
        rest = make_mlp_layers([width] + list(trunk) + [out], rngs, activation, activation_last_layer=False,
                               kernel_init=kernel_init)
        return [Concat(encs), *rest]

    actor = tower(actor_trunk_sizes, action_size * 2)
    critic = Sequential(tower(critic_trunk_sizes, 1))
    sampler = NormalTanhSampler(rngs, entropy_weight=entropy_weight, min_std=min_std, std_scale=std_scale)
    adapter = PPOAdapter(action=Sequential([*actor, sampler]), value=critic)
    if normalize_obs:
        return Sequential([Normalizer(sum(obs_sizes.values())), adapter])
    return adapter


def make_shared_trunk_actor_critic(obs_size: int, action_size: int, trunk_sizes: list[int],
                                   actor_head_sizes: list[int], critic_head_sizes: list[int], rngs: prng.Rngs,
                                   activation: Union[Callable, str] = feedforward.relu, normalize_obs: bool = True,
                                   entropy_weight: float = 1e-2, min_std: float = 1e-1,
                                   std_scale: float = 1.0) -> StatefulModule:
    """The shared-trunk network of the reference's composition tutorial (docs/tutorials/02_composition.rst,
    "shared body"): ``Sequential([Normalizer?, trunk MLP (activation after every layer), PPOAdapter(action =
    Sequential([actor head MLP, sampler]), value = critic head MLP)])`` - both ports read the trunk's features.
    Key-draw order: trunk, actor head, critic head, sampler."""
    if isinstance(activation, str):
        activation = {"swish": feedforward.swish, "tanh": feedforward.tanh, "relu": feedforward.relu}[activation]

    def kernel_init(key, shape):
        return prng.variance_scaling_uniform(key, shape[0], shape[1], 1.0)

    trunk = make_mlp([obs_size] + list(trunk_sizes), rngs, activation, activation_last_layer=True, kernel_init=kernel_init)
    width = trunk_sizes[-1]
    actor_head = make_mlp_layers([width] + list(actor_head_sizes) + [action_size * 2], rngs, activation,
                                 activation_last_layer=False, kernel_init=kernel_init)
    critic_head = make_mlp([width] + list(critic_head_sizes) + [1], rngs, activation, activation_last_layer=False,
                           kernel_init=kernel_init)
    sampler = NormalTanhSampler(rngs, entropy_weight=entropy_weight, min_std=min_std, std_scale=std_scale)
    adapter = PPOAdapter(action=Sequential([*actor_head, sampler]), value=critic_head)
    layers = [trunk, adapter]
    if normalize_obs:
        layers.insert(0, Normalizer(obs_size))
    return Sequential(layers)
