"""Host-side plan lowering (networks/plan.py::_lower_chain): a Concat of per-key Dense encoders
(containers.py:55-110) becomes block-diagonal layers over the concatenated observation vector.
Pure module-tree logic — no device needed."""
import numpy as np
import pytest

from nnx_ppo_b200 import prng
from nnx_ppo_b200.networks import factories, feedforward
from nnx_ppo_b200.networks.containers import Concat, Sequential
from nnx_ppo_b200.networks.plan import _lower_chain


def test_concat_of_encoders_lowers_to_block_diagonal_layers():
    nets = factories.make_dict_actor_critic({"proprio": 6, "target": 10}, 3, {"proprio": [8, 4], "target": [12, 5]},
                                            [16], [7, 7], prng.Rngs(0), normalize_obs=False)
    layers, keys, sizes = _lower_chain(nets.action.layers[:-1])
    assert keys == ["proprio", "target"] and sizes == [6, 10]
    assert [(l.in_features, l.out_features) for l in layers] == [(16, 20), (20, 9), (9, 16), (16, 6)]
    assert [(r0, c0, d.in_features, d.out_features) for r0, c0, d in layers[0].blocks] == [(0, 0, 6, 8), (6, 8, 10, 12)]
    assert [(r0, c0, d.in_features, d.out_features) for r0, c0, d in layers[1].blocks] == [(0, 0, 8, 4), (8, 4, 12, 5)]
    assert all(len(l.blocks) == 1 and l.blocks[0][:2] == (0, 0) for l in layers[2:])
    assert [l.activation_name for l in layers] == ["relu", "relu", "relu", "none"]
    # the critic tower splits the observation dict the same way
    _, ck, cs = _lower_chain(nets.value.layers)
    assert (ck, cs) == (keys, sizes)
    # a plain MLP chain has no key order
    plain = factories.make_mlp_actor_critic(5, 2, [8], [8], prng.Rngs(0), normalize_obs=False)
    layers, keys, sizes = _lower_chain(plain.action.layers[:-1])
    assert keys is None and sizes is None and [(l.in_features, l.out_features) for l in layers] == [(5, 8), (8, 4)]


def test_unsupported_chains_raise():
    r = prng.Rngs(0)
    enc = lambda sizes, act=feedforward.relu: factories.make_mlp(sizes, r, act, activation_last_layer=True)
    with pytest.raises(NotImplementedError, match="same depth"):
        _lower_chain([Concat(a=enc([4, 8]), b=enc([4, 8, 8]))])
    with pytest.raises(NotImplementedError, match="same activation"):
        _lower_chain([Concat(a=enc([4, 8]), b=enc([4, 8], feedforward.tanh))])
    with pytest.raises(NotImplementedError, match="Dense stacks"):
        _lower_chain([Concat(a=Sequential([Concat(x=enc([4, 8]))]))])
    with pytest.raises(NotImplementedError, match="unsupported layer"):
        _lower_chain([feedforward.Dense(4, 4, r), Concat(a=enc([4, 8]))])
    with pytest.raises(ValueError):
        Concat()
    with pytest.raises(ValueError):
        Concat({"a": enc([4, 8])}, b=enc([4, 8]))
