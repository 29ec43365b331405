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


# ------------------------------------------------------------------------------------------
# observation adapters (networks/utils.py: Flattener / Filter) — the reference's own unit tests
# (networks/utils_test.py:36-148) on torch CPU tensors
# ------------------------------------------------------------------------------------------
import torch                                                              # noqa: E402

from nnx_ppo_b200.networks.adapter import PPOAdapter                      # noqa: E402
from nnx_ppo_b200.networks.normalizer import Normalizer                   # noqa: E402
from nnx_ppo_b200.networks.plan import CompiledNet                        # noqa: E402
from nnx_ppo_b200.networks.utils import Filter, Flattener                 # noqa: E402


def test_filter_specs():
    out = Filter({"out": "a"})((), {"a": torch.ones(2, 3), "b": torch.zeros(2, 4)}).output
    assert torch.equal(out["out"], torch.ones(2, 3)) and "b" not in out
    obs = {"arm": {"proprio": torch.ones(2, 4), "target": torch.zeros(2, 5)}, "head": torch.full((2, 3), 5.0)}
    out = Filter({"p": ("arm", "proprio"), "z": "head", "calc": lambda o: o["head"] * 2})((), obs).output
    assert torch.equal(out["p"], torch.ones(2, 4)) and torch.equal(out["z"], obs["head"])
    assert torch.equal(out["calc"], torch.full((2, 3), 10.0)) and set(out) == {"p", "z", "calc"}
    with pytest.raises(TypeError):
        Filter({"x": 5})
    with pytest.raises(TypeError):
        Filter([("a", "b")])


def test_flattener_levels():
    x = {"b": torch.full((2, 5), 4.0), "a": torch.ones(2, 3)}
    flat = Flattener()((), x).output
    assert flat.shape == (2, 8) and torch.equal(flat[0], torch.tensor([1.0] * 3 + [4.0] * 5))   # jax.tree order: sorted keys
    assert Flattener()((), torch.ones(2, 7)).output.shape == (2, 7)
    assert Flattener()((), {"img": torch.ones(2, 3, 4)}).output.shape == (2, 12)
    nested = {"arm": {"proprio": torch.ones(2, 4), "target": torch.zeros(2, 8)}, "root": torch.full((2, 6), 3.0)}
    out = Flattener(preserve_levels=1)((), nested).output
    assert set(out) == {"arm", "root"} and out["arm"].shape == (2, 12) and out["root"].shape == (2, 6)
    same = Flattener(preserve_levels=1)((), x).output
    assert torch.equal(same["a"], x["a"]) and torch.equal(same["b"], x["b"])
    deep = {"arm": {"p": {"a": torch.ones(2, 3), "b": torch.zeros(2, 4)}, "t": torch.ones(2, 5)}}
    out = Flattener(preserve_levels=2)((), deep).output
    assert out["arm"]["p"].shape == (2, 7) and out["arm"]["t"].shape == (2, 5)
    with pytest.raises(ValueError):
        Flattener(preserve_levels=-1)
    with pytest.raises(TypeError):
        Flattener(preserve_levels=2)((), {"a": torch.ones(2, 3)})


def _adapter_net(normalize=True):
    base = factories.make_mlp_actor_critic(9, 2, [8], [8], prng.Rngs(0), normalize_obs=normalize)
    tail = list(base.layers) if normalize else [base]
    return Sequential([Filter({"p": ("arm", "proprio"), "h": "head"}), Flattener(), *tail])


def test_plan_recognises_leading_observation_adapters():
    """The plan compiler (no kernel is launched by it, so it runs on CPU tensors) strips leading
    Flattener / Filter layers, applies them as host plumbing and keeps the reference-shaped per-layer
    state / rollout_extras lists of the enclosing Sequential (containers.py:18-39)."""
    obs = {"arm": {"proprio": torch.arange(8.0).reshape(2, 4), "target": torch.zeros(2, 3)},
           "head": torch.full((2, 5), 7.0)}
    for normalize in (True, False):
        nets = _adapter_net(normalize)
        net = CompiledNet(nets, torch.device("cpu"))
        assert [type(m).__name__ for m in net.obs_adapters] == ["Filter", "Flattener"]
        assert net.plan.obs_dim == 9 and (net.normalizer is not None) == normalize
        flat = net.flat_obs(obs)
        assert torch.equal(flat, torch.cat([obs["head"], obs["arm"]["proprio"]], dim=-1))      # sorted keys: h, p
        state = nets.initialize_state(2)
        assert len(state) == (4 if normalize else 3) and state[0] == () and state[1] == ()
        ad_state, ad_extras = {"action": [(), ()], "value": [(), ()]}, {"action": [None, "raw"], "value": [None, None]}
        assert net.wrap((), ad_state, ()) == ([(), (), (), ad_state] if normalize else [(), (), ad_state])
        extras = net.wrap(flat, ad_extras, None)
        assert extras[:2] == [None, None] and extras[-1] is ad_extras and net.adapter_extras(extras) is ad_extras
        if normalize:
            assert extras[2] is flat
            seen = []                                                                  # Sequential zips layers with extras
            nets.layers[2].update_statistics = seen.append
            nets.update_statistics(extras)
            assert len(seen) == 1 and seen[0] is flat
    # adapters alone are not a network
    with pytest.raises(NotImplementedError):
        CompiledNet(Sequential([Flattener()]), torch.device("cpu"))


def test_plan_of_existing_topologies_is_unchanged():
    for nets, wrapped in ((factories.make_mlp_actor_critic(5, 2, [8], [8], prng.Rngs(0)), True),
                          (factories.make_mlp_actor_critic(5, 2, [8], [8], prng.Rngs(0), normalize_obs=False), False)):
        net = CompiledNet(nets, torch.device("cpu"))
        assert net.obs_adapters == []
        x = torch.ones(3, 5)
        assert net.flat_obs(x) is x
        ad = {"action": [None, "raw"], "value": [None, None]}
        assert net.wrap(x, ad, None) == ([x, ad] if wrapped else ad)
        assert net.wrap((), ad, ()) == ([(), ad] if wrapped else ad)
        assert net.adapter_extras(net.wrap(x, ad, None)) is ad
    d = factories.make_dict_actor_critic({"a": 4, "b": 6}, 2, {"a": [8], "b": [8]}, [16], [16], prng.Rngs(0))
    net = CompiledNet(d, torch.device("cpu"))
    o = {"b": torch.ones(2, 6), "a": torch.zeros(2, 4)}
    assert torch.equal(net.flat_obs(o), torch.cat([o["a"], o["b"]], dim=-1))                   # the Concat's key order
    plain = CompiledNet(factories.make_mlp_actor_critic(10, 2, [8], [8], prng.Rngs(0)), torch.device("cpu"))
    with pytest.raises(TypeError):
        plain.flat_obs(o)


# ------------------------------------------------------------------------------------------
# parameter arena layout and logical (oracle) order — the plan compiler on CPU tensors
# ------------------------------------------------------------------------------------------
def _check_arena(net, oracle_flat):
    logical = net.params_logical()
    assert logical.shape == oracle_flat.shape and np.array_equal(logical, oracle_flat)     # same init, same order
    for ch in (net.plan.actor, net.plan.critic):
        for l in range(ch.n_layers):
            assert ch.w_off[l] % 4 == 0 and ch.b_off[l] % 4 == 0                               # float4-aligned blocks
            assert ch.b_off[l] + ch.dims[l + 1] <= net.n_params
    # the logical index addresses every real parameter exactly once; padding and structural zeros are 0
    assert len(np.unique(net._logical_index)) == oracle_flat.size
    real = np.zeros(net.n_params, bool)
    real[net._logical_index] = True
    assert np.all(net.arena.numpy()[~real] == 0.0)
    if net.param_mask is not None:
        assert np.all(net.param_mask.numpy()[net._logical_index] == 1)
    # every module parameter is a VIEW into the arena: training mutates the user's network in place
    net.arena.add_(1.0)
    assert np.array_equal(net.params_logical(), oracle_flat + 1.0)
    p0 = net._logical_params[0]
    assert np.array_equal(p0.numpy().ravel(), (oracle_flat + 1.0)[:p0.numpy().size])
    net.load_params_logical(oracle_flat)
    assert np.array_equal(net.params_logical(), oracle_flat)


def test_arena_layout_mlp_and_dict_plans_match_oracle_order():
    from oracle import dictnet, nets as onets
    nets = factories.make_mlp_actor_critic(24, 5, [64, 32], [48], prng.Rngs(7), activation="tanh")
    onet = onets.make_mlp_actor_critic(24, 5, [64, 32], [48], seed=7, activation="tanh")
    net = CompiledNet(nets, torch.device("cpu"))
    assert (net.plan.obs_dim, net.plan.act_dim, net.plan.normalize) == (24, 5, 1) and net.param_mask is None
    _check_arena(net, onet.flat_params())
    sizes, enc = {"proprio": 6, "target": 10}, {"proprio": [8, 4], "target": [12, 5]}
    nets = factories.make_dict_actor_critic(sizes, 3, enc, [16], [7, 7], prng.Rngs(2))
    onet = dictnet.make_dict_actor_critic(sizes, 3, enc, [16], [7, 7], seed=2)
    net = CompiledNet(nets, torch.device("cpu"))
    assert net.plan.obs_dim == 16 and net.param_mask is not None
    # block-diagonal layers: the off-diagonal blocks are masked structural zeros
    n_struct_zero = 2 * ((6 * 12 + 10 * 8) + (8 * 5 + 12 * 4))                                # actor + critic towers
    assert int((net.param_mask == 0).sum()) == n_struct_zero
    assert float(net.arena[net.param_mask == 0].abs().max()) == 0.0
    _check_arena(net, onet.flat_params())


def test_arena_layout_recurrent_plan_matches_oracle_order():
    from nnx_ppo_b200.networks.rplan import RecurrentCompiledNet
    from oracle import recurrent as orec
    nets = factories.make_recurrent_actor_critic(16, 4, 32, 24, [48], prng.Rngs(1))
    onet = orec.make_recurrent_actor_critic(16, 4, [32], 24, [], [48], seed=1)
    net = RecurrentCompiledNet(nets, torch.device("cpu"))
    assert net.recurrent and net.obs_adapters == []
    assert np.array_equal(net.params_logical(), onet.flat_params())


def test_recurrent_plan_with_trainable_initial_state():
    """recurrent.py:85-87: two [H] Params, zeros; carry slot 0 <- initial_h, slot 1 <- initial_c; the arena
    holds them after the LSTM bias in the oracle's order; initialize_state / reset_state broadcast them."""
    from nnx_ppo_b200.networks.rplan import RecurrentCompiledNet
    from oracle import recurrent as orec
    nets = factories.make_recurrent_actor_critic(16, 4, 32, 32, [48], prng.Rngs(1), trainable_initial_state=True)
    onet = orec.make_recurrent_actor_critic(16, 4, [32], 32, [], [48], seed=1, trainable_initial_state=True)
    lstm = nets.layers[1].action.layers[1]
    assert lstm.trainable_initial_state and lstm.initial_h.shape == (32,) and not lstm.initial_h.numpy().any()
    lstm.initial_h.set(np.arange(32, dtype=np.float32))
    lstm.initial_c.set(-np.arange(32, dtype=np.float32))
    onet.init_c[:], onet.init_h[:] = np.arange(32), -np.arange(32)
    net = RecurrentCompiledNet(nets, torch.device("cpu"))
    lp = net.lplan
    assert lp.init_c_off > 0 and lp.init_c_off % 4 == 0 and lp.init_h_off == lp.init_c_off + 32
    assert net.lplan_step.init_c_off == 0 and net.lplan_step.w2_off == lp.w2_off
    assert np.array_equal(net.params_logical(), onet.flat_params())
    st = nets.initialize_state(3)
    c, h = net.get_carry(st)
    assert torch.equal(c, torch.arange(32.0).expand(3, 32)) and torch.equal(h, -torch.arange(32.0).expand(3, 32))
    rc, rh = net.reset_carry((torch.ones(3, 32), torch.ones(3, 32)), torch.tensor([True, False, True]))
    assert torch.equal(rc[0], torch.arange(32.0)) and torch.equal(rc[1], torch.ones(32)) and torch.equal(rh[2], -torch.arange(32.0))
    # the per-step FFMA kernels do not implement it: a plan that needs them is rejected up front
    bad = factories.make_recurrent_actor_critic(16, 4, 30, 24, [48], prng.Rngs(1), trainable_initial_state=True)
    with pytest.raises(NotImplementedError):
        RecurrentCompiledNet(bad, torch.device("cpu"))
