"""Shared-trunk networks and the routing containers - CPU checks.

* ``Parallel`` / ``Splitter`` follow nnx_ppo/networks/containers.py:115-218 (the cases of the reference's
  containers tests: dict outputs keyed by name, slices in keyword order, validation errors).
* the plan compiler lowers Sequential([Normalizer, trunk, PPOAdapter(...)]) to two chains that both START
  with the trunk, stored twice and tied;
* oracle/sharednet.py's analytic gradient (trunk = actor path + critic path) against float64 torch
  autograd of the same composite loss.
"""
import numpy as np
import pytest
import torch

from nnx_ppo_b200 import Rngs
from nnx_ppo_b200.networks import containers, factories, feedforward
from nnx_ppo_b200.networks.plan import CompiledNet
from oracle import env as oenv, nets as onets, ppo as oppo, prng as oprng, sharednet


def test_splitter_slices_in_keyword_order_and_ignores_excess():
    sp = containers.Splitter(mu=3, rho=2)
    x = torch.arange(2 * 7, dtype=torch.float32).reshape(2, 7)
    out = sp((), x)
    assert list(out.output) == ["mu", "rho"]
    assert torch.equal(out.output["mu"], x[:, 0:3]) and torch.equal(out.output["rho"], x[:, 3:5])
    assert out.next_state == () and out.rollout_extras is None and out.metrics == {}
    with pytest.raises(ValueError):
        containers.Splitter()
    with pytest.raises(ValueError):
        containers.Splitter(a=0)


def test_parallel_routes_one_input_to_named_components():
    par = containers.Parallel(head=containers.Splitter(a=2), tail=containers.Splitter(skip=2, b=1))
    x = torch.arange(8, dtype=torch.float32).reshape(2, 4)
    st = par.initialize_state(2)
    assert set(st) == {"head", "tail"}
    out = par(st, x)
    assert torch.equal(out.output["head"]["a"], x[:, :2]) and torch.equal(out.output["tail"]["b"], x[:, 2:3])
    assert set(out.next_state) == set(out.rollout_extras) == set(out.metrics) == {"head", "tail"}
    assert par.reset_state(st).keys() == st.keys()
    par.update_statistics(out.rollout_extras)
    par2 = containers.Parallel({"x": containers.Splitter(a=1)})
    assert list(par2.components) == ["x"]
    with pytest.raises(ValueError):
        containers.Parallel({"x": containers.Splitter(a=1)}, y=containers.Splitter(a=1))
    with pytest.raises(ValueError):
        containers.Parallel()


def _nets(seed=0, act="relu"):
    return (factories.make_shared_trunk_actor_critic(10, 3, [16, 12], [8], [6, 5], Rngs(seed), activation=act),
            sharednet.make_shared_trunk_actor_critic(10, 3, [16, 12], [8], [6, 5], seed=seed, activation=act))


def test_shared_trunk_plan_duplicates_and_ties_the_trunk():
    nets, onet = _nets()
    net = CompiledNet(nets, torch.device("cpu"))
    assert net.n_trunk_layers == 2 and net.plan.actor.n_layers == 4 and net.plan.critic.n_layers == 5
    assert [net.plan.actor.dims[i] for i in range(5)] == [10, 16, 12, 8, 6]
    assert [net.plan.critic.dims[i] for i in range(6)] == [10, 16, 12, 6, 5, 1]
    n_trunk = 10 * 16 + 16 + 16 * 12 + 12
    mask, tie = net.param_mask.numpy(), net.param_tie.numpy()
    assert (mask == 2).sum() == n_trunk and (tie >= 0).sum() == 2 * n_trunk
    idx = np.nonzero(tie >= 0)[0]
    assert np.array_equal(tie[tie[idx]], idx)                     # the tie is an involution
    assert np.all(mask[idx] + mask[tie[idx]] == 3)                # exactly one counted copy per pair
    # same initial values as the oracle, logical order [trunk, actor head, critic head]; both copies equal
    np.testing.assert_array_equal(net.params_logical(), onet.flat_params())
    a = net.arena.numpy()
    np.testing.assert_array_equal(a[idx], a[tie[idx]])
    assert nets.layers[-1].action.layers[-1].rng.count == onet.rng_count
    # state / extras pytrees follow the module tree: [normalizer, [trunk layers...], adapter]
    w = net.wrap("N", "A", "L")
    assert w[0] == "N" and w[1] == ["L", "L"] and w[2] == "A"
    # writing a trunk parameter from outside re-syncs the second copy
    k = nets.layers[1].layers[0].linear.kernel
    k.set(np.full(k.shape, 0.25, np.float32))
    a = net.arena.numpy()
    np.testing.assert_array_equal(a[idx], a[tie[idx]])
    p = net.params_logical()
    net.load_params_logical(p[::-1].copy())
    a = net.arena.numpy()
    np.testing.assert_array_equal(a[idx], a[tie[idx]])


def test_shared_trunk_rejects_what_it_cannot_lower():
    nets, _ = _nets()
    trunk = nets.layers[1]
    trunk.layers[-1].activation = None                            # a linear trunk output is not expressible
    trunk.layers[-1].activation_name = "none"
    with pytest.raises(NotImplementedError):
        CompiledNet(nets, torch.device("cpu"))
    nets2, _ = _nets()
    nets2.layers.insert(1, containers.Splitter(a=3))
    with pytest.raises(NotImplementedError):
        CompiledNet(nets2, torch.device("cpu"))


@pytest.mark.parametrize("act", ["relu", "tanh", "swish"])
def test_oracle_shared_trunk_gradient_matches_autograd(act):
    """float64 torch autograd of the full surrogate (actor + critic + entropy regulariser) through ONE trunk
    against the oracle's analytic float32 gradient; and against the unshared oracle for the head layers."""
    _, onet = _nets(3, act)
    O, A, B, T = 10, 3, 24, 6
    env = oenv.SyntheticEnv(O, A, max_len=5, term_thresh16=3000)
    st = env.reset(oprng.split(oprng.key(1), B))
    g = np.random.default_rng(0)
    onet.update_statistics(g.standard_normal((3, 20, O)).astype(np.float32))
    _, ro = oppo.unroll_env(env, st, onet, T, oprng.key(2))
    ro.loglik += (0.3 * g.standard_normal(ro.loglik.shape)).astype(np.float32)
    inds = np.arange(B, dtype=np.int32)
    base = onet.rng_count
    total, m, grads = oppo.ppo_loss_and_grads(onet, ro, inds, base)
    assert grads.size == onet.flat_params().size

    # the same loss as a torch graph over the LOGICAL parameters: the head gradients d_y / d_v come from the
    # oracle's loss head (pinned elsewhere); what is checked here is the chain rule through the shared trunk
    tdt = torch.float64
    shapes = [(ch.W[l].shape, ch.b[l].shape) for ch, l in onet._logical()]
    flat = torch.tensor(onet.flat_params(), dtype=tdt, requires_grad=True)
    Ws, bs, o = [], [], 0
    for ws, bsh in shapes:
        n = ws[0] * ws[1]; Ws.append(flat[o:o + n].reshape(ws)); o += n
        bs.append(flat[o:o + bsh[0]]); o += bsh[0]
    f = {"relu": torch.relu, "tanh": torch.tanh, "swish": torch.nn.functional.silu}[act]
    x = torch.tensor(onet.normalize_obs(ro.obs[:, inds].reshape(T * B, O)), dtype=tdt)
    nt, na = onet.n_trunk, onet.actor.n_layers
    h = x
    for l in range(nt):
        h = f(h @ Ws[l] + bs[l])
    ya = h
    for l in range(nt, na):
        ya = ya @ Ws[l] + bs[l]
        if l < na - 1:
            ya = f(ya)
    vc = h
    ncl = onet.critic.n_layers - nt
    for j in range(ncl):
        vc = vc @ Ws[na + j] + bs[na + j]
        if j < ncl - 1:
            vc = f(vc)
    surrogate = (ya * torch.tensor(m["d_y"], dtype=tdt)).sum() + (vc[:, 0] * torch.tensor(m["d_v"], dtype=tdt)).sum()
    surrogate.backward()
    ref = flat.grad.numpy()
    scale = np.abs(ref).max()
    assert np.abs(grads - ref).max() < 2e-5 * scale, (np.abs(grads - ref).max(), scale)
    n_trunk = sum(w[0] * w[1] + b[0] for w, b in shapes[:nt])
    assert np.abs(ref[:n_trunk]).max() > 0                         # the trunk does get gradient from both paths
