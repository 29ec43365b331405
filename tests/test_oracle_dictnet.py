"""CPU checks of oracle/dictnet.py (dict observations -> per-key encoders -> concat -> trunk;
reference containers.py:55-110): forward equals the explicit per-key computation, the analytic
backward equals float64 autograd, and the whole oracle ppo_step runs on it with exact bookkeeping."""
import numpy as np
import torch

from oracle import dictnet, env as oenv, ppo as oppo

F = np.float32
SIZES = {"proprio": 6, "target": 10}
ENC = {"proprio": [8, 4], "target": [12, 6]}


def test_encchain_forward_and_backward_match_autograd():
    net = dictnet.make_dict_actor_critic(SIZES, 3, ENC, [16], [16], seed=4, activation="tanh")
    ch = net.actor
    rng = np.random.default_rng(0)
    x = rng.normal(size=(9, 16)).astype(F)
    y, zs = ch.forward(x, keep=True)
    # explicit restatement: dict in, per-key stacks, concat
    xd = {"proprio": x[:, :6], "target": x[:, 6:]}
    outs, li = [], 0
    for k in SIZES:
        h = xd[k]
        for _ in ENC[k]:
            h = np.tanh(h @ ch.W[li] + ch.b[li]); li += 1
        outs.append(h)
    h = np.concatenate(outs, 1)
    h = np.tanh(h @ ch.W[li] + ch.b[li]); li += 1
    yr = h @ ch.W[li] + ch.b[li]
    assert np.allclose(y, yr, atol=1e-5)
    d_out = rng.normal(size=y.shape).astype(F)
    dWs, dbs = ch.backward(x, zs, d_out)
    Wt = [torch.tensor(w.astype(np.float64), requires_grad=True) for w in ch.W]
    bt = [torch.tensor(b.astype(np.float64), requires_grad=True) for b in ch.b]
    xt = torch.tensor(x.astype(np.float64))
    outs, li = [], 0
    for (c0, c1, idx) in ch.enc:
        hh = xt[:, c0:c1]
        for j in idx:
            hh = torch.tanh(hh @ Wt[j] + bt[j])
        outs.append(hh)
    hh = torch.cat(outs, 1)
    hh = torch.tanh(hh @ Wt[ch.trunk[0]] + bt[ch.trunk[0]])
    yy = hh @ Wt[ch.trunk[1]] + bt[ch.trunk[1]]
    (yy * torch.tensor(d_out.astype(np.float64))).sum().backward()
    for j in range(len(Wt)):
        assert np.allclose(dWs[j], Wt[j].grad.numpy(), atol=2e-5), j
        assert np.allclose(dbs[j], bt[j].grad.numpy(), atol=2e-5), j


def test_oracle_ppo_step_runs_on_dict_network():
    O, A, B, T, E, M = 16, 3, 16, 8, 2, 2
    net = dictnet.make_dict_actor_critic(SIZES, A, ENC, [16], [16], seed=1)
    e = oenv.SyntheticEnv(O, A, max_len=8, term_thresh16=3000)
    ts = oppo.new_training_state(e, net, B, 17)
    p0 = net.flat_params().copy()
    for _ in range(2):
        ts, m = oppo.ppo_step(e, ts, B, T, n_epochs=E, n_minibatches=M)
    assert all(np.isfinite(float(v)) for v in m.values())
    assert float(ts.steps_taken) == 2 * T * B == float(net.counter)
    assert ts.opt.count == 2 * E * M and np.abs(net.flat_params() - p0).max() > 0
