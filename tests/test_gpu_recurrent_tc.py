"""GPU parity of the tensor-core recurrent path (csrc/recurrent_tc.cu): the generic tcgen05 3xTF32 tile
kernels against float64 matmuls, and the whole-sequence forward / backward entry points
(b200ppo_lstm_seq_forward / _backward) against oracle/recurrent.py (replay from the minibatch's start carry
with the rollout's resets, BPTT fed with the oracle's d loss / d y).  SURVEY section 8 row a15."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nnx_ppo_b200 import _lib                                    # noqa: E402
from oracle import env as oenv, prng, recurrent as orec          # noqa: E402

from test_gpu_recurrent import _plan                             # noqa: E402

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def lib(cuda_device):
    from nnx_ppo_b200 import build
    build.build()
    return _lib.load()


@pytest.mark.parametrize("M,N,K,nt,slices", [(128, 64, 256, 0, 1), (300, 16, 64, 0, 1), (512, 1024, 64, 0, 1),
                                             (257, 6, 10, 0, 1), (512, 256, 1024, 1, 4), (200, 64, 1024, 1, 1),
                                             (130, 256, 16, 1, 1), (64, 40, 7, 1, 1)])
def test_generic_tile_gemm_matches_float64(cuda_device, lib, M, N, K, nt, slices):
    import torch
    g = np.random.default_rng(M + N + K)
    A = g.standard_normal((M, K)).astype(F)
    B = g.standard_normal((N, K) if nt else (K, N)).astype(F)
    dA, dB = torch.from_numpy(A).to(cuda_device), torch.from_numpy(B).to(cuda_device)
    C = torch.full((slices, M, N), 7.0, device=cuda_device)
    _lib.check(lib.b200ppo_rg_gemm_test(_lib.current_stream(), dA.data_ptr(), dB.data_ptr(), C.data_ptr(), M, N, K, nt, slices))
    torch.cuda.synchronize()
    ref = A.astype(np.float64) @ (B.astype(np.float64).T if nt else B.astype(np.float64))
    got = C.cpu().numpy().astype(np.float64).sum(axis=0)
    # error-compensated 3xTF32: fp32-level accuracy (tests/test_gpu_tensorcore.py measures 2-4x cuBLAS fp32)
    assert np.abs(got - ref).max() < 2e-5 * np.sqrt(K) * max(1.0, np.abs(ref).max() / np.sqrt(K))


@pytest.mark.parametrize("rows,Kdim,N,S", [(1000, 64, 64, 3), (4096, 320, 1024, 6), (555, 256, 16, 4), (97, 10, 7, 2),
                                           (16384, 64, 64, 64)])
def test_weight_gradient_tile_gemm_matches_float64(cuda_device, lib, rows, Kdim, N, S):
    import torch
    g = np.random.default_rng(rows + N)
    A = g.standard_normal((rows, Kdim)).astype(F)
    D = g.standard_normal((rows, N)).astype(F)
    dA, dD = torch.from_numpy(A).to(cuda_device), torch.from_numpy(D).to(cuda_device)
    scratch = torch.zeros(S * (Kdim + 1) * N + 64, device=cuda_device)
    outs = []
    for _ in range(2):
        W = torch.full((Kdim, N), 7.0, device=cuda_device)
        b = torch.full((N,), 7.0, device=cuda_device)
        _lib.check(lib.b200ppo_rg_tn_test(_lib.current_stream(), dA.data_ptr(), dD.data_ptr(), rows, Kdim, N, S,
                                          scratch.data_ptr(), W.data_ptr(), b.data_ptr()))
        torch.cuda.synchronize()
        outs.append((W.cpu().numpy(), b.cpu().numpy()))
    ref = A.astype(np.float64).T @ D.astype(np.float64)
    assert np.abs(outs[0][0] - ref).max() < 3e-5 * np.sqrt(rows)
    assert np.abs(outs[0][1] - D.astype(np.float64).sum(0)).max() < 3e-5 * np.sqrt(rows)
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])     # fixed-order sums


@pytest.mark.parametrize("cfg", [dict(O=64, A=8, B=40, T=16, P=64, H=256, act="relu", mb=list(range(0, 40, 2))),
                                 dict(O=16, A=4, B=37, T=20, P=32, H=32, act="relu", mb=list(range(37))),
                                 dict(O=10, A=3, B=300, T=9, P=12, H=16, act="tanh", mb=list(range(299, 10, -2))),
                                 dict(O=64, A=8, B=600, T=32, P=64, H=256, act="swish", mb=list(range(0, 600))),
                                 # trainable_initial_state: resets hand out the learned carry, which gets a gradient
                                 dict(O=16, A=4, B=300, T=20, P=32, H=32, act="tanh", mb=list(range(0, 300, 2)), init=True),
                                 dict(O=64, A=8, B=140, T=12, P=64, H=128, act="relu", mb=list(range(139, -1, -1)), init=True)])
@pytest.mark.parametrize("persist", [1, 0])
def test_sequence_replay_and_bptt_match_oracle(cuda_device, lib, cfg, persist):
    """persist = 1: the forward recurrence as one persistent launch (weights resident in shared memory, the carry
    handed between the CTAs of a row tile through arrival counters); 0: one launch per step."""
    import torch
    lib.b200ppo_lstm_set_persistent(persist)
    dev = cuda_device
    O, A, B, T, H, P = cfg["O"], cfg["A"], cfg["B"], cfg["T"], cfg["H"], cfg["P"]
    init = cfg.get("init", False)
    net = orec.make_recurrent_actor_critic(O, A, [P], H, [], [6], seed=3, activation=cfg["act"], trainable_initial_state=init)
    if init:                        # (the reference starts the learned carry at zero)
        gi = np.random.default_rng(4)
        net.init_c[:] = 0.5 * gi.standard_normal(H)
        net.init_h[:] = 0.5 * gi.standard_normal(H)
    e = oenv.SyntheticEnv(O, A, max_len=6, term_thresh16=9000)
    es = e.reset(prng.split(prng.key(5), B))
    es, carry, ro, start = orec.unroll_env(e, es, net, net.initialize_state(B), T, prng.key(11))
    net.update_statistics(ro.obs)                                 # a non-trivial normaliser
    assert ro.done.sum() > 0
    inds = np.asarray(cfg["mb"], np.int32)
    mb = len(inds)
    total, m, g_ref = orec.ppo_loss_and_grads(net, ro, start, inds, net.rng_count)
    plan, n_rec = _plan(net)
    assert lib.b200ppo_lstm_seq_supported(plan) == 1
    t = lambda a, dt=torch.float32: torch.tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    params = t(net.flat_params())
    done, ind_d = t(ro.done.astype(np.uint8), torch.uint8), t(inds, torch.int32)
    xs = net.normalize_obs(ro.obs[:, inds].reshape(T * mb, -1)).astype(F)          # what the critic pass leaves in xhat
    x = t(xs)
    Y = 2 * A
    ws = torch.zeros(int(lib.b200ppo_lstm_seq_workspace_floats(plan, T, mb)) + 64, device=dev)
    s = _lib.current_stream()
    # oracle replay
    cc, hh = start[0][inds].copy(), start[1][inds].copy()
    ys = []
    for k in range(T):
        cc, hh, yk, _ = orec.actor_step(net, cc, hh, xs.reshape(T, mb, -1)[k])
        ys.append(yk)
        cc, hh = net.reset_carry((cc, hh), ro.done[k, inds])
    y_ref = np.stack(ys)
    r_c = net.init_c if init else np.zeros(H, F)
    r_h = net.init_h if init else np.zeros(H, F)
    for raw in (False, True):      # replay form (normalised inputs) and rollout form (raw observations, normalised on load)
        c, h = t(start[0][inds]), t(start[1][inds])
        y = torch.zeros(T, mb, Y, device=dev)
        xin = t(ro.obs[:, inds].reshape(T * mb, -1)) if raw else x
        mean, std = (t(net.mean), t(net.norm_std())) if raw else (None, None)
        _lib.check(lib.b200ppo_lstm_seq_forward(s, plan, params.data_ptr(), _lib.ptr(mean), _lib.ptr(std), xin.data_ptr(),
                                                done.data_ptr(), ind_d.data_ptr(), B, c.data_ptr(), h.data_ptr(), T, mb,
                                                ws.data_ptr(), y.data_ptr(), 0 if raw else 1), "seq_forward")
        torch.cuda.synchronize()
        assert np.abs(y.cpu().numpy() - y_ref).max() < 3e-5 * max(1.0, np.abs(y_ref).max())
        assert np.abs(c.cpu().numpy() - cc).max() < 3e-5 and np.abs(h.cpu().numpy() - hh).max() < 3e-5
        d_last = ro.done[-1, inds]
        assert np.all(c.cpu().numpy()[d_last] == r_c) and np.all(h.cpu().numpy()[d_last] == r_h)     # reset: exact values
    # BPTT with the oracle's d loss / d y (the forward with keep_cache = 1 ran first in the loop above? no: last was raw)
    c, h = t(start[0][inds]), t(start[1][inds])
    _lib.check(lib.b200ppo_lstm_seq_forward(s, plan, params.data_ptr(), 0, 0, x.data_ptr(), done.data_ptr(), ind_d.data_ptr(),
                                            B, c.data_ptr(), h.data_ptr(), T, mb, ws.data_ptr(), y.data_ptr(), 1), "seq_forward")
    d_y = t(m["d_y"].reshape(T * mb, Y))
    outs = []
    for _ in range(2):
        grad = torch.full((plan.n_params,), 7.0, device=dev)
        _lib.check(lib.b200ppo_lstm_seq_backward(s, plan, params.data_ptr(), x.data_ptr(), d_y.data_ptr(), done.data_ptr(),
                                                 ind_d.data_ptr(), B, T, mb, ws.data_ptr(), grad.data_ptr()), "seq_backward")
        torch.cuda.synchronize()
        outs.append(grad.cpu().numpy().copy())
    ref = g_ref[:n_rec]
    scale = np.abs(ref).max()
    assert np.abs(outs[0][:n_rec] - ref).max() < 3e-4 * scale, (np.abs(outs[0][:n_rec] - ref).max(), scale)
    assert np.array_equal(outs[0], outs[1])                      # fixed-order sums: bit-reproducible
    assert np.all(outs[0][n_rec:] == 7.0)                        # the critic's slots are not touched
    assert lib.b200ppo_lstm_seq_num_launches(plan, T, mb, 0) == (1 if persist else T) + 6
    assert lib.b200ppo_lstm_seq_num_launches(plan, T, mb, 1) == 2 * T + 5 + (2 if init else 0)
    if init:                                                     # the learned carry's gradient is not negligible here
        o = plan.init_c_off
        assert np.abs(ref[o:o + 2 * H]).max() > 1e-3 * scale
        assert np.abs(outs[0][o:o + 2 * H] - ref[o:o + 2 * H]).max() < 3e-4 * np.abs(ref[o:o + 2 * H]).max()
    lib.b200ppo_lstm_set_persistent(0)
