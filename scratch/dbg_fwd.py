import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nnx_ppo_b200 import _lib, Rngs
from nnx_ppo_b200.algorithms import ppo
from nnx_ppo_b200.algorithms.engine import PPOEngine, AdamOptimizer
from nnx_ppo_b200.envs import SyntheticEnv
from nnx_ppo_b200.networks.plan import compile_network
from nnx_ppo_b200.networks.factories import make_mlp_actor_critic
lib = _lib.load()
dev = torch.device("cuda:0")
def run(O, A, ah, ch, B, T, act):
    nets = make_mlp_actor_critic(O, A, ah, ch, Rngs(1), activation=act)
    net = compile_network(nets)
    env = SyntheticEnv(O, A, max_len=24, term_thresh16=700)
    ts = ppo.new_training_state(env, nets, B, 17)
    eng = PPOEngine(net, env, ts.optimizer, B, T, 1, 2, 0.95, 0.99, 0.2, True, 1.0, use_graph=False)
    g = torch.Generator(device="cpu").manual_seed(0)
    eng.obs.copy_(torch.randn(T, B, O, generator=g) * 3)
    eng.next_obs_last.copy_(torch.randn(B, O, generator=g))
    eng.inds.copy_(torch.randperm(B, generator=g).to(torch.int32).reshape(1, B))
    net.normalizer.prepare(_lib.current_stream())
    mb = B // 2
    R = T * mb
    outs = {}
    for gemm in (0, 1):
        lib.b200ppo_set_gemm_mode(gemm)
        _lib.check(lib.b200ppo_update(_lib.current_stream(), net.plan, eng.hp, eng.bufs[0], T, B, mb, 0, 0, _lib.STAGE_FWD))
        torch.cuda.synchronize()
        def dbg(which, n):
            p = lib.b200ppo_update_debug_ptr(net.plan, T, mb, eng.ws.data_ptr(), which)
            off = (p - eng.ws.data_ptr()) // 4
            return eng.ws[off:off + n].cpu().numpy().copy()
        outs[gemm] = (dbg(2, R * 2 * A).reshape(R, 2 * A), dbg(1, R + mb))
    lib.b200ppo_set_gemm_mode(1)
    dy = np.abs(outs[0][0] - outs[1][0]); dv = np.abs(outs[0][1] - outs[1][1])
    print(O, A, ah, ch, B, T, act, "y err max", dy.max(), "per col", dy.max(0), "rows bad", (dy.max(1) > 1e-4).sum(), "/", R, "v err", dv.max())
    bad = np.nonzero(dy.max(1) > 1e-4)[0]
    if len(bad): print(" bad rows", bad[:20], "..", bad[-5:])
run(24, 5, [48, 40], [72], 70, 7, "swish")
run(24, 3, [48, 40], [72], 70, 7, "swish")
run(24, 5, [48], [72], 70, 7, "swish")
run(24, 5, [48, 48], [72], 70, 7, "swish")
run(24, 5, [48, 40], [32], 70, 7, "swish")
run(24, 5, [48, 40], [72], 96, 9, "swish")
run(24, 8, [64, 64], [72], 128, 8, "tanh")
run(64, 8, [64, 64, 64, 64], [256, 256], 256, 8, "tanh")
run(5, 1, [64, 64, 64, 64], [256, 256], 96, 30, "tanh")
