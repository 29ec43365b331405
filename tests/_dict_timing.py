"""BASELINE configs[3] at full size: dict obs {proprio 256, target 512} -> separate encoders -> trunk,
act = 21, n_envs = 8192, rollout 32, 4 epochs x 8 minibatches.  Prints ms / iteration (not a bench line)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from nnx_ppo_b200 import Rngs
from nnx_ppo_b200.algorithms import ppo
from nnx_ppo_b200.envs import SyntheticEnv
from nnx_ppo_b200.networks.factories import make_dict_actor_critic
from nnx_ppo_b200.networks.plan import compile_network
sizes = {"proprio": 256, "target": 512}
nets = make_dict_actor_critic(sizes, 21, {"proprio": [128, 64], "target": [128, 64]}, [256, 256], [256, 256], Rngs(0))
env = SyntheticEnv(768, 21, max_len=64, term_thresh16=512)
B, T = 8192, 32
ts = ppo.new_training_state(env, nets, B, 17)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ts, m = ppo.ppo_step(env, ts, B, T, 0.95, 0.99, 0.2, True, False, 4, 8)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"iter {it}: {dt*1e3:.1f} ms  ({B*T/dt/1e6:.2f} M samples/s)", {k: round(float(v), 5) for k, v in m.items() if 'mean' in k})
net = compile_network(nets)
assert float(net.arena[net.param_mask == 0].abs().max()) == 0.0 and bool(torch.isfinite(net.arena).all())
print("structural zeros intact, parameters finite; P =", net.n_params, "trainable =", int(net.param_mask.sum()))
