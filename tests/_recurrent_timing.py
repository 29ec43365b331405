import sys, time
sys.path.insert(0, '/root/repo')
import torch
from nnx_ppo_b200 import Rngs
from nnx_ppo_b200.algorithms import ppo
from nnx_ppo_b200.envs import SyntheticEnv
from nnx_ppo_b200.networks.factories import make_recurrent_actor_critic
O, A, B, T, E, M = 64, 8, 4096, 32, 4, 8
nets = make_recurrent_actor_critic(O, A, 64, 256, [256, 256], Rngs(0))
env = SyntheticEnv(O, A, max_len=64, term_thresh16=512)
ts = ppo.new_training_state(env, nets, B, 17)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ts, m = ppo.ppo_step(env, ts, B, T, 0.95, 0.99, 0.2, True, False, E, M)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"iter {it}: {dt*1e3:.1f} ms  ({B*T/dt/1e3:.1f} K samples/s)", {k: round(float(v), 5) for k, v in m.items() if 'mean' in k})
