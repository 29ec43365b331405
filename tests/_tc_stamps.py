import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
from nnx_ppo_b200 import Rngs, _lib, prng
from nnx_ppo_b200.algorithms import ppo
from nnx_ppo_b200.envs import SyntheticEnv
from nnx_ppo_b200.networks.factories import make_mlp_actor_critic
lib = _lib.load()
env = SyntheticEnv(64, 8); nets = make_mlp_actor_critic(64, 8, [64]*4, [256]*2, Rngs(0))
ts = ppo.new_training_state(env, nets, 4096, 17)
hyper = (4096, 32, 0.95, 0.99, 0.2, True, False, 4, 8)
for _ in range(2): ts, m = ppo.ppo_step(env, ts, *hyper)
eng = ppo._engine_for(env, ts, 4096, 32, 0.95, 0.99, 0.2, True, 4, 8, 1.0)
net = eng.net
buf = (ctypes.c_longlong * 128)()
def stamps(mask, name):
    s = _lib.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.b200ppo_update(s, net.plan, eng.hp, eng.bufs[0], 32, 4096, 512, 64, 0, mask))
    e1.record(); torch.cuda.synchronize()
    rc = lib.b200ppo_debug_timestamps(buf, 128); n = -(rc + 1000)
    t = np.array(buf[:n], dtype=np.int64); d = np.diff(t)
    print(name, 'us', round(e0.elapsed_time(e1)*1e3,1), 'n', n, 'total cycles', t[-1]-t[0]); print(d.tolist()); print('  probes:', list(buf[n:n+8]))
for _ in range(2): stamps(_lib.STAGE_FWD, 'fwd')
lib.b200ppo_set_gemm_mode(2)
stamps(_lib.STAGE_FWD, 'fwd-tf32x1')
stamps(_lib.STAGE_BWD, 'bwd-tf32x1(dw)')
lib.b200ppo_set_gemm_mode(1)
stamps(_lib.STAGE_GAE|_lib.STAGE_LOSS, 'gae+loss')
stamps(_lib.STAGE_BWD, 'bwd(dw last)')
for flag, nm in ((1,'dx'),):
    lib.b200ppo_debug_select(flag)
    stamps(_lib.STAGE_BWD, 'bwd(dw) '+nm)
for spin in (0,):
  for bx in (0, 5, 6, 8):
    lib.b200ppo_debug_select((bx << 8) | spin)
    stamps(_lib.STAGE_BWD, 'bwd(dw item %d spin %d)' % (bx, spin))
lib.b200ppo_debug_select(0)
