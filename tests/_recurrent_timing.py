"""BASELINE configs[2] at full size (profiling aid, not a bench line): ms / iteration through the public API and
CUDA-event timing of the pieces of one minibatch update (sequence forward / backward, MLP-path stages)."""
import sys
sys.path.insert(0, '/root/repo')
import time
import torch
from nnx_ppo_b200 import Rngs, _lib
from nnx_ppo_b200.algorithms import ppo
from nnx_ppo_b200.envs import SyntheticEnv
from nnx_ppo_b200.networks.factories import make_recurrent_actor_critic
from nnx_ppo_b200.networks.plan import compile_network
O, A, B, T, E, M = 64, 8, 4096, 32, 4, 8
nets = make_recurrent_actor_critic(O, A, 64, 256, [256, 256], Rngs(0))
env = SyntheticEnv(O, A, max_len=64, term_thresh16=512)
ts = ppo.new_training_state(env, nets, B, 17)
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ts, m = ppo.ppo_step(env, ts, B, T, 0.95, 0.99, 0.2, True, False, E, M)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"iter {it}: {dt*1e3:.1f} ms  ({B*T/dt/1e3:.1f} K samples/s)", {k: round(float(v), 5) for k, v in m.items() if 'mean' in k}, flush=True)
net = compile_network(nets)
eng = next(iter(net.engines.values()))
print("seq kernels:", eng.r_seq, "graph:", eng.r_graph is not None, "launches/iter:", eng.kernel_launches_per_iter)
if eng.r_seq:
    lib, lp, plan, mb = eng.lib, net.lplan, net.plan, eng.mb
    s = _lib.current_stream()
    arena = net.arena.data_ptr()
    ip = eng.inds.data_ptr()

    def timed(name, fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"  {name}: {e0.elapsed_time(e1) / reps * 1e3:.1f} us", flush=True)
    args = (s, plan, eng.hp, eng.bufs[0], T, B, mb, 2 * T, 0)
    timed("MLP-path FWD (critic + stand-in actor)", lambda: _lib.check(lib.b200ppo_update(*args, _lib.STAGE_FWD)))
    timed("seq_forward  (T=32, 512 rows)", lambda: _lib.check(lib.b200ppo_lstm_seq_forward(
        s, lp, arena, 0, 0, eng.r_xhat_ptr, eng.done.data_ptr(), ip, B, eng.r_c.data_ptr(), eng.r_h.data_ptr(), T, mb,
        eng.r_ws.data_ptr(), eng.r_y_ptr, 1)))
    timed("GAE + LOSS", lambda: _lib.check(lib.b200ppo_update(*args, _lib.STAGE_GAE | _lib.STAGE_LOSS)))
    timed("MLP-path BWD + RED", lambda: _lib.check(lib.b200ppo_update(*args, _lib.STAGE_BWD | _lib.STAGE_RED)))
    timed("seq_backward (T=32, 512 rows)", lambda: _lib.check(lib.b200ppo_lstm_seq_backward(
        s, lp, arena, eng.r_xhat_ptr, eng.r_dy_ptr, eng.done.data_ptr(), ip, B, T, mb, eng.r_ws.data_ptr(), eng.r_grad_ptr)))
    c, h = eng.r_carry
    mean_p, std_p = net.norm_ptrs()
    timed("rollout policy step (T=1, 4096 rows)", lambda: _lib.check(lib.b200ppo_lstm_seq_forward(
        s, lp, arena, mean_p, std_p, eng.r_env.obs.data_ptr(), 0, 0, B, c.data_ptr(), h.data_ptr(), 1, B,
        eng.r_ws.data_ptr(), eng.r_y.data_ptr(), 0)))
