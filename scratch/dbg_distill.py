import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from nnx_ppo_b200 import _lib
from nnx_ppo_b200.algorithms import distillation
from nnx_ppo_b200.envs import SyntheticEnv
from nnx_ppo_b200.networks.plan import compile_network
from oracle import distill as odistill, env as oenv
import test_gpu_distill as tg
dev = torch.device("cuda:0")
for gemm, act in ((1, "tanh"), (0, "tanh"), (1, "relu")):
    _lib.load().b200ppo_set_gemm_mode(gemm)
    O, A, B, T, E, M = 24, 3, 96, 9, 2, 2
    student, teacher, ostudent, oteacher = tg._nets(O, A, [48], [40, 40], act)
    teacher.eval()
    ekw = dict(max_len=24, term_thresh16=700)
    env, oe = SyntheticEnv(O, A, **ekw), oenv.SyntheticEnv(O, A, **ekw)
    ds = distillation.new_distillation_state(env, teacher, student, B, 17, learning_rate=3e-4)
    ods = odistill.new_distillation_state(oe, ostudent, B, 17)
    net = compile_network(student)
    for it in range(2):
        ds, m = distillation.distillation_step(env, teacher, ds, B, T, E, M)
        len(m)
        tr = {}
        ods, om = odistill.distillation_step(oe, oteacher, ods, B, T, n_epochs=E, n_minibatches=M, learning_rate=3e-4, trace=tr)
        eng = next(iter(net.engines.values()))
        print(gemm, act, it, "gpu", eng.metrics_host.numpy()[:, :3].tolist())
        print(gemm, act, it, "ora", tr["loss_per_update"].tolist())
        print("mu err", np.abs(eng.teacher_mu.cpu().numpy() - tr["teacher_mu"]).max(), "param err", np.abs(net.params_logical() - ostudent.flat_params()).max())
