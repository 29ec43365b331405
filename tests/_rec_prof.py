"""Profiling aid: one whole-sequence forward + backward of the recurrent actor at BASELINE configs[2] size
(T = 32, 512-row minibatch, obs 64, pre 64, hidden 256, act 8) and one rollout policy step (4096 rows), eager
launches on random data.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel launch list."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from nnx_ppo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
O, P, H, Y, T, mb, B = 64, 64, 256, 16, 32, 512, 4096
p = _lib.LstmPlan()
p.obs_dim, p.pre_dim, p.hidden, p.out_dim, p.act, p.normalize = O, P, H, Y, 1, 1
o = 0
p.w1_off = o; o += O * P
p.b1_off = o; o += P
p.wcat_off = o; o += (P + H) * 4 * H
p.bl_off = o; o += 4 * H
p.w2_off = o; o += H * Y
p.b2_off = o; o += Y
p.n_params = o
g = torch.Generator(device=dev).manual_seed(0)
params = 0.05 * torch.randn(o, device=dev, generator=g)
x = torch.randn(T * mb, O, device=dev, generator=g)
done = (torch.rand(T, B, device=dev, generator=g) < 0.02).to(torch.uint8)
inds = torch.randperm(B, device=dev)[:mb].to(torch.int32)
c, h = torch.zeros(mb, H, device=dev), torch.zeros(mb, H, device=dev)
y = torch.zeros(T * mb, Y, device=dev)
dy = 1e-3 * torch.randn(T * mb, Y, device=dev, generator=g)
grad = torch.zeros(o, device=dev)
ws = torch.zeros(int(lib.b200ppo_lstm_seq_workspace_floats(p, T, mb)) + 64, device=dev)
s = _lib.current_stream()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for r in range(reps):
    c.zero_(); h.zero_()
    ev[0].record()
    _lib.check(lib.b200ppo_lstm_seq_forward(s, p, params.data_ptr(), 0, 0, x.data_ptr(), done.data_ptr(), inds.data_ptr(), B,
                                            c.data_ptr(), h.data_ptr(), T, mb, ws.data_ptr(), y.data_ptr(), 1))
    ev[1].record()
    _lib.check(lib.b200ppo_lstm_seq_backward(s, p, params.data_ptr(), x.data_ptr(), dy.data_ptr(), done.data_ptr(),
                                             inds.data_ptr(), B, T, mb, ws.data_ptr(), grad.data_ptr()))
    ev[2].record()
    torch.cuda.synchronize()
    print(f"rep {r}: forward {ev[0].elapsed_time(ev[1]) * 1e3:.0f} us, backward {ev[1].elapsed_time(ev[2]) * 1e3:.0f} us", flush=True)
assert bool(torch.isfinite(grad).all()) and bool(torch.isfinite(y).all())

# the same calls inside a captured CUDA graph (how the engine runs them): per-call time without host launch gaps
def _graph_time(fn, n=10):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn(_lib.current_stream())
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(n):
                fn(_lib.current_stream())
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (2 * n)


if len(sys.argv) > 2 and sys.argv[2] == "graph":
    fwd = lambda s_: _lib.check(lib.b200ppo_lstm_seq_forward(s_, p, params.data_ptr(), 0, 0, x.data_ptr(), done.data_ptr(), inds.data_ptr(), B,
                                                             c.data_ptr(), h.data_ptr(), T, mb, ws.data_ptr(), y.data_ptr(), 1))
    bwd = lambda s_: _lib.check(lib.b200ppo_lstm_seq_backward(s_, p, params.data_ptr(), x.data_ptr(), dy.data_ptr(), done.data_ptr(),
                                                              inds.data_ptr(), B, T, mb, ws.data_ptr(), grad.data_ptr()))
    print(f"in-graph: seq_forward {_graph_time(fwd):.0f} us, seq_backward {_graph_time(bwd):.0f} us", flush=True)
    for TT in (1, 2, 4, 8, 16):
        f2 = lambda s_: _lib.check(lib.b200ppo_lstm_seq_forward(s_, p, params.data_ptr(), 0, 0, x.data_ptr(), done.data_ptr(), inds.data_ptr(), B,
                                                                c.data_ptr(), h.data_ptr(), TT, mb, ws.data_ptr(), y.data_ptr(), 1))
        b2 = lambda s_: _lib.check(lib.b200ppo_lstm_seq_backward(s_, p, params.data_ptr(), x.data_ptr(), dy.data_ptr(), done.data_ptr(),
                                                                 inds.data_ptr(), B, TT, mb, ws.data_ptr(), grad.data_ptr()))
        print(f"in-graph T={TT}: seq_forward {_graph_time(f2):.0f} us, seq_backward {_graph_time(b2):.0f} us", flush=True)
