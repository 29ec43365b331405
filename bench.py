#!/usr/bin/env python
"""Headline benchmark: PPO samples/sec (rollout policy + GAE + update) on the synthetic env.

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA engine)
    python bench.py --impl reference --steps K --warmup W    # CPU arm (oracle restatement)

A "step" is one full PPO iteration of BASELINE.json configs[1] on each GPU: fused 32-step rollout
of 4096 envs (obs 64, act 8, actor 4x64, critic 2x256, normalize_obs), permutation indices,
4 epochs x 8 minibatches of forward / GAE / loss / backward / Adam, Normalizer statistics.
`value` is device-timed (CUDA events) whole-job samples/sec with everything resident in HBM;
`e2e` is the same metric through the public `ppo_step` API including the per-iteration
host->device key block, the device->host metrics read and the host sync (the reference's
`throughput/train_sps` definition, ppo.py:191-214).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(obs=64, act=8, actor=[64] * 4, critic=[256] * 2, n_envs=4096, T=32, E=4, M=8,
           max_len=64, term_thresh16=512, seed=17, net_seed=0)
WORKLOAD = "configs[1]: synthetic env obs=64 act=8, MLP actor 4x64 / critic 2x256, n_envs=4096/GPU, rollout_length=32, 4 epochs x 8 minibatches"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def _p_mm(cfg):
    a = [cfg["obs"]] + cfg["actor"] + [2 * cfg["act"]]
    c = [cfg["obs"]] + cfg["critic"] + [1]
    pa = sum(x * y for x, y in zip(a[:-1], a[1:]))
    pc = sum(x * y for x, y in zip(c[:-1], c[1:]))
    pa_dx = sum(x * y for x, y in zip(a[1:-1], a[2:]))
    pc_dx = sum(x * y for x, y in zip(c[1:-1], c[2:]))
    return pa, pc, pa_dx, pc_dx


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """CPU arm: the reference's algorithm for this path on the host cores.  The reference itself
    (JAX/flax/optax) cannot be installed in this image, so this times the NumPy restatement
    (oracle/), kind="port", on a bounded sample of the same workload: 512 of the 4096 envs
    (same network, T, epochs and minibatch count; minibatch = 64 envs)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import env as oenv, nets as onets, ppo as oppo
    B = args.ref_envs
    cfg = CFG
    oe = oenv.SyntheticEnv(cfg["obs"], cfg["act"], cfg["max_len"], cfg["term_thresh16"])
    onet = onets.make_mlp_actor_critic(cfg["obs"], cfg["act"], cfg["actor"], cfg["critic"], seed=cfg["net_seed"])
    ots = oppo.new_training_state(oe, onet, B, cfg["seed"])
    for _ in range(args.warmup):
        ots, _ = oppo.ppo_step(oe, ots, B, cfg["T"], n_epochs=cfg["E"], n_minibatches=cfg["M"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ots, m = oppo.ppo_step(oe, ots, B, cfg["T"], n_epochs=cfg["E"], n_minibatches=cfg["M"])
    dt = time.perf_counter() - t0
    sps = B * cfg["T"] * args.steps / dt
    cores = os.cpu_count() or 1
    sample = f"{B} of {cfg['n_envs']} envs per step, full T/epochs/minibatch count, NumPy float32 (BLAS threads = host default)"
    line = {"impl": "reference", "metric": "PPO samples/sec (rollout policy + GAE + update)", "value": sps,
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "reference_arm_sample": sample},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference (JAX) not installable here: NumPy restatement of the same iteration, parity unpinned beyond the reference's KATs"}
    print(json.dumps(line), flush=True)


def time_stages(eng, ts_env_state, lib, _lib, torch):
    """Per-stage CUDA-event timing of the 32 updates of one iteration (eager launches on the
    current stream).  Returns {stage: total_ms} and leaves the training state consistent."""
    net = eng.net
    T, B, mb = eng.T, eng.B, eng.mb
    stages = [("fwd", _lib.STAGE_FWD), ("gae", _lib.STAGE_GAE), ("loss", _lib.STAGE_LOSS),
              ("bwd_dx", _lib.STAGE_BWD_DX), ("bwd_dw", _lib.STAGE_BWD_DW), ("red", _lib.STAGE_RED),
              ("adam", _lib.STAGE_ADAM)]
    tot = {k: 0.0 for k, _ in stages}
    s = _lib.current_stream()
    evs = []
    for u in range(eng.n_updates):
        off = 2 * T + u * 2 * (T + 1)
        row = []
        for name, mask in stages:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.b200ppo_update(s, net.plan, eng.hp, eng.bufs[u], T, B, mb, off, u, mask))
            e1.record()
            row.append((name, e0, e1))
        evs.append(row)
    torch.cuda.synchronize()
    for row in evs:
        for name, e0, e1 in row:
            tot[name] += e0.elapsed_time(e1)
    # the rollout launch alone, on a scratch copy of the env state (sampler counts are not advanced, so
    # the training state is untouched); median of 5
    scratch = type(ts_env_state)(ts_env_state.obs.clone(), ts_env_state.step_counter.clone(),
                                 ts_env_state.term_state.clone())
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng._enqueue_rollout(scratch)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    tot["rollout"] = sorted(times)[2]
    return tot


def run_own(args):
    import numpy as np
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from nnx_ppo_b200 import Rngs, _lib, prng
    from nnx_ppo_b200.algorithms import ppo
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks.factories import make_mlp_actor_critic
    lib = _lib.load()
    cfg = CFG
    env = SyntheticEnv(cfg["obs"], cfg["act"], cfg["max_len"], cfg["term_thresh16"])
    nets = make_mlp_actor_critic(cfg["obs"], cfg["act"], cfg["actor"], cfg["critic"], Rngs(cfg["net_seed"]))
    ts = ppo.new_training_state(env, nets, cfg["n_envs"], cfg["seed"])
    hyper = (cfg["n_envs"], cfg["T"], 0.95, 0.99, 0.2, True, False, cfg["E"], cfg["M"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up through the public API (iteration 0 eager, iteration 1 captures the CUDA graph)
    for _ in range(max(args.warmup, 3)):
        ts, metrics = ppo.ppo_step(env, ts, *hyper)
    eng = ppo._engine_for(env, ts, cfg["n_envs"], cfg["T"], 0.95, 0.99, 0.2, True, cfg["E"], cfg["M"], 1.0)
    samples_per_step = cfg["n_envs"] * cfg["T"] * world

    # ---- device-timed value: K iterations back to back, inputs resident in HBM ----
    clocks = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    key = ts.rng_key
    e0.record()
    for _ in range(args.steps):
        rk, nk = prng.split(key)
        eng.step(ts.env_states, rk, nk, fetch_metrics=False)
        key = nk
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop() if clocks is not None else None
    ts = ts.replace(rng_key=key, steps_taken=np.float32(ts.steps_taken + args.steps * cfg["n_envs"] * cfg["T"]))
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = samples_per_step * args.steps / (ms / 1e3)

    # ---- e2e: public API, per-step H2D (pinned key block) + D2H (metrics) + host sync ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ts, metrics = ppo.ppo_step(env, ts, *hyper)
    barrier()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / float(t_e.item())

    line = None
    if rank == 0:
        hbm, tf_peak, which = _peaks()
        line = {"metric": "PPO samples/sec (rollout policy + GAE + update)", "value": value, "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_envs": cfg["n_envs"] * world, "parallelism": f"dp{world} (env-sharded)",
                           "l2": "no flush: every update streams a ~150 MB working set (> 126 MB L2) that the iteration itself rewrites",
                           "collectives": ("none" if world == 1 else
                                           ("peer memory: epoch flags + P2P loads inside the GAE / loss / Adam kernels"
                                            if eng.p2p else "NCCL all-reduce x2 per update")),
                           "cuda_graph": eng.graph is not None, "done_rate": float(eng.done.float().mean()),
                           "truncation_rate": float(eng.trunc.float().mean())},
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": eng.h2d_bytes_per_step(),
                        "d2h_bytes_per_step": eng.d2h_bytes_per_step(), "api": "ppo.ppo_step",
                        "ms_per_step": 1e3 * float(t_e.item()) / args.steps},
                "gpu_launches": eng.kernel_launches_per_iter * args.steps,
                "clocks": clock_info,
                "final_metrics": {k: float(v) for k, v in metrics.items()}}
    # ---- roofline of the dominant kernel + cpu baseline: N = 1 only ----
    if world == 1 and not args.quick:
        # FFMA peak probe (fp32 CUDA-core ceiling; not in MEASURED_PEAKS.json)
        blocks, threads, iters = 148 * 8, 256, 20000
        sink = torch.zeros(blocks * threads, device="cuda")
        lib.b200ppo_ffma_peak(_lib.current_stream(), 1000, sink.data_ptr(), blocks, threads)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        lib.b200ppo_ffma_peak(_lib.current_stream(), iters, sink.data_ptr(), blocks, threads)
        f1.record()
        torch.cuda.synchronize()
        ffma_tf = blocks * threads * iters * 16 * 2 / (f0.elapsed_time(f1) * 1e-3) / 1e12
        # per-stage timing needs a fresh rollout in the buffers: run one API iteration, then re-time
        # its 32 updates stage by stage (this perturbs parameters once more; harmless for a bench)
        stage_ms = time_stages(eng, ts.env_states, lib, _lib, torch)
        eng.net.advance_rng(0)
        pa, pc, pa_dx, pc_dx = _p_mm(cfg)
        R, Rv, U = cfg["T"] * eng.mb, (cfg["T"] + 1) * eng.mb, eng.n_updates
        flops = {"fwd": 2.0 * (pa * R + pc * Rv), "bwd_dx": 2.0 * (pa_dx + pc_dx) * R, "bwd_dw": 2.0 * (pa + pc) * R}
        dom = max(flops, key=lambda k: stage_ms[k])
        ach = flops[dom] * U / (stage_ms[dom] * 1e-3) / 1e12
        mode = int(lib.b200ppo_set_gemm_mode(-1))
        tc = mode != 0
        kname = {"fwd": "upd_fwd", "bwd_dx": "upd_bwd_dx", "bwd_dw": "upd_bwd_dw"}[dom] + ("_tc_kernel" if tc else "_kernel")
        if tc and dom == "bwd_dw":
            kname = "upd_bwd_dw_tc2_kernel"       # bulk-copy + shared-memory transpose variant (every layer <= 256 wide)
        traffic = None                            # dram bytes per launch from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(kname)
        line["roofline"] = {"bound": "tensor", "kernel": kname,
                            "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                            "traffic": traffic, "peak_source": f"{which} bf16 tensor (sustained)",
                            "compute_path": ("tcgen05.mma kind::tf32, error-compensated 3xTF32 (3 MMAs per algorithmic product, fp32 accumulate in TMEM): "
                                             "achieved counts ALGORITHMIC flops; tensor-pipe work is 3x that" if mode == 1 else
                                             ("tcgen05.mma kind::tf32, plain TF32 (not fp32 parity)" if mode == 2 else "fp32 FFMA (CUDA cores)")),
                            # tf32 MMAs run at half the bf16 rate and every algorithmic product costs three of
                            # them: the ceiling of this formulation is peak / 6
                            "ceiling_3xtf32_tflops": tf_peak / 6.0, "frac_of_3xtf32_ceiling": ach / (tf_peak / 6.0),
                            "ffma_peak_tflops": ffma_tf, "frac_of_ffma_peak": ach / ffma_tf,
                            "flop_per_launch": flops[dom], "launch_ms": stage_ms[dom] / U,
                            "stage_ms_per_iteration": stage_ms,
                            "stage_tflops_algorithmic": {k: flops[k] * U / (stage_ms[k] * 1e-3) / 1e12 for k in flops}}
        # CPU baseline: the oracle on a bounded sample of the same workload
        from oracle import env as oenv, nets as onets, ppo as oppo
        Bs = args.ref_envs
        oe = oenv.SyntheticEnv(cfg["obs"], cfg["act"], cfg["max_len"], cfg["term_thresh16"])
        onet = onets.make_mlp_actor_critic(cfg["obs"], cfg["act"], cfg["actor"], cfg["critic"], seed=cfg["net_seed"])
        ots = oppo.new_training_state(oe, onet, Bs, cfg["seed"])
        ots, _ = oppo.ppo_step(oe, ots, Bs, cfg["T"], n_epochs=cfg["E"], n_minibatches=cfg["M"])
        n_cpu = 0
        t0 = time.perf_counter()
        while n_cpu < 3 or (time.perf_counter() - t0 < 10.0 and n_cpu < 40):
            ots, _ = oppo.ppo_step(oe, ots, Bs, cfg["T"], n_epochs=cfg["E"], n_minibatches=cfg["M"])
            n_cpu += 1
        cdt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": Bs * cfg["T"] * n_cpu / cdt, "unit": "samples/s", "cores": os.cpu_count(),
                                "kind": "port",
                                "sample": f"{n_cpu} iterations of {Bs} of {cfg['n_envs']} envs (same net/T/epochs/minibatch count), NumPy float32 oracle, {cdt:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down without running NCCL communicator destructors: destroying a communicator whose
        # kernels are still referenced by a captured CUDA graph hung the process at exit on B200
        # (driver 580.159, NCCL 2.28.9).  Everything is flushed and synchronised, so exit hard.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-envs", type=int, default=512)
    ap.add_argument("--quick", action="store_true", help="skip the roofline stage timing and the CPU baseline (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
