#!/usr/bin/env python
"""Headline benchmark: PPO samples/sec (rollout policy + GAE + update) on the synthetic env.

    python bench.py --gpus N --steps K --warmup W [--config mlp|cartpole_shapes|recurrent|dict]
    python bench.py --impl reference --steps K --warmup W [--config ...]

A "step" is one full PPO iteration on each GPU: T-step rollout of n_envs envs, permutation indices,
E epochs x M minibatches of forward / GAE / loss / backward / Adam, Normalizer statistics.
--config selects the BASELINE.json configuration (default `mlp` = configs[1], the one the metric is
quoted on; `cartpole_shapes` = configs[0] shapes on the synthetic env, `recurrent` = configs[2],
`dict` = configs[3]; configs[4] is `mlp` at --gpus 8).
`value` is device-timed (CUDA events) whole-job samples/sec with everything resident in HBM;
`e2e` is the same metric through the public `ppo_step` API including the per-iteration
host->device block (PRNG keys + hyper-parameters, pinned), the device->host metrics read and the
host sync (the reference's `throughput/train_sps` definition, ppo.py:191-214).
--impl reference: the real reference (nnx_ppo on the JAX CPU build) when `jax`, `flax`, `optax` and
`nnx_ppo` are importable (also from baseline/_ref); otherwise the NumPy restatement (oracle/), kind
"port".
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

CONFIGS = {
    "mlp": dict(obs=64, act=8, actor=[64] * 4, critic=[256] * 2, n_envs=4096, T=32, E=4, M=8,
                workload="configs[1]: synthetic env obs=64 act=8, MLP actor 4x64 / critic 2x256, n_envs=4096/GPU, "
                         "rollout_length=32, 4 epochs x 8 minibatches"),
    "cartpole_shapes": dict(obs=5, act=1, actor=[64] * 4, critic=[256] * 2, n_envs=1024, T=30, E=4, M=4,
                            workload="configs[0] shapes (CartpoleBalance: obs=5 act=1) on the synthetic env, MLP actor 4x64 / "
                                     "critic 2x256, n_envs=1024/GPU, rollout_length=30, 4 epochs x 4 minibatches (MJX absent)"),
    "recurrent": dict(obs=64, act=8, pre=64, hidden=256, critic=[256] * 2, n_envs=4096, T=32, E=4, M=8,
                      workload="configs[2]: recurrent actor Dense(64) -> LSTM(256) -> Dense (the reference has an LSTM, no GRU), "
                               "MLP critic 2x256, carry reset on done, carry replay across epochs, n_envs=4096/GPU, "
                               "rollout_length=32, 4 epochs x 8 minibatches"),
    "dict": dict(obs=768, act=21, obs_sizes={"proprio": 256, "target": 512}, enc={"proprio": [128, 64], "target": [128, 64]},
                 actor=[256, 256], critic=[256, 256], n_envs=8192, T=32, E=4, M=8,
                 workload="configs[3]: dict observations {proprio 256, target 512} -> per-key encoders 128-64 -> trunk 2x256, "
                          "act=21, n_envs=8192/GPU, rollout_length=32, 4 epochs x 8 minibatches"),
}
ENV_KW = dict(max_len=64, term_thresh16=512)
SEED, NET_SEED = 17, 0
METRIC = "PPO samples/sec (rollout policy + GAE + update)"


def _rollout_engine(lib, eng, cfg, recurrent) -> str:
    """Which rollout path ran (include/b200ppo.h, b200ppo_set_rollout_mode)."""
    if recurrent:
        return "per-step launches: tcgen05 sequence kernels with T = 1 + sampler + env (algorithms/recurrent.py)"
    mode = int(lib.b200ppo_set_rollout_mode(-1))
    n = int(lib.b200ppo_rollout_synth_num_launches(eng.net.plan, cfg["T"], cfg["n_envs"], 1))
    if n > 1:
        return (f"batched: {n} launches per rollout, one tcgen05 3xTF32 tile GEMM per layer and step over all envs, "
                "weights pre-split once per rollout (the network does not fit shared memory)")
    return ("fused persistent kernel, one launch: " +
            ("warp-level mma.sync m16n8k8 3xTF32 tiles, weights resident in shared memory" if mode >= 1 else "fp32 FFMA tiles"))


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def _chain_mm(sizes):
    return sum(x * y for x, y in zip(sizes[:-1], sizes[1:]))


def _flops(name, cfg, mb):
    """Algorithmic FLOPs per launch of the three GEMM stages of one update (real parameters only: the
    structural zeros of the block-diagonal encoder layers are not counted)."""
    T = cfg["T"]
    R, Rv = T * mb, (T + 1) * mb
    if name == "dict":
        def tower(trunk, out):
            enc = sum(_chain_mm([cfg["obs_sizes"][k]] + cfg["enc"][k]) for k in cfg["obs_sizes"])
            enc_first = sum(cfg["obs_sizes"][k] * cfg["enc"][k][0] for k in cfg["obs_sizes"])
            width = sum(cfg["enc"][k][-1] for k in cfg["obs_sizes"])
            tr = _chain_mm([width] + trunk + [out])
            return enc + tr, enc + tr - enc_first
        pa, pa_dx = tower(cfg["actor"], 2 * cfg["act"])
        pc, pc_dx = tower(cfg["critic"], 1)
    else:
        a = [cfg["obs"]] + cfg["actor"] + [2 * cfg["act"]]
        c = [cfg["obs"]] + cfg["critic"] + [1]
        pa, pc = _chain_mm(a), _chain_mm(c)
        pa_dx, pc_dx = _chain_mm(a[1:]), _chain_mm(c[1:])
    return {"fwd": 2.0 * (pa * R + pc * Rv), "bwd_dx": 2.0 * (pa_dx + pc_dx) * R, "bwd_dw": 2.0 * (pa + pc) * R}


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        self.nvml = None
        # NVML from a thread every ~2 ms: a 20-step timed region is ~0.1 s, less than nvidia-smi needs to start
        try:
            import threading
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(index)
            try:
                bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
                h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._h, self._nv = h, pynvml
            self._sm, self._bits, self._stop = [], 0, threading.Event()
            self._mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        self._sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        get = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
                        self._bits |= int(get(h))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.nvml = threading.Thread(target=loop, daemon=True)
            self.nvml.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.nvml.join(timeout=2)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            reasons = sorted(n for b, n in names.items() if self._bits & b)
            return {"sm_mhz": statistics.median(self._sm) if self._sm else None, "sm_max_mhz": self._mx,
                    "samples": len(self._sm), "reasons": reasons, "source": "nvml, sampled every ~2 ms inside the timed region"}
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


# ---------------------------------------------------------------------------------------------
# CPU arms
# ---------------------------------------------------------------------------------------------
def _oracle_runner(name, cfg, B):
    """(step_fn, label) for the NumPy restatement of `name` on B envs per step."""
    from oracle import env as oenv, nets as onets, ppo as oppo
    oe = oenv.SyntheticEnv(cfg["obs"], cfg["act"], **ENV_KW)
    kw = dict(n_epochs=cfg["E"], n_minibatches=cfg["M"])
    if name == "recurrent":
        from oracle import recurrent as orec
        onet = orec.make_recurrent_actor_critic(cfg["obs"], cfg["act"], [cfg["pre"]], cfg["hidden"], [], cfg["critic"], seed=NET_SEED)
        st = [orec.new_training_state(oe, onet, B, SEED)]

        def step():
            st[0], m = orec.ppo_step(oe, st[0], B, cfg["T"], **kw)
            return m
        return step
    if name == "dict":
        from oracle import dictnet
        onet = dictnet.make_dict_actor_critic(cfg["obs_sizes"], cfg["act"], cfg["enc"], cfg["actor"], cfg["critic"], seed=NET_SEED)
    else:
        onet = onets.make_mlp_actor_critic(cfg["obs"], cfg["act"], cfg["actor"], cfg["critic"], seed=NET_SEED)
    st = [oppo.new_training_state(oe, onet, B, SEED)]

    def step():
        st[0], m = oppo.ppo_step(oe, st[0], B, cfg["T"], **kw)
        return m
    return step


def _try_import_reference():
    """jax + flax + optax + nnx_ppo importable (also from baseline/_ref)?  Returns the modules or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        os.environ.setdefault("JAX_PLATFORMS", "cpu")
        import jax  # noqa: F401
        import flax  # noqa: F401
        import optax  # noqa: F401
        from nnx_ppo.algorithms import ppo as rppo  # noqa: F401
        return True
    except Exception:
        return None


def _jax_reference_runner(name, cfg, B):
    """The UNMODIFIED reference (`nnx_ppo.algorithms.ppo.ppo_step` under nnx.jit) on the synthetic env of
    SURVEY.md section 8(d) restated as a JAX env.  MLP configs only (the reference factory)."""
    import jax
    import jax.numpy as jp
    from flax import nnx
    from nnx_ppo.algorithms import ppo as rppo
    from nnx_ppo.algorithms.types import LoggingLevel
    from nnx_ppo.jax_dataclass import JaxDataclass
    from nnx_ppo.networks import factories
    import dataclasses as dc
    from oracle import env as oenv
    if name not in ("mlp", "cartpole_shapes"):
        raise NotImplementedError("reference arm on JAX: MLP configs only")
    O, A = cfg["obs"], cfg["act"]
    Wo_np, Wa_np = oenv.make_env_weights(O, A, 0)
    Wo, Wa = jp.asarray(Wo_np), jp.asarray(Wa_np)
    max_len, thr = ENV_KW["max_len"], ENV_KW["term_thresh16"]

    @dc.dataclass
    class S(JaxDataclass):
        obs: jax.Array
        reward: jax.Array
        done: jax.Array
        info: dict
        metrics: dict

    class Env:
        def reset(self, rng):
            k_base, k_cnt = jax.random.split(rng)
            kc = jax.random.key_data(k_cnt)
            return S(jax.random.normal(k_base, (O,)), jp.float32(0.0), jp.float32(0.0),
                     {"step_counter": jax.random.randint(k_cnt, (), 0, max_len // 2),
                      "term_state": (kc[0] ^ kc[1]).astype(jp.uint32), "truncated": jp.array(False)}, {})

        def step(self, s, a):
            obs = jp.tanh(s.obs @ Wo + a @ Wa)
            cnt = s.info["step_counter"] + 1
            term = s.info["term_state"] * jp.uint32(1664525) + jp.uint32(1013904223)
            truncated = cnt >= max_len
            done = ((term >> 16) < thr) | truncated
            return S(obs, -jp.mean(obs * obs), done.astype(jp.float32),
                     {"step_counter": cnt, "term_state": term, "truncated": truncated}, {})

    env = Env()
    nets = factories.make_mlp_actor_critic(O, A, cfg["actor"], cfg["critic"], nnx.Rngs(NET_SEED))
    st = [rppo.new_training_state(env, nets, B, SEED)]
    step_jit = nnx.jit(rppo.ppo_step, static_argnums=(0, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13))

    def step():
        st[0], m = step_jit(env, st[0], B, cfg["T"], 0.95, 0.99, 0.2, True, False, cfg["E"], cfg["M"], 1.0,
                            LoggingLevel.LOSSES, None)
        int(st[0].steps_taken)                       # the reference's per-iteration host sync (ppo.py:209)
        return m
    return step


def run_reference(args):
    """CPU arm on rank 0: the reference's own implementation when importable, else the oracle port, on a
    bounded sample of the workload (B envs of n_envs per step, everything else as configured), with all
    the host threads BLAS / XLA will use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):   # torchrun pins these to 1
        os.environ[v] = str(cores)
    name, cfg = args.config, CONFIGS[args.config]
    B = min(args.ref_envs, cfg["n_envs"])
    kind, step = "port", None
    if _try_import_reference():
        try:
            step = _jax_reference_runner(name, cfg, B)
            kind = "reference"
        except Exception as e:                               # noqa: BLE001
            sys.stderr.write(f"[bench] reference on JAX unavailable for this config ({e}); using the NumPy port\n")
    if step is None:
        step = _oracle_runner(name, cfg, B)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    sps = B * cfg["T"] * args.steps / dt
    sample = (f"{B} of {cfg['n_envs']} envs per step, full T / epochs / minibatch count, "
              + ("nnx_ppo.ppo_step under nnx.jit on the JAX CPU build" if kind == "reference"
                 else "NumPy float32 restatement (oracle/), BLAS threads = host cores"))
    line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "reference_arm_sample": sample},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if kind == "port":
        line["note"] = ("reference (JAX / flax / optax) not importable here: NumPy restatement of the same iteration; "
                        "pinned by the reference's GAE / Normalizer KATs, the rest parity-unpinned (DESIGN.md section 4)")
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------
def build_workload(name, cfg):
    from nnx_ppo_b200 import Rngs
    from nnx_ppo_b200.envs import SyntheticEnv
    from nnx_ppo_b200.networks import factories
    env = SyntheticEnv(cfg["obs"], cfg["act"], **ENV_KW)
    if name == "recurrent":
        nets = factories.make_recurrent_actor_critic(cfg["obs"], cfg["act"], cfg["pre"], cfg["hidden"], cfg["critic"], Rngs(NET_SEED))
    elif name == "dict":
        nets = factories.make_dict_actor_critic(cfg["obs_sizes"], cfg["act"], cfg["enc"], cfg["actor"], cfg["critic"], Rngs(NET_SEED))
    else:
        nets = factories.make_mlp_actor_critic(cfg["obs"], cfg["act"], cfg["actor"], cfg["critic"], Rngs(NET_SEED))
    return env, nets


def time_stages(eng, ts_env_state, lib, _lib, torch, local_only=False, graphs=True):
    """Per-stage CUDA-event timing of the E*M updates of one iteration.  `graphs` (single GPU): the E*M launches of a
    stage are captured into one CUDA graph and a replay is timed, so the figure is the stage's time on the device,
    kernel boundaries included, without the host's launch gaps (eager, a Python -> ctypes launch every ~15 us leaves the
    GPU idle between the events: the in-kernel globaltimer stamps of profiles/r2c_cta_times.log put the forward kernel
    at 39 us where eager events read 50-56).  With the peer exchange (N > 1) the launches run eagerly, once: every
    exchange epoch can be used only once.  `local_only`: the same launches without the peer exchange (world-size-1
    arithmetic on this rank's shard), so that the difference is what the cross-GPU synchronisation costs."""
    net = eng.net
    T, B, mb = eng.T, eng.B, eng.mb
    # "gae_loss": both stages in one call, as the iteration runs them (one fused launch, csrc/update.cu
    # upd_gae_loss_kernel); "gae" / "loss" alone are the two stand-alone kernels (idempotent recomputations)
    stages = [("fwd", _lib.STAGE_FWD), ("gae", _lib.STAGE_GAE), ("loss", _lib.STAGE_LOSS),
              ("gae_loss", _lib.STAGE_GAE | _lib.STAGE_LOSS),
              ("bwd_dx", _lib.STAGE_BWD_DX), ("bwd_dw", _lib.STAGE_BWD_DW), ("red_adam", _lib.STAGE_RED | _lib.STAGE_ADAM)]
    if eng.hp.world_size > 1 or local_only:
        # with the peer exchange every timed stage that waits for the other ranks also measures their launch skew:
        # keep the two exchange points of the real iteration (loss, Adam), not a third one
        stages = [st for st in stages if st[0] not in ("gae", "loss")]
    tot = {k: 0.0 for k, _ in stages}
    saved = None
    if local_only:
        saved = (eng.hp.world_size, [b.comm for b in eng.bufs])
        eng.hp.world_size = 1
        for b in eng.bufs:
            b.comm = 0

    def launch(u, mask):
        off = 2 * T + u * 2 * (T + 1)
        _lib.check(lib.b200ppo_update(_lib.current_stream(), net.plan, eng.hp, eng.bufs[u], T, B, mb, off, u, mask))

    if graphs and eng.hp.world_size == 1:
        for name, mask in stages:
            for u in range(eng.n_updates):            # eager once: every buffer a later stage reads exists
                launch(u, mask)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for u in range(eng.n_updates):
                    launch(u, mask)
            g.replay()
            torch.cuda.synchronize()
            reps = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                reps.append(e0.elapsed_time(e1))
            tot[name] = sorted(reps)[1]
            del g
    else:
        evs = []
        for u in range(eng.n_updates):
            row = []
            for name, mask in stages:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                launch(u, mask)
                e1.record()
                row.append((name, e0, e1))
            evs.append(row)
        torch.cuda.synchronize()
        for row in evs:
            for name, e0, e1 in row:
                tot[name] += e0.elapsed_time(e1)
    if saved is not None:
        eng.hp.world_size = saved[0]
        for b, c in zip(eng.bufs, saved[1]):
            b.comm = c
    if ts_env_state is not None and not local_only:
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
    from nnx_ppo_b200 import _lib, prng
    from nnx_ppo_b200.algorithms import ppo
    from nnx_ppo_b200.networks.plan import compile_network
    lib = _lib.load()
    name, cfg = args.config, CONFIGS[args.config]
    env, nets = build_workload(name, cfg)
    ts = ppo.new_training_state(env, nets, cfg["n_envs"], SEED)
    hyper = (cfg["n_envs"], cfg["T"], 0.95, 0.99, 0.2, True, False, cfg["E"], cfg["M"])
    net = compile_network(nets)
    recurrent = bool(net.recurrent)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up through the public API (iteration 0 eager, iteration 1 captures the CUDA graph)
    # at least W (>= 3) steps, then as many more as ~0.3 s of device work takes, so that the clocks have ramped (three
    # 5 ms steps after the idle time of import / build are not enough: the first timed steps then run below the final
    # clock).  The count is rank 0's, so that every rank runs the same number of (collective) iterations.
    n_warm = max(args.warmup, 3)
    for _ in range(n_warm):
        ts, metrics = ppo.ppo_step(env, ts, *hyper)
    float(metrics["losses/actor/mean"])
    barrier()
    t_w = time.perf_counter()
    ts, metrics = ppo.ppo_step(env, ts, *hyper)
    float(metrics["losses/actor/mean"])
    barrier()
    extra = torch.tensor([min(200, int(0.3 / max(time.perf_counter() - t_w, 1e-4)))], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.broadcast(extra, src=0)
    for _ in range(int(extra.item())):
        ts, metrics = ppo.ppo_step(env, ts, *hyper)
    float(metrics["losses/actor/mean"])
    n_warm += 1 + int(extra.item())
    eng = next(reversed(net.engines.values()))
    samples_per_step = cfg["n_envs"] * cfg["T"] * world

    # ---- device-timed value: K iterations back to back, inputs resident in HBM ----
    clocks = ClockSampler(local) if rank == 0 else None
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if recurrent:
        for _ in range(args.steps):                  # no device-only entry point on this path: the public call
            ts, metrics = ppo.ppo_step(env, ts, *hyper)
    else:
        key = ts.rng_key
        for _ in range(args.steps):
            rk, nk = prng.split(key)
            eng.step(ts.env_states, rk, nk, fetch_metrics=False)
            key = nk
        ts = ts.replace(rng_key=key, steps_taken=np.float32(ts.steps_taken + args.steps * cfg["n_envs"] * cfg["T"]))
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clock_info = clocks.stop() if clocks is not None else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    value = samples_per_step * args.steps / (ms / 1e3)

    # ---- e2e: public API, per-step H2D (pinned key + hyper-parameter block) + D2H (metrics) + host sync ----
    # ppo_step returns as soon as the iteration is enqueued (its metrics dict materialises on first read, like the
    # asynchronously dispatched arrays of the reference's jitted step); every step's metrics ARE read here, one step
    # late, so the host work of step k+1 overlaps the device work of step k.
    barrier()
    t0 = time.perf_counter()
    prev, read_back = None, 0.0
    for _ in range(args.steps):
        ts, metrics = ppo.ppo_step(env, ts, *hyper)
        if prev is not None:
            read_back += float(prev["losses/actor/mean"])
        prev = metrics
    read_back += float(prev["losses/actor/mean"])
    barrier()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = samples_per_step * args.steps / float(t_e.item())

    hbm, tf_peak, which = _peaks()
    line = {"metric": METRIC, "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "global_envs": cfg["n_envs"] * world, "parallelism": f"dp{world} (env-sharded)",
                       "l2": "no flush: every update streams a > 100 MB working set (around / above the 126 MB L2) that the iteration itself rewrites",
                       "collectives": ("none" if world == 1 else
                                       ("peer memory over NVLink inside the loss / Adam kernels: 8-byte {payload, epoch} words, "
                                         + ("gradient as reduce-scatter + all-gather (owner rank per 256-parameter block)"
                                            if world >= int(os.environ.get("B200PPO_P2P_2HOP", "4")) else
                                            "every rank pushes its gradient to every peer")
                                        if eng.p2p else "NCCL all-reduce x2 per update")),
                       "rollout_critic": "off: the fused rollout does not evaluate the critic (training replays it, ppo.py:425-446; "
                                         "value estimates are computed on request for logging)",
                       "rollout_engine": _rollout_engine(lib, eng, cfg, recurrent),
                       "pdl": int(lib.b200ppo_set_pdl(-1)),
                       "cuda_graph": getattr(eng, "graph", None) is not None, "done_rate": float(eng.done.float().mean()),
                       "truncation_rate": float(eng.trunc.float().mean())},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": eng.h2d_bytes_per_step(),
                    "d2h_bytes_per_step": eng.d2h_bytes_per_step(), "api": "ppo.ppo_step",
                    "read": "every step's loss metrics are read on the host inside the timed region, one step late (lazy metrics dict)",
                    "ms_per_step": 1e3 * float(t_e.item()) / args.steps},
            "gpu_launches": (getattr(eng, "kernel_launches_per_iter", 0) or 0) * args.steps,
            "clocks": clock_info,
            "final_metrics": {k: float(v) for k, v in metrics.items()}}
    # ---- roofline of the dominant kernel (every N); cpu baseline (N = 1) ----
    if not args.quick and not recurrent:
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
        # per-stage timing re-runs the E*M updates of the last rollout stage by stage (this perturbs the
        # parameters once more; harmless for a bench).  All ranks do it in lock step.
        barrier()
        stage_ms = time_stages(eng, ts.env_states, lib, _lib, torch)
        sync_us = None
        if world > 1 and eng.p2p:
            barrier()
            local_ms = time_stages(eng, None, lib, _lib, torch, local_only=True)
            sync_us = 1e3 * (sum(stage_ms[k] for k in local_ms) - sum(local_ms.values())) / eng.n_updates
        barrier()
        U = eng.n_updates
        flops = _flops(name, cfg, eng.mb)
        dom = max(flops, key=lambda k: stage_ms[k])
        ach = flops[dom] * U / (stage_ms[dom] * 1e-3) / 1e12
        mode = int(lib.b200ppo_set_gemm_mode(-1))
        tc = mode != 0
        kname = {"fwd": "upd_fwd", "bwd_dx": "upd_bwd_dx", "bwd_dw": "upd_bwd_dw"}[dom] + ("_tc_kernel" if tc else "_kernel")
        if tc and dom == "bwd_dw" and name != "dict":
            kname = "upd_bwd_dw_tc2_kernel"       # bulk-copy + shared-memory transpose variant (every layer <= 256 wide)
        traffic = None                            # dram bytes per launch from the committed ncu --set full capture
        for tfile in ("r2_traffic.json", "r1_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tfile)
            if name == "mlp" and os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(kname)
                break
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
                            "stage_timing": ("one CUDA-graph replay of the E*M launches of each stage (device time, no host launch gaps)"
                                             if world == 1 else "eager launches, one event pair per launch (includes host launch gaps)"),
                            "small_kernel_share": (stage_ms["gae_loss"] + stage_ms["red_adam"]) / (ms / args.steps),
                            "stage_tflops_algorithmic": {k: flops[k] * U / (stage_ms[k] * 1e-3) / 1e12 for k in flops}}
        if sync_us is not None:
            line["roofline"]["sync_us_per_update"] = sync_us
            line["roofline"]["sync_note"] = ("difference of eagerly launched stages with and without the peer exchange: an upper bound that "
                                             "includes the ranks' host launch skew (the in-graph cost is the N-GPU minus the 1-GPU iteration "
                                             "time over the 32 updates)")
    if not args.quick and recurrent:
        line["roofline"] = recurrent_roofline(eng, net, cfg, lib, _lib, torch, tf_peak, which, ms / args.steps)
    if world == 1 and not args.quick and rank == 0:
        # CPU baseline: the oracle on a bounded sample of the same workload
        Bs = min(args.ref_envs, cfg["n_envs"])
        step = _oracle_runner(name, cfg, Bs)
        step()
        n_cpu = 0
        t0 = time.perf_counter()
        while n_cpu < 2 or (time.perf_counter() - t0 < 10.0 and n_cpu < 40):
            step()
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


def recurrent_roofline(eng, net, cfg, lib, _lib, torch, tf_peak, which, ms_iter):
    """configs[2]: the dominant piece is the recurrent actor's replay (b200ppo_lstm_seq_forward) and its reverse
    pass (b200ppo_lstm_seq_backward): launch sequences, timed inside a captured CUDA graph on the engine's own
    buffers (the state of the last update), algorithmic flops of the Dense -> LSTM -> Dense actor over T x rows."""
    lp, T, mb, B = net.lplan, eng.T, eng.mb, eng.B
    O, P, H, Y = lp.obs_dim, lp.pre_dim, lp.hidden, lp.out_dim
    arena, ip = net.arena.data_ptr(), eng.inds.data_ptr()

    def fwd(s_):
        _lib.check(lib.b200ppo_lstm_seq_forward(s_, lp, arena, 0, 0, eng.r_xhat_ptr, eng.done.data_ptr(), ip, B,
                                                eng.r_c.data_ptr(), eng.r_h.data_ptr(), T, mb, eng.r_ws.data_ptr(), eng.r_y_ptr, 1))

    def bwd(s_):
        _lib.check(lib.b200ppo_lstm_seq_backward(s_, lp, arena, eng.r_xhat_ptr, eng.r_dy_ptr, eng.done.data_ptr(), ip, B,
                                                 T, mb, eng.r_ws.data_ptr(), eng.r_grad_ptr))

    def graph_us(fn, n=8):
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            fn(_lib.current_stream())
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for _ in range(n):
                    fn(_lib.current_stream())
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (2 * n)

    if not getattr(eng, "r_seq", False):
        return None
    us_f, us_b = graph_us(fwd), graph_us(bwd)
    R = T * mb
    f_fwd = 2.0 * R * (O * P + (P + H) * 4 * H + H * Y)
    f_bwd = 2.0 * f_fwd - 2.0 * R * O * P            # dX and dW of every layer; no input gradient for the first
    ach = (f_fwd + f_bwd) / ((us_f + us_b) * 1e-6) / 1e12
    U = eng.n_updates
    return {"bound": "tensor", "kernel": "b200ppo_lstm_seq_forward + b200ppo_lstm_seq_backward (per-step tcgen05 kernels + batched tile GEMMs)",
            "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None,
            "peak_source": f"{which} bf16 tensor (sustained)",
            "compute_path": "tcgen05.mma kind::tf32, error-compensated 3xTF32; achieved counts ALGORITHMIC flops",
            "ceiling_3xtf32_tflops": tf_peak / 6.0, "frac_of_3xtf32_ceiling": ach / (tf_peak / 6.0),
            "flop_per_launch": f_fwd + f_bwd, "launch_ms": (us_f + us_b) * 1e-3,
            "seq_forward_us": us_f, "seq_backward_us": us_b,
            "share_of_iteration": U * (us_f + us_b) * 1e-3 / ms_iter,
            "note": "latency bound: 2 x T dependent step launches per update on 4 row tiles (profiles/r2_recurrent_notes.md)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="mlp", choices=sorted(CONFIGS))
    ap.add_argument("--ref-envs", type=int, default=512)
    ap.add_argument("--quick", action="store_true", help="skip the roofline stage timing and the CPU baseline (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
