"""CPU-side checks: the C-ABI library builds/loads and exports exactly what include/b200ppo.h
declares; host-side key arithmetic and config mirror the reference (no compute calls here)."""
import ctypes
import dataclasses
import os
import re

import numpy as np
import pytest

from nnx_ppo_b200 import _lib, prng
from nnx_ppo_b200.algorithms import config, ppo
from nnx_ppo_b200.algorithms.types import LoggingLevel
from nnx_ppo_b200.networks import factories
from nnx_ppo_b200.networks.normalizer import Normalizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from nnx_ppo_b200 import build
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "b200ppo.h")).read()
    declared = set(re.findall(r"\b(b200ppo_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.b200ppo_version() == 100
    assert b"invalid" in lib.b200ppo_error_string(-1)


def test_struct_sizes_match_header_layout():
    assert ctypes.sizeof(_lib.Chain) == 4 + 4 + 9 * 4 + 4 + 8 * 8 + 8 * 8
    assert ctypes.sizeof(_lib.Plan) == 16 + 16 + 8 + 2 * ctypes.sizeof(_lib.Chain)
    assert ctypes.sizeof(_lib.HParams) == 52
    assert ctypes.sizeof(_lib.UpdateBufs) == 21 * 8
    assert ctypes.sizeof(_lib.SynthEnv) == 32


def test_header_is_plain_c_and_matches_the_ctypes_mirror(tmp_path):
    """The boundary is a C ABI: include/b200ppo.h must compile as C99 (no C++ / torch types), and the
    struct layouts a C caller sees must be the ones the ctypes binding (and the tests) use."""
    import subprocess
    structs = {"b200ppo_chain": _lib.Chain, "b200ppo_plan": _lib.Plan, "b200ppo_hparams": _lib.HParams,
               "b200ppo_update_bufs": _lib.UpdateBufs, "b200ppo_synth_env": _lib.SynthEnv,
               "b200ppo_lstm_plan": _lib.LstmPlan}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200ppo.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == ctypes.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, (cname, fname)


def test_argument_validation_without_gpu(lib):
    """Error behaviour of the ABI: bad arguments are rejected before any launch."""
    assert lib.b200ppo_gae(None, None, None, None, None, None, 4, 4, 0.9, 0.9, None) == -1
    assert lib.b200ppo_gae(None, None, None, None, None, None, 0, 4, 0.9, 0.9, None) == 0   # empty
    assert lib.b200ppo_permutation(None, None, 0, 4, None, None) == 0                        # empty
    assert lib.b200ppo_permutation(None, None, 8, 4, None, None) == -1
    assert lib.b200ppo_random_bits(None, 0, 0, -1, None) == -1
    plan = _lib.Plan()
    assert lib.b200ppo_update_workspace_bytes(plan, 4, 4) < 0


def test_host_prng_matches_jax_known_values():
    assert prng.split(prng.key(0)) == [(1797259609, 2579123966), (928981903, 3453687069)]
    assert prng.fold_in(prng.key(0), 1) == (928981903, 3453687069)
    r = prng.Rngs(0)
    assert r() == prng.fold_in(prng.key(0), 0) and r.count == 1


def test_config_fields_mirror_reference_defaults():
    """algorithms/config.py:11-68 — names and defaults are API."""
    p = config.PPOConfig()
    assert dataclasses.asdict(p) | {"logging_level": None} == {
        "n_envs": 256, "rollout_length": 20, "total_steps": 512_000, "gae_lambda": 0.95,
        "discounting_factor": 0.99, "clip_range": 0.2, "learning_rate": 1e-4,
        "normalize_advantages": True, "combine_advantages": False, "n_epochs": 4,
        "n_minibatches": 4, "critic_loss_weight": 1.0, "gradient_clipping": None,
        "weight_decay": None, "logging_level": None, "logging_percentiles": None}
    assert p.logging_level == LoggingLevel.LOSSES
    t = config.TrainConfig()
    assert t.seed == 17 and t.checkpoint_every_steps == 500_000
    assert config.EvalConfig().logging_percentiles == (0, 25, 50, 75, 100)
    assert ppo._should_run(50_000, 0, 50_000) and not ppo._should_run(49_999, 0, 50_000)
    assert not ppo._should_run(10, 0, 0)


def test_factory_topology_and_init_match_oracle():
    """factories.py:72-146: Sequential([Normalizer, PPOAdapter]); key-draw order of the init."""
    from oracle import nets as onets
    nets = factories.make_mlp_actor_critic(24, 5, [64, 64], [32], prng.Rngs(42), activation="tanh")
    assert isinstance(nets.layers[0], Normalizer)
    o = onets.make_mlp_actor_critic(24, 5, [64, 64], [32], seed=42, activation="tanh")
    actor = nets.layers[1].action.layers
    for l, (W, b) in zip(actor[:-1], zip(o.actor.W, o.actor.b)):
        assert np.array_equal(l.linear.kernel.numpy(), W) and np.array_equal(l.linear.bias.numpy(), b)
    assert actor[-1].rng.count == o.rng_count == 2 * 5
    bare = factories.make_mlp_actor_critic(4, 1, [8], [8], prng.Rngs(0), normalize_obs=False)
    assert type(bare).__name__ == "PPOAdapter"


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "nnx_ppo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has CUDA")
    from nnx_ppo_b200.envs import SyntheticEnv
    nets = factories.make_mlp_actor_critic(8, 2, [16], [16], prng.Rngs(0))
    with pytest.raises(_lib.B200PPOError):
        ppo.new_training_state(SyntheticEnv(8, 2), nets, 16, 0)


def test_reward_scaling_wrapper_scales_reset_and_step_rewards():
    """wrappers/reward_scaling_wrapper.py:8-28 on a batched (CPU tensor) env state."""
    import torch
    from nnx_ppo_b200.wrappers import RewardScalingWrapper

    @dataclasses.dataclass
    class S:
        obs: torch.Tensor
        reward: torch.Tensor
        done: torch.Tensor
        info: dict
        metrics: dict

    class Env:
        observation_size, action_size = 3, 2

        def reset(self, keys):
            B = keys.shape[0]
            return S(torch.zeros(B, 3), torch.full((B,), 2.0), torch.zeros(B), {}, {})

        def step(self, s, a):
            return S(s.obs + 1, a.sum(dim=1), s.done, s.info, s.metrics)

    w = RewardScalingWrapper(Env(), 0.25)
    assert (w.observation_size, w.action_size) == (3, 2)
    s0 = w.reset(torch.zeros(4, 2, dtype=torch.int32))
    assert torch.equal(s0.reward, torch.full((4,), 0.5))
    s1 = w.step(s0, torch.ones(4, 2))
    assert torch.equal(s1.reward, torch.full((4,), 0.5)) and torch.equal(s1.obs, torch.ones(4, 3))


def test_host_side_sizing_entry_points_for_the_bench_configuration(lib):
    """The ABI's host-only entry points (no launch) on the BASELINE configs[1] plan: workspace size,
    kernel launches per update (what `bench.py` reports as gpu_launches), scratch sizes, pointer
    helpers stay inside the workspace.  The plan compiler runs without a device."""
    import torch
    from nnx_ppo_b200.networks.plan import CompiledNet
    nets = factories.make_mlp_actor_critic(64, 8, [64] * 4, [256] * 2, prng.Rngs(0))
    net = CompiledNet(nets, torch.device("cpu"))
    assert net.n_params == 100_372
    T, mb = 32, 512
    ws = int(lib.b200ppo_update_workspace_bytes(net.plan, T, mb))
    assert 100e6 < ws < 130e6 and ws % 16 == 0                            # "~113 MB at cfg 2"
    assert int(lib.b200ppo_update_workspace_bytes(net.plan, T, 2 * mb)) > ws
    hp = _lib.HParams()
    hp.gamma, hp.lambda_, hp.clip_range, hp.critic_loss_weight = 0.99, 0.95, 0.2, 1.0
    hp.learning_rate, hp.adam_b1, hp.adam_b2, hp.adam_eps = 1e-4, 0.9, 0.999, 1e-8
    hp.weight_decay, hp.grad_clip, hp.normalize_advantages, hp.world_size = -1.0, -1.0, 1, 1
    first = int(lib.b200ppo_update_num_launches(net.plan, hp, T, mb, _lib.STAGE_ALL))
    later = int(lib.b200ppo_update_num_launches(net.plan, hp, T, mb, _lib.STAGE_ALL | _lib.STAGE_NO_PREP))
    assert (first, later) == (6, 5)            # prep + fwd, gae+loss (one fused launch), dX, dW, reduce+adam
    split = int(lib.b200ppo_update_num_launches(net.plan, hp, T, mb, _lib.STAGE_GAE)) + \
        int(lib.b200ppo_update_num_launches(net.plan, hp, T, mb, _lib.STAGE_LOSS))
    assert split == 2                          # the stages called one by one stay two kernels
    # one iteration: 32 updates + norm prepare, rollout, permutation, 2 stats passes, merge, finalize
    assert first + 31 * later + 7 == 168       # bench.py's gpu_launches at configs[1]
    hp.grad_clip = 0.5                         # clip_by_global_norm: reduce (+ exchange) + norm, then the clipped Adam
    assert int(lib.b200ppo_update_num_launches(net.plan, hp, T, mb, _lib.STAGE_ALL | _lib.STAGE_NO_PREP)) == 6
    assert int(lib.b200ppo_update_num_launches(net.plan, hp, T, mb, _lib.STAGE_RED)) == 1
    assert int(lib.b200ppo_update_num_launches(net.plan, hp, T, mb, _lib.STAGE_ADAM)) == 2
    base = 1 << 20
    for fn in (lib.b200ppo_update_adv_sums_ptr, lib.b200ppo_update_grad_ptr):
        p = int(fn(net.plan, T, mb, base))
        assert base <= p < base + ws and p % 16 == 0
    assert int(lib.b200ppo_permutation_scratch_bytes(4096, 4)) > 0
    assert int(lib.b200ppo_norm_scratch_bytes(64)) > 0
    assert int(lib.b200ppo_comm_bytes(net.plan, 8)) > 8 * 2 * 4 * net.n_params   # two parities x 8 rank slots


def test_rollout_engine_selection_is_host_logic(lib):
    """`b200ppo_rollout_synth_workspace_bytes` / `_num_launches`: networks whose weights fit an SM's shared memory keep
    the fused one-launch rollout (no workspace); BASELINE configs[3] (768-wide dict-observation encoders, 789 x 768
    env matrix) takes the batched per-step path once a workspace is offered.  Host-only (no launch)."""
    import torch
    from nnx_ppo_b200.networks.plan import CompiledNet
    mlp = CompiledNet(factories.make_mlp_actor_critic(64, 8, [64] * 4, [256] * 2, prng.Rngs(0)), torch.device("cpu"))
    assert int(lib.b200ppo_rollout_synth_workspace_bytes(mlp.plan, 4096)) == 0
    assert int(lib.b200ppo_rollout_synth_num_launches(mlp.plan, 32, 4096, 1)) == 1
    wide = CompiledNet(factories.make_dict_actor_critic({"proprio": 256, "target": 512}, 21, {"proprio": [128, 64], "target": [128, 64]},
                                                        [256, 256], [256, 256], prng.Rngs(0)), torch.device("cpu"))
    L = wide.plan.actor.n_layers
    assert L == 5
    ws = int(lib.b200ppo_rollout_synth_workspace_bytes(wide.plan, 8192))
    assert 80e6 < ws < 300e6 and ws % 4 == 0
    assert int(lib.b200ppo_rollout_synth_num_launches(wide.plan, 32, 8192, 1)) == (L + 1) + 1 + 32 * (L + 3)
    assert int(lib.b200ppo_rollout_synth_num_launches(wide.plan, 32, 8192, 0)) == 1        # no workspace: fused kernel
    assert int(lib.b200ppo_rollout_synth_workspace_bytes(wide.plan, 256)) == 0             # few envs: tiles would idle
    assert int(lib.b200ppo_rollout_synth_workspace_bytes(None, 8192)) == -1
    prev = lib.b200ppo_set_rollout_mode(0)                                                  # FFMA tiles: always fused
    try:
        assert int(lib.b200ppo_rollout_synth_workspace_bytes(wide.plan, 8192)) == 0
    finally:
        lib.b200ppo_set_rollout_mode(prev)


def test_host_key_arithmetic_equals_the_oracle_restatement():
    """nnx_ppo_b200/prng.py (host key flow of the product) against oracle/prng.py (NumPy restatement) on
    random keys: split, fold_in, uniform and the variance-scaling initializer draw."""
    from oracle import prng as oprng
    g = np.random.default_rng(5)
    for _ in range(20):
        k = (int(g.integers(0, 2 ** 32)), int(g.integers(0, 2 ** 32)))
        ok = np.array(k, np.uint32)
        n = int(g.integers(1, 9))
        assert prng.split(k, n) == [tuple(int(x) for x in r) for r in oprng.split(ok, n)]
        d = int(g.integers(0, 2 ** 32))
        assert prng.fold_in(k, d) == tuple(int(x) for x in oprng.fold_in(ok, d))
        shape = (int(g.integers(1, 7)), int(g.integers(1, 7)))
        assert np.array_equal(prng.uniform(k, shape, -1.0, 1.0), oprng.uniform(ok, shape, -1.0, 1.0))
        assert np.array_equal(prng.variance_scaling_uniform(k, *shape), oprng.variance_scaling_uniform(ok, *shape, 1.0))
    assert prng.key(2 ** 40 + 3) == (2 ** 8, 3) and prng.key(17) == (0, 17)


def test_rollout_and_policy_launchers_reject_before_launching(lib):
    """Error behaviour of the K1 launchers with a valid plan but unusable arguments: every check runs on
    the host before the first CUDA call, so the codes can be observed without a GPU (the fake pointers
    are never dereferenced on these paths)."""
    import torch
    from nnx_ppo_b200.networks.plan import CompiledNet
    EINVAL, ELIMIT = -1, -2
    net = CompiledNet(factories.make_mlp_actor_critic(12, 3, [16], [16], prng.Rngs(0)), torch.device("cpu"))
    env = _lib.SynthEnv()
    env.obs_dim, env.act_dim, env.max_len, env.term_thresh16 = 12, 3, 16, 100
    env.Wo, env.Wa = 0x10000, 0x10000 + 4 * 12 * 12
    p = 0x20000                                           # any non-null, 16-byte aligned address
    ok_args = [None, net.plan, env, p, p, p, p, p, 8, 32] + [p] * 11
    def rollout(**kw):
        a = list(ok_args)
        for i, v in kw.items():
            a[int(i[1:])] = v
        return lib.b200ppo_rollout_synth(*a)
    assert rollout(a3=None) == EINVAL                      # params
    assert rollout(a8=0) == EINVAL and rollout(a9=0) == EINVAL      # T, B
    assert rollout(a4=None) == EINVAL                      # normalising plan without statistics
    bad_env = _lib.SynthEnv()
    bad_env.obs_dim, bad_env.act_dim, bad_env.max_len, bad_env.term_thresh16 = 11, 3, 16, 100
    bad_env.Wo, bad_env.Wa = env.Wo, env.Wa
    assert rollout(a2=bad_env) == EINVAL                   # env / plan size mismatch
    env.Wa += 4
    assert rollout() == EINVAL                             # Wa must follow Wo contiguously
    env.Wa -= 4
    assert lib.b200ppo_eval_synth(None, net.plan, env, p, p, p, p, 1, 10, 32, p, p, p, p, p) == EINVAL   # mode: 0 or 2
    assert lib.b200ppo_eval_synth(None, net.plan, env, p, p, p, p, 2, 0, 32, p, p, p, p, p) == EINVAL    # L
    # the env half of a rollout step for policies evaluated elsewhere: same host-side checks
    nb = int(lib.b200ppo_synth_env_step_workspace_bytes(12, 3, 32))
    assert nb > 0 and nb % 4 == 0 and int(lib.b200ppo_synth_env_step_workspace_bytes(0, 3, 32)) == -1
    assert lib.b200ppo_synth_env_begin(None, env, 32, p, p, p, nb - 4) == EINVAL         # workspace too small
    assert lib.b200ppo_synth_env_begin(None, env, 32, None, p, p, nb) == EINVAL          # env observations
    assert lib.b200ppo_synth_env_begin(None, env, 0, p, p, p, nb) == EINVAL              # B
    step_ok = [None, env, p, 6, 0.1, 1.0, p, p, 0, 8, 32] + [p] * 11 + [p, nb]
    def env_step(**kw):
        a = list(step_ok)
        for i, v in kw.items():
            a[int(i[1:])] = v
        return lib.b200ppo_synth_env_step(*a)
    assert env_step(a2=None) == EINVAL                     # actor outputs
    assert env_step(a3=5) == EINVAL                        # ldy < 2 * act_dim
    assert env_step(a8=8) == EINVAL and env_step(a8=-1) == EINVAL      # t outside [0, T)
    assert env_step(a23=nb - 4) == EINVAL                  # workspace too small
    assert env_step(a22=None) == EINVAL                    # no workspace
    # tiles that cannot fit in shared memory even without the k-split scratch tile
    wide = CompiledNet(factories.make_mlp_actor_critic(6000, 3, [16], [16], prng.Rngs(0)), torch.device("cpu"))
    wenv = _lib.SynthEnv()
    wenv.obs_dim, wenv.act_dim, wenv.max_len, wenv.term_thresh16 = 6000, 3, 16, 100
    wenv.Wo, wenv.Wa = 0x10000, 0x10000 + 4 * 6000 * 6000
    a = list(ok_args)
    a[1], a[2] = wide.plan, wenv
    assert lib.b200ppo_rollout_synth(*a) == ELIMIT
    assert lib.b200ppo_eval_synth(None, wide.plan, wenv, p, p, p, p, 2, 10, 32, p, p, p, p, p) == ELIMIT
    assert lib.b200ppo_policy_step(None, wide.plan, p, p, p, p, 32, 0, p, 0, None, p, p, p, p, None, None) == ELIMIT
    assert lib.b200ppo_policy_step(None, net.plan, p, p, p, p, 0, 0, p, 0, None, p, p, p, p, None, None) == 0   # empty batch
    assert lib.b200ppo_policy_step(None, net.plan, p, p, p, p, 32, 1, p, 0, None, p, p, p, p, None, None) == EINVAL  # replay needs raw actions
    assert b"limit" in lib.b200ppo_error_string(ELIMIT).lower()


def test_update_launcher_rejects_before_launching(lib):
    """b200ppo_update's host-side argument checks (all before the first CUDA call)."""
    import torch
    from nnx_ppo_b200.networks.plan import CompiledNet
    EINVAL, EALIGN = -1, -3
    net = CompiledNet(factories.make_mlp_actor_critic(12, 3, [16], [16], prng.Rngs(0)), torch.device("cpu"))
    hp = _lib.HParams()
    hp.world_size = 1
    b = _lib.UpdateBufs()
    p = 0x40000
    for name, _ in _lib.UpdateBufs._fields_:
        if name not in ("comm", "param_mask"):
            setattr(b, name, p)
    call = lambda T=8, B=32, mb=16, u=0, hp_=hp, b_=b, plan=net.plan: lib.b200ppo_update(
        None, plan, hp_, b_, T, B, mb, 0, u, _lib.STAGE_ALL)
    assert call(T=0) == EINVAL and call(mb=0) == EINVAL and call(mb=64) == EINVAL and call(u=-1) == EINVAL
    assert call(plan=_lib.Plan()) != 0                                     # empty plan
    b.ws = p + 16
    assert call() == EALIGN                                                # workspace must be 256-byte aligned
    b.ws = p
    b.norm_mean = 0
    assert call() == EINVAL                                                # normalising plan without statistics
    b.norm_mean = p
    b.obs = 0
    assert call() == EINVAL
    b.obs = p
    hp.world_size = 0
    assert call() == EINVAL
    hp.world_size, hp.rank, b.comm, hp.grad_clip = 4, 4, p, -1.0
    assert call() == EINVAL                                                # rank out of range on the peer-memory path
    hp.rank = 1
    hp.world_size = _lib.MAX_RANKS + 1 if hasattr(_lib, "MAX_RANKS") else 17
    hp.grad_clip = -1.0
    assert call() == EINVAL


def test_dw_row_split_table_is_a_one_wave_cover(lib):
    """Host-only: the weight-gradient kernel's work table (greedy minimax over the measured per-stage costs, csrc/update.cu
    make_layout) for the four BASELINE plans: every item covers all rows in splits of a multiple of 32 rows, the splits
    fit one wave (<= one CTA per SM), nearly all SMs are used, and the costlier 256-column items get shorter row ranges
    than the 64-column ones."""
    import ctypes
    import torch
    import bench
    from nnx_ppo_b200.networks.plan import CompiledNet
    sms = int(lib.b200ppo_num_sms())
    for name in ("mlp", "cartpole_shapes", "dict"):
        cfg = bench.CONFIGS[name]
        _, nets = bench.build_workload(name, cfg)
        net = CompiledNet(nets, torch.device("cpu"))
        T, mb = cfg["T"], cfg["n_envs"] // cfg["M"]
        R = T * mb
        S = (ctypes.c_int32 * 32)()
        rps = (ctypes.c_int32 * 32)()
        n = int(lib.b200ppo_update_dw_splits(net.plan, T, mb, S, rps, 32))
        dims = []
        for ch in (net.plan.actor, net.plan.critic):
            for l in range(ch.n_layers):
                dims += [(ch.dims[l], ch.dims[l + 1])] * ((ch.dims[l] + 127) // 128)
        assert n == len(dims) and n >= 2
        tot = 0
        for i in range(n):
            assert S[i] >= 1 and rps[i] % 32 == 0 and S[i] * rps[i] >= R > (S[i] - 1) * rps[i], (name, i, S[i], rps[i])
            tot += S[i]
        assert sms - n <= tot <= sms, (name, tot, sms)
        wide = [rps[i] for i in range(n) if dims[i][1] > 128 and dims[i][0] % 4 == 0]
        narrow = [rps[i] for i in range(n) if dims[i][1] == 64 and dims[i][0] % 4 == 0]
        if wide and narrow:
            assert max(wide) < min(narrow), (name, wide, narrow)


def test_workspace_size_does_not_depend_on_the_kernel_structure_switches(lib):
    """b200ppo_set_update_paths changes the weight-gradient kernel's row-split table (host-only here); the partial-
    gradient buffer is sized for both tables, so flipping the switch on a live workspace cannot overrun it."""
    import torch
    from nnx_ppo_b200.networks.plan import CompiledNet
    nets = factories.make_mlp_actor_critic(64, 8, [64] * 4, [256] * 2, prng.Rngs(0))
    net = CompiledNet(nets, torch.device("cpu"))
    prev = lib.b200ppo_set_update_paths(-1, -1)
    try:
        sizes, tables = [], []
        for mn in (1, 0):
            lib.b200ppo_set_update_paths(-1, mn)
            sizes.append(int(lib.b200ppo_update_workspace_bytes(net.plan, 32, 512)))
            S = (ctypes.c_int32 * 32)()
            rps = (ctypes.c_int32 * 32)()
            n = int(lib.b200ppo_update_dw_splits(net.plan, 32, 512, S, rps, 32))
            tables.append(list(S[:n]))
        assert sizes[0] == sizes[1]
        assert tables[0] != tables[1]          # the tables themselves do differ (different per-stage costs)
    finally:
        lib.b200ppo_set_update_paths(prev & 1, (prev >> 1) & 1)
