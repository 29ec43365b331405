"""ctypes binding of libb200ppo.so (the C ABI declared in include/b200ppo.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, this module
raises.  torch is used only as the device allocator / stream provider.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200ppo.so")

MAX_LAYERS = 8
ACT_IDS = {"none": 0, None: 0, "relu": 1, "swish": 2, "silu": 2, "tanh": 3}

STAGE_FWD, STAGE_GAE, STAGE_LOSS, STAGE_BWD, STAGE_RED, STAGE_ADAM = 1, 2, 4, 8, 16, 32
STAGE_ALL = 63
STAGE_BWD_DX, STAGE_BWD_DW = 64, 128
STAGE_NO_PREP = 256
STAGE_NLL = 512
METRICS_STRIDE = 12   # B200PPO_METRICS_STRIDE
# layout of the device hyper-parameter block (B200PPO_HP_*)
HP_GAMMA, HP_LAMBDA, HP_CLIP_RANGE, HP_CRITIC_WEIGHT, HP_LEARNING_RATE = 0, 1, 2, 3, 4
HP_ADAM_B1, HP_ADAM_B2, HP_ADAM_EPS, HP_WEIGHT_DECAY, HP_GRAD_CLIP = 5, 6, 7, 8, 9
HP_FLOATS = 16


class Chain(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("act", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)),
                ("pad_", C.c_int32), ("w_off", C.c_int64 * MAX_LAYERS), ("b_off", C.c_int64 * MAX_LAYERS)]


class Plan(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("normalize", C.c_int32),
                ("pad_", C.c_int32), ("entropy_weight", C.c_float), ("min_std", C.c_float),
                ("std_scale", C.c_float), ("pad2_", C.c_float), ("n_params", C.c_int64),
                ("actor", Chain), ("critic", Chain)]


class HParams(C.Structure):
    _fields_ = [("gamma", C.c_float), ("lambda_", C.c_float), ("clip_range", C.c_float),
                ("critic_loss_weight", C.c_float), ("learning_rate", C.c_float),
                ("adam_b1", C.c_float), ("adam_b2", C.c_float), ("adam_eps", C.c_float),
                ("weight_decay", C.c_float), ("grad_clip", C.c_float),
                ("normalize_advantages", C.c_int32), ("world_size", C.c_int32), ("rank", C.c_int32)]


class UpdateBufs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "obs", "raw_action", "loglik_old", "reward", "done", "truncated", "next_obs_last", "inds",
        "norm_mean", "norm_std", "params", "adam_mu", "adam_nu", "rng_state", "metrics_out", "ws", "comm",
        "param_mask", "hparams_dev", "comm_epoch", "param_tie")]


class LstmPlan(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("pre_dim", C.c_int32), ("hidden", C.c_int32), ("out_dim", C.c_int32),
                ("act", C.c_int32), ("normalize", C.c_int32),
                ("w1_off", C.c_int64), ("b1_off", C.c_int64), ("wcat_off", C.c_int64), ("bl_off", C.c_int64),
                ("w2_off", C.c_int64), ("b2_off", C.c_int64), ("n_params", C.c_int64),
                ("init_c_off", C.c_int64), ("init_h_off", C.c_int64)]


class SynthEnv(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("max_len", C.c_int32),
                ("term_thresh16", C.c_int32), ("Wo", C.c_void_p), ("Wa", C.c_void_p)]


# every symbol include/b200ppo.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _u32, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
_PP, _HP, _BP, _EP = C.POINTER(Plan), C.POINTER(HParams), C.POINTER(UpdateBufs), C.POINTER(SynthEnv)
SYMBOLS = {
    "b200ppo_version": (C.c_int, []),
    "b200ppo_error_string": (C.c_char_p, [C.c_int]),
    "b200ppo_num_sms": (C.c_int, []),
    "b200ppo_random_bits": (C.c_int, [_vp, _u32, _u32, _i64, _vp]),
    "b200ppo_random_normal": (C.c_int, [_vp, _u32, _u32, _i64, _vp]),
    "b200ppo_permutation_scratch_bytes": (_i64, [_i32, _i32]),
    "b200ppo_permutation": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "b200ppo_gae": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _vp]),
    "b200ppo_norm_prepare": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "b200ppo_norm_scratch_bytes": (_i64, [_i32]),
    "b200ppo_norm_batch_stats": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "b200ppo_norm_merge": (C.c_int, [_vp, _vp, _i32, _f32, _i32, _vp, _vp, _vp]),
    "b200ppo_policy_workspace_bytes": (_i64, [_PP, _i32]),
    "b200ppo_policy_step": (C.c_int, [_vp, _PP, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _u32, _vp,
                                      _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200ppo_synth_reset": (C.c_int, [_vp, _EP, _vp, _i32, _vp, _vp, _vp]),
    "b200ppo_synth_init_keys": (C.c_int, [_vp, _u32, _u32, _i32, _vp]),
    "b200ppo_split_keys_dev": (C.c_int, [_vp, _vp, _i64, _i32, _vp]),
    "b200ppo_split_rows": (C.c_int, [_vp, _vp, _i32, _u32, _vp]),
    "b200ppo_rollout_synth": (C.c_int, [_vp, _PP, _EP, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200ppo_rollout_synth_ws": (C.c_int, [_vp, _PP, _EP, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                           _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "b200ppo_rollout_synth_workspace_bytes": (_i64, [_PP, _i32]),
    "b200ppo_synth_env_step_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "b200ppo_synth_env_begin": (C.c_int, [_vp, _EP, _i32, _vp, _vp, _vp, _i64]),
    "b200ppo_synth_env_step": (C.c_int, [_vp, _EP, _vp, _i32, C.c_float, C.c_float, _vp, _vp, _i32, _i32, _i32,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64]),
    "b200ppo_rollout_synth_num_launches": (C.c_int, [_PP, _i32, _i32, _i32]),
    "b200ppo_eval_synth": (C.c_int, [_vp, _PP, _EP, _vp, _vp, _vp, _vp, _i32, _i32, _i32,
                                     _vp, _vp, _vp, _vp, _vp]),
    "b200ppo_update_workspace_bytes": (_i64, [_PP, _i32, _i32]),
    "b200ppo_update": (C.c_int, [_vp, _PP, _HP, _BP, _i32, _i32, _i32, _u32, _i32, _i32]),
    "b200ppo_debug_cta_times": (C.c_int, [_vp, _i32]),
    "b200ppo_debug_select": (C.c_int, [C.c_int]),
    "b200ppo_debug_timestamps": (C.c_int, [_vp, _i32]),
    "b200ppo_set_gemm_mode": (C.c_int, [C.c_int]),
    "b200ppo_set_update_paths": (C.c_int, [C.c_int, C.c_int]),
    "b200ppo_update_dw_splits": (C.c_int, [_PP, _i32, _i32, _vp, _vp, _i32]),
    "b200ppo_set_rollout_mode": (C.c_int, [C.c_int]),
    "b200ppo_update_num_launches": (C.c_int, [_PP, _HP, _i32, _i32, _i32]),
    "b200ppo_update_adv_sums_ptr": (_vp, [_PP, _i32, _i32, _vp]),
    "b200ppo_update_grad_ptr": (_vp, [_PP, _i32, _i32, _vp]),
    "b200ppo_update_debug_ptr": (_vp, [_PP, _i32, _i32, _vp, _i32]),
    "b200ppo_comm_bytes": (_i64, [_PP, _i32]),
    "b200ppo_comm_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p)]),
    "b200ppo_comm_free": (C.c_int, [_vp]),
    "b200ppo_comm_ipc_get": (C.c_int, [_vp, C.c_char_p]),
    "b200ppo_comm_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "b200ppo_comm_ipc_close": (C.c_int, [_vp]),
    "b200ppo_lstm_cache_floats": (_i64, [C.POINTER(LstmPlan), _i32]),
    "b200ppo_lstm_step_fwd": (C.c_int, [_vp, C.POINTER(LstmPlan), _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "b200ppo_lstm_step_bwd": (C.c_int, [_vp, C.POINTER(LstmPlan), _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp]),
    "b200ppo_lstm_wgrad_scratch_floats": (_i64, [C.POINTER(LstmPlan), _i32]),
    "b200ppo_lstm_weight_grads": (C.c_int, [_vp, C.POINTER(LstmPlan), _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "b200ppo_lstm_seq_supported": (C.c_int, [C.POINTER(LstmPlan)]),
    "b200ppo_lstm_set_persistent": (C.c_int, [C.c_int]),
    "b200ppo_lstm_seq_workspace_floats": (_i64, [C.POINTER(LstmPlan), _i32, _i32]),
    "b200ppo_lstm_seq_num_launches": (C.c_int, [C.POINTER(LstmPlan), _i32, _i32, _i32]),
    "b200ppo_lstm_seq_forward": (C.c_int, [_vp, C.POINTER(LstmPlan), _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp,
                                           _i32, _i32, _vp, _vp, _i32]),
    "b200ppo_lstm_seq_backward": (C.c_int, [_vp, C.POINTER(LstmPlan), _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "b200ppo_rg_gemm_test": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32]),
    "b200ppo_rg_tn_test": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "b200ppo_sampler_step": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _u32, _vp, _vp, _vp, _vp, _vp]),
    "b200ppo_iter_finalize": (C.c_int, [_vp, _vp, _u32, _u32, _vp]),
    "b200ppo_set_pdl": (C.c_int, [C.c_int]),
    "b200ppo_tc_gemm_test": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32]),
    "b200ppo_tc_mn_test": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200ppo_tc_microbench": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32]),
    "b200ppo_ffma_peak": (C.c_int, [_vp, _i32, _vp, _i32, _i32]),
}


class B200PPOError(RuntimeError):
    pass


_lib = None


def load():
    """Load libb200ppo.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200PPOError(
            f"{LIB_PATH} is missing: build the sm_100a kernels first "
            "(python -m nnx_ppo_b200.build, or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().b200ppo_error_string(int(code)).decode()
        raise B200PPOError(f"{what or 'b200ppo call'} failed with code {code}: {msg}")


def ptr(t) -> int:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B200PPOError("b200ppo kernels need CUDA tensors; there is no CPU path")
        if t is not None and not t.is_contiguous():
            raise B200PPOError("b200ppo kernels need contiguous tensors")
