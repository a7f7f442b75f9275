"""ctypes binding of libdv3.so (include/dv3.h). There is no CPU fallback: importing this
module without the built library, or creating an engine without a CUDA device, raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdv3.so")
ABI_VERSION = 1


class Dv3Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "G", "D", "L", "cnn_mult", "Zc", "Zk", "A", "discrete", "image_obs", "obs_dim",
        "img_hw", "img_c", "B", "T", "H", "symlog_obs", "compute_mode", "world_size", "rank", "device",
        "num_curiosity_nets")]
    _fields_ += [("intrinsic_rewards_scale", C.c_float), ("reserved", C.c_int32 * 6)]


class Dv3TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ptr", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 6), ("kind", C.c_int32), ("group", C.c_int32)]


class Dv3Hparams(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "gamma", "lambda_", "entropy_scale", "return_normalization_decay", "wm_grad_clip", "actor_grad_clip",
        "critic_grad_clip", "wm_lr", "wm_eps", "ac_lr", "ac_eps", "critic_ema_decay")]
    _fields_ += [("train_actor", C.c_int32), ("critic_lr", C.c_float), ("critic_eps", C.c_float)]
    _fields_ += [(n, C.c_float) for n in ("disagree_grad_clip", "disagree_lr", "disagree_eps")]
    _fields_ += [("reserved", C.c_int32 * 2)]


# name -> (restype, argtypes); mirrors include/dv3.h one to one (tests/test_capi_symbols.py checks it)
_P, _I, _F, _U64, _I64, _SZ = C.c_void_p, C.c_int, C.c_float, C.c_uint64, C.c_int64, C.c_size_t
SIGNATURES = {
    "dv3_abi_version": (_I, []),
    "dv3_create": (_I, [C.POINTER(Dv3Config), C.POINTER(_P)]),
    "dv3_destroy": (_I, [_P]),
    "dv3_last_error": (C.c_char_p, [_P]),
    "dv3_default_hparams": (None, [C.POINTER(Dv3Hparams)]),
    "dv3_sync": (_I, [_P, _P]),
    "dv3_num_tensors": (_I, [_P]),
    "dv3_tensor_info": (_I, [_P, _I, C.POINTER(Dv3TensorDesc)]),
    "dv3_find_tensor": (_I, [_P, C.c_char_p, C.POINTER(Dv3TensorDesc)]),
    "dv3_write_tensor": (_I, [_P, C.c_char_p, _P, _SZ, _I, _P, _I]),
    "dv3_read_tensor": (_I, [_P, C.c_char_p, _P, _SZ, _I, _P, _I]),
    "dv3_group_buffer": (_I, [_P, _I, _I, C.POINTER(_P), C.POINTER(_I64)]),
    "dv3_fill_noise": (_I, [_P, _U64, _U64, _I, _P]),
    "dv3_initial_state": (_I, [_P, _P]),
    "dv3_observe_forward": (_I, [_P, _P]),
    "dv3_wm_loss_backward": (_I, [_P, _P]),
    "dv3_optimizer_step": (_I, [_P, _I, _F, _F, _F, _F, _P]),
    "dv3_train_world_model_step": (_I, [_P, C.POINTER(Dv3Hparams), _U64, _I, _P]),
    "dv3_forward_inference": (_I, [_P, _I, _U64, _P]),
    "dv3_dream_forward": (_I, [_P, _F, _I, _P]),
    "dv3_dream_forward_actions": (_I, [_P, _F, _I, _I, _P]),
    "dv3_replay_plan": (_I64, [_I, _I, _I, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P]),
    "dv3_replay_gather": (_I, [_P, _I, _I, _I64, _P, _I, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "dv3_ac_loss_backward": (_I, [_P, C.POINTER(Dv3Hparams), _I, _P]),
    "dv3_disagree_loss_backward": (_I, [_P, _P]),
    "dv3_critic_init_ema": (_I, [_P, _P]),
    "dv3_critic_update_ema": (_I, [_P, _F, _P]),
    "dv3_train_actor_critic_step": (_I, [_P, C.POINTER(Dv3Hparams), _U64, _I, _P]),
    "dv3_train_update": (_I, [_P, C.POINTER(Dv3Hparams), _U64, _I, _P]),
    "dv3_dp_piece": (_I, [_P, C.POINTER(Dv3Hparams), _I, _U64, _I, _P]),
    "dv3_get_inference_calls": (_I64, [_P]),
    "dv3_set_inference_calls": (_I, [_P, _I64]),
    "dv3_kernel_launch_count": (_I64, [_P]),
    "dv3_phase_report": (_I, [_P, _I, C.c_char_p, _SZ]),
    "dv3_scan_tc_schedule_check": (_I, [_I, _I, _I, _I, _I, _I]),
    "dv3_op_symlog": (_I, [_P, _P, _I64, _P]),
    "dv3_op_inverse_symlog": (_I, [_P, _P, _I64, _P]),
    "dv3_op_two_hot": (_I, [_P, _I64, _P, _P, _P, _P, _P, _P]),
    "dv3_op_sample_categorical": (_I, [_P, _P, _I64, _I, _P, _P]),
    "dv3_op_value_targets": (_I, [_P, _P, _P, _I, _I64, _F, _F, _P, _P]),
    "dv3_op_percentiles": (_I, [_P, _I64, _P, _P]),
    "dv3_op_gemm": (_I, [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, _P, _F, _I, _P]),
    "dv3_op_gemm_shadow": (_I, [_P, _I, _I, _P, _I, _P, _I, _I, _I, _I, _P, _F, _P]),
    "dv3_op_ln_silu_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I, _P]),
    "dv3_op_to_bf16": (_I, [_P, _I, _P, _I, _I64, _I, _P]),
    "dv3_op_gemm_bf16": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _F, _P]),
    "dv3_op_gemm_bf16_small": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _F, _P]),
    "dv3_op_gemm_bf16_wgrad": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _F, _P]),
    "dv3_op_ln_silu": (_I, [_P, _P, _P, _P, _I64, _I, _P]),
}

_lib = None


def load():
    """Loads libdv3.so once. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dreamer_v3_b200/csrc`. dreamer_v3_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.dv3_abi_version() != ABI_VERSION:
        raise ImportError("libdv3.so ABI version mismatch; rebuild")
    _lib = lib
    return lib
