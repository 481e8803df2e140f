"""ctypes binding of libtd3b200.so (the C ABI in include/td3_b200.h).

There is no CPU fallback: if the library is missing or does not load, every
product entry point raises.  ``load()`` itself works without a GPU (the library
links cudart statically), which is what the CPU-only tests use to check that
every declared symbol is exported.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("TD3_LIB_NAME", "libtd3b200.so"))   # TD3_LIB_NAME: debug builds

TD3_MAX_LINEAR = 8
TD3_MAX_SEGMENTS = 16
RNG_PHILOX, RNG_INJECTED = 0, 1
VARIANT_FEATURED, VARIANT_PARTICLES = 0, 1
NORM_NONE, NORM_LAYER, NORM_WEIGHT = 0, 1, 2
PRECISION_FP32, PRECISION_TF32 = 0, 1
PRECISIONS = {"fp32": PRECISION_FP32, "tf32": PRECISION_TF32}


class NetLayout(C.Structure):
    _fields_ = [
        ("n_linear", C.c_int32),
        ("dims", C.c_int32 * (TD3_MAX_LINEAR + 1)),
        ("w_off", C.c_int64 * TD3_MAX_LINEAR),
        ("wg_off", C.c_int64 * TD3_MAX_LINEAR),
        ("b_off", C.c_int64 * TD3_MAX_LINEAR),
        ("ln_g_off", C.c_int64 * TD3_MAX_LINEAR),
        ("ln_b_off", C.c_int64 * TD3_MAX_LINEAR),
        ("enc_hidden", C.c_int32),
        ("enc_out", C.c_int32),
        ("c1w_off", C.c_int64), ("c1b_off", C.c_int64),
        ("c2w_off", C.c_int64), ("c2b_off", C.c_int64),
        ("ln_in_g_off", C.c_int64), ("ln_in_b_off", C.c_int64),
        ("n_floats", C.c_int64),
    ]


class ParamSet(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("params", "target", "grad", "exp_avg", "exp_avg_sq")]


class AgentConfig(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("norm", C.c_int32), ("n_q", C.c_int32), ("state_dim", C.c_int32),
        ("action_dim", C.c_int32), ("n_particles", C.c_int32), ("particle_dim", C.c_int32),
        ("clamp_target_action", C.c_int32), ("n_agents", C.c_int32), ("precision", C.c_int32),
        ("max_action", C.c_float), ("discount", C.c_float), ("policy_noise", C.c_float), ("noise_clip", C.c_float),
        ("tau", C.c_double), ("lr_actor", C.c_double), ("lr_critic", C.c_double), ("beta1", C.c_double),
        ("beta2", C.c_double), ("adam_eps", C.c_double),
        ("policy_freq", C.c_int32), ("reserved1", C.c_int32),
        ("seed", C.c_uint64),
        ("actor", NetLayout), ("q", NetLayout),
    ]


class ReplayView(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("row_stride", C.c_int64), ("row_floats", C.c_int64), ("max_size", C.c_int64),
                ("size", C.c_int64), ("agent_stride", C.c_int64)]


_vp, _i32, _i64, _u64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double
_P = C.POINTER

# name -> (restype, argtypes); mirrors include/td3_b200.h one to one
SIGNATURES = {
    "td3_abi_version": (C.c_int, []),
    "td3_struct_sizes": (None, [_P(_i64)]),
    "td3_last_error": (C.c_char_p, []),
    "td3_device_info": (C.c_int, [_P(C.c_int), _P(C.c_int), _P(C.c_int), C.c_char_p, C.c_int]),
    "rb_add_rows": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp]),
    "rb_sample_indices": (C.c_int, [_P(ReplayView), _vp, _i64, _i32, _P(_i64), _P(_i64), _P(_vp), _P(_i64), _vp]),
    "rb_philox_indices": (C.c_int, [_vp, _i64, _i64, _u64, _u64, _u64, _vp]),
    "adam_polyak_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _f64, _f64, _f64, _f64, _f64, _vp]),
    "set_encoder_workspace_floats": (_i64, [_i64, _i64, _i64, _i64, _i64]),
    "set_encoder_fwd": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp]),
    "set_encoder_bwd": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64,
                                  _i32, _vp]),
    "set_encoder_fwd_bits": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp]),
    "set_encoder_bwd_fused": (C.c_int, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp,
                                        _i64, _vp]),
    "td3_gemm": (C.c_int, [_i64, _i64, _i64, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _i64, _vp, _i32, _i32, _vp]),
    "td3_agent_create": (C.c_int, [_P(AgentConfig), _P(_vp)]),
    "td3_agent_destroy": (C.c_int, [_vp]),
    "td3_agent_bind_params": (C.c_int, [_vp, _P(ParamSet), _P(ParamSet)]),
    "td3_agent_bind_state": (C.c_int, [_vp, _vp, _i64]),
    "td3_agent_bind_host_status": (C.c_int, [_vp, _vp]),
    "td3_agent_host_status_live": (C.c_int, [_vp]),
    "td3_agent_chain_active": (C.c_int, [_vp]),
    "td3_agent_params_changed": (C.c_int, [_vp]),
    "td3_agent_prepare": (C.c_int, [_vp, _P(ReplayView), _vp]),
    "dp_allreduce_grads": (C.c_int, [_vp, _P(_vp), _i32, _i64, _vp]),
    "td3_dp_bind_peers": (C.c_int, [_vp, _i32, _i32, _P(_vp), _P(_vp), _P(_vp)]),
    "td3_dp_set_fused_reduce": (C.c_int, [_vp, _i32]),
    "td3_agent_workspace_floats": (_i64, [_vp, _i64]),
    "td3_agent_plan": (C.c_int, [_vp, _i64, _vp, _i64, _vp]),
    "td3_agent_region": (C.c_int, [_vp, C.c_char_p, _P(_i64), _P(_i64)]),
    "td3_agent_set_global_batch": (C.c_int, [_vp, _i64, _i64]),
    "td3_train_n": (C.c_int, [_vp, _P(ReplayView), _i64, _i32, _i32, _i32, _vp]),
    "td3_debug_prefix_times": (C.c_int, [_vp, _P(ReplayView), _i32, _i32, _P(C.c_float), _P(_i32), _i32, _P(_i32)]),
    "td3_sample_batch": (C.c_int, [_vp, _P(ReplayView), _i32, _vp]),
    "td3_target_step": (C.c_int, [_vp, _vp]),
    "td3_critic_step": (C.c_int, [_vp, _i32, _vp]),
    "td3_critic_apply": (C.c_int, [_vp, _vp]),
    "td3_actor_step": (C.c_int, [_vp, _i32, _vp]),
    "td3_actor_apply": (C.c_int, [_vp, _vp]),
    "td3_actor_forward": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i64, _vp, _vp]),
    "td3_critic_forward": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp]),
    "td3_infer_b1": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, C.c_uint32, _i64, _vp]),
    "td3_infer_wait": (C.c_int, [_vp, _i32, C.c_uint32, _i64]),
    "td3_launch_count": (_i64, []),
}

_lib = None


def load():
    """dlopen libtd3b200.so and attach prototypes.  Raises RuntimeError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C td3_b200/csrc`. td3_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header and library out of sync
        fn.restype, fn.argtypes = res, args
    sizes = (_i64 * 4)()
    lib.td3_struct_sizes(sizes)
    mine = [C.sizeof(NetLayout), C.sizeof(ParamSet), C.sizeof(AgentConfig), C.sizeof(ReplayView)]
    if list(sizes) != mine:
        raise RuntimeError(f"ctypes struct mirrors out of sync with libtd3b200.so: C {list(sizes)} vs Python {mine}")
    if lib.td3_abi_version() != 1:
        raise RuntimeError("libtd3b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    """Map a td3_status to the exception the reference would have raised (SURVEY.md 8b 'Errors')."""
    if rc == 0:
        return
    msg = load().td3_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(f"libtd3b200 error {rc}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("td3_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    lib = load()
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    name = C.create_string_buffer(128)
    check(lib.td3_device_info(C.byref(sm), C.byref(maj), C.byref(mnr), name, 128))
    return lib


def stream_ptr():
    """The current CUDA stream of the current device as a void* (raw handle: no torch Stream object is built)."""
    import torch
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:                                   # older / newer torch without the private accessor
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
