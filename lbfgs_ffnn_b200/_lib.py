"""ctypes loader of libb200lbfgs.so (the C ABI declared in include/b200_lbfgs.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present when a
context is created, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200lbfgs.so")

OK = 0
ACT = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3}
PREC = {"fp32": 0, "tf32x3": 1, "tf32": 2}
LS = {"armijo": 0, "wolfe": 1}

LOSS_GRAD_FN = C.CFUNCTYPE(C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int)


class History(C.Structure):
    _fields_ = [("capacity", C.c_int), ("size", C.c_int), ("loss", C.POINTER(C.c_float)),
                ("grad_norm", C.POINTER(C.c_float)), ("time_ms", C.POINTER(C.c_float)), ("iterations", C.c_int),
                ("evaluations", C.c_long), ("launches", C.c_long)]


class LbfgsOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("tol", C.c_float), ("memory", C.c_int), ("max_line_iters", C.c_int),
                ("c1", C.c_float), ("rho", C.c_float), ("c2", C.c_float), ("linesearch", C.c_int),
                ("record_timing", C.c_int), ("shard_history", C.c_int)]


class GdOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("tol", C.c_float), ("lr", C.c_float), ("momentum", C.c_float),
                ("record_timing", C.c_int)]


class SgdOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("tol", C.c_float), ("lr", C.c_float), ("momentum", C.c_float),
                ("decay_rate", C.c_float), ("decay_step", C.c_int), ("batch_size", C.c_int), ("input_dim", C.c_int),
                ("output_dim", C.c_int), ("record_timing", C.c_int), ("sampling", C.c_int), ("seed", C.c_uint)]


class SlbfgsOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("tol", C.c_float), ("step_size", C.c_float), ("batch_size", C.c_int),
                ("memory", C.c_int), ("L", C.c_int), ("b_H", C.c_int), ("lam", C.c_float), ("epsilon", C.c_float),
                ("seed", C.c_uint), ("record", C.c_int), ("pair_eval", C.c_int), ("hvp_step_scale", C.c_float)]


# every symbol include/b200_lbfgs.h declares: name -> (restype, argtypes)
_vp, _i, _l, _sz, _f, _d = C.c_void_p, C.c_int, C.c_long, C.c_size_t, C.c_float, C.c_double
_pd, _pf = C.POINTER(C.c_double), C.POINTER(C.c_float)
SYMBOLS = {
    "b200_last_error": (C.c_char_p, []),
    "b200_abi_version": (_i, []),
    "b200_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "b200_ctx_destroy": (_i, [_vp]),
    "b200_ctx_synchronize": (_i, [_vp]),
    "b200_ctx_stream": (_vp, [_vp]),
    "b200_ctx_set_stream": (_i, [_vp, _vp]),
    "b200_ctx_device": (_i, [_vp]),
    "b200_ctx_profile": (_i, [_vp, _i]),
    "b200_ctx_profile_report": (_i, [_vp, C.c_char_p, _sz]),
    "b200_comm_unique_id": (_i, [_vp]),
    "b200_ctx_init_comm": (_i, [_vp, _vp, _i, _i]),
    "b200_ctx_rank": (_i, [_vp]),
    "b200_ctx_world": (_i, [_vp]),
    "b200_ctx_allreduce_f32": (_i, [_vp, _vp, _sz]),
    "b200_malloc": (_i, [C.POINTER(_vp), _sz]),
    "b200_free": (_i, [_vp]),
    "b200_memcpy_h2d": (_i, [_vp, _vp, _sz]),
    "b200_memcpy_d2h": (_i, [_vp, _vp, _sz]),
    "b200_memcpy_d2d": (_i, [_vp, _vp, _sz]),
    "b200_memset": (_i, [_vp, _i, _sz]),
    "b200_host_alloc_pinned": (_i, [C.POINTER(_vp), _sz]),
    "b200_host_free_pinned": (_i, [_vp]),
    "b200_net_create": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_vp)]),
    "b200_net_destroy": (_i, [_vp]),
    "b200_net_params_size": (_sz, [_vp]),
    "b200_net_output_size": (_i, [_vp]),
    "b200_net_bind_params": (_i, [_vp, C.c_uint]),
    "b200_net_params_data": (_vp, [_vp]),
    "b200_net_grads_data": (_vp, [_vp]),
    "b200_net_zero_grads": (_i, [_vp]),
    "b200_net_set_precision": (_i, [_vp, _i]),
    "b200_net_get_precision": (_i, [_vp]),
    "b200_net_set_l2": (_i, [_vp, _f]),
    "b200_net_set_global_batch": (_i, [_vp, _l]),
    "b200_net_quantize_input": (_i, [_vp, _vp, _l, C.POINTER(_i)]),
    "b200_net_clear_input_cache": (_i, [_vp]),
    "b200_net_forward": (_i, [_vp, _vp, _l]),
    "b200_net_loss_grad": (_i, [_vp, _vp, _vp, _l, _pf]),
    "b200_net_loss_grad_async": (_i, [_vp, _vp, _vp, _vp, _l, _vp, _vp]),
    "b200_net_copy_output_to_host": (_i, [_vp, _vp, _sz]),
    "b200_net_last_batch": (_i, [_vp]),
    "b200_net_copy_activation_to_host": (_i, [_vp, _i, _vp, _sz]),
    "b200_net_evaluate": (_i, [_vp, _vp, _vp, _l, _pd, _pd]),
    "b200_lbfgs_default_opts": (None, [C.POINTER(LbfgsOpts)]),
    "b200_lbfgs_solve": (_i, [_vp, _vp, LOSS_GRAD_FN, _vp, _i, _vp, _vp, _vp, _i, C.POINTER(LbfgsOpts),
                              C.POINTER(History)]),
    "b200_lbfgs_create": (_i, [_vp, _i, C.POINTER(LbfgsOpts), C.POINTER(_vp)]),
    "b200_lbfgs_run": (_i, [_vp, _vp, LOSS_GRAD_FN, _vp, _vp, _vp, _vp, _i, _i, C.POINTER(History)]),
    "b200_lbfgs_destroy": (_i, [_vp]),
    "b200_gd_default_opts": (None, [C.POINTER(GdOpts)]),
    "b200_gd_solve": (_i, [_vp, _vp, LOSS_GRAD_FN, _vp, _i, _vp, _vp, _vp, _i, C.POINTER(GdOpts), C.POINTER(History)]),
    "b200_sgd_default_opts": (None, [C.POINTER(SgdOpts)]),
    "b200_sgd_solve": (_i, [_vp, _vp, LOSS_GRAD_FN, _vp, _i, _vp, _vp, _vp, _i, C.POINTER(SgdOpts),
                            C.POINTER(History)]),
    "b200_slbfgs_default_opts": (None, [C.POINTER(SlbfgsOpts)]),
    "b200_slbfgs_solve": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, C.POINTER(SlbfgsOpts), C.POINTER(History)]),
    "b200_slbfgs_sample_stream": (_i, [C.c_uint, _l, _l, _i, _vp]),
    "b200_lbfgs_direction": (_i, [_vp, _sz, _i, _vp, _vp, _pf, _vp, _i, _vp, _pd]),
    "b200_vec_dot": (_i, [_vp, _vp, _vp, _sz, _pd]),
    "b200_vec_nrm2": (_i, [_vp, _vp, _sz, _pd]),
    "b200_vec_axpy": (_i, [_vp, _sz, _f, _vp, _vp]),
    "b200_vec_scal": (_i, [_vp, _sz, _f, _vp]),
    "b200_vec_trial_point": (_i, [_vp, _sz, _vp, _f, _vp, _vp]),
    "b200_convert_f64_to_f32": (_i, [_vp, _vp, _vp, _sz]),
    "b200_launch_count": (_l, []),
    "b200_debug_reload_env": (_i, []),
}

_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    """Load the shared library (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(or make -C lbfgs_ffnn_b200/csrc). This package has no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != OK:
        raise B200Error(f"libb200lbfgs status {status}: {lib().b200_last_error().decode()}")
