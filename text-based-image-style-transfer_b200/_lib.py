"""ctypes binding of libnst_b200.so (include/nst_b200.h).  No CPU path exists: loading fails loudly when
the library has not been built, and every compute entry fails with NST_ERR_DEVICE off sm_100."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnst_b200.so")

NST_MAX_CONV = 16
NST_LOSS_COUNT = 16
STOP_REASONS = {0: "", 1: "opt_entry", 2: "gtd", 3: "opt", 4: "step", 5: "loss", 6: "nonfinite"}


class NstError(RuntimeError):
    pass


class NstStatus(C.Structure):
    _fields_ = [("n_iter", C.c_int), ("func_evals", C.c_int), ("closure_calls", C.c_int), ("stop", C.c_int),
                ("hist_len", C.c_int), ("reserved", C.c_int),
                ("loss", C.c_double), ("prev_loss", C.c_double), ("t", C.c_double), ("H_diag", C.c_double),
                ("gtd", C.c_double), ("gmax", C.c_double), ("max_td", C.c_double),
                ("losses", C.c_float * NST_LOSS_COUNT)]


class NstLaunchTime(C.Structure):
    _fields_ = [("kind", C.c_int), ("layer", C.c_int), ("ms", C.c_float)]


KIND_NAMES = {1: "pixel", 2: "conv1_fwd", 3: "conv_fwd", 4: "gram", 5: "content", 6: "assemble", 7: "gram_bwd",
              8: "conv_dgrad", 9: "conv1_dgrad", 10: "lbfgs_pass1", 11: "lbfgs_reduce", 12: "lbfgs_control",
              13: "lbfgs_pass2"}

_P = C.c_void_p
_F3 = C.POINTER(C.c_float)
# name -> (restype, argtypes); mirrors include/nst_b200.h one to one
PROTOTYPES = {
    "nst_abi_version": (C.c_int, []),
    "nst_last_error": (C.c_char_p, []),
    "nst_device_check": (C.c_int, []),
    "nst_net_create": (C.c_int, [C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.c_int, _P]),
    "nst_net_destroy": (None, [_P]),
    "nst_plan_create": (C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "nst_plan_destroy": (None, [_P]),
    "nst_plan_bytes": (C.c_size_t, [_P]),
    "nst_plan_set_norm": (C.c_int, [_P, _F3, _F3]),
    "nst_plan_set_weights": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float]),
    "nst_plan_features": (C.c_int, [_P, _P, _P]),
    "nst_plan_tap_shape": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nst_plan_get_tap": (C.c_int, [_P, C.c_int, _P, _P]),
    "nst_plan_tap_gram": (C.c_int, [_P, C.c_int, _P, _P]),
    "nst_gram_chw": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "nst_style_mix_gram": (C.c_int, [_P, _P, C.c_int, C.c_float, _P, _P]),
    "nst_style_mix_chw": (C.c_int, [_P, _P, C.c_int, C.c_float, _P, _P]),
    "nst_plan_set_style_target": (C.c_int, [_P, C.c_int, _P, _P]),
    "nst_plan_set_content_target": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "nst_plan_channel_gate": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P, _P]),
    "nst_plan_set_edge_target": (C.c_int, [_P, _P, _P]),
    "nst_plan_eval": (C.c_int, [_P, _P, _P, _P, _P]),
    "nst_tv_edge": (C.c_int, [_P, _P, C.c_int, C.c_int, _F3, _F3, _P, _P]),
    "nst_edge_images": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "nst_normalize": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _F3, _F3, _P]),
    "nst_grayscale": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P]),
    "nst_mse": (C.c_int, [_P, _P, C.c_size_t, _P, _P]),
    "nst_total_variation": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "nst_channel_attention_chw": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P]),
    "nst_style_mix_tensors": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "nst_lbfgs_init": (C.c_int, [_P, _P, C.c_int, _P]),
    "nst_lbfgs_step": (C.c_int, [_P, _P]),
    "nst_lbfgs_prepare_graph": (C.c_int, [_P, _P]),
    "nst_lbfgs_status": (C.c_int, [_P, C.POINTER(NstStatus), _P]),
    "nst_lbfgs_get_x": (C.c_int, [_P, _P, _P]),
    "nst_lbfgs_trace": (C.c_int, [_P, _P, C.c_int, _P]),
    "nst_lbfgs_launches_per_step": (C.c_int, [_P]),
    "nst_lbfgs_partial_step": (C.c_int, [_P, C.c_int, _P]),
    "nst_plan_eval_timed": (C.c_int, [_P, _P, _P, C.POINTER(NstLaunchTime), C.c_int, _P]),
    "nst_lbfgs_step_timed": (C.c_int, [_P, C.POINTER(NstLaunchTime), C.c_int, _P]),
    "nst_lbfgs_step_timed_grouped": (C.c_int, [_P, C.POINTER(NstLaunchTime), C.c_int, _P]),
    "nst_run_frames_host": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_int, _P, _P, C.POINTER(_P),
                                      C.POINTER(C.c_int)]),
    "nst_plan_set_shared_gpu": (C.c_int, [_P, C.c_int]),
    "nst_batch_create": (C.c_int, [C.POINTER(_P), _P, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "nst_batch_size": (C.c_int, [_P]),
    "nst_batch_member": (_P, [_P, C.c_int]),
    "nst_lbfgs_freeze": (C.c_int, [_P, C.c_int, _P]),
    "nst_run_batch_host": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_int, _P, _P, _P, C.POINTER(C.c_int)]),
    "nst_mask_composite": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "nst_mask_gaussian_weights": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "nst_mip_split": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                C.POINTER(C.c_double), _P, _P]),
    "nst_mip_merge": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                C.POINTER(C.c_double), _P, _P]),
    "nst_video_assemble": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "nst_run_frame_host": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
}

# tuning aids exported by the instrumented build only (tools/build.py --instrument; -DNST_INSTRUMENT)
INSTR_LIB_PATH = os.path.join(_HERE, "libnst_b200_instr.so")
INSTR_PROTOTYPES = {
    "nst_plan_conv_phases": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_longlong), _P]),
    "nst_plan_timeline": (C.c_int, [_P, C.c_int, C.POINTER(C.c_ulonglong), _P]),
    "nst_lbfgs_ctl_clocks": (C.c_int, [_P, C.POINTER(C.c_longlong), _P]),
}

_lib = None


def load_instrumented():
    """tools/ only: binds the instrumented build (stamps, wait counters, timing experiments) in place of the product
    library for this process.  Must be called before anything else loads the library."""
    global _lib
    if _lib is not None:
        raise NstError("load_instrumented() must run before the product library is loaded")
    if not os.path.exists(INSTR_LIB_PATH):
        raise NstError("%s is missing: build it with `python tools/build.py --instrument`" % INSTR_LIB_PATH)
    lib = C.CDLL(INSTR_LIB_PATH)
    for name, (res, args) in list(PROTOTYPES.items()) + list(INSTR_PROTOTYPES.items()):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def load():
    """Loads the shared library and binds every symbol of the header.  Needs no GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NstError(
            "%s is missing: build it with `python tools/build.py` (or __graft_entry__.build()). "
            "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Raises NstError with the library's message for a negative return code; returns rc otherwise."""
    if rc < 0:
        msg = load().nst_last_error()
        raise NstError("nst_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
    return rc


def f3(values):
    return (C.c_float * 3)(*[float(v) for v in values])
