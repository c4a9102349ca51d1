"""ctypes binding of libptk_b200.so (the C ABI declared in include/ptk_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, the product
path raises.  (The CPU oracle lives under oracle/ and is only ever used by tests / bench baselines.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libptk_b200.so")

PTK_CELL_INT, PTK_CELL_HGRU = 0, 1
PTK_ACT_SOFTPLUS, PTK_ACT_TANH = 0, 1
PTK_CONV_FP32, PTK_CONV_BF16X3, PTK_CONV_BF16 = 0, 1, 2
PTK_INT_NPARAM = 30

EXPORTS = ("ptk_version", "ptk_last_error", "ptk_int_param_offsets", "ptk_int_saved_bytes",
           "ptk_int_workspace_bytes", "ptk_int_backward_workspace_bytes", "ptk_wgrad7", "ptk_wgrad7_workspace_bytes", "ptk_int_forward", "ptk_int_backward", "ptk_synth_pathtracker",
           "ptk_profile_start", "ptk_profile_stop", "ptk_launch_count", "ptk_conv7", "ptk_conv7_workspace_bytes",
           "ptk_convlstm_param_offsets", "ptk_convlstm_saved_bytes", "ptk_convlstm_workspace_bytes",
           "ptk_convlstm_forward", "ptk_convlstm_backward", "ptk_convlstm_input_grad",
           "ptk_bce_scores", "ptk_adam_step", "ptk_convlstm_backward_rbp", "ptk_convlstm_backward_penalty",
           "ptk_lstm_wide_op", "ptk_lstm_wide_op_workspace_bytes",
           "ptk_gru_param_offsets", "ptk_gru_saved_bytes", "ptk_gru_workspace_bytes", "ptk_gru_forward", "ptk_gru_backward")


class PtkDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("B", "T", "H", "W", "C", "K", "cell", "act", "no_inh", "conv_mode")]


class PtkLstmDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "T", "H", "W", "K", "conv_mode", "ckpt_ring", "rbp")]


class PtkGruDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "T", "H", "W", "C", "K", "conv_mode")]


class PtkError(RuntimeError):
    pass


_lib = None


def lib():
    """Load (once) and return the ctypes handle; raise loudly if the native library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PtkError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m pathtracker_models_b200._build`). "
            "There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64p, u64, i32, sz = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_uint64, ctypes.c_int32, ctypes.c_size_t
    dp = ctypes.POINTER(PtkDims)
    L.ptk_version.restype = ctypes.c_int
    L.ptk_version.argtypes = []
    L.ptk_last_error.restype = ctypes.c_char_p
    L.ptk_last_error.argtypes = []
    L.ptk_int_param_offsets.restype = ctypes.c_int
    L.ptk_int_param_offsets.argtypes = [dp, i64p]
    L.ptk_int_saved_bytes.restype = sz
    L.ptk_int_saved_bytes.argtypes = [dp]
    L.ptk_int_workspace_bytes.restype = sz
    L.ptk_int_workspace_bytes.argtypes = [dp]
    L.ptk_int_backward_workspace_bytes.restype = sz
    L.ptk_int_backward_workspace_bytes.argtypes = [dp]
    L.ptk_wgrad7_workspace_bytes.restype = sz
    L.ptk_wgrad7_workspace_bytes.argtypes = [i32, i32, i32, i32]
    L.ptk_wgrad7.restype = ctypes.c_int
    L.ptk_wgrad7.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp]
    L.ptk_int_forward.restype = ctypes.c_int
    L.ptk_int_forward.argtypes = [dp, vp, i64p, vp, vp, vp, vp, vp, vp, sz, vp]
    L.ptk_int_backward.restype = ctypes.c_int
    L.ptk_int_backward.argtypes = [dp, vp, i64p, vp, vp, vp, vp, vp, vp, sz, vp]
    L.ptk_synth_pathtracker.restype = ctypes.c_int
    L.ptk_synth_pathtracker.argtypes = [u64, u64, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    L.ptk_conv7_workspace_bytes.restype = sz
    L.ptk_conv7_workspace_bytes.argtypes = [i32, i32, i32]
    L.ptk_conv7.restype = ctypes.c_int
    L.ptk_conv7.argtypes = [i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    ldp = ctypes.POINTER(PtkLstmDims)
    L.ptk_convlstm_param_offsets.restype = ctypes.c_int
    L.ptk_convlstm_param_offsets.argtypes = [ldp, i64p]
    L.ptk_convlstm_saved_bytes.restype = sz
    L.ptk_convlstm_saved_bytes.argtypes = [ldp]
    L.ptk_convlstm_workspace_bytes.restype = sz
    L.ptk_convlstm_workspace_bytes.argtypes = [ldp]
    L.ptk_convlstm_forward.restype = ctypes.c_int
    L.ptk_convlstm_forward.argtypes = [ldp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.ptk_convlstm_backward.restype = ctypes.c_int
    L.ptk_convlstm_backward.argtypes = [ldp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.ptk_convlstm_backward_penalty.restype = ctypes.c_int
    L.ptk_convlstm_backward_penalty.argtypes = [ldp, vp, vp, vp, vp, vp, ctypes.c_int32, vp, vp, vp, sz, vp]
    L.ptk_convlstm_input_grad.restype = ctypes.c_int
    L.ptk_convlstm_input_grad.argtypes = [ldp, vp, vp, vp, sz, vp]
    L.ptk_bce_scores.restype = ctypes.c_int
    L.ptk_bce_scores.argtypes = [i32, vp, vp, vp, vp, vp]
    L.ptk_adam_step.restype = ctypes.c_int
    L.ptk_adam_step.argtypes = [ctypes.c_int64, vp, vp, vp, vp, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                ctypes.c_float, ctypes.c_int64, i32, vp]
    L.ptk_convlstm_backward_rbp.restype = ctypes.c_int
    L.ptk_convlstm_backward_rbp.argtypes = [ldp, vp, vp, vp, vp, i32, vp, vp, vp, sz, vp]
    L.ptk_lstm_wide_op_workspace_bytes.restype = sz
    L.ptk_lstm_wide_op_workspace_bytes.argtypes = [i32]
    L.ptk_lstm_wide_op.restype = ctypes.c_int
    L.ptk_lstm_wide_op.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    gdp = ctypes.POINTER(PtkGruDims)
    L.ptk_gru_param_offsets.restype = ctypes.c_int
    L.ptk_gru_param_offsets.argtypes = [gdp, i64p]
    L.ptk_gru_saved_bytes.restype = sz
    L.ptk_gru_saved_bytes.argtypes = [gdp]
    L.ptk_gru_workspace_bytes.restype = sz
    L.ptk_gru_workspace_bytes.argtypes = [gdp]
    L.ptk_gru_forward.restype = ctypes.c_int
    L.ptk_gru_forward.argtypes = [gdp, vp, i64p, vp, vp, vp, vp, sz, vp]
    L.ptk_gru_backward.restype = ctypes.c_int
    L.ptk_gru_backward.argtypes = [gdp, vp, i64p, vp, vp, vp, vp, vp, vp, sz, vp]
    L.ptk_profile_start.restype = None
    L.ptk_profile_start.argtypes = []
    L.ptk_profile_stop.restype = ctypes.c_int
    L.ptk_profile_stop.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]
    L.ptk_launch_count.restype = ctypes.c_longlong
    L.ptk_launch_count.argtypes = []
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().ptk_last_error().decode("utf-8", "replace")
        raise PtkError(f"{what} failed (code {rc}): {msg}")


def param_offsets(dims):
    arr = (ctypes.c_int64 * (PTK_INT_NPARAM + 1))()
    check(lib().ptk_int_param_offsets(ctypes.byref(dims), arr), "ptk_int_param_offsets")
    return list(arr)


PROF_KINDS = ("conv_fwd", "conv_dgrad", "wgrad", "pointwise_fwd", "pointwise_bwd", "bn_apply", "misc", "gate_wgrad")


def profile_start():
    lib().ptk_profile_start()


def profile_stop():
    """-> {kind: (total_ms, launches)} measured with CUDA events on the launching stream."""
    ms = (ctypes.c_double * len(PROF_KINDS))()
    cnt = (ctypes.c_longlong * len(PROF_KINDS))()
    lib().ptk_profile_stop(ms, cnt)
    return {k: (ms[i], cnt[i]) for i, k in enumerate(PROF_KINDS)}


def launch_count():
    return int(lib().ptk_launch_count())
