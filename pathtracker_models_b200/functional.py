"""torch.autograd bridge to the C ABI: raw device pointers + the current CUDA stream go to
ptk_int_forward / ptk_int_backward (include/ptk_b200.h); torch only owns memory and streams."""
import ctypes

import torch
from torch.autograd.function import once_differentiable

from . import _lib


class IntConfig:
    """static configuration of one recurrent circuit instance (maps onto PtkDims)."""

    def __init__(self, C=32, K=7, cell=_lib.PTK_CELL_INT, act=_lib.PTK_ACT_SOFTPLUS, no_inh=False,
                 conv_mode=_lib.PTK_CONV_FP32):
        self.C, self.K, self.cell, self.act, self.no_inh, self.conv_mode = C, K, cell, act, bool(no_inh), conv_mode

    def dims(self, x):
        B, ch, T, H, W = x.shape
        if ch != 3:
            raise ValueError(f"expected x of shape [B,3,T,H,W], got {tuple(x.shape)}")
        return _lib.PtkDims(B, T, H, W, self.C, self.K, self.cell, self.act, int(self.no_inh), self.conv_mode)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _strides(x):
    return (ctypes.c_int64 * 5)(*x.stride())


def _check_inputs(x, flat):
    if not x.is_cuda:
        raise _lib.PtkError("pathtracker_models_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    if x.dtype != torch.float32 or flat.dtype != torch.float32:
        raise _lib.PtkError("the recurrent core computes in float32; got x=%s params=%s" % (x.dtype, flat.dtype))
    if flat.device != x.device:
        raise _lib.PtkError("parameters and input are on different devices")


def _buffers(L, dims, device, want_saved, backward=False):
    ws_bytes = (L.ptk_int_backward_workspace_bytes if backward else L.ptk_int_workspace_bytes)(ctypes.byref(dims))
    if ws_bytes == 0:
        _lib.check(1, "ptk_int_workspace_bytes")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    saved = None
    if want_saved:
        saved = torch.empty(L.ptk_int_saved_bytes(ctypes.byref(dims)), dtype=torch.uint8, device=device)
    return ws, ws_bytes, saved


def int_forward_raw(x, flat, cfg, want_saved, testmode=False):
    """one call of ptk_int_forward on x.device's current stream. returns logits[B], saved, gates, states"""
    _check_inputs(x, flat)
    L = _lib.lib()
    dims = cfg.dims(x)
    B, _, T, H, W = x.shape
    flat = flat.contiguous()
    with torch.cuda.device(x.device):
        ws, ws_bytes, saved = _buffers(L, dims, x.device, want_saved)
        logits = torch.empty(B, dtype=torch.float32, device=x.device)
        gates = states = None
        if testmode:
            gates = torch.empty(B, T, cfg.C, H, W, dtype=torch.float32, device=x.device)
            states = torch.empty(B, T, 1, H, W, dtype=torch.float32, device=x.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        rc = L.ptk_int_forward(ctypes.byref(dims), _ptr(x), _strides(x), _ptr(flat), _ptr(logits), _ptr(saved),
                               _ptr(gates), _ptr(states), _ptr(ws), ws_bytes, stream)
        _lib.check(rc, "ptk_int_forward")
    return logits, saved, gates, states


class _IntFunction(torch.autograd.Function):
    """logits = InT/FFhGRU time loop (x, flat parameters); backward = hand-written BPTT kernels."""

    @staticmethod
    def forward(ctx, x, flat, cfg, testmode=False):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        logits, saved, gates, states = int_forward_raw(x, flat, cfg, want_saved=need, testmode=testmode)
        if testmode and saved is not None:
            # testmode writes the attention gates straight into the [B,T,C,H,W] output instead of the checkpoint
            # buffer; the reverse loop reads them from the checkpoint layout [T][B,C,H,W] (ptk_api.cu:bind_saved:
            # ckI, ckE (T+1 each), Y0, Y1, A, ...), so mirror them there (saliency path, viz_model_att.py:189-196)
            B, _, T, H, W = x.shape
            S = B * cfg.C * H * W
            a_off = (2 * (T + 1) + 2 * T) * S
            saved.view(torch.float32)[a_off:a_off + T * S].view(T, B, cfg.C, H, W).copy_(gates.permute(1, 0, 2, 3, 4))
        ctx.cfg = cfg
        ctx.saved_buf = saved
        ctx.save_for_backward(x, flat)
        if not testmode:
            return logits
        ctx.mark_non_differentiable(gates, states)
        return logits, gates, states

    @staticmethod
    @once_differentiable
    def backward(ctx, dlogits, *_unused):
        x, flat = ctx.saved_tensors
        cfg = ctx.cfg
        L = _lib.lib()
        dims = cfg.dims(x)
        flat = flat.contiguous()
        dlogits = dlogits.contiguous().float()
        with torch.cuda.device(x.device):
            ws, ws_bytes, _ = _buffers(L, dims, x.device, False, backward=True)
            dflat = torch.empty_like(flat)
            dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
            stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            rc = L.ptk_int_backward(ctypes.byref(dims), _ptr(x), _strides(x), _ptr(flat), _ptr(ctx.saved_buf),
                                    _ptr(dlogits), _ptr(dflat), _ptr(dx), _ptr(ws), ws_bytes, stream)
            _lib.check(rc, "ptk_int_backward")
        ctx.saved_buf = None
        return dx, dflat, None, None


def int_apply(x, flat, cfg):
    return _IntFunction.apply(x, flat, cfg)


def synth_pathtracker(B, T, H=32, W=32, n_distractors=14, speed=1, seed=1234, first_index=0, device="cuda"):
    """on-device synthetic PathTracker batch: (x [B,3,T,H,W] float32 in {0,1}, y [B,1] float32)."""
    L = _lib.lib()
    device = torch.device(device)
    with torch.cuda.device(device):
        x = torch.empty(B, 3, T, H, W, dtype=torch.float32, device=device)
        y = torch.empty(B, dtype=torch.float32, device=device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        rc = L.ptk_synth_pathtracker(seed, first_index, B, T, H, W, n_distractors, speed, _ptr(x), _ptr(y), stream)
        _lib.check(rc, "ptk_synth_pathtracker")
    return x, y.view(B, 1)


def conv7(x, weight, conv_mode=_lib.PTK_CONV_FP32, transpose=False, out=None, want_stats=False):
    """one 7x7 horizontal-connection convolution through ptk_conv7 (test / measurement entry)."""
    L = _lib.lib()
    B, Cc, H, W = x.shape
    x, weight = x.contiguous(), weight.contiguous()
    with torch.cuda.device(x.device):
        accumulate = out is not None
        if out is None:
            out = torch.empty_like(x)
        stats = torch.empty(2, Cc, dtype=torch.float32, device=x.device) if want_stats else None
        nbytes = L.ptk_conv7_workspace_bytes(B, H, W)
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=x.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        rc = L.ptk_conv7(B, H, W, conv_mode, int(transpose), int(accumulate), _ptr(x), _ptr(weight), _ptr(out),
                         _ptr(stats), _ptr(ws), nbytes, stream)
        _lib.check(rc, "ptk_conv7")
    return (out, stats) if want_stats else out


def wgrad7(inp, dy, conv_mode=_lib.PTK_CONV_BF16X3):
    """the 7x7 weight gradient summed over T (input, output-gradient) pairs through ptk_wgrad7 (test / measurement
    entry).  inp, dy: [T,B,32,H,W] -> dW [32,32,7,7]."""
    L = _lib.lib()
    T, B, Cc, H, W = inp.shape
    inp, dy = inp.contiguous().float(), dy.contiguous().float()
    with torch.cuda.device(inp.device):
        out = torch.empty(Cc, Cc, 7, 7, dtype=torch.float32, device=inp.device)
        nbytes = L.ptk_wgrad7_workspace_bytes(T, B, H, W)
        if nbytes == 0:
            _lib.check(1, "ptk_wgrad7_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=inp.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(inp.device).cuda_stream)
        rc = L.ptk_wgrad7(T, B, H, W, conv_mode, _ptr(inp), _ptr(dy), _ptr(out), _ptr(ws), nbytes, stream)
        _lib.check(rc, "ptk_wgrad7")
    return out


def lstm_wide_op(op, x, weights, aux=None):
    """one wide operator of the ConvLSTM time loops on the tcgen05 kernels (ptk_lstm_wide_op; test / measurement entry).
    weights [4,25,25,K,K] (Wh of the gates i,f,c,o).  op 'conv': x [B,25,H,W] (+ aux [B,100,H,W]) -> [B,100,H,W];
    'transpose': x [B,100,H,W] -> [B,25,H,W]; 'wgrad': x = h [B,25,H,W], aux = dY [B,100,H,W] -> [4,25,25,K,K]."""
    L = _lib.lib()
    code = {"conv": 0, "transpose": 1, "wgrad": 2}[op]
    B, _, H, W = x.shape
    K = weights.shape[-1]
    x, weights = x.contiguous().float(), weights.contiguous().float()
    aux = aux.contiguous().float() if aux is not None else None
    with torch.cuda.device(x.device):
        shape = {0: (B, 100, H, W), 1: (B, 25, H, W), 2: tuple(weights.shape)}[code]
        out = torch.empty(shape, dtype=torch.float32, device=x.device)
        nbytes = L.ptk_lstm_wide_op_workspace_bytes(K)
        if nbytes == 0:
            _lib.check(1, "ptk_lstm_wide_op_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        rc = L.ptk_lstm_wide_op(B, H, W, K, code, _ptr(x), _ptr(aux), _ptr(weights), _ptr(out), _ptr(ws), nbytes, stream)
        _lib.check(rc, "ptk_lstm_wide_op")
    return out
