"""Drop-in for models/convlstm.py of the reference (ConvLSTMCell :57-90, ConvLSTM :93-173): same
constructors, parameter names and forward(x, epoch, itr, target, criterion, testmode) contract; the
Gabor stem, the T cell steps, the BatchNorm + 1x1 head, BPTT and the Jacobian penalty run in
libptk_b200 (ptk_convlstm_forward / ptk_convlstm_backward)."""
import ctypes
import math
import os

import numpy as np
import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable
from torch.nn import init

from . import _lib
from .functional import _ptr


class ConvLSTMCell(nn.Module):
    """parameter container of the cell (models/convlstm.py:57-82); executed inside the CUDA time loop."""

    def __init__(self, input_channels, hidden_channels, kernel_size):
        super().__init__()
        self.input_channels, self.hidden_channels, self.kernel_size = input_channels, hidden_channels, kernel_size
        self.num_features = 4
        self.padding = int((kernel_size - 1) / 2)
        for g in "ifco":            # registration (and RNG) order of the reference: Wx<g> then Wh<g>
            setattr(self, "Wx" + g, nn.Conv2d(input_channels, hidden_channels, kernel_size, 1, self.padding, bias=True))
            setattr(self, "Wh" + g, nn.Conv2d(hidden_channels, hidden_channels, kernel_size, 1, self.padding, bias=False))
        self.Wci = self.Wcf = self.Wco = None

    def forward(self, *a, **k):
        raise RuntimeError("the cell is executed inside the fused CUDA time loop; call the model, not the cell")


def _find_gabor(gabor, gabor_path):
    if gabor is not None:
        return torch.as_tensor(gabor, dtype=torch.float32)
    for cand in (gabor_path, os.environ.get("PTK_GABOR_PATH"), "utils/gabor_serre.npy"):   # convlstm.py:105
        if cand and os.path.exists(cand):
            return torch.from_numpy(np.load(cand)).float()
    raise FileNotFoundError("ConvLSTM needs the Gabor bank 'utils/gabor_serre.npy' (models/convlstm.py:105): run from "
                            "the reference root, set PTK_GABOR_PATH, or pass gabor=/gabor_path=")


class _ConvLSTMFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flat, dims, want_penalty, rbp_iters=None, want_states=False, penalty_in_graph=False, owner=None):
        L = _lib.lib()
        B, _, H, W = x.shape
        flat = flat.contiguous()
        x = x.contiguous()
        with torch.cuda.device(x.device):
            ws_bytes = L.ptk_convlstm_workspace_bytes(ctypes.byref(dims))
            sv_bytes = L.ptk_convlstm_saved_bytes(ctypes.byref(dims))
            if ws_bytes == 0 or sv_bytes == 0:
                _lib.check(1, "ptk_convlstm_workspace_bytes")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            saved = torch.empty(sv_bytes, dtype=torch.uint8, device=x.device)
            out = torch.empty(B, 2, H, W, dtype=torch.float32, device=x.device)
            pen = torch.empty(B, 25, H, W, dtype=torch.float32, device=x.device) if want_penalty else None
            stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            rc = L.ptk_convlstm_forward(ctypes.byref(dims), _ptr(x), _ptr(flat), _ptr(out), _ptr(pen), _ptr(saved),
                                        _ptr(ws), ws_bytes, stream)
            _lib.check(rc, "ptk_convlstm_forward")
        ctx.dims, ctx.saved_buf, ctx.rbp_iters = dims, saved, rbp_iters
        ctx.penalty_in_graph = bool(penalty_in_graph and want_penalty)
        ctx.owner = owner            # the module instance that receives `last_rbp_terms` (per module, not process-wide)
        ctx.set_materialize_grads(False)   # an unused output arrives as None: no penalty term in the total -> the ordinary backward
        ctx.out_shape = (B, 2, H, W)
        if want_states:      # testmode of the rbp branch: h_0 .. h_T (models/convlstm.py:127-135); hs is the first saved block
            n = B * 25 * H * W
            states = saved[: (dims.T + 1) * n * 4].view(torch.float32).view(dims.T + 1, B, 25, H, W).clone()
        else:
            states = torch.empty(0, device=x.device)
        ctx.save_for_backward(x, flat)
        if pen is None:
            pen = torch.ones(1, dtype=torch.float32, device=x.device)    # convlstm.py:151 placeholder
        if ctx.penalty_in_graph:      # jacobian_penalty=True: the penalty carries a graph (create_graph=True, convlstm.py:157-160)
            ctx.mark_non_differentiable(states)
        else:
            ctx.mark_non_differentiable(pen, states)
        return out, pen, states

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dpen, _dstates):
        x, flat = ctx.saved_tensors
        L = _lib.lib()
        dims = ctx.dims
        with torch.cuda.device(x.device):
            ws_bytes = L.ptk_convlstm_workspace_bytes(ctypes.byref(dims))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            dflat = torch.empty_like(flat)
            dout = torch.zeros(ctx.out_shape, dtype=torch.float32, device=x.device) if dout is None else dout.contiguous().float()
            stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            if ctx.penalty_in_graph and dpen is not None:
                # loss + <dpen, penalty>: the penalty's double back-propagation in closed form (ptk_convlstm_backward_penalty)
                terms = torch.zeros(1, dtype=torch.float32, device=x.device)
                dpen = dpen.contiguous().float()
                rc = L.ptk_convlstm_backward_penalty(ctypes.byref(dims), _ptr(x), _ptr(flat), _ptr(ctx.saved_buf), _ptr(dout),
                                                     _ptr(dpen), int(ctx.rbp_iters or 0), _ptr(dflat), _ptr(terms), _ptr(ws),
                                                     ws_bytes, stream)
                _lib.check(rc, "ptk_convlstm_backward_penalty")
                if ctx.rbp_iters is not None and ctx.owner is not None:
                    ctx.owner.last_rbp_terms = terms
            elif ctx.rbp_iters is None:
                rc = L.ptk_convlstm_backward(ctypes.byref(dims), _ptr(x), _ptr(flat), _ptr(ctx.saved_buf), _ptr(dout),
                                             _ptr(dflat), _ptr(ws), ws_bytes, stream)
                _lib.check(rc, "ptk_convlstm_backward")
            else:
                terms = torch.zeros(1, dtype=torch.float32, device=x.device)
                rc = L.ptk_convlstm_backward_rbp(ctypes.byref(dims), _ptr(x), _ptr(flat), _ptr(ctx.saved_buf), _ptr(dout),
                                                 int(ctx.rbp_iters), _ptr(dflat), _ptr(terms), _ptr(ws), ws_bytes, stream)
                _lib.check(rc, "ptk_convlstm_backward_rbp")
                if ctx.owner is not None:
                    ctx.owner.last_rbp_terms = terms          # device tensor; read it only when you want the sync
            dx = None
            if ctx.needs_input_grad[0]:       # the reference is differentiable w.r.t. the image: dx = conv0^T (d loss / d u)
                dx = torch.empty_like(x)
                rc = L.ptk_convlstm_input_grad(ctypes.byref(dims), _ptr(flat), _ptr(dx), _ptr(ws), ws_bytes, stream)
                _lib.check(rc, "ptk_convlstm_input_grad")
        ctx.saved_buf = None
        return dx, dflat, None, None, None, None, None, None


class ConvLSTM(nn.Module):
    """models/convlstm.py:93-173: grad_method 'bptt' (full back-propagation through time) or 'rbp' (recurrent
    back-propagation: Neumann series on the last step, `dummyhgru` :9-54, truncated at `num_iter` terms)."""

    def __init__(self, timesteps=8, filt_size=15, num_iter=50, exp_name='exp1', jacobian_penalty=False,
                 grad_method='bptt', gabor=None, gabor_path=None):
        super().__init__()
        self.timesteps, self.num_iter, self.exp_name = timesteps, num_iter, exp_name
        self.jacobian_penalty, self.grad_method = jacobian_penalty, grad_method
        if grad_method not in ('bptt', 'rbp'):
            raise ValueError(f"grad_method must be 'bptt' or 'rbp', got {grad_method!r}")   # the reference silently skips the loop
        self.conv0 = nn.Conv2d(1, 25, kernel_size=7, padding=3)
        self.conv0.weight.data = _find_gabor(gabor, gabor_path).reshape(25, 1, 7, 7).clone()
        self.unit1 = ConvLSTMCell(25, 25, filt_size)
        self.bn = nn.BatchNorm2d(25, eps=1e-03, track_running_stats=False)
        self.conv6 = nn.Conv2d(25, 2, kernel_size=1)
        init.xavier_normal_(self.conv6.weight)
        init.constant_(self.conv6.bias, math.log((1 - 0.01) / 0.01))
        # the two per-step convolutions of the time loops (h -> gates and its transpose): 3-pass split-bf16 tcgen05
        # implicit GEMMs (default) or strict-fp32 CUDA cores; same parity gate (tests/test_convlstm_gpu.py).  _lib.PTK_CONV_BF16 =
        # the bf16 variant (one bf16 product per k-step, gradients ~1e-2: opt-in, not the parity path)
        self.conv_mode = _lib.PTK_CONV_BF16X3
        self.last_rbp_terms = None      # rbp: accepted Neumann terms of the latest backward (device tensor, 1 float)

    def flat_parameters(self):
        # registration order == C ABI order; by attribute (not parameters()) so that nn.DataParallel replicas work
        ts = [self.conv0.weight, self.conv0.bias]
        for g in "ifco":
            wx, wh = getattr(self.unit1, "Wx" + g), getattr(self.unit1, "Wh" + g)
            ts += [wx.weight, wx.bias, wh.weight]
        ts += [self.bn.weight, self.bn.bias, self.conv6.weight, self.conv6.bias]
        return torch.cat([t.reshape(-1) for t in ts])

    def forward(self, x, epoch, itr, target, criterion, testmode=False):
        if not x.is_cuda:
            raise _lib.PtkError("pathtracker_models_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        B, ch, H, W = x.shape
        if ch != 1:
            raise ValueError(f"expected x of shape [B,1,H,W], got {tuple(x.shape)}")
        rbp = self.grad_method == 'rbp'
        # recurrent back-propagation reads only the last two steps: constant-memory checkpoint ring (models/convlstm.py:124-135
        # runs the first T-1 steps under no_grad for exactly that reason); testmode returns every hidden state, so it keeps them
        dims = _lib.PtkLstmDims(B, self.timesteps, H, W, self.unit1.kernel_size, int(self.conv_mode), int(rbp and not testmode),
                                int(rbp))
        from .parallel import flat_gradient_hook
        output, pen, states = _ConvLSTMFunction.apply(x.float(), flat_gradient_hook(self, self.flat_parameters()), dims, bool(self.training),
                                                      int(self.num_iter) if rbp else None, bool(rbp and testmode),
                                                      bool(self.jacobian_penalty), self)
        loss = criterion(output, target)                                # convlstm.py:148
        if testmode:                                                     # the bptt branch never fills `states`
            return output, (list(states.unbind(0)) if rbp else []), loss
        return output, pen, loss
