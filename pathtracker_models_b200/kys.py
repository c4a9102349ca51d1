"""Drop-in for models/kys.py of the reference (ConvGRUCell :7-48, GRU :70-135; `--model gru`, utils/engine.py:147-153):
same constructors, parameter names, state_dict order and init RNG order; `forward(x[, testmode])` ->
`(logits [B,1], jv_penalty [1])` / `(logits, states [B,T,1,H,W], gates [B,T,C,H,W])`.

The whole of GRU.forward (:104-135) -- stem, the T ConvGRU steps and the readout head (1x1 conv, cat with the goal
channel, 5x5 conv, global mean, Linear) -- forward and BPTT runs in libptk_b200 (ptk_gru_forward / ptk_gru_backward);
all 14 parameter tensors travel in one flat vector."""
import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd.function import once_differentiable

from . import _lib
from .functional import _ptr, _strides


class ConvGRUCell(nn.Module):
    """parameter container of the cell (models/kys.py:8-34); executed inside the CUDA time loop."""

    def __init__(self, input_dim, hidden_dim, kernel_size, padding_mode='zeros'):
        super().__init__()
        if padding_mode != 'zeros':
            raise NotImplementedError("only padding_mode='zeros' is built (the only one models/kys.py:GRU uses)")
        self.hidden_dim = hidden_dim
        ks = kernel_size if isinstance(kernel_size, (list, tuple)) else (kernel_size, kernel_size)
        pad = ks[0] // 2, ks[1] // 2
        self.kernel_size = ks[0]
        self.conv_reset = nn.Conv2d(input_dim + hidden_dim, hidden_dim, ks, padding=pad)
        self.conv_update = nn.Conv2d(input_dim + hidden_dim, hidden_dim, ks, padding=pad)
        self.conv_state_new = nn.Conv2d(input_dim + hidden_dim, hidden_dim, ks, padding=pad)

    def forward(self, *a, **k):
        raise RuntimeError("the cell is executed inside the fused CUDA time loop; call the model, not the cell")


class _GruFunction(torch.autograd.Function):
    """logits (and the per-step checkpoints) = stem + ConvGRU time loop + readout; backward = hand-written BPTT."""

    @staticmethod
    def forward(ctx, x, flat, dims, want_dumps):
        L = _lib.lib()
        B, _, T, H, W = x.shape
        C = dims.C
        flat = flat.contiguous()
        with torch.cuda.device(x.device):
            ws_bytes = L.ptk_gru_workspace_bytes(ctypes.byref(dims))
            sv_bytes = L.ptk_gru_saved_bytes(ctypes.byref(dims))
            if ws_bytes == 0 or sv_bytes == 0:
                _lib.check(1, "ptk_gru_workspace_bytes")
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            saved = torch.empty(sv_bytes, dtype=torch.uint8, device=x.device)
            logits = torch.empty(B, 1, dtype=torch.float32, device=x.device)
            stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            rc = L.ptk_gru_forward(ctypes.byref(dims), _ptr(x), _strides(x), _ptr(flat), _ptr(logits), _ptr(saved), _ptr(ws),
                                   ws_bytes, stream)
            _lib.check(rc, "ptk_gru_forward")
        ctx.dims, ctx.saved_buf = dims, saved
        ctx.save_for_backward(x, flat)
        if want_dumps:      # saved layout (ptk_api.cu:gru_carve): h [T+1][B,C,P], r [T][B,C,P], ...
            S = B * C * H * W
            sv = saved.view(torch.float32)
            hs = sv[S:(T + 1) * S].view(T, B, C, H, W).permute(1, 0, 2, 3, 4).contiguous()
            rs = sv[(T + 1) * S:(2 * T + 1) * S].view(T, B, C, H, W).permute(1, 0, 2, 3, 4).contiguous()
        else:
            hs = rs = torch.empty(0, device=x.device)
        ctx.mark_non_differentiable(hs, rs)
        return logits, hs, rs

    @staticmethod
    @once_differentiable
    def backward(ctx, dlogits, _dhs, _drs):
        x, flat = ctx.saved_tensors
        L = _lib.lib()
        dims = ctx.dims
        with torch.cuda.device(x.device):
            ws_bytes = L.ptk_gru_workspace_bytes(ctypes.byref(dims))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            dflat = torch.empty_like(flat)
            dx = torch.empty(x.shape, dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
            dlogits = dlogits.contiguous().float()
            stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            rc = L.ptk_gru_backward(ctypes.byref(dims), _ptr(x), _strides(x), _ptr(flat), _ptr(ctx.saved_buf), _ptr(dlogits),
                                    _ptr(dflat), _ptr(dx), _ptr(ws), ws_bytes, stream)
            _lib.check(rc, "ptk_gru_backward")
        ctx.saved_buf = None
        return dx, dflat, None, None


class GRU(nn.Module):
    def __init__(self, dimensions, timesteps=8, kernel_size=15, jacobian_penalty=False, grad_method='bptt'):
        super().__init__()
        self.timesteps, self.jacobian_penalty, self.grad_method = timesteps, jacobian_penalty, grad_method
        self.hgru_size = dimensions
        if dimensions != 64 or kernel_size != 7:
            raise NotImplementedError(f"only dimensions=64, kernel_size=7 are built (utils/engine.py:147-153 passes exactly "
                                      f"these); got {dimensions}, {kernel_size}")
        # construction (and RNG) order of models/kys.py:80-102
        self.bn = nn.BatchNorm3d(dimensions, eps=1e-03, track_running_stats=False)     # unused by forward, as in the reference
        self.preproc = nn.Conv3d(3, dimensions, kernel_size=1, padding=0)
        self.unit1 = ConvGRUCell(input_dim=dimensions, hidden_dim=dimensions, kernel_size=kernel_size)
        self.readout_conv = nn.Conv2d(dimensions, 1, 1)
        self.target_conv = nn.Conv2d(2, 1, 5, padding=2)
        torch.nn.init.zeros_(self.target_conv.bias)
        self.readout_dense = nn.Linear(1, 1)
        self.nl = F.softplus
        # cell convolutions: 3-pass split-bf16 tcgen05 blocks (default) or strict-fp32 CUDA cores; same parity gate
        self.conv_mode = _lib.PTK_CONV_BF16X3

    def flat_parameters(self):
        """the 14 tensors of stem, cell and head in the C ABI's order (by attribute: works on DataParallel replicas)."""
        u = self.unit1
        ts = [self.preproc.weight, self.preproc.bias, u.conv_reset.weight, u.conv_reset.bias, u.conv_update.weight,
              u.conv_update.bias, u.conv_state_new.weight, u.conv_state_new.bias, self.readout_conv.weight,
              self.readout_conv.bias, self.target_conv.weight, self.target_conv.bias, self.readout_dense.weight,
              self.readout_dense.bias]
        return torch.cat([t.reshape(-1) for t in ts])

    def forward(self, x, testmode=False):
        if not x.is_cuda:
            raise _lib.PtkError("pathtracker_models_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != 3:
            raise ValueError(f"expected x of shape [B,3,T,H,W], got {tuple(x.shape)}")
        x = x.float()
        B, _, T, H, W = x.shape
        dims = _lib.PtkGruDims(B, T, H, W, self.hgru_size, self.unit1.kernel_size, int(self.conv_mode))
        from .parallel import flat_gradient_hook
        output, hs, rs = _GruFunction.apply(x, flat_gradient_hook(self, self.flat_parameters()), dims, bool(testmode))
        if testmode:
            with torch.no_grad():
                states = self.readout_conv(hs.reshape(B * T, self.hgru_size, H, W)).view(B, T, 1, H, W)
            return output, states, rs
        return output, torch.ones(1, dtype=torch.float32, device=x.device)      # kys.py:135 placeholder penalty
