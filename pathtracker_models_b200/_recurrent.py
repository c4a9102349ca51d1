"""Host-side mirror of the reference's recurrent modules: parameter containers with the same
names / shapes / init order as models/InT.py:58-208 and models/ffhgru_hierarchy.py:59-209, whose
forward() hands the whole time loop to the CUDA library (no per-step PyTorch ops).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init

from . import _lib
from .functional import IntConfig, int_apply, int_forward_raw

# canonical flat-parameter order == PTK_P_* in include/ptk_b200.h
_GATES = ("a_w", "a_u", "i_w", "i_u", "e_w", "e_u")
PARAM_ORDER = (["preproc.weight", "preproc.bias", "unit1.w_exc", "unit1.w_inh", "unit1.alpha", "unit1.mu",
                "unit1.gamma", "unit1.kappa"]
               + [f"unit1.{g}_gate.{s}" for g in _GATES for s in ("weight", "bias")]
               + ["unit1.bn.0.weight", "unit1.bn.0.bias", "unit1.bn.1.weight", "unit1.bn.1.bias",
                  "readout_conv.weight", "readout_conv.bias", "target_conv.weight", "target_conv.bias",
                  "readout_dense.weight", "readout_dense.bias"])


def _act_code(nl):
    if nl in (F.softplus,) or getattr(nl, "__name__", "") == "softplus":
        return _lib.PTK_ACT_SOFTPLUS
    if nl in (torch.tanh, F.tanh) or getattr(nl, "__name__", "") == "tanh":
        return _lib.PTK_ACT_TANH
    raise NotImplementedError(f"nl={nl!r}: the CUDA circuit implements F.softplus and tanh (utils/engine.py:77-146)")


class _CellParams(nn.Module):
    """Parameters of one recurrent cell.  `order` reproduces the RNG consumption order of the
    reference constructor so that torch.manual_seed(s); Cell(...) is bit-identical to it."""

    def __init__(self, hidden_size, kernel_size, timesteps, order, no_inh=False, init_w=None, lesions=(),
                 scalars=("alpha", "mu", "gamma", "kappa", "w")):
        super().__init__()
        self.hidden_size, self.kernel_size, self.timesteps = hidden_size, kernel_size, timesteps
        self.padding = self.h_padding = kernel_size // 2
        self.use_attention, self.no_inh, self.batchnorm = True, no_inh, True
        C, K = hidden_size, kernel_size
        mk = lambda: nn.Conv2d(C, C, 1)
        self.a_w_gate, self.a_u_gate = mk(), mk()
        init.orthogonal_(self.a_w_gate.weight)
        init.orthogonal_(self.a_u_gate.weight)
        init.constant_(self.a_w_gate.bias, 1.)
        init.constant_(self.a_u_gate.bias, 1.)
        self.i_w_gate, self.i_u_gate, self.e_w_gate, self.e_u_gate = mk(), mk(), mk(), mk()
        self.w_exc = nn.Parameter(torch.empty(C, C, K, K))
        if not no_inh:
            self.w_inh = nn.Parameter(torch.empty(C, C, K, K))
        for name in order:                       # InT: exc then inh; hGRU: inh then exc
            if hasattr(self, name):
                init.orthogonal_(getattr(self, name))
        for n in scalars:                        # registration order == state_dict order of the reference
            setattr(self, n, nn.Parameter(torch.empty(C, 1, 1)))
        self.bn = nn.ModuleList([nn.BatchNorm2d(C, eps=1e-03, affine=True, track_running_stats=False)
                                 for _ in range(2)])
        for g in (self.i_w_gate, self.i_u_gate, self.e_w_gate, self.e_u_gate):
            init.orthogonal_(g.weight)
        for bn in self.bn:
            init.constant_(bn.weight, 0.1)
        with torch.no_grad():
            if not no_inh:
                self.alpha.fill_(1.)
                self.mu.fill_(0.)
            else:                                # reference leaves these uninitialised (InT.py:112-114)
                self.alpha.fill_(0.)
                self.mu.fill_(0.)
            self.gamma.fill_(0.)
            self.kappa.fill_(1.)
            self.w.fill_(1. if init_w is None else init_w)   # dead parameter (InT.py:100,118)
            for g, src in ((self.i_w_gate, self.a_w_gate), (self.e_w_gate, self.a_w_gate),
                           (self.i_u_gate, self.a_u_gate), (self.e_u_gate, self.a_u_gate)):
                g.bias.copy_(-src.bias)          # InT.py:121-125
        for n in lesions:                        # frozen at init value, not zeroed (InT.py:132-143)
            getattr(self, n).requires_grad = False

    def forward(self, *a, **k):
        raise RuntimeError("the cell is executed inside the fused CUDA time loop; call the model, not the cell")


def _attr_by_path(module, path):
    obj = module
    for part in path.split("."):
        obj = obj[int(part)] if part.isdigit() else getattr(obj, part, None)
        if obj is None:
            return None
    return obj


class _RecurrentModel(nn.Module):
    """stem + cell + readout container; forward = one call into the CUDA library."""

    _cell_code = _lib.PTK_CELL_INT

    def _build(self, dimensions, timesteps, kernel_size, cell, nl, jacobian_penalty, grad_method):
        self.timesteps = timesteps
        self.jacobian_penalty = jacobian_penalty
        self.grad_method = grad_method
        self.hgru_size = dimensions
        self.preproc = nn.Conv3d(3, dimensions, kernel_size=1)
        self.unit1 = cell()
        self.readout_conv = nn.Conv2d(dimensions, 1, 1)
        self.target_conv = nn.Conv2d(2, 1, 5, padding=2)
        init.zeros_(self.target_conv.bias)
        self.readout_dense = nn.Linear(1, 1)
        self.nl = nl
        # 7x7 convolutions: 3-pass split-bf16 tcgen05 tiles (default) or strict-fp32 CUDA cores; both pass the
        # same parity gate against the fp64 reference (tests/test_parity_gpu.py)
        self.conv_mode = _lib.PTK_CONV_BF16X3
        # grad_method / jacobian_penalty are stored and otherwise ignored, exactly as the reference does
        # (models/InT.py:189-190: InT.forward has no rbp branch and returns a constant penalty, :242-245)

    def _config(self):
        return IntConfig(C=self.hgru_size, K=self.unit1.kernel_size, cell=self._cell_code, act=_act_code(self.nl),
                         no_inh=self.unit1.no_inh, conv_mode=self.conv_mode)

    def flat_parameters(self):
        """the 30 live parameters, concatenated in the C ABI's canonical order (differentiable)."""
        # looked up by attribute path, not named_parameters(): nn.DataParallel replicas (mainclean.py:132-134) hold
        # their broadcast copies as plain tensor attributes and report no parameters at all
        parts = []
        for n in PARAM_ORDER:
            t = _attr_by_path(self, n)
            if t is not None:
                parts.append(t.reshape(-1))
            else:   # InT(no_inh=True) has no w_inh (models/InT.py:91-93): the slot stays, filled with zeros
                ref = _attr_by_path(self, "unit1.w_exc")
                parts.append(torch.zeros(ref.numel(), dtype=ref.dtype, device=ref.device))
        return torch.cat(parts)

    def forward(self, x, testmode=False):
        cfg = self._config()
        if x.dtype != torch.float32:
            x = x.float()
        flat = self.flat_parameters()
        B = x.shape[0]
        if testmode:   # models/InT.py:230-233,244: (output, states [B,T,1,H,W], gates [B,T,C,H,W])
            if torch.is_grad_enabled() and (x.requires_grad or flat.requires_grad):
                # the logits stay differentiable (saliency maps: imgs.requires_grad = True; model_step(test=True);
                # score.backward(); imgs.grad -- viz_model_att.py:189-196); states / gates are plain outputs
                from .functional import _IntFunction
                logits, gates, states = _IntFunction.apply(x, flat, cfg, True)
                return logits.view(B, 1), states, gates
            with torch.no_grad():
                logits, _, gates, states = int_forward_raw(x, flat, cfg, want_saved=False, testmode=True)
            return logits.view(B, 1), states, gates
        if getattr(self, "_dp_enabled", False) and flat.requires_grad:
            from .parallel import average_flat_gradient_
            group = self._dp_group
            flat.register_hook(lambda g: average_flat_gradient_(g.contiguous(), group))
        logits = int_apply(x, flat, cfg)
        jv_penalty = torch.ones(1, dtype=torch.float32, device=x.device)   # models/InT.py:243
        return logits.view(B, 1), jv_penalty
