"""Loss / metrics / optimizer either side of the recurrent path, on device.

`bce_with_logits_and_scores(output, target)` replaces `criterion(output, target)` + `acc_scores(target, output)`
(mainclean.py:189-198, utils/misc_functions.py:12-45): one kernel, no `.item()` round trips; returns the
differentiable mean loss and a device tensor `[bal_acc*100, precision, recall, f1]`.

`FusedAdam(model.parameters(), lr)` replaces `torch.optim.Adam(model.parameters(), lr=args.lr)` +
`optimizer.zero_grad()` (mainclean.py:157,205-206): the parameters are re-homed as views of ONE flat fp32 buffer
(values unchanged, `state_dict()` / `named_parameters()` unaffected), so an update is two launches
(gradient gather, fused Adam) instead of ~10 per tensor.  `param_groups[0]['lr']` is read at every step, so
`torch.optim.lr_scheduler.StepLR` (mainclean.py:160) keeps working."""
import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _BceScores(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        if not logits.is_cuda:
            raise _lib.PtkError("bce_with_logits_and_scores needs CUDA tensors (there is no CPU fallback)")
        z = logits.detach().reshape(-1).float().contiguous()
        y = target.detach().reshape(-1).float().contiguous()
        if z.numel() != y.numel():
            raise ValueError("output and target must hold the same number of elements")
        out = torch.empty(5, dtype=torch.float32, device=z.device)
        dz = torch.empty_like(z)
        with torch.cuda.device(z.device):
            _lib.check(_lib.lib().ptk_bce_scores(z.numel(), _ptr(z), _ptr(y), _ptr(dz), _ptr(out), _stream(z.device)),
                       "ptk_bce_scores")
        ctx.save_for_backward(dz)
        ctx.shape = logits.shape
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, gloss, _gscores):
        (dz,) = ctx.saved_tensors
        return (dz * gloss).reshape(ctx.shape), None


def bce_with_logits_and_scores(output, target):
    """-> (loss, scores) with scores = [bal_acc*100, precision, recall, f1] (device tensor, no sync)."""
    loss, out = _BceScores.apply(output, target)
    return loss, out[1:]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam takes one parameter group (mainclean.py:157 uses one)")
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 for p in ps):
            raise _lib.PtkError("FusedAdam needs fp32 CUDA parameters (there is no CPU fallback)")
        self._ps = ps
        self._sizes = [p.numel() for p in ps]
        n = sum(self._sizes)
        dev = ps[0].device
        self._flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in ps:          # re-home every parameter as a view of the flat buffer (values unchanged)
                k = p.numel()
                self._flat[off:off + k].copy_(p.reshape(-1))
                p.data = self._flat[off:off + k].view(p.shape)
                off += k
        self._m = torch.zeros_like(self._flat)
        self._v = torch.zeros_like(self._flat)
        self._g = torch.zeros_like(self._flat)
        self._steps = [0] * len(ps)

    def add_param_group(self, param_group):
        if getattr(self, "_flat", None) is not None:
            raise ValueError("FusedAdam re-homes its parameters into one flat buffer at construction; build a new optimizer "
                             "instead of adding a parameter group")
        super().add_param_group(param_group)

    def state_dict(self):
        """torch.optim.Optimizer.state_dict() plus the fused state (moments and per-parameter step counts), so that a resumed
        run continues Adam's bias correction instead of silently restarting it."""
        sd = super().state_dict()
        sd["fused"] = {"exp_avg": self._m.clone(), "exp_avg_sq": self._v.clone(), "steps": list(self._steps)}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        fused = state_dict.pop("fused", None)
        super().load_state_dict(state_dict)
        if fused is not None:
            if fused["exp_avg"].numel() != self._m.numel() or len(fused["steps"]) != len(self._steps):
                raise ValueError("FusedAdam.load_state_dict: fused state does not match this optimizer's parameters")
            self._m.copy_(fused["exp_avg"]); self._v.copy_(fused["exp_avg_sq"])
            self._steps = [int(v) for v in fused["steps"]]

    @torch.no_grad()
    def step(self, closure=None, zero_grad=False):
        """One Adam update; `zero_grad=True` folds `optimizer.zero_grad()` (grads set to None) into it.
        Like torch.optim.Adam, a parameter whose `.grad` is None is skipped entirely (moments and its own step
        counter untouched): the flat vector is updated in runs of consecutive parameters that share a step count
        (one launch when every parameter has a gradient; the dead `unit1.w` splits it in two)."""
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        views = self._g.split(self._sizes)
        srcs, dsts = [], []
        runs, off = [], 0       # [first element, count, step]
        base = self._flat.data_ptr()
        for i, (p, v) in enumerate(zip(self._ps, views)):
            k = self._sizes[i]
            if p.data_ptr() != base + 4 * off:
                raise RuntimeError("FusedAdam: a parameter's storage was re-assigned after construction (p.data = ...); it no "
                                   "longer lives in the optimizer's flat buffer and would silently stop being updated")
            if p.grad is not None:
                srcs.append(p.grad.reshape(-1)); dsts.append(v)
                self._steps[i] += 1
                if runs and runs[-1][0] + runs[-1][1] == off and runs[-1][2] == self._steps[i]:
                    runs[-1][1] += k
                else:
                    runs.append([off, k, self._steps[i]])
            off += k
        if not srcs:
            return loss
        torch._foreach_copy_(dsts, srcs)
        dev = self._flat.device
        L = _lib.lib()
        with torch.cuda.device(dev):
            st = _stream(dev)
            for first, count, stepno in runs:
                q = lambda t: ctypes.c_void_p(t.data_ptr() + 4 * first)
                rc = L.ptk_adam_step(count, q(self._flat), q(self._g), q(self._m), q(self._v), float(g["lr"]),
                                     float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), stepno, 0, st)
                _lib.check(rc, "ptk_adam_step")
        if zero_grad:
            for p in self._ps:
                p.grad = None
        return loss
