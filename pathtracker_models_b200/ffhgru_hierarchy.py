"""Drop-in for the hGRU relative of the reference (models/ffhgru_hierarchy.py:59-276):
identical to InT except that the cell's `gated_inhibition` is the attention gate (:147)."""
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from ._recurrent import _CellParams, _RecurrentModel


class hConvGRUCell(_CellParams):
    """parameters of the hGRU cell (models/ffhgru_hierarchy.py:59-133)."""

    def __init__(self, hidden_size, kernel_size, timesteps, batchnorm=True, grad_method='bptt', use_attention=False):
        if not use_attention:
            raise NotImplementedError("use_attention=False is unrunnable in the reference (ffhgru_hierarchy.py:148-152)")
        super().__init__(hidden_size, kernel_size, timesteps, order=("w_inh", "w_exc"), init_w=1.,
                         scalars=("alpha", "gamma", "kappa", "w", "mu"))   # ffhgru_hierarchy.py:101-105


class FFhGRU(_RecurrentModel):
    """models/ffhgru_hierarchy.py:176-276."""

    _cell_code = _lib.PTK_CELL_HGRU

    def __init__(self, dimensions, timesteps=8, kernel_size=15, jacobian_penalty=False, grad_method='bptt'):
        super().__init__()
        # unused BatchNorm3d kept for state_dict compatibility (ffhgru_hierarchy.py:186)
        self.bn = nn.BatchNorm3d(dimensions, eps=1e-03, track_running_stats=False)
        self._build(dimensions, timesteps, kernel_size,
                    lambda: hConvGRUCell(hidden_size=dimensions, kernel_size=kernel_size, use_attention=True,
                                         timesteps=timesteps),
                    F.softplus, jacobian_penalty, grad_method)
