"""Drop-in for models/InT.py of the reference: same constructors, parameter names and
forward(x[, testmode]) contract (models/InT.py:58-245), CUDA time loop underneath."""
import torch.nn.functional as F

from . import _lib
from ._recurrent import _CellParams, _RecurrentModel


class rCell(_CellParams):
    """parameters of the Index-and-Track cell (models/InT.py:58-143)."""

    def __init__(self, hidden_size, kernel_size, timesteps, batchnorm=True, grad_method='bptt', use_attention=False,
                 no_inh=False, lesion_alpha=False, lesion_gamma=False, lesion_mu=False, lesion_kappa=False):
        if not use_attention:
            raise NotImplementedError("InT always builds its cell with use_attention=True (models/InT.py:196)")
        lesions = [n for n, f in (("alpha", lesion_alpha), ("mu", lesion_mu), ("gamma", lesion_gamma),
                                  ("kappa", lesion_kappa)) if f]
        super().__init__(hidden_size, kernel_size, timesteps, order=("w_exc", "w_inh"), no_inh=no_inh,
                         lesions=lesions)


class InT(_RecurrentModel):
    """models/InT.py:182-245."""

    _cell_code = _lib.PTK_CELL_INT

    def __init__(self, dimensions, timesteps=8, kernel_size=15, jacobian_penalty=False, grad_method='bptt',
                 no_inh=False, lesion_alpha=False, lesion_mu=False, lesion_gamma=False, lesion_kappa=False,
                 nl=F.softplus):
        super().__init__()
        self._build(dimensions, timesteps, kernel_size,
                    lambda: rCell(hidden_size=dimensions, kernel_size=kernel_size, use_attention=True,
                                  no_inh=no_inh, lesion_alpha=lesion_alpha, lesion_mu=lesion_mu,
                                  lesion_gamma=lesion_gamma, lesion_kappa=lesion_kappa, timesteps=timesteps),
                    nl, jacobian_penalty, grad_method)
