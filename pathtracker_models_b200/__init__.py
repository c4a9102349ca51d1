"""pathtracker_models_b200 -- B200-native (sm_100a) recurrent core of PathTracker's InT / hGRU /
ConvLSTM models behind the reference's nn.Module interface.  See DESIGN.md."""
from . import _lib  # noqa: F401
from . import InT, convlstm, ffhgru_hierarchy, kys  # noqa: F401
from .functional import synth_pathtracker  # noqa: F401


def build(force=False, verbose=False):
    from ._build import build as _b
    return _b(force=force, verbose=verbose)
