"""symlog / inverse_symlog with the reference's signatures (utils/symlog.py:9-28), computed by libdv3."""
import torch

from .. import ops


def _dev(x):
    t = torch.as_tensor(x)
    return t.cuda() if not t.is_cuda else t


def symlog(x):
    t = _dev(x)
    return ops.symlog(t).reshape(t.shape)


def inverse_symlog(y):
    t = _dev(y)
    return ops.inverse_symlog(t).reshape(t.shape)
