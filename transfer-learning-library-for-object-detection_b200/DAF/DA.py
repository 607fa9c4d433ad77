"""Gradient-reversal layer of the DA heads (lib/DAF/DA.py:19-33): identity forward,
backward ``-alpha * grad`` in one kernel (the reference launches neg() and mul())."""
from tlod_b200.autograd import GradReverse, grad_reverse as _grad_reverse


class GRLayer(object):
    """``GRLayer.apply(x)`` with the reference's fixed alpha = 0.1."""

    @staticmethod
    def apply(x):
        return GradReverse.apply(x, 0.1, None)


def grad_reverse(x, alpha=0.1, row_weight=None):
    return _grad_reverse(x, alpha, row_weight)
