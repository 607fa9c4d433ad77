"""``RoICropFunction()(input1, input2)`` -- the reference's instance-style call
(lib/model/roi_crop/functions/roi_crop.py:7-21) on top of the static autograd Function."""
from tlod_b200.autograd import RoICropFunction as _Fn


class RoICropFunction(object):
    def __call__(self, input1, input2):
        return _Fn.apply(input1, input2)

    forward = __call__
