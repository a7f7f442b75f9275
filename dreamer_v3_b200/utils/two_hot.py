"""two_hot with the reference's signature (utils/two_hot.py:9-78), computed by libdv3.
Only the configuration the reference actually uses on the hot path (255 buckets, [-20, 20]) is native."""
import torch

from .. import ops


def two_hot(value, num_buckets: int = 255, lower_bound: float = -20.0, upper_bound: float = 20.0):
    if (num_buckets, float(lower_bound), float(upper_bound)) != (255, -20.0, 20.0):
        raise NotImplementedError("libdv3 implements the 255-bucket [-20, 20] two-hot used by the reward and critic heads")
    t = torch.as_tensor(value)
    t = t.cuda() if not t.is_cuda else t
    return ops.two_hot(t)[4]
