"""The reference's loss functions (losses/world_model_losses.py, critic_loss.py, actor_loss.py) are fused into
libdv3's loss kernels (csrc/dv3_loss.cu) and run inside train_*_one_step; `compute_value_targets` is also
exposed stand-alone because callers / tests use it directly (critic_loss.py:92-139)."""
import torch

from .. import ops


def compute_value_targets(*, rewards_t0_to_H_BxT, intrinsic_rewards_t1_to_H_BxT=None, continues_t0_to_H_BxT,
                          value_predictions_t0_to_H_BxT, gamma, lambda_):
    if intrinsic_rewards_t1_to_H_BxT is not None:
        raise NotImplementedError("curiosity is out of scope")
    dev = lambda x: torch.as_tensor(x, dtype=torch.float32).cuda()
    return ops.value_targets(dev(rewards_t0_to_H_BxT), dev(continues_t0_to_H_BxT), dev(value_predictions_t0_to_H_BxT),
                             float(gamma), float(lambda_))
