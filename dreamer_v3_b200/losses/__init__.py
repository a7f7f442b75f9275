"""The reference's loss functions (losses/world_model_losses.py, critic_loss.py, actor_loss.py) are fused into
libdv3's loss kernels (csrc/dv3_loss.cu) and run inside train_*_one_step; `compute_value_targets` is also
exposed stand-alone because callers / tests use it directly (critic_loss.py:92-139)."""
import torch

from .. import ops


def compute_value_targets(*, rewards_t0_to_H_BxT, intrinsic_rewards_t1_to_H_BxT=None, continues_t0_to_H_BxT,
                          value_predictions_t0_to_H_BxT, gamma, lambda_):
    dev = lambda x: torch.as_tensor(x, dtype=torch.float32).cuda()
    r = dev(rewards_t0_to_H_BxT)
    if intrinsic_rewards_t1_to_H_BxT is not None:  # critic_loss.py:118-120 (the first reward is never used)
        r = torch.cat([r[:1], r[1:] + dev(intrinsic_rewards_t1_to_H_BxT)], dim=0)
    return ops.value_targets(r, dev(continues_t0_to_H_BxT), dev(value_predictions_t0_to_H_BxT),
                             float(gamma), float(lambda_))
