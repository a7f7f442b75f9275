"""Learning sanity on the GPU (the reference's only integration check is 'the world model can memorise trivially
predictable data', project_log.txt:205-207): repeated updates on one fixed synthetic batch through the public API
(CUDA-graph replay, noise drawn on the device) must drive the world-model losses down and stay finite."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("discrete,mode", [(True, 0), (False, 1)])
def test_losses_decrease_on_fixed_batch(discrete, mode):
    from dreamer_v3_b200.models.components import MLP, VectorDecoder
    from dreamer_v3_b200.models.dreamer_model import DreamerModel
    from dreamer_v3_b200.models.world_model import WorldModel
    from dreamer_v3_b200.spec import Box, Discrete
    from dreamer_v3_b200.training.train_one_step import train_actor_and_critic_one_step, train_world_model_one_step
    from oracle.dreamer_oracle import make_synthetic_sample
    B, T, H = 8, 16, 5
    space = Discrete(3) if discrete else Box(-1, 1, (2,))
    wm = WorldModel(model_dimension="XXS", action_space=space, batch_length_T=T, encoder=MLP(model_dimension="XXS"),
                    decoder=VectorDecoder(model_dimension="XXS", observation_space=Box(-1, 1, (6,))), compute_mode=mode)
    dm = DreamerModel(model_dimension="XXS", action_space=space, batch_size_B=B, batch_length_T=T, horizon_H=H, world_model=wm)
    dm.critic.init_ema()
    sample = make_synthetic_sample(dm.engine.cfg, seed=0)
    hist = []
    for step in range(150):
        r = train_world_model_one_step(sample=sample, batch_size_B=B, batch_length_T=T, grad_clip=1000.0, world_model=wm)
        a = train_actor_and_critic_one_step(
            world_model_forward_train_outs=r["WORLD_MODEL_forward_train_outs"],
            is_terminated=torch.as_tensor(sample["is_terminated"]).reshape(-1), horizon_H=H, gamma=0.997, lambda_=0.95,
            actor_grad_clip=100.0, critic_grad_clip=100.0, disagree_grad_clip=100.0, dreamer_model=dm,
            entropy_scale=3e-4, return_normalization_decay=0.99)
        if step % 10 == 0 or step == 149:
            hist.append([float(r["WORLD_MODEL_L_total"]), float(r["WORLD_MODEL_L_decoder"]), float(r["WORLD_MODEL_L_reward"]),
                         float(a["CRITIC_L_total"]), float(a["ACTOR_L_total"])])
    hist = np.array(hist)
    assert np.isfinite(hist).all(), hist
    assert hist[-1, 0] < 0.8 * hist[0, 0], hist[:, 0]          # total world-model loss
    assert hist[-1, 1] < 0.7 * hist[0, 1], hist[:, 1]          # decoder SSE
    assert hist[-1, 2] < hist[0, 2], hist[:, 2]                # two-hot reward CE moves off log(255)
    assert dm.critic.optimizer.iterations == 150 and wm.optimizer.iterations == 150
    dm.engine.close()
