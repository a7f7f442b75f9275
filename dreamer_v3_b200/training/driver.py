"""The training part of the reference's driver loop (run_experiment.py:349-472: sample env steps, add them to the replay
buffer, update until `replayed_steps / env_steps >= training_ratio`; :692-714 save; :224-233, :489-509 resume) on the
B200 engine. The environment runner, evaluation and TensorBoard summaries of run_experiment.py are out of scope
(SURVEY.md section 2): the caller passes `collect`, any callable that acts in its environments (typically through
`DreamerModel.forward_inference`) and returns finished / ongoing `Episode` chunks like `EnvRunnerV2.sample` does
(env_runner_v2.py:221-330).

Everything on the update path stays on the device: `EpisodeReplayBuffer.sample` gathers the `[B, T, ...]` batch straight
into the engine's input tensors (csrc/dv3_replay.cu), and each of the two update steps is one CUDA-graph replay.
"""
import os

import torch

from ..utils import checkpoint
from .train_one_step import train_actor_and_critic_one_step, train_world_model_one_step

# run_experiment.py:131-148, 245-252
DEFAULTS = dict(training_ratio=1024, world_model_grad_clip=1000.0, critic_grad_clip=100.0, actor_grad_clip=100.0,
                disagree_grad_clip=100.0, discount_gamma=0.997, gae_lambda=0.95, entropy_scale=3e-4,
                return_normalization_decay=0.99, train_critic=True, train_actor=True)


def _engine_inputs(engine):
    t = engine.tensor
    return {"obs": t("in.obs"), "rewards": t("in.rewards"), "is_first": t("in.is_first"), "is_terminated": t("in.is_terminated")}


def update_from_buffer(*, dreamer_model, buffer, batch_size_B, batch_length_T, horizon_H, **hyper):
    """One sub-iteration of run_experiment.py:404-472: draw a batch, train the world model, then critic and actor."""
    h = dict(DEFAULTS, **hyper)
    e = dreamer_model.engine
    cfg = e.cfg
    out = _engine_inputs(e)
    if not cfg.discrete:
        out["actions"] = e.tensor("in.actions")
    sample = buffer.sample(batch_size_B=batch_size_B, batch_length_T=batch_length_T, out=out)
    if cfg.discrete:  # stored as ints, fed one-hot (run_experiment.py:419-423)
        sample["actions_ints"] = sample["actions"]
        sample["actions"] = torch.nn.functional.one_hot(sample["actions"].long(), cfg.A).to(torch.float32)
    wm_results = train_world_model_one_step(
        sample=sample, batch_size_B=batch_size_B, batch_length_T=batch_length_T, grad_clip=h["world_model_grad_clip"],
        world_model=dreamer_model.world_model)
    ac_results = None
    if h["train_critic"]:
        ac_results = train_actor_and_critic_one_step(
            world_model_forward_train_outs=wm_results["WORLD_MODEL_forward_train_outs"],
            is_terminated=sample["is_terminated"].reshape(-1), horizon_H=horizon_H, gamma=h["discount_gamma"],
            lambda_=h["gae_lambda"], actor_grad_clip=h["actor_grad_clip"], critic_grad_clip=h["critic_grad_clip"],
            disagree_grad_clip=h["disagree_grad_clip"], dreamer_model=dreamer_model, entropy_scale=h["entropy_scale"],
            return_normalization_decay=h["return_normalization_decay"], train_actor=h["train_actor"],
            use_curiosity=dreamer_model.use_curiosity)
    return wm_results, ac_results


def run_training(*, dreamer_model, buffer, collect, num_iterations, batch_size_B, batch_length_T, horizon_H,
                 checkpoint_path=None, model_save_frequency_main_iters=0, resume_from=None, counters=None, on_iteration=None,
                 **hyper):
    """run_experiment.py:349-472 + :692-714. `collect(iteration)` -> list of Episode chunks of the env steps taken since the
    last call. Returns the counters the reference keeps (total_env_steps, total_replayed_steps, total_train_steps,
    iteration). With `resume_from` the models, optimizers, replay buffer and counters are restored first (:224-233, :489-509);
    `counters` (a dict this function returned) continues an earlier call in the same process."""
    h = dict(DEFAULTS, **hyper)
    c = dict(total_env_steps=0, total_replayed_steps=0, total_train_steps=0, iteration=0)
    c.update(counters or {})
    critic_ema_initialised = c["total_train_steps"] > 0
    if resume_from is not None:
        c.update(checkpoint.load_checkpoint(resume_from, dreamer_model, buffer=buffer))
        critic_ema_initialised = True    # the EMA weights came with the critic's weights
    start = c["iteration"]
    for iteration in range(start, start + int(num_iterations)):
        # ---- environment steps (:349-392), at least one batch worth of timesteps in the buffer before training ----
        env_steps = 0
        while True:
            episodes = collect(iteration)
            buffer.add(episodes=episodes)
            env_steps += sum(len(ep) for ep in episodes)
            if buffer.get_num_timesteps() >= batch_size_B * batch_length_T:
                break
        c["total_env_steps"] += env_steps
        # ---- updates until the training ratio is met (:404-472) ----
        replayed, sub_iter, last = 0, 0, (None, None)
        while replayed / env_steps < h["training_ratio"]:
            if h["train_critic"] and not critic_ema_initialised:   # :439-454
                dreamer_model.critic.init_ema()
                critic_ema_initialised = True
            last = update_from_buffer(dreamer_model=dreamer_model, buffer=buffer, batch_size_B=batch_size_B,
                                      batch_length_T=batch_length_T, horizon_H=horizon_H, **hyper)
            replayed += batch_size_B * batch_length_T
            sub_iter += 1
            c["total_train_steps"] += 1
        c["total_replayed_steps"] += replayed
        c["iteration"] = iteration + 1
        if on_iteration is not None:
            on_iteration(iteration, dict(c), last[0], last[1])
        # ---- save (:692-714): <checkpoint_path>/<iteration>/{world_model,actor,critic,...}.npz, buffer.npz, counters ----
        if checkpoint_path and model_save_frequency_main_iters and (iteration + 1) % model_save_frequency_main_iters == 0:
            checkpoint.save_checkpoint(os.path.join(checkpoint_path, str(iteration + 1)), dreamer_model, buffer=buffer, **c)
    return c
