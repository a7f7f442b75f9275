"""train_world_model_one_step / train_actor_and_critic_one_step with the reference's signatures and
result-dict keys (training/train_one_step.py:16-252). Every tensor in the results is a view of an
engine-owned device buffer (valid until the next update); nothing is synchronised here.

Data parallelism: when torch.distributed is initialised with world_size > 1 (and the models were built with
that world_size), the flat gradient buffers are all-reduced (sum; the engine already divides the losses by
the global batch) between the backward pass and the optimizer step, and the value targets are all-gathered
before the percentile (the only statistic that does not decompose over the batch)."""
import os

import torch
import torch.distributed as dist


def _dp(engine):
    return engine.world_size > 1 and dist.is_available() and dist.is_initialized()


# Gradient all-reduce overlapped with the backward pass: worth its extra graph launch once the collective is
# bandwidth-bound; below this many bytes of head-network gradients (XS: 2 MB, XL: 390 MB) one collective after the backward
# pass is faster. DV3_DP_OVERLAP=1 / 0 forces it on / off.
_OVERLAP_MIN_BYTES = 32 << 20


def _overlap(engine):
    env = os.environ.get("DV3_DP_OVERLAP")
    if env is not None:
        return env not in ("", "0")
    return engine.group_buffer("world_model_heads", "grad").numel() * 4 >= _OVERLAP_MIN_BYTES


def train_world_model_one_step(*, sample, batch_size_B, batch_length_T, grad_clip, world_model):
    world_model._configure(B=int(batch_size_B))
    e = world_model.engine
    assert int(batch_length_T) == e.cfg.T
    e.set_sample(sample)
    opt = world_model.optimizer
    if _dp(e):
        # data parallel: graph-replayed pieces with the NCCL collectives between them
        e.hp.wm_grad_clip, e.hp.wm_lr, e.hp.wm_eps = float(grad_clip), opt.learning_rate, opt.epsilon
        seed = 0 if world_model._explicit_noise else world_model.noise_seed + 104729 * e.rank
        if _overlap(e):
            # three buckets in finalisation order: the head networks' gradients (the tail of the flat buffer) are final after
            # piece 5, the recurrent part's after piece 6; their all-reduces run on NCCL's stream next to the pieces that
            # follow (which only write the buckets in front), the small encoder bucket closes the pass
            e.dp_piece(5, seed)
            w1 = dist.all_reduce(e.group_buffer("world_model_heads", "grad"), async_op=True)
            e.dp_piece(6)
            w2 = dist.all_reduce(e.group_buffer("world_model_recurrent", "grad"), async_op=True)
            e.dp_piece(7)
            dist.all_reduce(e.group_buffer("world_model_encoder", "grad"))
            w1.wait()
            w2.wait()
        else:
            e.dp_piece(0, seed)
            dist.all_reduce(e.group_buffer("world_model", "grad"))
        e.dp_piece(1)
    else:
        # single process: forward + losses + BPTT + clip/Adam replayed as ONE CUDA graph (noise drawn inside it)
        e.hp.wm_grad_clip, e.hp.wm_lr, e.hp.wm_eps = float(grad_clip), opt.learning_rate, opt.epsilon
        e.train_world_model_step(noise_seed=0 if world_model._explicit_noise else world_model.noise_seed)
    t = e.tensor
    s = t("wm.scalars")
    stats = t("opt.world_model.stats")
    L_dec, L_rew, L_con = t("wm.L_decoder_B_T"), t("wm.L_reward_B_T"), t("wm.L_continue_B_T")
    return {
        "WORLD_MODEL_forward_train_outs": world_model._forward_outs(),
        "WORLD_MODEL_learned_initial_h": world_model.initial_h,
        "WORLD_MODEL_L_decoder_B_T": L_dec, "WORLD_MODEL_L_decoder": s[0],
        "WORLD_MODEL_L_reward_B_T": L_rew, "WORLD_MODEL_L_reward": s[1],
        "WORLD_MODEL_L_continue_B_T": L_con, "WORLD_MODEL_L_continue": s[2],
        "WORLD_MODEL_L_prediction_B_T": L_dec + L_rew + L_con, "WORLD_MODEL_L_prediction": s[0] + s[1] + s[2],
        "WORLD_MODEL_L_dynamics_B_T": t("wm.L_dynamics_B_T"), "WORLD_MODEL_L_dynamics": s[4],
        "WORLD_MODEL_L_representation_B_T": t("wm.L_representation_B_T"), "WORLD_MODEL_L_representation": s[5],
        "WORLD_MODEL_L_total_B_T": t("wm.L_total_B_T"), "WORLD_MODEL_L_total": s[6],
        "WORLD_MODEL_gradients_maxabs": stats[1],
        "WORLD_MODEL_gradients_clipped_by_glob_norm_maxabs": stats[2],
    }


def train_actor_and_critic_one_step(*, world_model_forward_train_outs, is_terminated, horizon_H, gamma, lambda_,
                                    actor_grad_clip, critic_grad_clip, disagree_grad_clip, dreamer_model,
                                    entropy_scale, return_normalization_decay, train_actor=True, use_curiosity=False):
    e = dreamer_model.engine
    if bool(use_curiosity) != (e.cfg.num_curiosity_nets > 0):
        raise ValueError("use_curiosity must match the DreamerModel's (dreamer_model.py:62-67 builds the disagree nets)")
    if use_curiosity and _dp(e):
        raise ValueError("curiosity is single-process (the intrinsic rewards are centred over the whole dream batch)")
    hp = e.hp
    if use_curiosity:
        dopt = dreamer_model.disagree_nets.optimizer
        hp.disagree_grad_clip, hp.disagree_lr, hp.disagree_eps = float(disagree_grad_clip), dopt.learning_rate, dopt.epsilon
    hp.gamma, hp.lambda_, hp.entropy_scale = float(gamma), float(lambda_), float(entropy_scale)
    hp.return_normalization_decay = float(return_normalization_decay)
    hp.train_actor = int(bool(train_actor))
    t = e.tensor
    aopt, copt = dreamer_model.actor.optimizer, dreamer_model.critic.optimizer
    hz = t("wm.hz")
    h_in = world_model_forward_train_outs["h_states_BxT"]
    own_states = torch.is_tensor(h_in) and h_in.is_cuda and h_in.data_ptr() == hz.data_ptr()
    if _dp(e) and own_states and horizon_H == e.cfg.H:
        wm = dreamer_model.world_model
        e.write("in.is_terminated", torch.as_tensor(is_terminated).reshape(e.cfg.B, e.cfg.T))
        hp.actor_grad_clip, hp.critic_grad_clip = float(actor_grad_clip), float(critic_grad_clip)
        hp.ac_lr, hp.ac_eps, hp.critic_ema_decay = aopt.learning_rate, aopt.epsilon, dreamer_model.critic.ema_decay
        hp.critic_lr, hp.critic_eps = copt.learning_rate, copt.epsilon
        seed = 0 if wm._explicit_noise else wm.noise_seed + 104729 * e.rank
        e.dp_piece(2, seed)
        dist.all_gather_into_tensor(t("ac.value_targets_all").reshape(-1), t("ac.value_targets").reshape(-1))
        e.dp_piece(3)
        # actor and critic gradients are adjacent in memory: one collective for both
        dist.all_reduce(e.group_buffer("actor_critic" if train_actor else "critic", "grad"))
        e.dp_piece(4)
        return _ac_results(e, dreamer_model._dream_outs(), train_actor)
    if not _dp(e) and own_states and horizon_H == e.cfg.H:
        # single process, dreaming from the engine's own last observe pass: the whole actor/critic update (dream, value
        # targets, losses, gradients, clip/Adam x2, critic EMA) is ONE CUDA-graph replay
        wm = dreamer_model.world_model
        e.write("in.is_terminated", torch.as_tensor(is_terminated).reshape(e.cfg.B, e.cfg.T))
        hp.actor_grad_clip, hp.critic_grad_clip = float(actor_grad_clip), float(critic_grad_clip)
        hp.ac_lr, hp.ac_eps, hp.critic_ema_decay = aopt.learning_rate, aopt.epsilon, dreamer_model.critic.ema_decay
        hp.critic_lr, hp.critic_eps = copt.learning_rate, copt.epsilon
        e.train_actor_critic_step(noise_seed=0 if wm._explicit_noise else wm.noise_seed)
        return _ac_results(e, dreamer_model._dream_outs(), train_actor, use_curiosity)
    dream_data = dreamer_model.dream_trajectory(
        start_states={"h": world_model_forward_train_outs["h_states_BxT"],
                      "z": world_model_forward_train_outs["z_states_BxT"]},
        start_is_terminated=is_terminated, timesteps_H=horizon_H, gamma=gamma)
    if _dp(e):
        e.ac_loss_backward(1)
        dist.all_gather_into_tensor(t("ac.value_targets_all").reshape(-1), t("ac.value_targets").reshape(-1))
        e.ac_loss_backward(2)
        dist.all_reduce(e.group_buffer("actor_critic" if train_actor else "critic", "grad"))
    else:
        e.ac_loss_backward(3)
    if train_actor:
        e.optimizer_step("actor", float(actor_grad_clip), aopt.learning_rate, aopt.epsilon)
    # critic Adam + update_ema() fused (train_one_step.py:235-250)
    e.optimizer_step("critic", float(critic_grad_clip), copt.learning_rate, copt.epsilon, dreamer_model.critic.ema_decay)
    if use_curiosity:  # train_one_step.py:180-181, 209-212, 226-246
        e.disagree_loss_backward()
        e.optimizer_step("disagree", hp.disagree_grad_clip, hp.disagree_lr, hp.disagree_eps)
    return _ac_results(e, dream_data, train_actor, use_curiosity)


def _ac_results(e, dream_data, train_actor, use_curiosity=False):
    t = e.tensor
    s = t("ac.scalars")
    results = {
        "VALUE_TARGETS_symlog_H_B": t("ac.VALUE_TARGETS_symlog_H_B"),
        "CRITIC_L_total": s[0], "CRITIC_L_total_H_B": t("ac.CRITIC_L_total_H_B"),
        "CRITIC_L_neg_logp_of_value_targets_H_B": t("ac.CRITIC_L_neg_logp_of_value_targets_H_B"),
        "CRITIC_L_neg_logp_of_value_targets": s[1],
        "CRITIC_L_slow_critic_regularization_H_B": t("ac.CRITIC_L_slow_critic_regularization_H_B"),
        "CRITIC_L_slow_critic_regularization": s[2],
        "VALUE_TARGETS_H_B": t("ac.value_targets"),
        "dream_data": dream_data,
    }
    cst = t("opt.critic.stats")
    results["CRITIC_gradients_maxabs"] = cst[1]
    results["CRITIC_gradients_clipped_by_glob_norm_maxabs"] = cst[2]
    if train_actor:
        ast = t("opt.actor.stats")
        pct = t("ac.pct_ema")
        results.update({
            "ACTOR_L_total_H_B": t("ac.ACTOR_L_total_H_B"), "ACTOR_L_total": s[3],
            "ACTOR_logp_actions_dreamed_H_B": t("ac.ACTOR_logp_actions_dreamed_H_B"),
            "ACTOR_scaled_value_targets_H_B": t("ac.ACTOR_scaled_value_targets_H_B"),
            "ACTOR_value_targets_pct95_ema": pct[1], "ACTOR_value_targets_pct5_ema": pct[0],
            "ACTOR_action_entropy_H_B": t("ac.ACTOR_action_entropy_H_B"), "ACTOR_action_entropy": s[4],
            "ACTOR_L_neglogp_reinforce_term_H_B": t("ac.ACTOR_L_neglogp_reinforce_term_H_B"),
            "ACTOR_L_neglogp_reinforce_term": s[5],
            "ACTOR_L_neg_entropy_term_H_B": t("ac.ACTOR_L_neg_entropy_term_H_B"), "ACTOR_L_neg_entropy_term": s[6],
            "ACTOR_gradients_maxabs": ast[1], "ACTOR_gradients_clipped_by_glob_norm_maxabs": ast[2],
        })
    if use_curiosity:  # train_one_step.py:201-208, 230-231
        ds, dst = t("disagree.scalars"), t("opt.disagree.stats")
        results.update({
            "DISAGREE_L_total": ds[0],
            "DISAGREE_intrinsic_rewards_H_B": dream_data["rewards_intrinsic_t1_to_H_B"],
            "DISAGREE_intrinsic_rewards": ds[1],
            "DISAGREE_gradients_maxabs": dst[1], "DISAGREE_gradients_clipped_by_glob_norm_maxabs": dst[2],
        })
    return results
