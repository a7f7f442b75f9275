"""DreamerModel with the reference's Python surface (models/dreamer_model.py:20-303), backed by libdv3."""
import torch

from .actor_network import ActorNetwork
from .critic_network import CriticNetwork
from .disagree_networks import DisagreeNetworks


class _ActionDist:
    """View of one dreamed step's action distribution exposing the members the losses / summaries use
    (actor_loss.py:38,53,65): `.logits` (= log of the unimixed probs) for Discrete, `.loc/.scale` for Box."""

    def __init__(self, discrete, probs=None, loc=None, scale=None):
        self.discrete, self.probs, self.loc, self.scale = discrete, probs, loc, scale

    @property
    def logits(self):
        return torch.log(self.probs)

    def entropy(self):
        if self.discrete:
            p = self.probs / self.probs.sum(-1, keepdim=True)
            return -(p * torch.log(p)).sum(-1)
        return (1.4189385332046727 + torch.log(self.scale)).sum(-1)

    def log_prob(self, a):
        if self.discrete:
            return (a * self.logits).sum(-1)
        z = (a - self.loc) / self.scale
        return (-0.5 * z * z - torch.log(self.scale) - 0.9189385332046727).sum(-1)


class DreamerModel:
    def __init__(self, *, model_dimension="XS", action_space, batch_size_B, batch_length_T, horizon_H, world_model,
                 use_curiosity=False, intrinsic_rewards_scale=0.1, world_size=1, rank=0):
        self.model_dimension = model_dimension
        self.action_space = action_space
        self.use_curiosity = bool(use_curiosity)
        self.batch_size_B, self.batch_length_T, self.horizon_H = batch_size_B, batch_length_T, horizon_H
        self.world_model = world_model
        world_model._configure(B=batch_size_B, H=horizon_H, world_size=world_size, rank=rank,
                               use_curiosity=self.use_curiosity, intrinsic_rewards_scale=float(intrinsic_rewards_scale))
        get = lambda: self.world_model.engine
        self.actor = ActorNetwork(action_space=action_space, model_dimension=model_dimension, engine_getter=get)
        self.critic = CriticNetwork(model_dimension=model_dimension, engine_getter=get)
        # dreamer_model.py:61-67
        self.disagree_nets = DisagreeNetworks(
            model_dimension=model_dimension, intrinsic_rewards_scale=intrinsic_rewards_scale, engine_getter=get) if self.use_curiosity else None

    @property
    def engine(self):
        return self.world_model.engine

    def forward_train(self, observations, actions, is_first, training=None):
        return self.world_model.forward_train(observations, actions, is_first)

    def get_initial_state(self):
        """dreamer_model.py:136-144"""
        s = self.world_model.get_initial_state()
        s["a"] = torch.zeros((1, self.engine.cfg.A), dtype=torch.float32, device=self.engine.device)
        return s

    def forward_inference(self, previous_states, observations, is_first, training=None):
        """dreamer_model.py:113-128 (called per env step by env_runner_v2.py:274-282): world-model step + actor sample.
        Returns (actions, {"h", "z", "a"})."""
        wm = self.world_model
        cfg = self.engine.cfg
        rows, hz = wm._run_inference(previous_states, observations, is_first)
        actions = self.engine.tensor("inf.a")[:rows].clone()
        return actions, {"h": hz[:, :cfg.G].clone(), "z": hz[:, cfg.G:].unflatten(1, (cfg.Zc, cfg.Zk)).clone(), "a": actions}

    def dream_trajectory(self, *, start_states, start_is_terminated, timesteps_H, gamma):
        """dreamer_model.py:147-303. Returns views of engine-owned, time-major buffers."""
        e = self.engine
        cfg = e.cfg
        assert timesteps_H == cfg.H, f"engine was built for horizon {cfg.H}"
        h, z = start_states["h"], start_states["z"]
        hz = e.tensor("wm.hz")
        own = torch.is_tensor(h) and h.is_cuda and h.data_ptr() == hz.data_ptr() and torch.is_tensor(z) and z.is_cuda \
            and z.data_ptr() == hz[:, cfg.G:].data_ptr()
        if not own:
            e.write("in.dream_h0", torch.as_tensor(h).reshape(cfg.N, cfg.G))
            e.write("in.dream_z0", torch.as_tensor(z).reshape(cfg.N, cfg.Z))
        e.write("in.is_terminated", torch.as_tensor(start_is_terminated).reshape(cfg.B, cfg.T))
        self.world_model._draw_noise(2)
        e.dream_forward(gamma=gamma, start_from_observe=own)
        return self._dream_outs()

    def _dream_outs(self):
        e = self.engine
        cfg = e.cfg
        dhz = e.tensor("dream.hz")
        acts = e.tensor("dream.actions")
        if cfg.discrete:
            probs = e.tensor("dream.actor_probs")
            dists = [_ActionDist(True, probs=probs[i]) for i in range(cfg.H + 1)]
        else:
            loc = torch.tanh(e.tensor("dream.actor_out"))
            scale = 0.9 * torch.sigmoid(e.tensor("dream.actor_std_out") + 2.0) + 0.1
            dists = [_ActionDist(False, loc=loc[i], scale=scale[i]) for i in range(cfg.H + 1)]
        ret = {
            "h_states_t0_to_H_B": dhz[..., :cfg.G],
            "z_states_prior_t0_to_H_B": dhz[..., cfg.G:].unflatten(2, (cfg.Zc, cfg.Zk)),
            "rewards_dreamed_t0_to_H_B": e.tensor("dream.rewards"),
            "continues_dreamed_t0_to_H_B": e.tensor("dream.continues"),
            "actions_dreamed_t0_to_H_B": acts,
            "actions_dreamed_distributions_t0_to_H_B": dists,
            "values_dreamed_t0_to_H_B": e.tensor("dream.values"),
            "values_symlog_dreamed_logits_t0_to_HxB": e.tensor("dream.value_logits"),
            "v_symlog_dreamed_ema_t0_to_H_B": e.tensor("dream.values_symlog_ema"),
            "dream_loss_weights_t0_to_H_B": e.tensor("dream.loss_weights"),
        }
        if self.use_curiosity:  # dreamer_model.py:296-298
            ret["rewards_intrinsic_t1_to_H_B"] = e.tensor("dream.rewards_intrinsic")[1:]
            ret["z_predicted_probs_N_HxB"] = list(e.tensor("disagree.z_predicted_probs"))
        if cfg.discrete:
            ret["actions_ints_dreamed_t0_to_H_B"] = e.tensor("dream.action_indices")
        return ret

    def dream_trajectory_with_burn_in(self, *, start_states, timesteps_burn_in, timesteps_H, observations, actions,
                                      use_sampled_actions_in_dream=False, use_random_actions_in_dream=False,
                                      random_actions=None):
        """dreamer_model.py:306-413 (evaluation path, run_experiment.py:629-644): `timesteps_burn_in` posterior steps on
        the given observations / actions (one `forward_inference` kernel sequence each), then `timesteps_H` prior steps
        with the actor's, the given, or random actions (`dv3_dream_forward_actions`). Rows = observations.shape[0] <= B.
        `random_actions` [H, rows] ints makes use_random_actions_in_dream reproducible."""
        assert not (use_sampled_actions_in_dream and use_random_actions_in_dream)
        e = self.engine
        cfg = e.cfg
        assert timesteps_H == cfg.H, f"engine was built for horizon {cfg.H}"
        dev = e.device
        observations = torch.as_tensor(observations)
        actions = torch.as_tensor(actions).to(dev, torch.float32)
        rows = int(observations.shape[0])
        states = dict(start_states)
        for i in range(timesteps_burn_in):
            first = torch.full((rows,), 1.0 if i == 0 else 0.0)
            states = self.world_model.forward_inference(states, observations[:, i], first)
            states["a"] = actions[:, i]
        given = e.tensor("in.dream_actions")
        given.zero_()
        given[0, :rows] = states["a"]
        if use_sampled_actions_in_dream:
            given[1:, :rows] = actions[:, timesteps_burn_in:timesteps_burn_in + timesteps_H].transpose(0, 1)
        elif use_random_actions_in_dream:
            if random_actions is None:
                random_actions = torch.randint(0, cfg.A, (timesteps_H, rows))
            given[1:, :rows] = torch.nn.functional.one_hot(torch.as_tensor(random_actions).long(), cfg.A).to(dev, torch.float32)
        h0, z0 = e.tensor("in.dream_h0"), e.tensor("in.dream_z0")
        h0.zero_(); z0.zero_()
        h0[:rows] = torch.as_tensor(states["h"]).to(dev).reshape(rows, cfg.G)
        z0[:rows] = torch.as_tensor(states["z"]).to(dev).reshape(rows, cfg.Z)
        self.world_model._draw_noise(2)
        e.dream_forward_actions(2 if (use_sampled_actions_in_dream or use_random_actions_in_dream) else 1,
                                start_from_observe=False)
        dhz = e.tensor("dream.hz")[:, :rows]
        a_t0_to_H = e.tensor("dream.actions")[:, :rows]
        ret = {
            "h_states_t0_to_H_B": dhz[..., :cfg.G],
            "z_states_prior_t0_to_H_B": dhz[..., cfg.G:].unflatten(2, (cfg.Zc, cfg.Zk)),
            # inverse_symlog(reward_predictor) / continue_predictor mode on every dreamed state incl. t0 (:374-392)
            "rewards_dreamed_t0_to_H_B": e.tensor("dream.rewards")[:, :rows],
            "continues_dreamed_t0_to_H_B": (e.tensor("dream.continue_logits")[:, :rows] > 0).to(torch.float32),
        }
        key = ("actions_sampled_t0_to_H_B" if use_sampled_actions_in_dream
               else "actions_random_t0_to_H_B" if use_random_actions_in_dream else "actions_dreamed_t0_to_H_B")
        ret[key] = a_t0_to_H
        if cfg.discrete:
            ret[key.replace("actions_", "actions_ints_", 1)] = torch.argmax(a_t0_to_H, dim=-1)
        return ret
