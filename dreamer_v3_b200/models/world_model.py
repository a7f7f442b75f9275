"""WorldModel with the reference's Python surface (models/world_model.py:22-346), backed by libdv3."""
import numpy as np
import torch

from ..engine import Engine
from ..spec import EngineConfig, action_space_info, init_params, param_spec


class _Normal:
    """Minimal view of MultivariateNormalDiag(loc, 1) -- only `.loc` is consumed (world_model_losses.py:36)."""

    def __init__(self, loc):
        self.loc = loc


class _Bernoulli:
    """Minimal view of tfp Bernoulli(logits) (continue_predictor.py:41-46; world_model_losses.py:74)."""

    def __init__(self, logits):
        self.logits = logits

    def mode(self):
        return (self.logits > 0).to(torch.float32)

    def log_prob(self, x):
        return -torch.nn.functional.binary_cross_entropy_with_logits(self.logits, x, reduction="none")


class _OptimizerInfo:
    def __init__(self, engine_getter, group, learning_rate, epsilon):
        self._e, self.group, self.learning_rate, self.epsilon = engine_getter, group, learning_rate, epsilon

    @property
    def iterations(self):
        return int(self._e().tensor(f"opt.{self.group}.step").item())

    def get_weights(self):
        """Keras-2 `optimizer.get_weights()` (run_experiment.py:706): [iterations, m..., v...]."""
        from ..utils import checkpoint
        return checkpoint.get_optimizer_weights(self._e(), self.group)

    def set_weights(self, weights):
        from ..utils import checkpoint
        checkpoint.set_optimizer_weights(self._e(), self.group, weights)


class _Variables:
    """Named zero-copy views of one optimizer group's parameters."""

    def __init__(self, engine_getter, group, prefix=""):
        self._e, self.group, self.prefix = engine_getter, group, prefix

    def names(self):
        e = self._e()
        return [pi.name for pi in param_spec(e.cfg) if pi.group == self.group and pi.name.startswith(self.prefix)]

    def __iter__(self):
        e = self._e()
        return iter([e.tensor(n) for n in self.names()])

    def __len__(self):
        return len(self.names())


class _SubModel:
    def __init__(self, engine_getter, group, prefix):
        self.trainable_variables = _Variables(engine_getter, group, prefix)


class WorldModel:
    def __init__(self, *, model_dimension="XS", action_space, batch_length_T=64, encoder, decoder,
                 num_gru_units=None, symlog_obs=True, seed=0, compute_mode=0):
        self.model_dimension = model_dimension
        self.batch_length_T = batch_length_T
        self.symlog_obs = symlog_obs
        self.action_space = action_space
        self.encoder, self.decoder = encoder, decoder
        self.num_gru_units_override = num_gru_units
        self.seed = seed
        self.compute_mode = compute_mode
        self._engine = None
        self._engine_args = dict(B=None, H=15, world_size=1, rank=0, use_curiosity=False, intrinsic_rewards_scale=0.1)
        self._noise_step = 0
        self._explicit_noise = False
        self.noise_seed = 0x5EED + seed
        get = lambda: self.engine
        self.optimizer = _OptimizerInfo(get, "world_model", 1e-4, 1e-8)  # world_model.py:124
        self.trainable_variables = _Variables(get, "world_model")
        self.sequence_model = _SubModel(get, "world_model", "seq.")
        self.dynamics_predictor = _SubModel(get, "world_model", "dyn.")
        self.reward_predictor = _SubModel(get, "world_model", "rew.")
        self.continue_predictor = _SubModel(get, "world_model", "cont.")
        self.posterior_mlp = _SubModel(get, "world_model", "post.mlp")
        self.posterior_representation_layer = _SubModel(get, "world_model", "post.repr")

    # ------------------------------------------------------------------ engine management
    def _make_config(self, B, H):
        discrete, A = action_space_info(self.action_space)
        image = bool(getattr(self.decoder, "image", False))
        return EngineConfig.make(
            self.model_dimension, A=A, discrete=discrete, image_obs=image,
            obs_dim=getattr(self.decoder, "obs_dim", 4), img_c=getattr(self.decoder, "img_c", 3),
            B=B, T=self.batch_length_T, H=H, symlog_obs=self.symlog_obs, num_gru_units=self.num_gru_units_override,
            use_curiosity=self._engine_args["use_curiosity"],
            intrinsic_rewards_scale=self._engine_args["intrinsic_rewards_scale"])

    def _configure(self, B=None, H=None, world_size=None, rank=None, use_curiosity=None, intrinsic_rewards_scale=None):
        a = self._engine_args
        for k, v in (("B", B), ("H", H), ("world_size", world_size), ("rank", rank), ("use_curiosity", use_curiosity),
                     ("intrinsic_rewards_scale", intrinsic_rewards_scale)):
            if v is not None:
                if self._engine is not None and a[k] != v:
                    raise RuntimeError(f"engine already built with {k}={a[k]}, cannot change to {v}")
                a[k] = v

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            a = self._engine_args
            if a["B"] is None:
                raise RuntimeError("batch size unknown: call forward_train / train_world_model_one_step first "
                                   "or construct a DreamerModel(batch_size_B=...)")
            cfg = self._make_config(a["B"], a["H"])
            self._engine = Engine(cfg, compute_mode=self.compute_mode, world_size=a["world_size"], rank=a["rank"])
            self._engine.load_params(init_params(cfg, seed=self.seed, mode="keras"))
        return self._engine

    def get_weights(self):
        """`model.get_weights()` as saved by run_experiment.py:705 (order: utils/checkpoint.py)."""
        from ..utils import checkpoint
        return checkpoint.get_weights(self.engine, "world_model")

    def set_weights(self, weights):
        from ..utils import checkpoint
        checkpoint.set_weights(self.engine, "world_model", weights)

    @property
    def num_gru_units(self):
        return self._make_config(1, 1).G

    @property
    def initial_h(self):
        return self.engine.tensor("initial_h")

    def set_noise(self, u_post=None, u_dyn=None, act_noise=None):
        """Supply the sampling noise explicitly (parity runs); otherwise it is drawn on the device."""
        self._explicit_noise = True
        self.engine.set_noise(u_post, u_dyn, act_noise)

    def _draw_noise(self, which):
        if not self._explicit_noise:
            self._noise_step += 1
            self.engine.fill_noise(self.noise_seed, self._noise_step, which)

    # ------------------------------------------------------------------ reference API
    def get_initial_state(self):
        """world_model.py:127-134 -> {"h": [1,G], "z": [1,Zc,Zk]}"""
        e = self.engine
        e.initial_state()
        cfg = e.cfg
        return {"h": e.tensor("wm.h0").reshape(1, cfg.G).clone(), "z": e.tensor("wm.z0").reshape(1, cfg.Zc, cfg.Zk).clone()}

    def forward_train(self, observations, actions, is_first, training=None):
        """world_model.py:183-326. Returns views of engine-owned buffers (valid until the next call)."""
        B = int(is_first.shape[0])
        self._configure(B=B)
        e = self.engine
        e.set_sample({"obs": observations, "actions": actions, "is_first": is_first})
        self._draw_noise(1)
        e.observe_forward()
        return self._forward_outs()

    def _forward_outs(self):
        e = self.engine
        cfg = e.cfg
        hz = e.tensor("wm.hz")
        obs = e.tensor("wm.obs_symlog") if cfg.symlog_obs else e.tensor("in.obs")
        cont_logits = e.tensor("wm.continue_logits")
        dist = _Bernoulli(cont_logits)
        return {
            "sampled_obs_symlog_BxT": obs if cfg.symlog_obs else obs.reshape((cfg.N,) + tuple(obs.shape[2:])),
            "obs_distribution_BxT": _Normal(e.tensor("wm.obs_loc")),
            "reward_logits_BxT": e.tensor("wm.reward_logits"),
            "rewards_BxT": e.tensor("wm.rewards"),
            "continue_distribution_BxT": dist,
            "continues_BxT": dist.mode(),
            "z_posterior_probs_BxT": e.tensor("wm.z_posterior_probs"),
            "z_prior_probs_BxT": e.tensor("wm.z_prior_probs"),
            "h_states_BxT": hz[:, :cfg.G],
            "z_states_BxT": hz[:, cfg.G:].unflatten(1, (cfg.Zc, cfg.Zk)),
        }

    def _run_inference(self, previous_states, observations, is_first):
        """One acting step on the engine for rows = observations.shape[0] <= B parallel envs."""
        e = self.engine
        cfg = e.cfg
        rows = int(torch.as_tensor(is_first).reshape(-1).shape[0])
        if rows > cfg.B:
            raise ValueError(f"forward_inference supports up to B={cfg.B} rows, got {rows}")

        def put(name, value, width):
            dst = e.tensor(name).reshape(cfg.B, -1)[:rows]
            src = value if torch.is_tensor(value) else torch.from_numpy(np.ascontiguousarray(value))
            dst.copy_(src.reshape(rows, width).to(torch.float32), non_blocking=True)

        put("inf.obs", observations, cfg.obs_flat)
        put("inf.is_first", is_first, 1)
        put("inf.prev_h", previous_states["h"], cfg.G)
        put("inf.prev_z", previous_states["z"], cfg.Z)
        a = previous_states.get("a")
        if a is None:
            a = torch.zeros((rows, cfg.A))
        put("inf.prev_a", a, cfg.A)
        if self._explicit_noise:
            e.forward_inference(rows, 0)
        else:
            self._noise_step += 1
            e.forward_inference(rows, self.noise_seed + 7919 * self._noise_step)
        hz = e.tensor("inf.hz")[:rows]
        return rows, hz

    def set_inference_noise(self, u_z, act_noise):
        """Explicit noise for the acting step (parity runs): u_z [rows, Zc], act_noise [rows] or [rows, A]."""
        self._explicit_noise = True
        e = self.engine
        rows = int(np.asarray(u_z).shape[0])
        e.tensor("inf.u_z")[:rows].copy_(torch.as_tensor(np.asarray(u_z), dtype=torch.float32))
        an = torch.as_tensor(np.asarray(act_noise), dtype=torch.float32)
        e.tensor("inf.act_noise")[:rows].copy_(an.reshape(e.tensor("inf.act_noise")[:rows].shape))

    def forward_inference(self, previous_states, observations, is_first, training=None):
        """world_model.py:141-180: masks the previous states with the learned initial state where is_first, runs one
        sequence-model step and draws the posterior z from the observation. Returns {"h": [rows,G], "z": [rows,Zc,Zk]}."""
        cfg = self.engine.cfg
        rows, hz = self._run_inference(previous_states, observations, is_first)
        return {"h": hz[:, :cfg.G].clone(), "z": hz[:, cfg.G:].unflatten(1, (cfg.Zc, cfg.Zk)).clone()}

    __call__ = forward_inference
