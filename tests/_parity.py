"""Shared helper of the GPU parity tests: runs the oracle (torch CPU, fp64 by default) and the CUDA
engine on the same seeded inputs / weights / sampling noise and reports per-tensor deviations."""
import numpy as np
import torch

from dreamer_v3_b200.engine import Engine
from dreamer_v3_b200.spec import EngineConfig, init_params
from oracle.dreamer_oracle import Oracle, make_synthetic_sample


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def to_np(x):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class FixtureOracle:
    """Replays a `tests/golden/ref_*.npz` fixture (outputs of the reference's own code, see
    tests/golden/make_reference_golden.py) through the result-dict layout of `Oracle`, so that `ParityRun.step`
    compares the CUDA engine against reference-generated vectors with the same code that compares it against the
    live oracle."""

    def __init__(self, cfg, z):
        self.cfg, self.z, self.step_no = cfg, z, 0
        grp = lambda pre: {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}  # noqa: E731
        self._grp = grp
        self.params0 = {k: v.astype(np.float32) for k, v in grp("param0/").items()}
        self.sample = grp("sample/")
        self.params = dict(self.params0)
        self.last_noise = {}

    group_names = Oracle.group_names

    @property
    def P(self):
        return self.params

    def numpy_params(self):
        return {k: np.asarray(v, dtype=np.float32) for k, v in self.params.items()}

    def train_world_model_one_step(self, sample, rng=None, margin=0.0):
        import oracle.numerics as nx
        cfg = self.cfg
        self.step_no += 1
        g = self.g = self._grp(f"s{self.step_no}/")
        T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))  # noqa: E731
        self.last_noise = {"u_post": g["noise/u_post"], "u_dyn": g["noise/u_dyn"],
                           ("u_act" if cfg.discrete else "eps_act"): g["noise/act"]}
        zi = g["wm/z_indices_BxT"]
        fwd = {
            "h_states_BxT": T(g["wm/h_states_BxT"]), "z_indices_BxT": zi,
            "z_states_BxT": T(np.eye(cfg.Zk)[zi]),
            "z_posterior_probs_BxT": T(g["wm/z_posterior_probs_BxT"]), "z_prior_probs_BxT": T(g["wm/z_prior_probs_BxT"]),
            "reward_logits_BxT": T(g["wm/reward_logits_BxT"]), "rewards_BxT": T(g["wm/rewards_BxT"]),
            "continue_distribution_BxT.logits": T(g["wm/continue_logits"]),
            "obs_distribution_BxT.loc": T(g["wm/obs_loc"]),
        }
        res = {k[3:]: T(v) for k, v in g.items() if k.startswith("wm/WORLD_MODEL_")}
        res["WORLD_MODEL_forward_train_outs"] = fwd
        # integer bucket index of the symlog'd replay rewards (two_hot.py:46-49; float32 arithmetic)
        res["reward_two_hot_k"] = nx.two_hot_indices(nx.symlog(self.sample["rewards"].reshape(-1)))[0]
        names = self.group_names("world_model")
        res["grads"] = {n: T(g["grad/" + n]) for n in names}
        res["grad_global_norm"] = float(np.sqrt(sum((g["grad/" + n].astype(np.float64) ** 2).sum() for n in names)))
        for n in names:
            self.params[n] = g["param/" + n]
        return res

    def train_actor_and_critic_one_step(self, h, z, is_terminated, H=15, rng=None, margin=0.0):
        cfg, g = self.cfg, self.g
        T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))  # noqa: E731
        dream = {k[6:]: T(v) for k, v in g.items() if k.startswith("dream/") and "indices" not in k and "ints" not in k}
        dream["z_indices_t1_to_H_B"] = g["dream/z_indices_t1_to_H_B"]
        if cfg.discrete:
            dream["actions_ints_dreamed_t0_to_H_B"] = g["dream/actions_ints_dreamed_t0_to_H_B"]
        ac = {k[3:]: T(v) for k, v in g.items() if k.startswith("ac/")}
        ac["dream_data"] = dream
        groups = ("actor", "critic") + (("disagree",) if cfg.num_curiosity_nets > 0 else ())
        names = [n for grp in groups for n in self.group_names(grp)]
        ac["grads"] = {n: T(g["grad/" + n]) for n in names}
        for grp in groups:
            ac[grp + "_grad_global_norm"] = float(np.sqrt(sum(
                (g["grad/" + n].astype(np.float64) ** 2).sum() for n in self.group_names(grp))))
        for n in names + self.group_names("critic_ema"):
            self.params[n] = g["param/" + n]
        return ac


class ParityRun:
    """One oracle-vs-engine comparison of a full train update (world model + actor/critic)."""

    def __init__(self, cfg: EngineConfig, seed=0, init_mode="random", oracle_dtype=torch.float64, compute_mode=0,
                 margin=1e-3, steps=1, sync_grads=True, sample_fn=None, fixture=None, forced=False):
        self.cfg = cfg
        self.report = {}      # name -> rel err (float tensors)
        self.exact = {}       # name -> number of mismatching integers
        if fixture is not None:  # replay of reference-generated golden vectors instead of a live oracle
            self.oracle = FixtureOracle(cfg, fixture)
            self.params0, self.sample = self.oracle.params0, self.oracle.sample
        else:
            self.params0 = init_params(cfg, seed=seed + 1, mode=init_mode)
            self.sample = make_synthetic_sample(cfg, seed=seed + 2)
            if sample_fn is not None:
                sample_fn(self.sample)
            self.oracle = Oracle(cfg, self.params0, dtype=oracle_dtype)
        self.rng = np.random.default_rng(seed + 3)
        self.engine = Engine(cfg, compute_mode=compute_mode)
        self.engine.load_params(self.params0)
        self.margin = margin
        self.sync_grads = sync_grads
        # forced = True (reduced-precision device modes): the device run goes first on plain uniforms, the oracle is then
        # given the indices the device drew (Oracle._draw, forced mode) and records how far each uniform lies outside the
        # oracle's own CDF bucket; every tensor stays comparable and the flips are bounded by the caller instead of waived
        self.forced = forced
        self.forced_violation, self.forced_flips = {}, {}
        for _ in range(steps):
            self.step()

    def cmp(self, name, got, want):
        self.report[name] = max(self.report.get(name, 0.0), rel_err(to_np(got), to_np(want)))

    def cmp_scalar(self, name, got, want, terms):
        got, want = float(to_np(got)), float(to_np(want))
        scale = max(abs(want), float(np.abs(np.asarray(terms, dtype=np.float64)).mean())) + 1e-12
        self.report[name] = max(self.report.get(name, 0.0), abs(got - want) / scale)

    def cmp_delta(self, name, new, old, want_new, want_old):
        """Adam update parity. Both sides hold fp32 weights, so a delta is only resolved to ~1 ulp of the
        weight; that representational floor is subtracted before comparing against the tolerance."""
        new, old, want_new, want_old = (to_np(x).astype(np.float64) for x in (new, old, want_new, want_old))
        d_got, d_want = new - old, want_new - want_old
        scale = np.abs(d_want).max() + 1e-30
        floor = 2.0 * np.finfo(np.float32).eps * np.abs(want_new).max()
        err = max(0.0, np.abs(d_got - d_want).max() - floor) / scale
        self.report[name] = max(self.report.get(name, 0.0), float(err))

    def cmp_exact(self, name, got, want):
        got, want = to_np(got).astype(np.int64), to_np(want).astype(np.int64)
        assert got.shape == want.shape, (name, got.shape, want.shape)
        self.exact[name] = self.exact.get(name, 0) + int((got != want).sum())

    def step(self):
        cfg, o, e = self.cfg, self.oracle, self.engine
        B, T, N, H, G = cfg.B, cfg.T, cfg.N, cfg.H, cfg.G
        self.prev_params = o.numpy_params()
        # ---------------- world model ----------------
        if self.forced:
            u_post = self.rng.random((T, B, cfg.Zc), dtype=np.float32)
            e.set_sample(self.sample)
            e.set_noise(u_post=u_post)
            e.observe_forward()
            zi = e.tensor("wm.z_indices").cpu().numpy().reshape(B, T, cfg.Zc)
            o.set_forced({"u_post": [zi[:, tt] for tt in range(T)]})
            res = o.train_world_model_one_step(self.sample, u_post=u_post)
            self._collect_forced()
        else:
            res = o.train_world_model_one_step(self.sample, rng=self.rng, margin=self.margin)
            u_post = o.last_noise["u_post"]
        fwd = res["WORLD_MODEL_forward_train_outs"]
        if not self.forced:  # (forced: the device pass above is the one compared - split-K atomics make a rerun differ in the last ulp)
            e.set_sample(self.sample)
            e.set_noise(u_post=u_post)
            e.observe_forward()
        t = e.tensor
        if "encoder_out_TB" in fwd:  # not part of forward_train's result dict (absent from reference fixtures)
            self.cmp("wm.enc_out", t("wm.enc_out"), fwd["encoder_out_TB"].transpose(0, 1).reshape(N, -1))
        self.cmp("wm.h_states", t("wm.hz")[:, :G], fwd["h_states_BxT"])
        self.cmp_exact("wm.z_indices", t("wm.z_indices"), fwd["z_indices_BxT"])
        self.cmp("wm.z_states", t("wm.hz")[:, G:], fwd["z_states_BxT"].reshape(N, -1))
        self.cmp("wm.z_posterior_probs", t("wm.z_posterior_probs"), fwd["z_posterior_probs_BxT"])
        self.cmp("wm.z_prior_probs", t("wm.z_prior_probs"), fwd["z_prior_probs_BxT"])
        self.cmp("wm.reward_logits", t("wm.reward_logits"), fwd["reward_logits_BxT"])
        self.cmp("wm.rewards", t("wm.rewards"), fwd["rewards_BxT"])
        self.cmp("wm.continue_logits", t("wm.continue_logits"), fwd["continue_distribution_BxT.logits"])
        self.cmp("wm.obs_loc", t("wm.obs_loc"), fwd["obs_distribution_BxT.loc"])
        e.wm_loss_backward()
        for k_e, k_o in (("wm.L_decoder_B_T", "WORLD_MODEL_L_decoder_B_T"), ("wm.L_reward_B_T", "WORLD_MODEL_L_reward_B_T"),
                         ("wm.L_continue_B_T", "WORLD_MODEL_L_continue_B_T"), ("wm.L_dynamics_B_T", "WORLD_MODEL_L_dynamics_B_T"),
                         ("wm.L_representation_B_T", "WORLD_MODEL_L_representation_B_T"), ("wm.L_total_B_T", "WORLD_MODEL_L_total_B_T")):
            self.cmp(k_e, t(k_e), res[k_o])
        self.cmp("wm.L_total", t("wm.scalars")[6], res["WORLD_MODEL_L_total"])
        self.cmp_exact("wm.reward_two_hot_k", t("wm.reward_two_hot_k"), res["reward_two_hot_k"])
        for n, g in res["grads"].items():
            self.cmp("grad." + n, t("grad." + n), g)
        before = {n: t(n).detach().cpu().numpy().copy() for n in o.group_names("world_model")}
        if self.sync_grads:  # keep both trajectories in lock-step; the gradients were compared above
            for n, g in res["grads"].items():
                e.write("grad." + n, g.to(torch.float32))
        e.optimizer_step("world_model", 1000.0, 1e-4, 1e-8)
        self.cmp("wm.grad_global_norm", t("opt.world_model.stats")[0], res["grad_global_norm"])
        op = o.numpy_params()
        for n in o.group_names("world_model"):
            self.cmp("param." + n, t(n), op[n])
            if self.sync_grads:
                self.cmp_delta("adam_delta." + n, t(n), before[n], op[n], self.prev_params[n])
        # ---------------- actor / critic ----------------
        if self.forced:
            u_dyn = self.rng.random((H, N, cfg.Zc), dtype=np.float32)
            act_noise = (self.rng.random((H + 1, N), dtype=np.float32) if cfg.discrete
                         else self.rng.standard_normal((H + 1, N, cfg.A)).astype(np.float32))
            e.set_noise(u_dyn=u_dyn, act_noise=act_noise)
            # Dream from the engine's own observe buffers, the way a train update does (that is the entry the persistent rollout
            # kernels take: they need the class indices of the start states, wm.z_indices), with the oracle's h written over the
            # engine's so that both dream from identical start states (the z indices are already identical: forced draws).
            e.write("wm.hz", torch.cat([fwd["h_states_BxT"].reshape(N, -1), fwd["z_states_BxT"].reshape(N, -1)], dim=1).to(torch.float32))
            e.dream_forward(start_from_observe=True)
            zi = e.tensor("dream.z_indices").cpu().numpy().reshape(H, N, cfg.Zc)
            forced = {"u_dyn": list(zi), "continues": e.tensor("dream.continues").cpu().numpy().reshape(-1)}
            if cfg.discrete:
                forced["u_act"] = list(e.tensor("dream.action_indices").cpu().numpy().reshape(H + 1, N))
            o.set_forced(forced)
            ac = o.train_actor_and_critic_one_step(fwd["h_states_BxT"], fwd["z_states_BxT"], self.sample["is_terminated"],
                                                   H=H, u_dyn=u_dyn, act_noise=act_noise)
            self._collect_forced()
            o.set_forced(None)
        else:
            ac = o.train_actor_and_critic_one_step(fwd["h_states_BxT"], fwd["z_states_BxT"], self.sample["is_terminated"],
                                                   H=H, rng=self.rng, margin=self.margin)
            u_dyn, act_noise = o.last_noise["u_dyn"], o.last_noise["u_act" if cfg.discrete else "eps_act"]
        dream = ac["dream_data"]
        if not self.forced:
            e.set_noise(u_dyn=u_dyn, act_noise=act_noise)
            # the oracle dreams from its own (pre-update) h/z; feed exactly those to the engine
            e.write("in.dream_h0", fwd["h_states_BxT"].to(torch.float32))
            e.write("in.dream_z0", fwd["z_states_BxT"].reshape(N, -1).to(torch.float32))
            e.dream_forward(start_from_observe=False)
        dhz = t("dream.hz")
        self.cmp("dream.h_states", dhz[..., :G], dream["h_states_t0_to_H_B"])
        self.cmp_exact("dream.z_indices", t("dream.z_indices"), dream["z_indices_t1_to_H_B"])
        self.cmp("dream.actions", t("dream.actions"), dream["actions_dreamed_t0_to_H_B"])
        if cfg.discrete:
            self.cmp_exact("dream.action_indices", t("dream.action_indices"), dream["actions_ints_dreamed_t0_to_H_B"])
        self.cmp("dream.rewards", t("dream.rewards"), dream["rewards_dreamed_t0_to_H_B"])
        self.cmp_exact("dream.continue_flags", t("dream.continues"), dream["continues_dreamed_t0_to_H_B"])
        self.cmp("dream.continues", t("dream.continues"), dream["continues_dreamed_t0_to_H_B"])
        self.cmp("dream.values", t("dream.values"), dream["values_dreamed_t0_to_H_B"])
        self.cmp("dream.value_logits", t("dream.value_logits"), dream["values_symlog_dreamed_logits_t0_to_HxB"])
        self.cmp("dream.values_symlog_ema", t("dream.values_symlog_ema"), dream["v_symlog_dreamed_ema_t0_to_H_B"])
        self.cmp("dream.loss_weights", t("dream.loss_weights"), dream["dream_loss_weights_t0_to_H_B"])
        cur = cfg.num_curiosity_nets > 0
        if cur:  # Plan2Explore disagree nets (models/disagree_networks.py:47-95)
            zp = dream["z_predicted_probs_N_HxB"]
            zp = np.stack([to_np(p) for p in zp], 0) if isinstance(zp, (list, tuple)) else to_np(zp)
            self.cmp("dream.z_predicted_probs", t("disagree.z_predicted_probs"), zp.reshape(cfg.num_curiosity_nets, (H + 1) * N, cfg.Zc, cfg.Zk))
            self.cmp("dream.rewards_intrinsic", t("dream.rewards_intrinsic")[1:], dream["rewards_intrinsic_t1_to_H_B"])
        e.ac_loss_backward(3)
        self.cmp("ac.value_targets", t("ac.value_targets"), ac["VALUE_TARGETS_H_B"])
        for k in ("CRITIC_L_total_H_B", "CRITIC_L_neg_logp_of_value_targets_H_B", "CRITIC_L_slow_critic_regularization_H_B",
                  "VALUE_TARGETS_symlog_H_B", "ACTOR_L_total_H_B", "ACTOR_logp_actions_dreamed_H_B",
                  "ACTOR_scaled_value_targets_H_B", "ACTOR_action_entropy_H_B", "ACTOR_L_neglogp_reinforce_term_H_B",
                  "ACTOR_L_neg_entropy_term_H_B"):
            self.cmp("ac." + k, t("ac." + k), ac[k])
        self.cmp("ac.CRITIC_L_total", t("ac.scalars")[0], ac["CRITIC_L_total"])
        # the actor loss is a mean of signed terms (centred advantages): its error is measured against the mean magnitude
        # of its terms, not against the cancelled sum
        self.cmp_scalar("ac.ACTOR_L_total", t("ac.scalars")[3], ac["ACTOR_L_total"], to_np(ac["ACTOR_L_total_H_B"]))
        self.cmp("ac.pct_ema", t("ac.pct_ema"), np.array([ac["ACTOR_value_targets_pct5_ema"], ac["ACTOR_value_targets_pct95_ema"]]))
        if cur:  # losses/disagree_loss.py:11-41, train_one_step.py:201-212
            e.disagree_loss_backward()
            self.cmp("ac.DISAGREE_L_total", t("disagree.scalars")[0], ac["DISAGREE_L_total"])
            self.cmp_scalar("ac.DISAGREE_intrinsic_rewards", t("disagree.scalars")[1], ac["DISAGREE_intrinsic_rewards"],
                            to_np(dream["rewards_intrinsic_t1_to_H_B"]))
        for n, g in ac["grads"].items():
            self.cmp("grad." + n, t("grad." + n), g)
        before = {n: t(n).detach().cpu().numpy().copy()
                  for g_ in ("actor", "critic") + (("disagree",) if cur else ()) for n in o.group_names(g_)}
        if self.sync_grads:
            for n, g in ac["grads"].items():
                e.write("grad." + n, g.to(torch.float32))
        e.optimizer_step("actor", 100.0, 3e-5, 1e-5)
        e.optimizer_step("critic", 100.0, 3e-5, 1e-5, 0.98)
        self.cmp("ac.actor_grad_global_norm", t("opt.actor.stats")[0], ac["actor_grad_global_norm"])
        self.cmp("ac.critic_grad_global_norm", t("opt.critic.stats")[0], ac["critic_grad_global_norm"])
        if cur:
            e.optimizer_step("disagree", 100.0, 1e-4, 1e-5)  # run_experiment.py:252, disagree_networks.py:43
            self.cmp("ac.disagree_grad_global_norm", t("opt.disagree.stats")[0], ac["disagree_grad_global_norm"])
        op = o.numpy_params()
        for grp in ("actor", "critic", "critic_ema") + (("disagree",) if cur else ()):
            for n in o.group_names(grp):
                self.cmp("param." + n, t(n), op[n])
                if self.sync_grads and grp != "critic_ema":
                    self.cmp_delta("adam_delta." + n, t(n), before[n], op[n], self.prev_params[n])
        torch.cuda.synchronize()

    def _collect_forced(self):
        o = self.oracle
        for k, v in o.forced_violation.items():
            self.forced_violation[k] = max(self.forced_violation.get(k, 0.0), v)
        for k, v in o.forced_flips.items():
            self.forced_flips[k] = self.forced_flips.get(k, 0) + v

    def worst(self, k=12):
        return sorted(self.report.items(), key=lambda kv: -kv[1])[:k]

    def close(self):
        self.engine.close()
