"""ORACLE (test infrastructure, not product code) -- PyTorch-CPU restatement of the DreamerV3
training hot path of sven1977/dreamer_v3 (TensorFlow 2 / Keras / tfp), with autograd
providing the reference gradients.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; `dreamer_v3_b200` (the product) never does and fails loudly
when its CUDA library is missing.

PARITY PIN: the reference ships no tests or fixtures for this path, and TensorFlow / Keras / tfp
(versions not pinned by the reference; ~TF 2.10-2.12, tfp 0.18-0.20 inferred) cannot be installed in
this image.  The reference's OWN Python files are therefore executed, unmodified, over a TF-API shim
(tests/golden/tfshim/, PyTorch-CPU backed) by tests/golden/make_reference_golden.py, which calls
`train_world_model_one_step` / `train_actor_and_critic_one_step` as run_experiment.py does and commits
the outputs (tests/golden/ref_*.npz).  tests/test_reference_golden.py holds this oracle to those
vectors: 1e-9 relative in float64, every sampled index exact, over two consecutive updates
(forward tensors, losses, dream, lambda-returns, every gradient, Adam / EMA results).
That pins the reference's COMPOSITION (masking, wrap-around action index, concatenation orders,
stop-gradients, loss weighting, clipping, optimizer / EMA order).  Still UNPINNED: TensorFlow's own
arithmetic of the primitives (Dense, LayerNormalization eps, GRU gate order / reset_after, conv
padding, tfp formulas, Keras Adam), which shim and oracle both restate from the documented defaults
(SURVEY.md appendix A); `utils/two_hot.py:23-37` additionally pins `oracle/numerics.py:two_hot`.

Deliberate, documented deviations (they make results reproducible, not different):
  * sampling noise is an explicit input (uniforms for categoricals, normals for Box actions)
    and categorical draws use the float64 inverse CDF of `oracle/numerics.py:sample_categorical`;
  * RNG draws whose results the reference discards are not made.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import numerics as nx

NUM_BUCKETS = 255
LN_EPS = 1e-3  # Keras LayerNormalization default epsilon


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def symlog(x):  # utils/symlog.py:9-12
    return torch.sign(x) * torch.log(torch.abs(x) + 1.0)


def inverse_symlog(y):  # utils/symlog.py:15-28
    return torch.sign(y) * (torch.exp(torch.abs(y)) - 1.0)


class Oracle:
    """Holds parameters + optimizer state; methods mirror the reference call graph."""

    def __init__(self, cfg, params, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.P = {k: _t(v, dtype).clone().requires_grad_(not k.startswith("critic_ema."))
                  for k, v in params.items()}
        self.adam = {g: {"t": 0, "m": {}, "v": {}} for g in ("world_model", "actor", "critic", "disagree")}
        self.ema_pct5 = float("nan")  # actor_network.py:31-36
        self.ema_pct95 = float("nan")
        self.buckets = torch.linspace(-20.0, 20.0, NUM_BUCKETS, dtype=dtype)
        self.last_noise = {}
        # float type of the symlog -> two-hot target arithmetic: float32 like the reference (the bucket-index
        # contract); np.float64 only to compare against the float64 reference fixtures at 1e-9
        self.th_float = np.float32

    # ------------------------------------------------------------------ groups
    def set_forced(self, forced):
        """See `_draw`. forced: dict noise-key -> per-call index arrays (+ "continues": [(H+1)*N] flags), or None."""
        self.forced = forced
        self._forced_calls, self.forced_violation, self.forced_flips = {}, {}, {}

    def group_names(self, group):
        def grp(n):
            if n.startswith("actor."):
                return "actor"
            if n.startswith("critic_ema."):
                return "critic_ema"
            if n.startswith("critic."):
                return "critic"
            if n.startswith("disagree."):
                return "disagree"
            return "world_model"
        return [n for n in self.P if grp(n) == group]

    def numpy_params(self):
        return {k: v.detach().numpy().astype(np.float32) for k, v in self.P.items()}

    # ------------------------------------------------------------------ layers
    def mlp(self, prefix, x, L):
        """components/mlp.py:69-84: L x (Dense no-bias -> LayerNorm(eps 1e-3) -> SiLU)."""
        P = self.P
        for l in range(L):
            w = P[f"{prefix}.mlp{l}.w"]
            x = x @ w
            x = F.layer_norm(x, (w.shape[1],), P[f"{prefix}.mlp{l}.g"], P[f"{prefix}.mlp{l}.b"], LN_EPS)
            x = F.silu(x)
        return x

    def dense(self, prefix, x):
        return x @ self.P[f"{prefix}.w"] + self.P[f"{prefix}.bias"]

    def _draw(self, probs, u, key, rng, margin):
        """Categorical draw with explicit (or recorded, margin-safe) uniforms.

        Forced mode (`self.forced[key]` = list of index arrays, one per call): the draw is *given* (the indices a
        reduced-precision device run sampled from the same uniforms) and what is recorded instead is how far each
        uniform lies outside the CDF bucket of its forced index under THIS oracle's probabilities
        (`self.forced_violation[key]`, 0 when the oracle would have drawn the same index). Tests of the bf16 path bound
        that distance, so a flipped draw is checked instead of waived and every downstream tensor stays comparable."""
        p_np = probs.detach().to(torch.float32).numpy()
        forced = getattr(self, "forced", None)
        if forced is not None and key in forced:
            n_call = self._forced_calls.get(key, 0)
            self._forced_calls[key] = n_call + 1
            idx = np.asarray(forced[key][n_call]).astype(np.int64).reshape(p_np.shape[:-1])
            cdf = np.cumsum(p_np.astype(np.float64), axis=-1)
            cdf = cdf / cdf[..., -1:]
            hi = np.take_along_axis(cdf, idx[..., None], -1)[..., 0]
            lo = np.where(idx > 0, np.take_along_axis(cdf, np.maximum(idx - 1, 0)[..., None], -1)[..., 0], 0.0)
            uu = np.asarray(u, dtype=np.float64).reshape(idx.shape)
            viol = np.maximum(np.maximum(lo - uu, uu - hi), 0.0)
            viol = np.where(idx == p_np.shape[-1] - 1, np.maximum(lo - uu, 0.0), viol)  # last bucket is open above
            self.forced_violation[key] = max(self.forced_violation.get(key, 0.0), float(viol.max()) if viol.size else 0.0)
            self.forced_flips[key] = self.forced_flips.get(key, 0) + int((nx.sample_categorical(p_np, np.asarray(u, dtype=np.float32)) != idx).sum())
            return idx
        if u is None:
            assert rng is not None, f"noise {key!r} missing and no rng given"
            u = rng.random(p_np.shape[:-1], dtype=np.float32)
            if margin > 0:
                for _ in range(100):
                    bad = nx.cdf_margin(p_np, u) < margin
                    if not bad.any():
                        break
                    u[bad] = rng.random(int(bad.sum()), dtype=np.float32)
            self.last_noise.setdefault(key, []).append(u.copy())
        else:
            u = np.asarray(u, dtype=np.float32)
        return nx.sample_categorical(p_np, u)

    def representation(self, prefix, x, u, key=None, rng=None, margin=0.0, sample=True):
        """components/representation_layer.py:48-113."""
        cfg = self.cfg
        logits = self.dense(prefix, x).reshape(-1, cfg.Zc, cfg.Zk)
        probs = torch.softmax(logits, dim=-1)
        probs = 0.99 * probs + 0.01 * (1.0 / cfg.Zk)  # :87
        if not sample:
            return None, probs, None
        idx = self._draw(probs, u, key, rng, margin)
        onehot = F.one_hot(torch.as_tensor(idx, dtype=torch.long), cfg.Zk).to(self.dtype)
        z = onehot + probs - probs.detach()  # :108-110 straight-through
        return z, probs, idx

    def sequence_model(self, z, a, h):
        """components/sequence_model.py:56-75 (+ Keras GRU reset_after=True, gate order z,r,h)."""
        P, G = self.P, self.cfg.G
        x = torch.cat([z.reshape(z.shape[0], -1), a], dim=-1)
        u = self.mlp("seq.pre", x, 1)
        gx = u @ P["seq.gru.kernel"] + P["seq.gru.bias"][0]
        gh = h @ P["seq.gru.recurrent"] + P["seq.gru.bias"][1]
        xz, xr, xh = gx[:, :G], gx[:, G:2 * G], gx[:, 2 * G:]
        hz, hr, hh = gh[:, :G], gh[:, G:2 * G], gh[:, 2 * G:]
        zg = torch.sigmoid(xz + hz)
        rg = torch.sigmoid(xr + hr)
        c = torch.tanh(xh + rg * hh)
        return zg * h + (1.0 - zg) * c

    def encoder(self, x):
        cfg, P = self.cfg, self.P
        if not cfg.image_obs:
            return self.mlp("enc", x, cfg.L)  # run_experiment.py:190-193
        # components/cnn_atari.py:78-86; x is NHWC
        out = x
        for l in range(4):
            w = P[f"enc.conv{l}.w"]  # [kh,kw,cin,cout]
            y = F.conv2d(out.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=2, padding=1)
            y = y.permute(0, 2, 3, 1)
            y = F.layer_norm(y, (w.shape[3],), P[f"enc.conv{l}.g"], P[f"enc.conv{l}.b"], LN_EPS)
            out = F.silu(y)
        assert out.shape[1] == 4 and out.shape[2] == 4
        return out.reshape(out.shape[0], -1)

    def decoder(self, hz):
        cfg, P = self.cfg, self.P
        if not cfg.image_obs:
            return self.dense("dec.out", self.mlp("dec", hz, cfg.L))  # vector_decoder.py:34-60
        # components/conv_transpose_atari.py:94-141
        out = self.dense("dec.dense", hz).reshape(-1, 4, 4, 8 * cfg.cnn_mult)
        for l in range(3):
            w = P[f"dec.convT{l}.w"]  # [kh,kw,out,in]
            y = F.conv_transpose2d(out.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), stride=2, padding=1)
            y = y.permute(0, 2, 3, 1)
            y = F.layer_norm(y, (w.shape[2],), P[f"dec.convT{l}.g"], P[f"dec.convT{l}.b"], LN_EPS)
            out = F.silu(y)
        w = P["dec.out.w"]
        y = F.conv_transpose2d(out.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), bias=P["dec.out.bias"],
                               stride=2, padding=1).permute(0, 2, 3, 1)
        y = y + 0.5  # :121
        return y.reshape(y.shape[0], -1)

    def bucket_head(self, prefix, x):
        """components/reward_predictor_layer.py:68-112 -> (expected value, logits)."""
        logits = self.dense(prefix, x)
        probs = torch.softmax(logits, dim=-1)
        return (probs * self.buckets).sum(-1), logits

    def reward_predictor(self, hz):
        return self.bucket_head("rew.head", self.mlp("rew", hz, self.cfg.L))

    def continue_predictor(self, hz):
        logits = self.dense("cont.out", self.mlp("cont", hz, self.cfg.L)).squeeze(-1)
        return (logits > 0).to(self.dtype), logits  # Bernoulli mode, continue_predictor.py:41-46

    def critic(self, hz, use_ema=False):
        pre = "critic_ema" if use_ema else "critic"
        return self.bucket_head(f"{pre}.head", self.mlp(pre, hz, self.cfg.L))

    def actor(self, hz, noise, key=None, rng=None, margin=0.0):
        """models/actor_network.py:63-118. hz must already be stop-gradient'ed by the caller."""
        cfg = self.cfg
        out = self.dense("actor.out", self.mlp("actor", hz, cfg.L))
        if cfg.discrete:
            probs = torch.softmax(out, dim=-1)
            probs = 0.99 * probs + 0.01 * (1.0 / cfg.A)  # :89
            idx = self._draw(probs, noise, key, rng, margin)
            onehot = F.one_hot(torch.as_tensor(idx, dtype=torch.long), cfg.A).to(self.dtype)
            a = onehot + (probs - probs.detach())
            return a, {"probs": probs, "idx": idx}
        s = self.dense("actor.std_out", self.mlp("actor.std", hz, cfg.L))
        std = (1.0 - 0.1) * torch.sigmoid(s + 2.0) + 0.1  # :105-109
        loc = torch.tanh(out)
        if noise is None:
            noise = rng.standard_normal(tuple(loc.shape)).astype(np.float32)
            self.last_noise.setdefault(key, []).append(noise.copy())
        a = loc + std * _t(noise, self.dtype)  # reparameterised Normal sample, :111-114
        return a, {"loc": loc, "std": std}

    # ------------------------------------------------------------------ world model
    def get_initial_state(self):
        """models/world_model.py:127-134."""
        h = torch.tanh(self.P["initial_h"]).unsqueeze(0)
        _, probs, _ = self.representation("dyn.repr", self.mlp("dyn", h, 1), None, sample=False)
        z = F.one_hot(torch.argmax(probs, dim=-1), self.cfg.Zk).to(self.dtype)
        return {"h": h, "z": z.detach()}

    def forward_inference(self, prev_h, prev_z, prev_a, obs, is_first, u_z=None, act_noise=None, rng=None, margin=0.0):
        """models/world_model.py:141-180 (+ compute_posterior_z :329-342) followed by the actor, i.e.
        DreamerModel.forward_inference (models/dreamer_model.py:113-128). All inputs are [B, ...] (no time axis)."""
        cfg, dt = self.cfg, self.dtype
        prev_h, prev_z, prev_a = _t(prev_h, dt), _t(prev_z, dt).reshape(-1, cfg.Zc, cfg.Zk), _t(prev_a, dt)
        obs, f = _t(obs, dt), _t(is_first, dt).reshape(-1, 1)
        init = self.get_initial_state()
        h_in = prev_h * (1.0 - f) + init["h"] * f
        z_in = prev_z * (1.0 - f.unsqueeze(-1)) + init["z"] * f.unsqueeze(-1)
        a_in = prev_a * (1.0 - f)
        h = self.sequence_model(z_in, a_in, h_in)
        x = symlog(obs) if cfg.symlog_obs else obs
        enc = self.encoder(x)
        p = self.mlp("post", torch.cat([enc, h], dim=-1), 1)
        z, probs, idx = self.representation("post.repr", p, u_z, "inf_u_z", rng, margin)
        hz = torch.cat([h, z.reshape(z.shape[0], -1)], -1)
        a, dist = self.actor(hz.detach(), act_noise, "inf_act_noise", rng, margin)
        return {"h": h.detach(), "z": z.detach(), "a": a.detach(), "z_idx": idx, "z_probs": probs.detach(), "dist": dist}

    def forward_train(self, obs, actions, is_first, u_post=None, rng=None, margin=0.0):
        """models/world_model.py:183-326. obs [B,T,...], actions [B,T,A] (one-hot / float),
        is_first [B,T]; u_post [T,B,Zc] uniforms (or drawn from `rng` and recorded)."""
        cfg, dt = self.cfg, self.dtype
        obs, actions, is_first = _t(obs, dt), _t(actions, dt), _t(is_first, dt)
        B, T = obs.shape[0], obs.shape[1]
        if cfg.symlog_obs:
            obs = symlog(obs)
        obs_f = obs.reshape((B * T,) + tuple(obs.shape[2:]))
        enc = self.encoder(obs_f).reshape(B, T, -1).transpose(0, 1)  # [T,B,E]
        init = self.get_initial_state()
        h0, z0 = init["h"], init["z"]
        act_tm = actions.transpose(0, 1)
        first_tm = is_first.transpose(0, 1)
        h_prev, z_prev = h0.repeat(B, 1), z0.repeat(B, 1, 1)
        hs, zs, posts, priors, zidx = [], [], [], [], []
        self.last_noise.pop("u_post", None)
        for t in range(T):
            f = first_tm[t].unsqueeze(-1)
            h_tm1 = h_prev * (1.0 - f) + h0 * f  # :246-247
            z_tm1 = z_prev * (1.0 - f.unsqueeze(-1)) + z0 * f.unsqueeze(-1)  # :249-250
            a_tm1 = act_tm[t - 1] * (1.0 - f)  # :253 (t=0 wraps to T-1, as in the reference)
            h_t = self.sequence_model(z_tm1, a_tm1, h_tm1)  # :256
            p = self.mlp("post", torch.cat([enc[t], h_t], dim=-1), 1)  # :259-260
            z_t, post, idx = self.representation(
                "post.repr", p, None if u_post is None else u_post[t], "u_post", rng, margin)
            _, prior, _ = self.representation("dyn.repr", self.mlp("dyn", h_t, 1), None, sample=False)  # :272
            hs.append(h_t); zs.append(z_t); posts.append(post); priors.append(prior); zidx.append(idx)
            h_prev, z_prev = h_t, z_t
        fold = lambda lst: torch.stack(lst, dim=1).reshape((B * T,) + tuple(lst[0].shape[1:]))
        h_BxT, z_BxT = fold(hs), fold(zs)
        hz = torch.cat([h_BxT, z_BxT.reshape(B * T, -1)], dim=-1)
        obs_loc = self.decoder(hz)
        rewards, reward_logits = self.reward_predictor(hz)
        continues, cont_logits = self.continue_predictor(hz)
        if "u_post" in self.last_noise:
            self.last_noise["u_post"] = np.stack(self.last_noise["u_post"], 0)
        return {
            "sampled_obs_symlog_BxT": obs_f,
            "obs_distribution_BxT.loc": obs_loc,
            "reward_logits_BxT": reward_logits,
            "rewards_BxT": rewards,
            "continue_distribution_BxT.logits": cont_logits,
            "continues_BxT": continues,
            "z_posterior_probs_BxT": fold(posts),
            "z_prior_probs_BxT": fold(priors),
            "h_states_BxT": h_BxT,
            "z_states_BxT": z_BxT,
            "z_indices_BxT": np.stack(zidx, axis=1).reshape(B * T, cfg.Zc),
            "encoder_out_TB": enc,
        }

    def world_model_losses(self, fwd, rewards, is_terminated, B, T):
        """losses/world_model_losses.py:15-151 + training/train_one_step.py:48-77."""
        dt = self.dtype
        obs = fwd["sampled_obs_symlog_BxT"].reshape(B * T, -1)
        L_dec = ((fwd["obs_distribution_BxT.loc"] - obs) ** 2).sum(-1).reshape(B, T)  # :36
        r_sym = nx.symlog(np.asarray(rewards, dtype=np.float32).reshape(-1), self.th_float)  # :45-47
        th = _t(nx.two_hot(r_sym, ft=self.th_float), dt)  # :49
        logp = fwd["reward_logits_BxT"] - torch.logsumexp(fwd["reward_logits_BxT"], -1, keepdim=True)
        L_rew = -(logp * th).sum(-1).reshape(B, T)  # :57-59
        cont = 1.0 - _t(is_terminated, dt).reshape(-1)
        l = fwd["continue_distribution_BxT.logits"]
        L_con = -(cont * F.logsigmoid(l) + (1.0 - cont) * F.logsigmoid(-l)).reshape(B, T)  # :73-74

        def kl(p, q):  # tfp kl_divergence(Independent(OneHotCategorical(probs=p)), ... q), summed over Zc x Zk
            lp = torch.log_softmax(torch.log(p), -1)
            lq = torch.log_softmax(torch.log(q), -1)
            return (torch.softmax(torch.log(p), -1) * (lp - lq)).sum((-1, -2))

        post, prior = fwd["z_posterior_probs_BxT"], fwd["z_prior_probs_BxT"]
        kl_dyn = kl(post.detach(), prior)
        kl_rep = kl(post, prior.detach())
        one = torch.ones_like(kl_dyn)
        L_dyn = torch.where(kl_dyn > 1.0, kl_dyn, one).reshape(B, T)  # free bits, :125-147
        L_rep = torch.where(kl_rep > 1.0, kl_rep, one).reshape(B, T)
        L_pred = L_dec + L_rew + L_con
        L_tot_BT = 1.0 * L_pred + 0.5 * L_dyn + 0.1 * L_rep  # train_one_step.py:74
        return {
            "WORLD_MODEL_L_decoder_B_T": L_dec, "WORLD_MODEL_L_decoder": L_dec.mean(),
            "WORLD_MODEL_L_reward_B_T": L_rew, "WORLD_MODEL_L_reward": L_rew.mean(),
            "WORLD_MODEL_L_continue_B_T": L_con, "WORLD_MODEL_L_continue": L_con.mean(),
            "WORLD_MODEL_L_prediction_B_T": L_pred, "WORLD_MODEL_L_prediction": L_pred.mean(),
            "WORLD_MODEL_L_dynamics_B_T": L_dyn, "WORLD_MODEL_L_dynamics": L_dyn.mean(),
            "WORLD_MODEL_L_representation_B_T": L_rep, "WORLD_MODEL_L_representation": L_rep.mean(),
            "WORLD_MODEL_L_total_B_T": L_tot_BT, "WORLD_MODEL_L_total": L_tot_BT.mean(),
            "reward_two_hot_k": nx.two_hot_indices(r_sym, ft=self.th_float)[0],
        }

    # ------------------------------------------------------------------ optimizer
    def _apply(self, group, names, grads, clip, lr, eps):
        """tf.clip_by_global_norm + Keras Adam (train_one_step.py:80-86)."""
        st = self.adam[group]
        gn = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
        scale = clip / max(gn, clip)
        st["t"] += 1
        t = st["t"]
        lr_t = lr * math.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        clipped = []
        with torch.no_grad():
            for n, g in zip(names, grads):
                g = g * scale
                clipped.append(g)
                m = st["m"].setdefault(n, torch.zeros_like(g))
                v = st["v"].setdefault(n, torch.zeros_like(g))
                m += (g - m) * (1.0 - 0.9)
                v += (g * g - v) * (1.0 - 0.999)
                self.P[n] -= lr_t * m / (torch.sqrt(v) + eps)
        return gn, clipped

    def train_world_model_one_step(self, sample, u_post=None, grad_clip=1000.0, rng=None, margin=0.0,
                                   apply=True):
        """training/train_one_step.py:17-126. `sample` = dict(obs, actions, rewards, is_first,
        is_terminated) of numpy arrays shaped [B,T,...]."""
        B, T = np.asarray(sample["is_first"]).shape
        fwd = self.forward_train(sample["obs"], sample["actions"], sample["is_first"], u_post, rng, margin)
        res = self.world_model_losses(fwd, sample["rewards"], sample["is_terminated"], B, T)
        names = self.group_names("world_model")
        grads = torch.autograd.grad(res["WORLD_MODEL_L_total"], [self.P[n] for n in names], allow_unused=True)
        grads = [g if g is not None else torch.zeros_like(self.P[n]) for n, g in zip(names, grads)]
        res["grads"] = {n: g.detach().clone() for n, g in zip(names, grads)}
        res["WORLD_MODEL_gradients_maxabs"] = max(float(g.abs().max()) for g in grads)
        if apply:
            gn, clipped = self._apply("world_model", names, grads, grad_clip, 1e-4, 1e-8)  # world_model.py:124
            res["grad_global_norm"] = gn
            res["WORLD_MODEL_gradients_clipped_by_glob_norm_maxabs"] = max(float(g.abs().max()) for g in clipped)
        res["WORLD_MODEL_forward_train_outs"] = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in fwd.items()}
        return res

    # ------------------------------------------------------------------ dream
    def dream_trajectory(self, h, z, start_is_terminated, H, gamma, u_dyn=None, act_noise=None, rng=None,
                         margin=0.0):
        """models/dreamer_model.py:147-303. h [N,G], z [N,Zc,Zk] (one-hot), start_is_terminated [N];
        u_dyn [H,N,Zc]; act_noise = u_act [H+1,N] (Discrete) or eps_act [H+1,N,A] (Box)."""
        cfg, dt = self.cfg, self.dtype
        h, z = _t(h, dt), _t(z, dt)
        N = h.shape[0]
        akey = "u_act" if cfg.discrete else "eps_act"
        for k in ("u_dyn", akey):
            self.last_noise.pop(k, None)

        def act(h, z, i):
            hz = torch.cat([h.detach(), z.detach().reshape(N, -1)], -1)  # :179-180, :200-201
            return self.actor(hz, None if act_noise is None else act_noise[i], akey, rng, margin)

        hs, zs, acts, dists, zidx = [h], [z], [], [], []
        a, d = act(h, z, 0)
        acts.append(a); dists.append(d)
        for i in range(H):
            h = self.sequence_model(z, a, h)  # :191
            zz, _, idx = self.representation(  # :195
                "dyn.repr", self.mlp("dyn", h, 1), None if u_dyn is None else u_dyn[i], "u_dyn", rng, margin)
            z = zz
            a, d = act(h, z, i + 1)  # :199
            hs.append(h); zs.append(z); acts.append(a); dists.append(d); zidx.append(idx)
        h_HB = torch.stack(hs, 0)
        z_HB = torch.stack(zs, 0)
        hz = torch.cat([h_HB.reshape((H + 1) * N, -1), z_HB.reshape((H + 1) * N, -1)], -1)
        a_HB = torch.stack(acts, 0)
        r_sym, _ = self.reward_predictor(hz)
        r_HB = inverse_symlog(r_sym).reshape(H + 1, N)  # :219-224
        c, c_logit = self.continue_predictor(hz)
        forced = getattr(self, "forced", None)
        if forced is not None and "continues" in forced:  # Bernoulli mode 1[logit > 0] given; record |logit| where it differs
            cf = _t(np.asarray(forced["continues"]).reshape(-1), dt).clone()
            cf[:N] = c.reshape(-1)[:N]  # step 0 is replaced by 1 - start_is_terminated below (:252-255), not a prediction
            diff = (cf != c.reshape(-1))
            self.forced_violation["continues"] = float(c_logit.detach().reshape(-1)[diff].abs().max()) if bool(diff.any()) else 0.0
            self.forced_flips["continues"] = int(diff.sum())
            c = cf
        c_HB = c.reshape(H + 1, N)
        c_HB = torch.cat([1.0 - _t(start_is_terminated, dt).reshape(1, N), c_HB[1:]], 0)  # :252-255
        w_HB = torch.cumprod(gamma * c_HB, dim=0) / gamma  # :261
        v_sym, v_logits = self.critic(hz)
        v_HB = inverse_symlog(v_sym).reshape(H + 1, N)
        v_ema_sym, _ = self.critic(hz, use_ema=True)
        for k in ("u_dyn", akey):
            if k in self.last_noise:
                self.last_noise[k] = np.stack(self.last_noise[k], 0)
        out = {
            "h_states_t0_to_H_B": h_HB,
            "z_states_prior_t0_to_H_B": z_HB,
            "rewards_dreamed_t0_to_H_B": r_HB,
            "continues_dreamed_t0_to_H_B": c_HB,
            "actions_dreamed_t0_to_H_B": a_HB,
            "actions_dreamed_distributions_t0_to_H_B": dists,
            "values_dreamed_t0_to_H_B": v_HB,
            "values_symlog_dreamed_logits_t0_to_HxB": v_logits,
            "v_symlog_dreamed_ema_t0_to_H_B": v_ema_sym.reshape(H + 1, N),
            "dream_loss_weights_t0_to_H_B": w_HB,
            "z_indices_t1_to_H_B": np.stack(zidx, 0) if zidx else None,
        }
        if getattr(cfg, "num_curiosity_nets", 0) > 0:  # dreamer_model.py:226-240, :296-298
            cur = self.compute_intrinsic_rewards(hz, a_HB.reshape((H + 1) * N, -1))
            out["rewards_intrinsic_t1_to_H_B"] = cur["rewards_intrinsic"].reshape(H + 1, N)[1:]
            out["z_predicted_probs_N_HxB"] = cur["z_predicted_probs_N_HxB"]
        if cfg.discrete:
            out["actions_ints_dreamed_t0_to_H_B"] = torch.argmax(a_HB, -1)
        return out

    # ------------------------------------------------------------------ curiosity (Plan2Explore disagreement)
    def disagree_forward_train(self, hz, a):
        """models/disagree_networks.py:78-95: every net = MLP(L layers) -> RepresentationLayer probs on
        stop_gradient([h, z, a])."""
        cfg = self.cfg
        x = torch.cat([hz, a], -1).detach()
        return [self.representation(f"disagree.net{n}.repr", self.mlp(f"disagree.net{n}", x, cfg.L), None, sample=False)[1]
                for n in range(cfg.num_curiosity_nets)]

    def compute_intrinsic_rewards(self, hz, a):
        """models/disagree_networks.py:47-76: population stddev over the nets of every class probability, mean over the
        Zc*Zk probabilities, minus the mean over all rows (the reference's `#TEST` lines), times the scale."""
        cfg = self.cfg
        probs = self.disagree_forward_train(hz, a)
        p = torch.stack(probs, 0).reshape(len(probs), hz.shape[0], -1)
        std = torch.sqrt(((p - p.mean(0, keepdim=True)) ** 2).mean(0))  # tf.math.reduce_std
        s = std.mean(-1)
        s = s - s.mean()
        return {"rewards_intrinsic": s * cfg.intrinsic_rewards_scale, "z_predicted_probs_N_HxB": probs}

    def disagree_loss(self, dream):
        """losses/disagree_loss.py:11-41 (no loss weights, as in the reference)."""
        z = dream["z_states_prior_t0_to_H_B"].detach()
        Hp1, N = z.shape[0], z.shape[1]
        target = z[1:].reshape(Hp1 - 1, N, -1)
        loss = 0.0
        for p in dream["z_predicted_probs_N_HxB"]:
            pred = p.reshape(Hp1, N, -1)[:-1]
            loss = loss + ((pred - target) ** 2).sum(-1).mean()
        return loss

    def dream_trajectory_with_burn_in(self, start_states, timesteps_burn_in, timesteps_H, observations, actions,
                                      u_burn, u_dyn, act_noise=None, use_sampled_actions_in_dream=False):
        """models/dreamer_model.py:306-413 (evaluation path, run_experiment.py:629-644). start_states {"h","z","a"}
        [B,...]; observations [B, >=burn_in, ...]; actions [B, burn_in (+H), A]; u_burn [burn_in, B, Zc] uniforms of the
        posterior draws, u_dyn [H, B, Zc] of the prior draws, act_noise [H, B(, A)] for the actor samples a_1..a_H
        (unused with use_sampled_actions_in_dream)."""
        cfg, dt = self.cfg, self.dtype
        B = np.asarray(observations).shape[0]
        actions = _t(actions, dt)
        h, z, a = _t(start_states["h"], dt), _t(start_states["z"], dt), _t(start_states["a"], dt)
        dummy = np.full((B,) if cfg.discrete else (B, cfg.A), 0.5, np.float32)  # the actor's sample is not used in burn-in
        for i in range(timesteps_burn_in):
            first = np.full((B,), 1.0 if i == 0 else 0.0, np.float32)  # :326-330
            st = self.forward_inference(h, z, a, np.asarray(observations)[:, i], first, u_z=u_burn[i], act_noise=dummy)
            h, z, a = st["h"], st["z"], actions[:, i]  # :331
        hs, zs, acts, zidx = [h], [z], [a], []
        for j in range(timesteps_H):
            h = self.sequence_model(z, a, h)  # :338-342
            z, _, idx = self.representation("dyn.repr", self.mlp("dyn", h, 1), u_dyn[j], "burn_u_dyn")  # :344
            if use_sampled_actions_in_dream:
                a = actions[:, timesteps_burn_in + j]  # :347-348
            else:
                hz = torch.cat([h, z.reshape(B, -1)], -1)
                a, _ = self.actor(hz.detach(), act_noise[j], "burn_act")  # :354
            hs.append(h); zs.append(z); acts.append(a); zidx.append(idx)
        h_HB, z_HB, a_HB = torch.stack(hs, 0), torch.stack(zs, 0), torch.stack(acts, 0)
        hz = torch.cat([h_HB.reshape(-1, cfg.G), z_HB.reshape(-1, cfg.Z)], -1)
        r_sym, _ = self.reward_predictor(hz)
        c, _ = self.continue_predictor(hz)
        key = "actions_sampled_t0_to_H_B" if use_sampled_actions_in_dream else "actions_dreamed_t0_to_H_B"
        out = {
            "h_states_t0_to_H_B": h_HB.detach(), "z_states_prior_t0_to_H_B": z_HB.detach(),
            "rewards_dreamed_t0_to_H_B": inverse_symlog(r_sym).reshape(-1, B).detach(),  # :383
            "continues_dreamed_t0_to_H_B": c.reshape(-1, B).detach(),
            key: a_HB.detach(), "z_indices_t1_to_H_B": np.stack(zidx, 0),
        }
        if cfg.discrete:
            out[key.replace("actions_", "actions_ints_", 1)] = torch.argmax(a_HB, -1)
        return out

    @staticmethod
    def compute_value_targets(r, c, v, gamma, lambda_, intrinsic_t1_to_H=None):
        """losses/critic_loss.py:92-139 (differentiable restatement)."""
        discount = c[1:] * gamma
        r1 = r[1:] if intrinsic_t1_to_H is None else r[1:] + intrinsic_t1_to_H  # :118-120
        inter = r1 + discount * (1.0 - lambda_) * v[1:]
        Rs = [v[-1]]
        for t in reversed(range(discount.shape[0])):
            Rs.append(inter[t] + discount[t] * lambda_ * Rs[-1])
        return torch.stack(list(reversed(Rs))[:-1], 0)

    def critic_loss(self, dream, value_targets):
        """losses/critic_loss.py:15-89."""
        dt = self.dtype
        Hp1, N = dream["rewards_dreamed_t0_to_H_B"].shape
        H = Hp1 - 1
        vt = value_targets.detach()
        vt_sym = nx.symlog(vt.numpy().astype(self.th_float).reshape(-1), self.th_float)
        th = _t(nx.two_hot(vt_sym, ft=self.th_float), dt).reshape(H, N, NUM_BUCKETS)
        logits = dream["values_symlog_dreamed_logits_t0_to_HxB"].reshape(Hp1, N, NUM_BUCKETS)[:-1]
        logp = logits - torch.logsumexp(logits, -1, keepdim=True)
        L_val = -(logp * th).sum(-1)
        ema = dream["v_symlog_dreamed_ema_t0_to_H_B"].detach()[:-1]
        th_ema = _t(nx.two_hot(ema.numpy().astype(self.th_float).reshape(-1), ft=self.th_float), dt).reshape(H, N, NUM_BUCKETS)
        L_reg = -(logp * th_ema).sum(-1)
        L_HB = (L_val + L_reg) * dream["dream_loss_weights_t0_to_H_B"].detach()[:-1]  # :73
        return {
            "VALUE_TARGETS_symlog_H_B": _t(vt_sym, dt).reshape(H, N),
            "CRITIC_L_total": L_HB.mean(),
            "CRITIC_L_total_H_B": L_HB,
            "CRITIC_L_neg_logp_of_value_targets_H_B": L_val,
            "CRITIC_L_neg_logp_of_value_targets": L_val.mean(),
            "CRITIC_L_slow_critic_regularization_H_B": L_reg,
            "CRITIC_L_slow_critic_regularization": L_reg.mean(),
        }

    def actor_loss(self, dream, value_targets, entropy_scale, decay):
        """losses/actor_loss.py:12-138."""
        cfg = self.cfg
        # compute_scaled_value_targets (:99-138); percentile EMA side effect
        vt_np = value_targets.detach().to(torch.float32).numpy()
        p5, p95 = float(nx.percentile_nearest(vt_np, 5)), float(nx.percentile_nearest(vt_np, 95))
        if math.isnan(self.ema_pct5):
            self.ema_pct5, self.ema_pct95 = p5, p95
        else:
            f32 = np.float32
            self.ema_pct5 = float(f32(decay) * f32(self.ema_pct5) + f32(1.0 - decay) * f32(p5))
            self.ema_pct95 = float(f32(decay) * f32(self.ema_pct95) + f32(1.0 - decay) * f32(p95))
        offset = self.ema_pct5
        invscale = max(1e-8, self.ema_pct95 - self.ema_pct5)
        v = dream["values_dreamed_t0_to_H_B"][:-1]
        scaled = (value_targets - offset) / invscale - (v - offset) / invscale
        dists = dream["actions_dreamed_distributions_t0_to_H_B"][:-1]
        actions = dream["actions_dreamed_t0_to_H_B"].detach()[:-1]
        if cfg.discrete:
            logp_all = torch.stack([torch.log(d["probs"]) for d in dists], 0)  # dist.logits = log(unimix probs)
            logp = (actions * logp_all).sum(-1)  # :40-43
            logp_loss = logp * scaled.detach()  # :45-47
            # OneHotCategorical(logits=log p).entropy() = -sum softmax(l) log_softmax(l)
            ent = torch.stack([-(torch.softmax(torch.log(d["probs"]), -1) *
                                 torch.log_softmax(torch.log(d["probs"]), -1)).sum(-1) for d in dists], 0)
        else:
            logp = torch.stack([
                (-0.5 * ((actions[i] - d["loc"]) / d["std"]) ** 2 - torch.log(d["std"]) - 0.5 * math.log(2 * math.pi)).sum(-1)
                for i, d in enumerate(dists)], 0)
            logp_loss = scaled  # :57 (gradient flows through the dynamics)
            ent = torch.stack([(0.5 + 0.5 * math.log(2 * math.pi) + torch.log(d["std"])).sum(-1) for d in dists], 0)
        L_reinforce = -logp_loss
        L_ent = -entropy_scale * ent
        L_HB = (L_reinforce + L_ent) * dream["dream_loss_weights_t0_to_H_B"].detach()[:-1]
        return {
            "ACTOR_L_total_H_B": L_HB, "ACTOR_L_total": L_HB.mean(),
            "ACTOR_logp_actions_dreamed_H_B": logp,
            "ACTOR_scaled_value_targets_H_B": scaled,
            "ACTOR_value_targets_pct95_ema": self.ema_pct95,
            "ACTOR_value_targets_pct5_ema": self.ema_pct5,
            "ACTOR_action_entropy_H_B": ent, "ACTOR_action_entropy": ent.mean(),
            "ACTOR_L_neglogp_reinforce_term_H_B": L_reinforce,
            "ACTOR_L_neglogp_reinforce_term": L_reinforce.mean(),
            "ACTOR_L_neg_entropy_term_H_B": L_ent, "ACTOR_L_neg_entropy_term": L_ent.mean(),
        }

    def train_actor_and_critic_one_step(self, h_BxT, z_BxT, is_terminated, H=15, gamma=0.997, lambda_=0.95,
                                        actor_grad_clip=100.0, critic_grad_clip=100.0, entropy_scale=3e-4,
                                        return_normalization_decay=0.99, u_dyn=None, act_noise=None, rng=None,
                                        margin=0.0, apply=True, disagree_grad_clip=100.0):
        """training/train_one_step.py:130-252 (use_curiosity = cfg.num_curiosity_nets > 0)."""
        use_curiosity = getattr(self.cfg, "num_curiosity_nets", 0) > 0
        dream = self.dream_trajectory(h_BxT, z_BxT, np.asarray(is_terminated).reshape(-1), H, gamma,
                                      u_dyn, act_noise, rng, margin)
        vt = self.compute_value_targets(dream["rewards_dreamed_t0_to_H_B"], dream["continues_dreamed_t0_to_H_B"],
                                        dream["values_dreamed_t0_to_H_B"], gamma, lambda_,
                                        dream["rewards_intrinsic_t1_to_H_B"] if use_curiosity else None)
        res = self.critic_loss(dream, vt)
        res.update(self.actor_loss(dream, vt, entropy_scale, return_normalization_decay))
        res["VALUE_TARGETS_H_B"] = vt
        res["dream_data"] = dream
        an, cn = self.group_names("actor"), self.group_names("critic")
        ag = torch.autograd.grad(res["ACTOR_L_total"], [self.P[n] for n in an], retain_graph=True, allow_unused=True)
        cg = torch.autograd.grad(res["CRITIC_L_total"], [self.P[n] for n in cn], retain_graph=use_curiosity, allow_unused=True)
        ag = [g if g is not None else torch.zeros_like(self.P[n]) for n, g in zip(an, ag)]
        cg = [g if g is not None else torch.zeros_like(self.P[n]) for n, g in zip(cn, cg)]
        res["grads"] = {n: g.detach().clone() for n, g in list(zip(an, ag)) + list(zip(cn, cg))}
        if use_curiosity:  # train_one_step.py:180-181, 201-212
            dn = self.group_names("disagree")
            res["DISAGREE_L_total"] = self.disagree_loss(dream)
            res["DISAGREE_intrinsic_rewards_H_B"] = dream["rewards_intrinsic_t1_to_H_B"]
            res["DISAGREE_intrinsic_rewards"] = dream["rewards_intrinsic_t1_to_H_B"].mean()
            dg = torch.autograd.grad(res["DISAGREE_L_total"], [self.P[n] for n in dn], allow_unused=True)
            dg = [g if g is not None else torch.zeros_like(self.P[n]) for n, g in zip(dn, dg)]
            res["grads"].update({n: g.detach().clone() for n, g in zip(dn, dg)})
            res["DISAGREE_gradients_maxabs"] = max(float(g.abs().max()) for g in dg)
        res["ACTOR_gradients_maxabs"] = max(float(g.abs().max()) for g in ag)
        res["CRITIC_gradients_maxabs"] = max(float(g.abs().max()) for g in cg)
        if apply:
            res["actor_grad_global_norm"], _ = self._apply("actor", an, ag, actor_grad_clip, 3e-5, 1e-5)
            res["critic_grad_global_norm"], _ = self._apply("critic", cn, cg, critic_grad_clip, 3e-5, 1e-5)
            if use_curiosity:  # disagree_networks.py:43: Adam(1e-4, eps 1e-5)
                res["disagree_grad_global_norm"], cl = self._apply("disagree", dn, dg, disagree_grad_clip, 1e-4, 1e-5)
                res["DISAGREE_gradients_clipped_by_glob_norm_maxabs"] = max(float(g.abs().max()) for g in cl)
            self.update_ema()
        return res

    def update_ema(self, decay=0.98):
        """models/critic_network.py:93-100."""
        with torch.no_grad():
            for n in self.group_names("critic"):
                e = "critic_ema." + n[len("critic."):]
                self.P[e].copy_(decay * self.P[e] + (1.0 - decay) * self.P[n])


def make_synthetic_sample(cfg, seed=0):
    """Synthetic replay batch of BASELINE.md section 2 for a config (shapes from cfg)."""
    rng = np.random.default_rng(seed)
    B, T, A = cfg.B, cfg.T, cfg.A
    if cfg.image_obs:
        obs = rng.integers(0, 256, size=(B, T, cfg.img_hw, cfg.img_hw, cfg.img_c), dtype=np.uint8)
        obs = (obs.astype(np.float32) / 128.0 - 1.0).astype(np.float32)  # env_runner_v2.py:53-54
    else:
        obs = (rng.standard_normal((B, T, cfg.obs_dim)) * 2.0).astype(np.float32)
    if cfg.discrete:
        ai = rng.integers(0, A, size=(B, T))
        actions = np.eye(A, dtype=np.float32)[ai]
    else:
        actions = rng.uniform(-1.0, 1.0, size=(B, T, A)).astype(np.float32)
    is_first = (rng.random((B, T)) < 1.0 / 50.0).astype(np.float32)
    is_first[:, 0] = 1.0
    is_terminated = np.zeros((B, T), np.float32)
    is_terminated[:, :-1] = is_first[:, 1:] * (rng.random((B, T - 1)) < 0.9)
    rewards = np.where(is_first > 0, 0.0, 1.0).astype(np.float32)
    rewards = (rewards * rng.choice([1.0, 1.0, 0.5, -1.0, 3.7], size=(B, T))).astype(np.float32)
    return {"obs": obs, "actions": actions, "rewards": rewards, "is_first": is_first,
            "is_terminated": is_terminated}
