// Actor-critic half of the hot path: dream_trajectory (dreamer_model.py:147-303), lambda-returns, critic and
// actor losses, and their gradients (train_one_step.py:130-252). Dream tensors are time-major [H+1, N, ...]
// (row = i*N + n) as in the reference. Heads (reward / continue / critic / EMA critic) run once over all
// (H+1)*N rows after the rollout; the actor has to run inside it because a_i feeds h_{i+1}.
#include "dv3_engine.h"
#include "dv3_dream_tc.h"

#include <cstdlib>
#include <cstring>

namespace dv3 {

// the critic has its own Adam in the reference (critic_network.py:58); 0 = same as the actor's
static inline float critic_lr(const dv3_hparams* hp) { return hp->critic_lr > 0.f ? hp->critic_lr : hp->ac_lr; }
static inline float critic_eps(const dv3_hparams* hp) { return hp->critic_eps > 0.f ? hp->critic_eps : hp->ac_eps; }

void Engine::actor_fwd(long long row_off) {
  // models/actor_network.py:63-118 on rows [row_off, row_off + N) of the dream
  const float* x = dr_hz + (size_t)row_off * (G + Z);
  const void* x16 = dr_hz16 ? (const void*)((const unsigned short*)dr_hz16 + (size_t)row_off * (G + Z)) : nullptr;
  float* out = actor_logits + (size_t)row_off * A;
  auto mean_branch = [&] {
    mlp_fwd(actor_mlp, actor_act, x, G + Z, N, row_off, x16);
    mm(actor_act.act.back() + (size_t)row_off * D, D, false, actor_out.w, A, false, out, A, N, A, D, actor_out.bias);
  };
  if (discrete) {
    mean_branch();
    disc_actor_fwd(cur, out, in_act_noise + row_off, actor_probs + (size_t)row_off * A, dr_a_idx + row_off,
                   dr_act + (size_t)row_off * A, A, N, A);
  } else {
    float* sraw = actor_s + (size_t)row_off * A;
    run_branches({mean_branch, [&] {
                    mlp_fwd(actor_std_mlp, actor_std_act, x, G + Z, N, row_off, x16);
                    mm(actor_std_act.act.back() + (size_t)row_off * D, D, false, actor_std_out.w, A, false, sraw, A, N, A, D, actor_std_out.bias);
                  }});
    box_actor_fwd(cur, out, sraw, in_act_noise + (size_t)row_off * A, dr_act + (size_t)row_off * A, A, N, A);
  }
}

// The fused path: the actor's contractions (everything up to the last hidden layer's pre-activation) ...
bool Engine::actor_step_applicable(int i) const {
  static const bool off = std::getenv("DV3_NO_ACTOR_STEP") != nullptr;
  const bool given = dream_action_mode == 2 || (dream_action_mode == 1 && i == 0);
  return !off && !given && actor_step_supported(D, A);
}
void Engine::actor_step_mlps(long long row_off) {
  const float* x = dr_hz + (size_t)row_off * (G + Z);
  const void* x16 = dr_hz16 ? (const void*)((const unsigned short*)dr_hz16 + (size_t)row_off * (G + Z)) : nullptr;
  // (the first layer's pre-activations were zeroed for all dream rows at the start of the rollout: see dream_forward)
  if (discrete) {
    mlp_fwd(actor_mlp, actor_act, x, G + Z, N, row_off, x16, true, true);
  } else {
    run_branches({[&] { mlp_fwd(actor_mlp, actor_act, x, G + Z, N, row_off, x16, true, true); },
                  [&] { mlp_fwd(actor_std_mlp, actor_std_act, x, G + Z, N, row_off, x16, true, true); }});
  }
}
// ... and its tail: LN+SiLU of that layer, the output layer, the sample and (with_pre) the pre-GRU layer of step i + 1
int Engine::actor_step_tail(int i, bool with_pre) {
  const size_t off = (size_t)i * N;
  const LnLayer& Wx = pre_mlp.layers[0];
  ActorStepArgs a{};
  a.rows = N; a.D = D; a.A = A; a.discrete = discrete ? 1 : 0; a.do_pre = with_pre ? 1 : 0;
  auto fill = [&](ActorStepNet& n, const Mlp& w, MlpAct& act, const DenseLayer& head) {
    const size_t l = w.layers.size() - 1;
    n.pre = act.pre[l] + off * D; n.act = act.act[l] + off * D;
    n.act16 = N >= 128 && !act.act16.empty() && act.act16[l] ? (void*)((unsigned short*)act.act16[l] + off * D) : nullptr;
    n.mean = act.mean[l] + off; n.rstd = act.rstd[l] + off; n.g = w.layers[l].g; n.b = w.layers[l].b;
    n.out_w = head.w; n.out_b = head.bias;
  };
  fill(a.net[0], actor_mlp, actor_act, actor_out);
  if (!discrete) fill(a.net[1], actor_std_mlp, actor_std_act, actor_std_out);
  a.noise = in_act_noise + off * (discrete ? 1 : A);
  a.logits = actor_logits + off * A; a.s_raw = actor_s ? actor_s + off * A : nullptr;
  a.probs = actor_probs ? actor_probs + off * A : nullptr; a.a_idx = dr_a_idx ? dr_a_idx + off : nullptr;
  a.act = dr_act + off * A;
  if (with_pre) {
    a.W_pre_a = Wx.w + (size_t)Z * D; a.pre_g = Wx.g; a.pre_b = Wx.b;
    a.x_pre = dr_x_pre + off * D; a.x_mean = dr_x_mean + off; a.x_rstd = dr_x_rstd + off; a.u = dr_u + off * D;
    a.u16 = dr_u16 ? (void*)((unsigned short*)dr_u16 + off * D) : nullptr;
  }
  const cudaError_t ce = actor_step(cur, a);
  if (ce != cudaSuccess) return fail(std::string("actor step launch: ") + cudaGetErrorString(ce));
  return 0;
}

// Action of dream step i: the actor's sample, or the caller's (dream_trajectory_with_burn_in, dreamer_model.py:306-413:
// a_0 = the last burn-in action; with use_sampled_actions / use_random_actions every action is given).
void Engine::dream_action(int i) {
  const long long off = (long long)i * N;
  if (dream_action_mode == 2 || (dream_action_mode == 1 && i == 0))
    cudaMemcpyAsync(dr_act + (size_t)off * A, in_dream_actions + (size_t)off * A, (size_t)N * A * 4, cudaMemcpyDeviceToDevice, cur);
  else
    actor_fwd(off);
}

// The H dreamed steps as ONE persistent kernel (dv3_dream_tc.cu): tensor-core mode, XS / S layer sizes, actor-sampled actions.
bool Engine::dream_rollout_tc(cudaStream_t s, int start_from_observe) {
  // Two persistent rollout kernels, both opt-in: measured on B200 neither beats the per-kernel path inside the replayed CUDA graph
  // (XS, us per dreamed step: per-kernel path ~64; DV3_DREAM_TC=1, grid-wide stage kernel of dv3_dream_tc.cu: 65 - 7 - 8 grid-wide
  // stages of 6 - 9 us each; DV3_DREAM_RR=1, row-resident kernel of dv3_dream_rr.cu: 52 (Discrete) - 61 (Box) - a CTA ingests the
  // step's 1.6 MB of weights at 68 GB/s, 21 us, its 5632 HMMA take 14.7 us, and with 8 warps per SM the epilogues are latency
  // chains). Both are parity-tested against the oracle.
  const bool opt_in = std::getenv("DV3_DREAM_TC") != nullptr;
  const bool rr = !opt_in && start_from_observe && std::getenv("DV3_DREAM_RR") != nullptr && std::getenv("DV3_DREAM_RR")[0] != 0 && dream_rr_supported(N, G, D, Z, Zc, Zk, A, L);
  if (!(opt_in || rr) || dream_action_mode != 0 || !use_persistent || dr_hz16 == nullptr) return false;
  if (opt_in && dream_tc_ctas <= 0) return false;
  const LnLayer& Wx = pre_mlp.layers[0];
  const LnLayer& Wd = dyn_mlp.layers[0];
  DreamTcArgs a{};
  a.N = N; a.H = H; a.G = G; a.D = D; a.Z = Z; a.Zc = Zc; a.Zk = Zk; a.A = A; a.L = L; a.discrete = discrete ? 1 : 0;
  auto fill = [&](DreamTcMlp& m, const Mlp& w, MlpAct& act) -> bool {
    for (int l = 0; l < L; ++l) {
      m.pre[l] = act.pre[l]; m.act[l] = act.act[l]; m.mean[l] = act.mean[l]; m.rstd[l] = act.rstd[l];
      m.act16[l] = act.act16.empty() ? nullptr : act.act16[l];
      m.g[l] = w.layers[l].g; m.b[l] = w.layers[l].b;
      m.wt[l] = shadow_t(w.layers[l].w, &m.ld_wt[l]);
      if (m.wt[l] == nullptr) return false;
    }
    return true;
  };
  if (!fill(a.actor, actor_mlp, actor_act)) return false;
  if (!discrete && !fill(a.actor_std, actor_std_mlp, actor_std_act)) return false;
  a.actor_out_w = actor_out.w; a.actor_out_b = actor_out.bias; a.std_out_w = actor_std_out.w; a.std_out_b = actor_std_out.bias;
  a.W_pre_a = Wx.w + (size_t)Z * D;
  a.pre_g = Wx.g; a.pre_b = Wx.b; a.gru_b = gru_b; a.dyn_g = Wd.g; a.dyn_b = Wd.b; a.zp_bias = dyn_repr.bias;
  a.Wt_pre_z = shadow_t(Wx.w, &a.ld_pre_z); a.Wt_U = shadow_t(gru_U, &a.ld_U); a.Wt_K = shadow_t(gru_K, &a.ld_K);
  a.Wt_dyn = shadow_t(Wd.w, &a.ld_dyn); a.Wt_zp = shadow_t(dyn_repr.w, &a.ld_zp);
  if (!a.Wt_pre_z || !a.Wt_U || !a.Wt_K || !a.Wt_dyn || !a.Wt_zp) return false;
  if (dr_u16 == nullptr || dr_q_act16 == nullptr) return false;   // the kernel's contractions read the bf16 twins through TMA
  for (int l = 0; l + 1 < L; ++l)
    if (a.actor.act16[l] == nullptr || (!discrete && a.actor_std.act16[l] == nullptr)) return false;
  a.act_noise = in_act_noise; a.u_dyn = in_u_dyn;
  a.hz = dr_hz; a.hz16 = dr_hz16;
  a.actor_logits = actor_logits; a.actor_s = actor_s; a.actor_probs = actor_probs; a.a_idx = dr_a_idx; a.dr_act = dr_act;
  a.x_pre = dr_x_pre; a.x_mean = dr_x_mean; a.x_rstd = dr_x_rstd; a.u = dr_u; a.u16 = dr_u16;
  a.gx = dr_gx; a.gh = dr_gh; a.zpart = sN_D0; a.gx_raw = sN_3G0;
  a.q_pre = dr_q_pre; a.q_mean = dr_q_mean; a.q_rstd = dr_q_rstd; a.q_act = dr_q_act; a.q_act16 = dr_q_act16;
  a.prior_logits = dr_prior_logits; a.prior_probs = dr_prior_probs; a.z_idx = dr_z_idx;
  a.bar = scan_bar + 48;
  a.prof = std::getenv("DV3_SCAN_STAMPS") ? scan_stamps + 2048 : nullptr;
  if (rr) {
    DreamRrExtra x{};
    x.Wn_pre = shadow_n(Wx.w, &x.ld_n_pre);
    x.Wn_act[0] = shadow_n(actor_mlp.layers[0].w, &x.ld_n_act[0]);
    x.Wn_act[1] = discrete ? x.Wn_act[0] : shadow_n(actor_std_mlp.layers[0].w, &x.ld_n_act[1]);
    if (discrete) x.ld_n_act[1] = x.ld_n_act[0];
    x.z0_idx = z_idx;
    if (!x.Wn_pre || !x.Wn_act[0] || !x.Wn_act[1] || x.ld_n_pre % 8 != 0 || x.ld_n_act[0] % 8 != 0 || x.ld_n_act[1] % 8 != 0) return false;
    const cudaError_t ce = launch_dream_rr(s, a, x);
    if (ce != cudaSuccess) { err = std::string("dream rollout (row-resident) launch: ") + cudaGetErrorString(ce); dream_rollout_failed = true; return false; }
    return true;
  }
  const cudaError_t ce = launch_dream_tc(s, a, dream_tc_ctas);
  if (ce != cudaSuccess) { err = std::string("dream rollout (tensor-core) launch: ") + cudaGetErrorString(ce); return false; }
  return true;
}

int Engine::dream_forward(float gamma, int start_from_observe, cudaStream_t s) {
  cur = s;
  set_ln_fast(cfg.compute_mode == 1);
  last_gamma = gamma;
  const LnLayer& Wx = pre_mlp.layers[0];
  const LnLayer& Wd = dyn_mlp.layers[0];
  const int W = G + Z;
  refresh_shadows(s);  // the world model was just updated (run_experiment.py:426-459), actor / critic at the end of the last step
  // start states: row n = b*T + t of the observe pass (run_experiment.py:458-460)
  if (start_from_observe) {
    cudaMemcpyAsync(dr_hz, hz, (size_t)N * W * 4, cudaMemcpyDeviceToDevice, s);
  } else {
    copy_2d(s, in_dream_h0, G, dr_hz, W, N, G);
    copy_2d(s, in_dream_z0, Z, dr_hz + G, W, N, Z);
  }
  unsigned short* hz16p = reinterpret_cast<unsigned short*>(dr_hz16);   // bf16 twin of dream.hz, written by the producers below
  auto h16 = [&](size_t row, int col) -> void* { return hz16p ? (void*)(hz16p + row * W + col) : nullptr; };
  auto t16 = [&](void* base, size_t off) -> void* { return base ? (void*)((unsigned short*)base + off) : nullptr; };
  if (hz16p) to_bf16(s, dr_hz, W, hz16p, W, N, W);
  // Accumulating contractions of the rollout add into rows zeroed here, once for all steps: at the XS sizes the K = Z and
  // K = G + Z contractions are split-K, and a zero-fill node in front of each would sit on the step's dependency chain
  const bool fused_any = actor_step_applicable(1);
  cudaMemsetAsync(dr_x_pre, 0, (size_t)H * N * D * 4, s);
  if (fused_any) {
    cudaMemsetAsync(actor_act.pre[0], 0, (size_t)R * actor_mlp.layers[0].out * 4, s);
    if (!discrete) cudaMemsetAsync(actor_std_act.pre[0], 0, (size_t)R * actor_std_mlp.layers[0].out * 4, s);
  }
  phase_mark("dream.begin");
  dream_rollout_failed = false;
  if (dream_rollout_tc(s, start_from_observe)) {
    phase_mark("dream.rollout_tc");
  } else if (dream_rollout_failed) {
    return fail(err);   // a persistent rollout kernel that applies but cannot launch is an error, not a reason to fall back
  } else
  for (int i = 1; i <= H; ++i) {
    const size_t prev = (size_t)(i - 1) * N, curr = (size_t)i * N, st = (size_t)(i - 1) * N;  // st: per-step buffers
    const float* hz_p = dr_hz + prev * W;
    float* hz_c = dr_hz + curr * W;
    // Everything that only needs [h, z]_{i-1} runs concurrently: the actor (its action a_{i-1} enters the sequence model
    // through the small a-part of the pre-GRU layer only), the z-part of that layer, and h_{i-1} . U.
    const bool fused = actor_step_applicable(i - 1);
    run_branches({[&] { if (fused) actor_step_mlps((long long)prev); else dream_action(i - 1); },
                  [&] { mm(hz_p + G, W, false, Wx.w, D, false, dr_x_pre + st * D, D, N, D, Z, nullptr, 1.f, h16(prev, G)); },
                  [&] { mm(hz_p, W, false, gru_U, 3 * G, false, dr_gh + st * 3 * G, 3 * G, N, 3 * G, G, gru_b + 3 * G, 0.f, h16(prev, 0)); }});
    phase_mark("dream.actor|z|gh");
    if (fused) {
      // actor tail, a . W_pre[Z:] and the pre-GRU LayerNorm in one launch
      if (int rc = actor_step_tail(i - 1, true)) return rc;
    } else {
      mm(dr_act + prev * A, A, false, Wx.w + (size_t)Z * D, D, false, dr_x_pre + st * D, D, N, D, A, nullptr, 1.f);
      ln_silu_fwd(s, dr_x_pre + st * D, D, Wx.g, Wx.b, dr_u + st * D, D, dr_x_mean + st, dr_x_rstd + st, 1, N, D, t16(dr_u16, st * D));
    }
    mm(dr_u + st * D, D, false, gru_K, 3 * G, false, dr_gx + st * 3 * G, 3 * G, N, 3 * G, D, gru_b, 0.f, t16(dr_u16, st * D));
    phase_mark("dream.pre+gx");
    gru_fwd(s, dr_gx + st * 3 * G, 3 * G, dr_gh + st * 3 * G, 3 * G, hz_p, W, hz_c, W, N, G, h16(curr, 0));
    phase_mark("dream.gru");
    // prior z_i ~ dynamics_predictor(h_i), straight-through sample
    mm(hz_c, W, false, Wd.w, D, false, dr_q_pre + st * D, D, N, D, G, nullptr, 0.f, h16(curr, 0));
    ln_silu_fwd(s, dr_q_pre + st * D, D, Wd.g, Wd.b, dr_q_act + st * D, D, dr_q_mean + st, dr_q_rstd + st, 1, N, D, t16(dr_q_act16, st * D));
    mm(dr_q_act + st * D, D, false, dyn_repr.w, Z, false, dr_prior_logits + st * Z, Z, N, Z, D, dyn_repr.bias, 0.f, t16(dr_q_act16, st * D));
    repr_fwd(s, dr_prior_logits + st * Z, Z, dr_prior_probs + st * Z, Z, in_u_dyn + st * Zc, Zc, dr_z_idx + st * Zc, Zc,
             hz_c + G, W, N, Zc, Zk, 1, h16(curr, G));
    phase_mark("dream.prior");
  }
  // heads on all (H+1)*N rows: four independent chains, run concurrently (together with the actor of the last dreamed state,
  // whose action is never fed back)
  run_branches({
      [&] {
        mlp_fwd(critic_mlp, critic_act, dr_hz, W, R, 0, dr_hz16);
        mm(critic_act.act.back(), D, false, critic_head.w, DV3_NB, false, critic_logits, DV3_NB, R, DV3_NB, D, critic_head.bias, 0.f, critic_act.act16.back());
        bucket_expectation(cur, critic_logits, DV3_NB, v_E, R);
      },
      [&] {
        mlp_fwd(rew_mlp, dr_rew_act, dr_hz, W, R, 0, dr_hz16);
        mm(dr_rew_act.act.back(), D, false, rew_head.w, DV3_NB, false, dr_rew_logits, DV3_NB, R, DV3_NB, D, rew_head.bias, 0.f, dr_rew_act.act16.back());
        bucket_expectation(cur, dr_rew_logits, DV3_NB, dr_rew_E, R);
      },
      [&] {
        mlp_fwd(cont_mlp, dr_cont_act, dr_hz, W, R, 0, dr_hz16);
        mm(dr_cont_act.act.back(), D, false, cont_out.w, 1, false, dr_cont_logit, 1, R, 1, D, cont_out.bias);
      },
      [&] {
        mlp_fwd(ema_mlp, ema_act, dr_hz, W, R, 0, dr_hz16);
        mm(ema_act.act.back(), D, false, ema_head.w, DV3_NB, false, ema_logits, DV3_NB, R, DV3_NB, D, ema_head.bias, 0.f, ema_act.act16.back());
        bucket_expectation(cur, ema_logits, DV3_NB, ema_E, R);
      },
      [&] {
        if (actor_step_applicable(H)) { actor_step_mlps((long long)H * N); actor_step_tail(H, false); }
        else dream_action(H);
      }});
  phase_mark("dream.heads");
  if (!dis.empty()) {   // dreamer_model.py:226-240: needs the actions of all H + 1 dreamed states
    disagree_forward();
    phase_mark("dream.disagree");
  }
  // real-space rewards / continues / values, loss weights (dreamer_model.py:219-280) and lambda-returns
  AcArgs a{};
  a.H = H; a.N = N; a.A = A; a.discrete = discrete; a.gamma = gamma; a.lambda_ = last_lambda;
  a.r_intr = dis.empty() ? nullptr : dis_ri;
  a.rew_E = dr_rew_E; a.cont_logit = dr_cont_logit; a.v_E = v_E; a.ema_E = ema_E; a.is_terminated = in_is_term;
  a.r = ac_r; a.c = ac_c; a.w = ac_w; a.v = ac_v; a.targets = ac_targets;
  ac_value_targets(s, a);
  phase_mark("dream.targets");
  return 0;
}

int Engine::ac_loss_backward(const dv3_hparams* hp, int phase, cudaStream_t s) {
  cur = s;
  set_ln_fast(cfg.compute_mode == 1);
  const int W = G + Z;
  const long long HN = (long long)H * N;
  AcArgs a{};
  a.H = H; a.N = N; a.A = A; a.discrete = discrete;
  a.gamma = hp->gamma; a.lambda_ = hp->lambda_; a.entropy_scale = hp->entropy_scale; a.decay = hp->return_normalization_decay;
  last_lambda = hp->lambda_;
  a.inv_count = inv_world / (float)HN;
  a.r_intr = dis.empty() ? nullptr : dis_ri;
  a.rew_E = dr_rew_E; a.cont_logit = dr_cont_logit; a.v_E = v_E; a.ema_E = ema_E; a.is_terminated = in_is_term;
  a.r = ac_r; a.c = ac_c; a.w = ac_w; a.v = ac_v; a.targets = ac_targets;
  if (phase & 1) {
    ac_value_targets(s, a);  // dreamer_model.py:219-280 tail + critic_loss.py:92-139
    if (cfg.world_size == 1) cudaMemcpyAsync(ac_targets_all, ac_targets, (size_t)HN * 4, cudaMemcpyDeviceToDevice, s);
  }
  if (!(phase & 2)) return 0;

  // percentile EMAs over the (all-gathered) value targets (actor_loss.py:110-129): part of actor_loss, which the reference
  // only calls when the actor is trained (train_one_step.py:176-189) - a critic-only update leaves them untouched
  if (hp->train_actor) percentile_ema(s, ac_targets_all, HN * cfg.world_size, hp->return_normalization_decay, pct_ema, pct_raw, nullptr);

  AcLossArgs l{};
  l.base = a; l.ema = pct_ema; l.critic_logits = critic_logits; l.dcritic_logits = dcritic_logits;
  l.actor_logits = actor_logits; l.action_idx = dr_a_idx; l.dactor_logits = dactor_logits;
  l.actor_mu = actor_logits; l.actor_s = actor_s; l.actions = dr_act; l.dmu = dmu; l.ds = ds; l.dR = dR; l.dV = dV;
  l.L_critic_HB = L_critic_HB; l.L_val_HB = L_val_HB; l.L_reg_HB = L_reg_HB; l.targets_symlog = targets_symlog;
  l.L_actor_HB = L_actor_HB; l.logp_HB = logp_HB; l.scaled_HB = scaled_HB; l.ent_HB = ent_HB;
  l.L_reinf_HB = L_reinf_HB; l.L_ent_HB = L_ent_HB; l.scalars = ac_scalars;
  phase_mark("ac.begin");
  ac_losses(s, l);
  phase_mark("ac.losses");

  // ---- critic: gradient only reaches the fast critic's own weights (train_one_step.py:197-200). The chain is independent
  // of everything the actor needs, so it runs on an auxiliary stream next to the whole actor backward. ----
  auto critic_chain = [&] {
    cudaMemsetAsync(grads[2], 0, (size_t)numel[2] * 4, cur);
    dense_bwd(critic_head, critic_act.act.back(), D, dcritic_logits, DV3_NB, sN_D2, D, 0.f, true, HN);
    mlp_bwd(critic_mlp, critic_act, dr_hz, W, sN_D2, sN_D3, nullptr, 0, true, HN, 0, dr_hz16);
  };
  if (!hp->train_actor) { critic_chain(); return 0; }
  side_run(critic_chain);
  actor_backward(hp, a);
  phase_mark("ac.actor_mlps_bwd");
  side_join();
  phase_mark("ac.join_critic");
  return 0;
}

// Actor gradients: REINFORCE through the actor MLP (Discrete) or BPTT through rewards / values / the RSSM steps (Box).
void Engine::actor_backward(const dv3_hparams* hp, const AcArgs& a) {
  cudaStream_t s = cur;
  const int W = G + Z;
  const long long HN = (long long)H * N;
  {
    (void)hp;
  }
  cudaMemsetAsync(grads[1], 0, (size_t)numel[1] * 4, s);
  if (!discrete) {
    // Box actions: the loss is -scaled advantages with gradient (actor_loss.py:57) -> BPTT through rewards,
    // values and the 15 RSSM steps into the reparameterised action samples.
    const LnLayer& Wx = pre_mlp.layers[0];
    const LnLayer& Wd = dyn_mlp.layers[0];
    ac_returns_bwd(s, a, dR, dV, drew_E, dv_E);
    cudaMemsetAsync(d_dr_hz, 0, (size_t)R * W * 4, s);
    bucket_expectation_bwd(s, dr_rew_logits, DV3_NB, drew_E, dr_dlogits255, DV3_NB, R);
    dense_bwd(rew_head, nullptr, 0, dr_dlogits255, DV3_NB, sN_D0, D, 0.f, false, R);
    mlp_bwd(rew_mlp, dr_rew_act, dr_hz, W, sN_D0, sN_D1, d_dr_hz, W, false, R);
    bucket_expectation_bwd(s, critic_logits, DV3_NB, dv_E, dr_dlogits255, DV3_NB, R);
    dense_bwd(critic_head, nullptr, 0, dr_dlogits255, DV3_NB, sN_D0, D, 0.f, false, R);
    mlp_bwd(critic_mlp, critic_act, dr_hz, W, sN_D0, sN_D1, d_dr_hz, W, false, R);
    phase_mark("ac.heads_dx");
    // per-step slices (rows N .. of the (H+1) N-row scratch; rows 0 .. N stay the per-step du scratch below), zeroed once: the
    // split-K contraction of each step accumulates into its slice instead of waiting for a zero-fill node of its own
    cudaMemsetAsync(sN_D0 + (size_t)N * D, 0, (size_t)H * N * D * 4, s);
    for (int i = H; i >= 1; --i) {
      const size_t prev = (size_t)(i - 1) * N, curr = (size_t)i * N, st = (size_t)(i - 1) * N;
      float* dq = sN_D0 + (st + N) * D;   // d q_act of this step
      float* d_c = d_dr_hz + curr * W;
      float* d_p = d_dr_hz + prev * W;
      const float* hz_p = dr_hz + prev * W;
      // z_i = straight-through sample of prior(h_i)
      repr_bwd(s, dr_prior_logits + st * Z, Z, d_c + G, W, nullptr, 0, sN_Z, Z, N, Zc, Zk, twin_of(sN_Z));
      mm(sN_Z, Z, false, dyn_repr.w, Z, true, dq, D, N, D, Z, nullptr, 1.f, twin_of(sN_Z));
      ln_silu_bwd(s, dq, D, dr_q_pre + st * D, D, dr_q_mean + st, dr_q_rstd + st, Wd.g, Wd.b, sN_D1, D, nullptr, nullptr, 1, N, D, twin_of(sN_D1));
      mm(sN_D1, D, false, Wd.w, D, true, d_c, W, N, G, D, nullptr, 1.f, twin_of(sN_D1));
      // h_i = GRU(pre([z_{i-1}, a_{i-1}]), h_{i-1})
      gru_bwd(s, d_c, W, dr_gx + st * 3 * G, 3 * G, dr_gh + st * 3 * G, 3 * G, hz_p, W, sN_3G0, 3 * G, sN_3G1, 3 * G, d_p, W, 1, N, G,
              twin_of(sN_3G0), twin_of(sN_3G1));
      run_branches({[&] {
                      mm(sN_3G0, 3 * G, false, gru_K, 3 * G, true, sN_D0, D, N, D, 3 * G, nullptr, 0.f, twin_of(sN_3G0));
                      float* dxp = dr_dx_pre + st * D;   // kept per step: the action gradients are formed for all steps at once below
                      void* dxp16 = twin_of(dr_dx_pre) ? (void*)((unsigned short*)twin_of(dr_dx_pre) + st * D) : nullptr;
                      ln_silu_bwd(cur, sN_D0, D, dr_x_pre + st * D, D, dr_x_mean + st, dr_x_rstd + st, Wx.g, Wx.b, dxp, D, nullptr, nullptr, 1, N, D, dxp16);
                      mm(dxp, D, false, Wx.w, D, true, d_p + G, W, N, Z, D, nullptr, 1.f, dxp16);
                    },
                    [&] { mm(sN_3G1, 3 * G, false, gru_U, 3 * G, true, d_p, W, N, G, 3 * G, nullptr, 1.f, twin_of(sN_3G1)); }});
      phase_mark("ac.bptt_step");
    }
    // a_{i-1} enters step i only through the a-part of the pre-GRU layer and the actor inputs are stop-gradient'ed, so the
    // action gradients of all H steps are one contraction + one reparameterisation backward after the loop
    mm(dr_dx_pre, D, false, Wx.w + (size_t)Z * D, D, true, dr_dact, A, (int)HN, A, D);
    box_actor_bwd(s, actor_logits, actor_s, in_act_noise, dr_dact, A, dmu, ds, HN, A);
    phase_mark("ac.action_grads");
  }
  // actor MLPs: inputs were stop-gradient'ed (dreamer_model.py:179-180), so only weight gradients; the mean and the std
  // network (Box) are independent chains
  if (!discrete) {
    run_branches({[&] {
                    dense_bwd(actor_out, actor_act.act.back(), D, dactor_logits, A, sN_D0, D, 0.f, true, HN);
                    mlp_bwd(actor_mlp, actor_act, dr_hz, W, sN_D0, sN_D1, nullptr, 0, true, HN, 0, dr_hz16);
                  },
                  [&] {
                    dense_bwd(actor_std_out, actor_std_act.act.back(), D, ds, A, sN_D4, D, 0.f, true, HN);
                    mlp_bwd(actor_std_mlp, actor_std_act, dr_hz, W, sN_D4, sN_D5, nullptr, 0, true, HN, 0, dr_hz16);
                  }});
  } else {
    dense_bwd(actor_out, actor_act.act.back(), D, dactor_logits, A, sN_D0, D, 0.f, true, HN);
    mlp_bwd(actor_mlp, actor_act, dr_hz, W, sN_D0, sN_D1, nullptr, 0, true, HN, 0, dr_hz16);
  }
}

// Curiosity forward (models/disagree_networks.py:47-95): every net = MLP(L) -> RepresentationLayer probabilities on
// stop_gradient([h, z, a]) of all (H+1) N dreamed rows; intrinsic reward = mean over the Zc Zk probabilities of their
// population stddev over the nets, centred over all rows, times the scale (row block 0 is computed like the reference does
// and dropped by the consumers: rewards_intrinsic_t1_to_H_B, dreamer_model.py:236-238).
void Engine::disagree_forward() {
  cudaStream_t s = cur;
  const int W = G + Z, Nd = (int)dis.size();
  copy_2d(s, dr_hz, W, dis_in, dis_ld, R, W);
  copy_2d(s, dr_act, A, dis_in + W, dis_ld, R, A);
  void* in16 = R >= 128 ? twin_of(dis_in) : nullptr;   // one bf16 copy of the shared input for the first layer of every net
  if (in16) to_bf16(s, dis_in, dis_ld, in16, dis_ld, R, dis_ld);
  // the nets are independent chains: four of them in flight (main stream + the auxiliary ones)
  auto net_fwd = [&](int n) {
    DisagreeNet& d = dis[n];
    mlp_fwd(d.mlp, d.act, dis_in, dis_ld, R, 0, in16);
    mm(d.act.act.back(), D, false, d.repr.w, Z, false, d.logits, Z, R, Z, D, d.repr.bias, 0.f, d.act.act16.empty() ? nullptr : d.act.act16.back());
    repr_fwd(cur, d.logits, Z, d.probs, Z, nullptr, 0, nullptr, 0, nullptr, 0, R, Zc, Zk, 0);
  };
  std::vector<std::function<void()>> groups;
  for (int g = 0; g < 4 && g < Nd; ++g)
    groups.push_back([&, g] { for (int n = g; n < Nd; n += 4) net_fwd(n); });
  run_branches(groups);
  disagree_std_rows(s, dis_probs, (size_t)R * Z, Nd, R, Z, dis_std);
  disagree_intrinsic(s, dis_std, R, N, cfg.intrinsic_rewards_scale, dis_ri, dis_scalars);
}

// disagree_loss (losses/disagree_loss.py:11-41) and its gradients (train_one_step.py:209-212): per net the squared error of
// the probabilities predicted at t < H against the dreamed prior z of t + 1, mean over [H, N], summed over the nets; the
// inputs are stop-gradient'ed, so only weight gradients. Time-major rows: t < H are the first H N rows, t + 1 is N rows on.
int Engine::disagree_loss_backward(cudaStream_t s) {
  if (dis.empty()) return fail("disagree_loss_backward: the engine was created without curiosity nets");
  cur = s;
  set_ln_fast(cfg.compute_mode == 1);
  const int W = G + Z, Nd = (int)dis.size();
  const long long HN = (long long)H * N;
  cudaMemsetAsync(grads[kDisagree], 0, (size_t)numel[kDisagree] * 4, s);
  for (int n = 0; n < Nd; ++n) {
    DisagreeNet& d = dis[n];
    disagree_loss_rows(s, d.probs, dr_hz + (size_t)N * W + G, W, HN, Z, 1.0f / (float)HN, dis_dprobs, dis_mse + (size_t)n * HN);
    repr_bwd(s, d.logits, Z, nullptr, 0, dis_dprobs, Z, dis_dlogits, Z, (int)HN, Zc, Zk);
    dense_bwd(d.repr, d.act.act.back(), D, dis_dlogits, Z, dis_s0, D, 0.f, true, HN);
    mlp_bwd(d.mlp, d.act, dis_in, dis_ld, dis_s0, dis_s1, nullptr, 0, true, HN, 0, HN >= 128 ? twin_of(dis_in) : nullptr);
  }
  disagree_loss_total(s, dis_mse, (long long)Nd * HN, HN, dis_scalars);
  phase_mark("ac.disagree_bwd");
  return 0;
}

int Engine::train_ac(const dv3_hparams* hp, cudaStream_t s) {
  int rc = dream_forward(hp->gamma, 1, s);
  if (rc) return rc;
  rc = ac_loss_backward(hp, 3, s);
  if (rc) return rc;
  if (!dis.empty()) {   // train_one_step.py:180-181, 209-212, 226-246
    rc = disagree_loss_backward(s);
    if (!rc) rc = optimizer_step(kDisagree, hp->disagree_grad_clip, hp->disagree_lr, hp->disagree_eps, -1.f, s);
    if (rc) return rc;
  }
  if (hp->train_actor) {
    rc = optimizer_step(1, hp->actor_grad_clip, hp->ac_lr, hp->ac_eps, -1.f, s);
    if (rc) return rc;
  }
  return optimizer_step(2, hp->critic_grad_clip, critic_lr(hp), critic_eps(hp), hp->critic_ema_decay, s);
}

template <class F>
int Engine::run_graphed(GraphSlot& slot, const dv3_hparams* hp, uint64_t seed, cudaStream_t s, F body) {
  cudaStream_t run = s;
  const bool hop = (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread);
  if (hop) {
    if (own_stream == nullptr) {
      if (cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed");
      cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming);
    }
    run = own_stream;
    cudaEventRecord(ev_in, s);
    cudaStreamWaitEvent(run, ev_in, 0);
  }
  if (slot.exec == nullptr || slot.seed != seed || std::memcmp(&slot.hp, hp, sizeof(*hp)) != 0) {
    if (slot.exec) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    const long long before = g_launches;
    cudaError_t e = cudaStreamBeginCapture(run, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return fail(std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
    int rc = body(run);
    e = cudaStreamEndCapture(run, &graph);
    slot.launches = g_launches - before;
    g_launches = before;
    if (rc || e != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      if (rc) return rc;
      return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(&slot.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    slot.seed = seed;
    slot.hp = *hp;
  }
  cudaError_t e = cudaGraphLaunch(slot.exec, run);
  if (e != cudaSuccess) return fail(std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
  g_launches += slot.launches;
  if (hop) {
    cudaEventRecord(ev_out, run);
    cudaStreamWaitEvent(s, ev_out, 0);
  }
  return 0;
}

// train_world_model_one_step / train_actor_and_critic_one_step as one CUDA-graph replay each. seed != 0: the step's
// sampling noise is drawn on the device inside the graph.
int Engine::train_wm_graph(const dv3_hparams* hp, uint64_t seed, cudaStream_t s) {
  return run_graphed(slot_wm, hp, seed, s, [&](cudaStream_t st) -> int {
    if (seed != 0) { bump_counter(st, noise_step); fill_noise(seed, 0, 1, true, st); }
    return train_wm(hp, st);
  });
}
int Engine::train_ac_graph(const dv3_hparams* hp, uint64_t seed, cudaStream_t s) {
  return run_graphed(slot_ac, hp, seed, s, [&](cudaStream_t st) -> int {
    if (seed != 0) fill_noise(seed, 0, 2, true, st);
    return train_ac(hp, st);
  });
}

// Data-parallel update = five graph-replayed pieces with the NCCL collectives of train_one_step.py between them:
//   0: noise + observe forward + world-model losses / BPTT      -> all-reduce(world_model grads)
//   1: world-model clip + Adam
//   2: noise + dream rollout + value targets                    -> all-gather(value targets)
//   3: percentiles, critic / actor losses and gradients          -> all-reduce(critic grads), all-reduce(actor grads)
//   4: actor and critic clip + Adam (+ critic EMA)
// Overlapped variant of piece 0 (large models), three gradient buckets in finalisation order: 5 = noise + observe forward +
// losses + head backward + reverse-time scan -> the caller starts all-reduce(head-network grads, dv3_group_buffer group 6)
// asynchronously; 6 = recurrent weight gradients, running next to that collective -> async all-reduce(group 7); 7 = encoder
// backward, next to both -> all-reduce(encoder grads, group 8).
int Engine::dp_piece(const dv3_hparams* hp, int piece, uint64_t seed, int use_graph, cudaStream_t s) {
  if (piece < 0 || piece > 7) return fail("dp_piece: piece must be in [0, 7]");
  auto body = [&](cudaStream_t st) -> int {
    int rc = 0;
    switch (piece) {
      case 0:
        if (seed != 0) { bump_counter(st, noise_step); fill_noise(seed, 0, 1, true, st); }
        rc = observe_forward(st);
        if (!rc) rc = wm_loss_backward(st);
        break;
      case 1: rc = optimizer_step(0, hp->wm_grad_clip, hp->wm_lr, hp->wm_eps, -1.f, st); break;
      case 2:
        if (seed != 0) fill_noise(seed, 0, 2, true, st);
        rc = dream_forward(hp->gamma, 1, st);
        if (!rc) rc = ac_loss_backward(hp, 1, st);
        break;
      case 3: rc = ac_loss_backward(hp, 2, st); break;
      case 5:   // piece 0 cut once more, so that the all-reduce of the head gradients runs under piece 6
        if (seed != 0) { bump_counter(st, noise_step); fill_noise(seed, 0, 1, true, st); }
        rc = observe_forward(st);
        if (!rc) rc = wm_loss_backward(st, 1);
        break;
      case 6: rc = wm_loss_backward(st, 2); break;
      case 7: rc = wm_loss_backward(st, 4); break;
      default:
        if (hp->train_actor) rc = optimizer_step(1, hp->actor_grad_clip, hp->ac_lr, hp->ac_eps, -1.f, st);
        if (!rc) rc = optimizer_step(2, hp->critic_grad_clip, critic_lr(hp), critic_eps(hp), hp->critic_ema_decay, st);
    }
    return rc;
  };
  if (!use_graph) return body(s);
  return run_graphed(slot_piece[piece], hp, seed, s, body);
}

int Engine::train_update(const dv3_hparams* hp, uint64_t seed, int use_graph, cudaStream_t s) {
  auto body = [&](cudaStream_t st) -> int {
    if (seed != 0) {
      bump_counter(st, noise_step);
      fill_noise(seed, 0, 3, true, st);
    }
    int rc = train_wm(hp, st);
    if (rc) return rc;
    return train_ac(hp, st);
  };
  if (!use_graph) return body(s);
  // The legacy default stream cannot be captured: run the graph on an engine-owned stream, ordered after / before
  // the caller's stream with events.
  cudaStream_t run = s;
  const bool hop = (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread);
  if (hop) {
    if (own_stream == nullptr) {
      if (cudaStreamCreateWithFlags(&own_stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed");
      cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_out, cudaEventDisableTiming);
    }
    run = own_stream;
    cudaEventRecord(ev_in, s);
    cudaStreamWaitEvent(run, ev_in, 0);
  }
  if (graph_exec == nullptr || graph_seed != seed || std::memcmp(&graph_hp, hp, sizeof(*hp)) != 0) {
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    const long long before = g_launches;
    cudaError_t e = cudaStreamBeginCapture(run, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) return fail(std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
    int rc = body(run);
    e = cudaStreamEndCapture(run, &graph);
    graph_launches = g_launches - before;
    g_launches = before;
    if (rc || e != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      if (rc) return rc;
      return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    }
    e = cudaGraphInstantiate(&graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    graph_seed = seed;
    graph_hp = *hp;
  }
  cudaError_t e = cudaGraphLaunch(graph_exec, run);
  if (e != cudaSuccess) return fail(std::string("cudaGraphLaunch: ") + cudaGetErrorString(e));
  g_launches += graph_launches;
  if (hop) {
    cudaEventRecord(ev_out, run);
    cudaStreamWaitEvent(s, ev_out, 0);
  }
  return 0;
}

}  // namespace dv3
