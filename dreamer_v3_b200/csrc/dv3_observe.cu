// World-model half of the hot path: RSSM observe scan forward (world_model.py:183-326), the world-model
// losses (world_model_losses.py) and BPTT through the scan (replaces tape.gradient, train_one_step.py:80).
//
// Layout: every per-timestep tensor is kept B-major [B, T, width] exactly like the reference's folded
// "BxT" outputs, so a timestep slice is a strided view (row stride T*width) and nothing is transposed.
// Work that does not sit on the recurrence is hoisted out of the T loop into N = B*T-row contractions:
//   forward : encoder part of the posterior MLP, action part of the pre-GRU layer, the whole prior net
//   backward: every weight gradient (one [N]-row contraction each after the reverse loop)
#include "dv3_engine.h"
#include "dv3_scan.h"

#include <cstdlib>

namespace dv3 {

int Engine::initial_state(cudaStream_t s) {
  cur = s;
  set_ln_fast(cfg.compute_mode == 1);
  const LnLayer& Wd = dyn_mlp.layers[0];
  tanh_fwd(s, initial_h, h0, G);  // world_model.py:128
  mm(h0, G, false, Wd.w, D, false, q0_pre, D, 1, D, G);
  ln_silu_fwd(s, q0_pre, D, Wd.g, Wd.b, q0_act, D, nullptr, nullptr, 1, 1, D);
  mm(q0_act, D, false, dyn_repr.w, Z, false, z0_logits, Z, 1, Z, D, dyn_repr.bias);
  repr_fwd(s, z0_logits, Z, nullptr, 0, nullptr, 0, z0_idx, Zc, z0, Z, 1, Zc, Zk, 2);  // mode, not a sample (:130-132)
  return 0;
}

int Engine::observe_forward(cudaStream_t s) {
  cur = s;
  set_ln_fast(cfg.compute_mode == 1);
  const LnLayer& Wp = post_mlp.layers[0];
  const LnLayer& Wx = pre_mlp.layers[0];
  const LnLayer& Wd = dyn_mlp.layers[0];
  const int ldhz = T * (G + Z);

  refresh_shadows(s);
  if (cfg.symlog_obs) symlog_fwd(s, in_obs, obs_sym, (long long)N * OF);  // world_model.py:206-207
  encoder_fwd();
  scan_prepare(s, in_actions, in_is_first, a_in, wf, B, T, A);
  mm(enc_out, E, false, Wp.w, D, false, p_pre, D, N, D, E);              // hoisted: enc_t . W_post[:E]
  mm(a_in, A, false, Wx.w + (size_t)Z * D, D, false, actW, D, N, D, A);  // hoisted: a_{t-1} . W_pre[Z:]
  initial_state(s);
  phase_mark("obs.enc+prep");

  // measured on B200: the persistent kernel wins up to size L; at XL (GRU 4096) weight streaming dominates and the
  // per-op path with split-K GEMMs is faster
  auto scan_persistent = [&]() -> int {
    ScanFwdParams sp{};
    sp.B = B; sp.T = T; sp.G = G; sp.D = D; sp.Z = Z; sp.Zc = Zc; sp.Zk = Zk; sp.E = E;
    sp.W_pre = Wx.w; sp.pre_g = Wx.g; sp.pre_b = Wx.b; sp.gru_K = gru_K; sp.gru_U = gru_U; sp.gru_b = gru_b;
    sp.W_post = Wp.w; sp.post_g = Wp.g; sp.post_b = Wp.b; sp.W_zq = post_repr.w; sp.b_zq = post_repr.bias;
    sp.is_first = in_is_first; sp.actW = actW; sp.u_post = in_u_post; sp.h0 = h0; sp.z0_idx = z0_idx;
    sp.h_in = h_in; sp.z_in = z_in; sp.x_pre = x_pre; sp.x_mean = x_mean; sp.x_rstd = x_rstd; sp.u_act = u_act;
    sp.gates = gx; sp.gh = gh; sp.hz = hz; sp.p_pre = p_pre; sp.p_mean = p_mean; sp.p_rstd = p_rstd;
    sp.p_act = p_act; sp.post_logits = post_logits; sp.post_probs = post_probs; sp.z_idx = z_idx;
    sp.stamps = std::getenv("DV3_SCAN_STAMPS") ? scan_stamps : nullptr;
    sp.bar = scan_bar;
    cudaError_t ce = scan_cluster_supported(sp) ? launch_observe_scan_cluster(s, sp, cfg.compute_mode == 1) : launch_observe_scan_fwd(s, sp, scan_ctas);
    if (ce != cudaSuccess) return fail(std::string("observe scan launch: ") + cudaGetErrorString(ce));
    return 0;
  };
  // tensor-core mode, S ... XL: one persistent kernel streams the bf16 weight shadows through tcgen05 (dv3_scan_tc.cu)
  ScanTcWeights tw{};
  bool use_tc = false;
  if (use_persistent && scan_tc_ctas > 0) {
    if (scan_tc_supported(B, G, D, Z, Zk)) {
      tw.Kt = shadow_t(gru_K, &tw.ld_Kt);
      tw.Ut = shadow_t(gru_U, &tw.ld_Ut);
      tw.Wpost_h_t = shadow_t(Wp.w + (size_t)E * D, &tw.ld_Wpost_h_t);
      tw.Wzq_t = shadow_t(post_repr.w, &tw.ld_Wzq_t);
      use_tc = tw.Kt && tw.Ut && tw.Wpost_h_t && tw.Wzq_t;
    }
  }
  if (use_tc) {
    ScanFwdParams sp{};
    sp.B = B; sp.T = T; sp.G = G; sp.D = D; sp.Z = Z; sp.Zc = Zc; sp.Zk = Zk; sp.E = E;
    sp.W_pre = Wx.w; sp.pre_g = Wx.g; sp.pre_b = Wx.b; sp.gru_K = gru_K; sp.gru_U = gru_U; sp.gru_b = gru_b;
    sp.W_post = Wp.w; sp.post_g = Wp.g; sp.post_b = Wp.b; sp.W_zq = post_repr.w; sp.b_zq = post_repr.bias;
    sp.is_first = in_is_first; sp.actW = actW; sp.u_post = in_u_post; sp.h0 = h0; sp.z0_idx = z0_idx;
    sp.h_in = h_in; sp.z_in = z_in; sp.x_pre = x_pre; sp.x_mean = x_mean; sp.x_rstd = x_rstd; sp.u_act = u_act;
    sp.gates = gx; sp.gh = gh; sp.hz = hz; sp.p_pre = p_pre; sp.p_mean = p_mean; sp.p_rstd = p_rstd;
    sp.p_act = p_act; sp.post_logits = post_logits; sp.post_probs = post_probs; sp.z_idx = z_idx;
    sp.bar = scan_bar;
    sp.stamps = std::getenv("DV3_SCAN_STAMPS") ? scan_stamps : nullptr;
    cudaError_t ce = launch_observe_scan_tc(s, sp, tw, gx_acc, scan_tc_ctas);
    if (ce != cudaSuccess) return fail(std::string("observe scan (tensor-core) launch: ") + cudaGetErrorString(ce));
  } else if (use_persistent && scan_ctas > 0 && G <= 2048) {
    if (int rc = scan_persistent()) return rc;
  } else
  for (int t = 0; t < T; ++t) {
    const float* hz_prev = t > 0 ? hz + (size_t)(t - 1) * (G + Z) : nullptr;
    scan_mask_inputs(s, in_is_first + t, T, hz_prev, ldhz, hz_prev ? hz_prev + G : nullptr, ldhz, h0, z0,
                     actW + (size_t)t * D, T * D, h_in + (size_t)t * G, T * G, z_in + (size_t)t * Z, T * Z,
                     x_pre + (size_t)t * D, T * D, B, G, Z, D);
    // sequence model (sequence_model.py:56-75)
    mm(z_in + (size_t)t * Z, T * Z, false, Wx.w, D, false, x_pre + (size_t)t * D, T * D, B, D, Z, nullptr, 1.f);
    ln_silu_fwd(s, x_pre + (size_t)t * D, T * D, Wx.g, Wx.b, u_act + (size_t)t * D, T * D, x_mean + t, x_rstd + t, T, B, D);
    mm(u_act + (size_t)t * D, T * D, false, gru_K, 3 * G, false, gx + (size_t)t * 3 * G, T * 3 * G, B, 3 * G, D, gru_b);
    mm(h_in + (size_t)t * G, T * G, false, gru_U, 3 * G, false, gh + (size_t)t * 3 * G, T * 3 * G, B, 3 * G, G, gru_b + 3 * G);
    gru_fwd(s, gx + (size_t)t * 3 * G, T * 3 * G, gh + (size_t)t * 3 * G, T * 3 * G, h_in + (size_t)t * G, T * G,
            hz + (size_t)t * (G + Z), ldhz, B, G);
    // posterior (world_model.py:259-269)
    mm(hz + (size_t)t * (G + Z), ldhz, false, Wp.w + (size_t)E * D, D, false, p_pre + (size_t)t * D, T * D, B, D, G, nullptr, 1.f);
    ln_silu_fwd(s, p_pre + (size_t)t * D, T * D, Wp.g, Wp.b, p_act + (size_t)t * D, T * D, p_mean + t, p_rstd + t, T, B, D);
    mm(p_act + (size_t)t * D, T * D, false, post_repr.w, Z, false, post_logits + (size_t)t * Z, T * Z, B, Z, D, post_repr.bias);
    repr_fwd(s, post_logits + (size_t)t * Z, T * Z, post_probs + (size_t)t * Z, T * Z, in_u_post + (size_t)t * B * Zc, Zc,
             z_idx + (size_t)t * Zc, T * Zc, hz + (size_t)t * (G + Z) + G, ldhz, B, Zc, Zk, 1);
  }
  // After the scan, four independent chains over the folded B*T rows run concurrently (fork/join on auxiliary streams):
  // the prior net (world_model.py:272; its sample is discarded by the reference) and the three heads (:299-309).
  phase_mark("obs.scan");
  if (hz16 != nullptr) to_bf16(s, hz, G + Z, hz16, G + Z, N, G + Z);  // bf16 twin of [h, z] for the N-row head contractions
  run_branches({
      [&] { decoder_fwd(); },
      [&] {
        mm(hz, G + Z, false, Wd.w, D, false, q_pre, D, N, D, G, nullptr, 0.f, hz16);
        ln_silu_fwd(cur, q_pre, D, Wd.g, Wd.b, q_act, D, q_mean, q_rstd, 1, N, D);
        mm(q_act, D, false, dyn_repr.w, Z, false, prior_logits, Z, N, Z, D, dyn_repr.bias);
        repr_fwd(cur, prior_logits, Z, prior_probs, Z, nullptr, 0, nullptr, 0, nullptr, 0, N, Zc, Zk, 0);
      },
      [&] {
        mlp_fwd(rew_mlp, rew_act, hz, G + Z, N, 0, hz16);
        mm(rew_act.act.back(), D, false, rew_head.w, DV3_NB, false, rew_logits, DV3_NB, N, DV3_NB, D, rew_head.bias, 0.f, rew_act.act16.back());
        bucket_expectation(cur, rew_logits, DV3_NB, rew_E, N);
      },
      [&] {
        mlp_fwd(cont_mlp, cont_act, hz, G + Z, N, 0, hz16);
        mm(cont_act.act.back(), D, false, cont_out.w, 1, false, cont_logit, 1, N, 1, D, cont_out.bias);
      }});
  phase_mark("obs.heads");
  return 0;
}

// phase bits: 1 = losses, head backward, reverse-time scan: after it the gradients of the head networks (reward, continue,
// decoder - the tail of the flat world-model gradient buffer, dv3_group_buffer group 6) are final; 2 = the weight gradients of
// the recurrent part (posterior, GRU, pre-GRU layer, initial state: the middle of the buffer, group 7); 4 = the encoder
// backward (the front, group 8). A data-parallel caller starts the all-reduce of each bucket as soon as it is final and runs
// it under the next phase (train_one_step.py of the package). 7 = everything.
int Engine::wm_loss_backward(cudaStream_t s, int phase) {
  cur = s;
  set_ln_fast(cfg.compute_mode == 1);
  const LnLayer& Wp = post_mlp.layers[0];
  const LnLayer& Wx = pre_mlp.layers[0];
  const LnLayer& Wd = dyn_mlp.layers[0];
  const int ldhz = T * (G + Z);
  if (phase & 1) {
  cudaMemsetAsync(grads[0], 0, (size_t)numel[0] * 4, s);

  WmLossArgs a{};
  a.B = B; a.T = T; a.Z = Z; a.Zc = Zc; a.Zk = Zk; a.obs_flat = OF;
  a.inv_count = inv_world / (float)N;
  a.obs = obs_sym; a.loc = obs_loc; a.rew_logits = rew_logits; a.rewards = in_rewards; a.cont_logit = cont_logit;
  a.is_terminated = in_is_term; a.post = post_probs; a.prior = prior_probs;
  a.L_dec = L_dec; a.L_rew = L_rew; a.L_con = L_con; a.L_dyn = L_dyn; a.L_rep = L_rep; a.L_tot = L_tot;
  a.scalars = wm_scalars; a.dloc = dloc; a.drew_logits = drew_logits; a.dcont_logit = dcont_logit;
  a.dpost = dpost; a.dprior = dprior; a.rew_k = rew_k;
  phase_mark("wmb.begin");
  wm_losses(s, a);
  phase_mark("wmb.losses");

  // heads -> d_hz
  cudaMemsetAsync(d_hz, 0, (size_t)N * (G + Z) * 4, s);
  decoder_bwd();
  dense_bwd(rew_head, rew_act.act.back(), D, drew_logits, DV3_NB, sN_D0, D, 0.f, true, N);
  mlp_bwd(rew_mlp, rew_act, hz, G + Z, sN_D0, sN_D1, d_hz, G + Z, true, N, 0, hz16);
  dense_bwd(cont_out, cont_act.act.back(), D, dcont_logit, 1, sN_D0, D, 0.f, true, N);
  mlp_bwd(cont_mlp, cont_act, hz, G + Z, sN_D0, sN_D1, d_hz, G + Z, true, N, 0, hz16);
  // prior (dynamics predictor) on all rows: KL gradient w.r.t. prior probs -> d_hz (h part)
  repr_bwd(s, prior_logits, Z, nullptr, 0, dprior, Z, dprior_logits, Z, N, Zc, Zk);
  dense_bwd(dyn_repr, q_act, D, dprior_logits, Z, sN_D0, D, 0.f, true, N);
  ln_silu_bwd(s, sN_D0, D, q_pre, D, q_mean, q_rstd, Wd.g, Wd.b, dq_pre, D, Wd.dg, Wd.db, 1, N, D);
  mm(hz, G + Z, true, dq_pre, D, false, Wd.dw, D, G, D, N, nullptr, 1.f);
  mm(dq_pre, D, false, Wd.w, D, true, d_hz, G + Z, N, G, D, nullptr, 1.f);

  phase_mark("wmb.heads_bwd");
  // reverse-time scan: only input gradients on the recurrence
  // (as in the forward pass: at XL the weights no longer fit L2 and the per-op path on the weight-streaming kernels of
  //  dv3_gemv.cu is faster - C5 62.0 vs 72.7 ms per update; at L the persistent kernel still wins, 40.6 vs 42.8 ms)
  ScanTcWeightsBwd twb{};
  bool use_tc_bwd = false;
  if (use_persistent && scan_tc_ctas > 0 && scan_tc_supported(B, G, D, Z, Zk) && std::getenv("DV3_NO_SCAN_TC_BWD") == nullptr) {
    twb.Wzq = shadow_n(post_repr.w, &twb.ld_Wzq);
    twb.Wpost_h = shadow_n(Wp.w + (size_t)E * D, &twb.ld_Wpost_h);
    twb.U = shadow_n(gru_U, &twb.ld_U);
    twb.K = shadow_n(gru_K, &twb.ld_K);
    twb.Wpre_z = shadow_n(Wx.w, &twb.ld_Wpre_z);
    use_tc_bwd = twb.Wzq && twb.Wpost_h && twb.U && twb.K && twb.Wpre_z;
  }
  if (use_tc_bwd) {
    // tensor-core mode, S ... XL: one persistent kernel streams the plain bf16 weight shadows through tcgen05 (dv3_scan_tc.cu)
    cudaMemsetAsync(dp_act_all, 0, (size_t)N * D * 4, s);
    cudaMemsetAsync(du_all, 0, (size_t)N * D * 4, s);
    ScanBwdParams bp{};
    bp.B = B; bp.T = T; bp.G = G; bp.D = D; bp.Z = Z; bp.Zc = Zc; bp.Zk = Zk;
    bp.post_g = Wp.g; bp.post_b = Wp.b; bp.pre_g = Wx.g; bp.pre_b = Wx.b;
    bp.is_first = in_is_first; bp.post_logits = post_logits; bp.dpost = dpost; bp.p_pre = p_pre; bp.p_mean = p_mean;
    bp.p_rstd = p_rstd; bp.gates = gx; bp.gh = gh; bp.h_in = h_in; bp.x_pre = x_pre; bp.x_mean = x_mean; bp.x_rstd = x_rstd;
    bp.d_hz = d_hz; bp.dlogits = dpost_logits; bp.dp_act = dp_act_all; bp.dp_pre = dp_pre; bp.dh_tot = dh_tot;
    bp.dgx = dgx; bp.dgh = dgh; bp.dh_in = dh_in; bp.du = du_all; bp.dx_pre = dx_pre; bp.bar = scan_bar + 32;
    bp.stamps = std::getenv("DV3_SCAN_STAMPS") ? scan_stamps : nullptr;
    cudaError_t ce = launch_observe_scan_tc_bwd(s, bp, twb, scan_tc_ctas);
    if (ce != cudaSuccess) return fail(std::string("observe scan bwd (tensor-core) launch: ") + cudaGetErrorString(ce));
    ln_silu_bwd(s, dp_act_all, D, p_pre, D, p_mean, p_rstd, Wp.g, Wp.b, nullptr, 0, Wp.dg, Wp.db, 1, N, D);
    ln_silu_bwd(s, du_all, D, x_pre, D, x_mean, x_rstd, Wx.g, Wx.b, nullptr, 0, Wx.dg, Wx.db, 1, N, D);
  } else if (use_persistent && cfg.compute_mode == 1 && [&] {
               // tensor-core mode, XS: the 16-CTA cluster kernel (dv3_scan_cluster.cu) with bf16 weight rows resident on chip
               ScanBwdParams bp{};
               bp.B = B; bp.T = T; bp.G = G; bp.D = D; bp.Z = Z; bp.Zc = Zc; bp.Zk = Zk;
               bp.W_zq = post_repr.w; bp.W_post_h = Wp.w + (size_t)E * D; bp.gru_U = gru_U; bp.gru_K = gru_K; bp.W_pre = Wx.w;
               if (!scan_bwd_cluster_supported(bp)) return false;
               bp.post_g = Wp.g; bp.post_b = Wp.b; bp.pre_g = Wx.g; bp.pre_b = Wx.b;
               bp.is_first = in_is_first; bp.post_logits = post_logits; bp.dpost = dpost; bp.p_pre = p_pre; bp.p_mean = p_mean;
               bp.p_rstd = p_rstd; bp.gates = gx; bp.gh = gh; bp.h_in = h_in; bp.x_pre = x_pre; bp.x_mean = x_mean; bp.x_rstd = x_rstd;
               bp.d_hz = d_hz; bp.dlogits = dpost_logits; bp.dp_act = dp_act_all; bp.dp_pre = dp_pre; bp.dh_tot = dh_tot;
               bp.dgx = dgx; bp.dgh = dgh; bp.dh_in = dh_in; bp.du = du_all; bp.dx_pre = dx_pre;
               bp.stamps = std::getenv("DV3_SCAN_STAMPS") ? scan_stamps : nullptr;
               const cudaError_t ce = launch_observe_scan_bwd_cluster(s, bp);
               if (ce != cudaSuccess) { fail(std::string("observe scan bwd (cluster) launch: ") + cudaGetErrorString(ce)); return false; }
               return true;
             }()) {
    ln_silu_bwd(s, dp_act_all, D, p_pre, D, p_mean, p_rstd, Wp.g, Wp.b, nullptr, 0, Wp.dg, Wp.db, 1, N, D);
    ln_silu_bwd(s, du_all, D, x_pre, D, x_mean, x_rstd, Wx.g, Wx.b, nullptr, 0, Wx.dg, Wx.db, 1, N, D);
  } else if (use_persistent && scan_bwd_ctas > 0 && G <= 2048) {
    transpose_f32(s, post_repr.w, Wt_zq, D, Z);
    transpose_f32(s, Wp.w + (size_t)E * D, Wt_post_h, G, D);
    transpose_f32(s, gru_U, Ut, G, 3 * G);
    transpose_f32(s, gru_K, Kt, D, 3 * G);
    transpose_f32(s, Wx.w, Wt_pre_z, Z, D);
    ScanBwdParams bp{};
    bp.B = B; bp.T = T; bp.G = G; bp.D = D; bp.Z = Z; bp.Zc = Zc; bp.Zk = Zk;
    bp.Wt_zq = Wt_zq; bp.Wt_post_h = Wt_post_h; bp.Ut = Ut; bp.Kt = Kt; bp.Wt_pre_z = Wt_pre_z;
    bp.post_g = Wp.g; bp.post_b = Wp.b; bp.pre_g = Wx.g; bp.pre_b = Wx.b;
    bp.is_first = in_is_first; bp.post_logits = post_logits; bp.dpost = dpost; bp.p_pre = p_pre; bp.p_mean = p_mean;
    bp.p_rstd = p_rstd; bp.gates = gx; bp.gh = gh; bp.h_in = h_in; bp.x_pre = x_pre; bp.x_mean = x_mean; bp.x_rstd = x_rstd;
    bp.d_hz = d_hz; bp.dlogits = dpost_logits; bp.dp_act = dp_act_all; bp.dp_pre = dp_pre; bp.dh_tot = dh_tot;
    bp.dgx = dgx; bp.dgh = dgh; bp.dh_in = dh_in; bp.du = du_all; bp.dx_pre = dx_pre; bp.bar = scan_bar + 32;
    bp.stamps = std::getenv("DV3_SCAN_STAMPS") ? scan_stamps : nullptr;
    cudaError_t ce = launch_observe_scan_bwd(s, bp, scan_bwd_ctas);
    if (ce != cudaSuccess) return fail(std::string("observe scan bwd launch: ") + cudaGetErrorString(ce));
    // LayerNorm gain / offset gradients of the two in-loop layers, one N-row pass each
    ln_silu_bwd(s, dp_act_all, D, p_pre, D, p_mean, p_rstd, Wp.g, Wp.b, nullptr, 0, Wp.dg, Wp.db, 1, N, D);
    ln_silu_bwd(s, du_all, D, x_pre, D, x_mean, x_rstd, Wx.g, Wx.b, nullptr, 0, Wx.dg, Wx.db, 1, N, D);
  } else {
  // per-op path. At the sizes where it is the default (XL) the contractions stream their weights (dv3_gemv.cu) and the
  // [K, N] layout streams 1.7x faster than the [N, K] one: use the transposed copies (one transposition per update).
  const bool tp = G > 2048 || std::getenv("DV3_SCAN_BWD_TP") != nullptr;   // (env: lets the small test sizes cover this branch)
  if (tp) {
    transpose_f32(s, post_repr.w, Wt_zq, D, Z);
    transpose_f32(s, Wp.w + (size_t)E * D, Wt_post_h, G, D);
    transpose_f32(s, gru_U, Ut, G, 3 * G);
    transpose_f32(s, gru_K, Kt, D, 3 * G);
    transpose_f32(s, Wx.w, Wt_pre_z, Z, D);
  }
  for (int t = T - 1; t >= 0; --t) {
    float* dhz_t = d_hz + (size_t)t * (G + Z);
    // posterior representation layer: straight-through dz + KL dprobs -> dlogits
    repr_bwd(s, post_logits + (size_t)t * Z, T * Z, dhz_t + G, ldhz, dpost + (size_t)t * Z, T * Z,
             dpost_logits + (size_t)t * Z, T * Z, B, Zc, Zk);
    if (tp) mm(dpost_logits + (size_t)t * Z, T * Z, false, Wt_zq, D, false, sB_D0, D, B, D, Z);
    else mm(dpost_logits + (size_t)t * Z, T * Z, false, post_repr.w, Z, true, sB_D0, D, B, D, Z);
    ln_silu_bwd(s, sB_D0, D, p_pre + (size_t)t * D, T * D, p_mean + t, p_rstd + t, Wp.g, Wp.b, dp_pre + (size_t)t * D,
                T * D, Wp.dg, Wp.db, T, B, D);
    if (tp) mm(dp_pre + (size_t)t * D, T * D, false, Wt_post_h, G, false, dhz_t, ldhz, B, G, D, nullptr, 1.f);
    else mm(dp_pre + (size_t)t * D, T * D, false, Wp.w + (size_t)E * D, D, true, dhz_t, ldhz, B, G, D, nullptr, 1.f);
    // GRU
    gru_bwd(s, dhz_t, ldhz, gx + (size_t)t * 3 * G, T * 3 * G, gh + (size_t)t * 3 * G, T * 3 * G, h_in + (size_t)t * G,
            T * G, dgx + (size_t)t * 3 * G, T * 3 * G, dgh + (size_t)t * 3 * G, T * 3 * G, dh_in + (size_t)t * G, T * G, 0, B, G);
    if (tp) {
      mm(dgh + (size_t)t * 3 * G, T * 3 * G, false, Ut, G, false, dh_in + (size_t)t * G, T * G, B, G, 3 * G, nullptr, 1.f);
      mm(dgx + (size_t)t * 3 * G, T * 3 * G, false, Kt, D, false, sB_D0, D, B, D, 3 * G);
    } else {
      mm(dgh + (size_t)t * 3 * G, T * 3 * G, false, gru_U, 3 * G, true, dh_in + (size_t)t * G, T * G, B, G, 3 * G, nullptr, 1.f);
      mm(dgx + (size_t)t * 3 * G, T * 3 * G, false, gru_K, 3 * G, true, sB_D0, D, B, D, 3 * G);
    }
    ln_silu_bwd(s, sB_D0, D, x_pre + (size_t)t * D, T * D, x_mean + t, x_rstd + t, Wx.g, Wx.b, dx_pre + (size_t)t * D,
                T * D, Wx.dg, Wx.db, T, B, D);
    if (t > 0) {
      if (tp) mm(dx_pre + (size_t)t * D, T * D, false, Wt_pre_z, Z, false, sB_Z, Z, B, Z, D);
      else mm(dx_pre + (size_t)t * D, T * D, false, Wx.w, D, true, sB_Z, Z, B, Z, D);
      float* dprev = d_hz + (size_t)(t - 1) * (G + Z);
      scan_mask_grads(s, in_is_first + t, T, dh_in + (size_t)t * G, T * G, dprev, ldhz, B, G);
      scan_mask_grads(s, in_is_first + t, T, sB_Z, Z, dprev + G, ldhz, B, Z);
    }
  }
  }
  phase_mark("wmb.scan_bwd");
  }   // phase & 1
  if (phase & 2) {
  // weight gradients, one N-row contraction each; they are independent of one another -> four concurrent chains.
  // In tensor-core mode both operands of each get a bf16 copy first (the scan kernels write fp32 only), so the
  // contractions run on the TMA-fed MN-major path.
  auto tw = [&](const float* src, int cols_) -> const void* {
    void* t = twin_of(src);
    if (t != nullptr) to_bf16(cur, src, cols_, t, cols_, N, cols_);
    return t;
  };
  const void* enc16 = image ? (const void*)enc_conv[3].act16 : (enc_act.act16.empty() ? nullptr : (const void*)enc_act.act16.back());
  run_branches({
      [&] {
        const void* dpl16 = tw(dpost_logits, Z);
        const void* dpp16 = tw(dp_pre, D);
        mm(p_act, D, true, dpost_logits, Z, false, post_repr.dw, Z, D, Z, N, nullptr, 1.f, dpl16 ? tw(p_act, D) : nullptr, dpl16);
        if (post_repr.dbias) colsum_acc(cur, dpost_logits, Z, N, Z, post_repr.dbias);
        mm(enc_out, E, true, dp_pre, D, false, Wp.dw, D, E, D, N, nullptr, 1.f, dpp16 ? enc16 : nullptr, dpp16);
        mm(hz, G + Z, true, dp_pre, D, false, Wp.dw + (size_t)E * D, D, G, D, N, nullptr, 1.f, dpp16 ? hz16 : nullptr, dpp16);
      },
      [&] {
        const void* dgx16 = tw(dgx, 3 * G);
        mm(u_act, D, true, dgx, 3 * G, false, d_gru_K, 3 * G, D, 3 * G, N, nullptr, 1.f, dgx16 ? tw(u_act, D) : nullptr, dgx16);
        colsum_acc(cur, dgx, 3 * G, N, 3 * G, d_gru_b);
      },
      [&] {
        const void* dgh16 = tw(dgh, 3 * G);
        mm(h_in, G, true, dgh, 3 * G, false, d_gru_U, 3 * G, G, 3 * G, N, nullptr, 1.f, dgh16 ? tw(h_in, G) : nullptr, dgh16);
        colsum_acc(cur, dgh, 3 * G, N, 3 * G, d_gru_b + 3 * G);
      },
      [&] {
        const void* dxp16 = tw(dx_pre, D);
        mm(z_in, Z, true, dx_pre, D, false, Wx.dw, D, Z, D, N, nullptr, 1.f, dxp16 ? tw(z_in, Z) : nullptr, dxp16);
        mm(a_in, A, true, dx_pre, D, false, Wx.dw + (size_t)Z * D, D, A, D, N, nullptr, 1.f);
        // learned initial state: every is_first row (and every row at t = 0) starts from tanh(initial_h)
        mm(wf, 1, true, dh_in, G, false, dh0, G, 1, G, N);
        tanh_bwd_acc(cur, dh0, h0, d_initial_h, G);
      }});
  phase_mark("wmb.wgrads");
  }   // phase & 2
  if (!(phase & 4)) return 0;
  // encoder
  mm(dp_pre, D, false, Wp.w, D, true, d_enc_out, E, N, E, D);
  encoder_bwd(d_enc_out);
  phase_mark("wmb.encoder_bwd");
  return 0;
}

// ---------------------------------------------------------------------------------------------------- acting step
// DreamerModel.forward_inference (dreamer_model.py:113-128): one RSSM step for `rows` parallel envs + actor sample.
int Engine::forward_inference(int rows, uint64_t noise_seed, cudaStream_t s) {
  if (rows < 1 || rows > B) return fail("forward_inference: rows must be in [1, B]");
  cur = s;
  set_ln_fast(cfg.compute_mode == 1);
  const LnLayer& Wp = post_mlp.layers[0];
  const LnLayer& Wx = pre_mlp.layers[0];
  const int W = G + Z;
  if (noise_seed != 0) {
    ++inf_calls;
    philox_fill(s, inf_u_z, (long long)rows * Zc, noise_seed, 0x1000 + inf_calls * 2, nullptr, 0);
    philox_fill(s, inf_act_noise, discrete ? rows : (long long)rows * A, noise_seed, 0x1001 + inf_calls * 2, nullptr, discrete ? 0 : 1);
  }
  refresh_shadows(s);
  initial_state(s);
  // mask previous states with the learned initial state where is_first (world_model.py:160-172); actions are zeroed
  infer_mask(s, inf_is_first, inf_prev_h, inf_prev_z, inf_prev_a, h0, z0, inf_h_in, inf_z_in, inf_a_in, rows, G, Z, A);
  // sequence model
  mm(inf_z_in, Z, false, Wx.w, D, false, inf_x_pre, D, rows, D, Z);
  mm(inf_a_in, A, false, Wx.w + (size_t)Z * D, D, false, inf_x_pre, D, rows, D, A, nullptr, 1.f);
  ln_silu_fwd(s, inf_x_pre, D, Wx.g, Wx.b, inf_u, D, nullptr, nullptr, 1, rows, D);
  mm(inf_u, D, false, gru_K, 3 * G, false, inf_gx, 3 * G, rows, 3 * G, D, gru_b);
  mm(inf_h_in, G, false, gru_U, 3 * G, false, inf_gh, 3 * G, rows, 3 * G, G, gru_b + 3 * G);
  gru_fwd(s, inf_gx, 3 * G, inf_gh, 3 * G, inf_h_in, G, inf_hz, W, rows, G);
  // posterior z from the observation (compute_posterior_z, world_model.py:329-342)
  const float* x = inf_obs;
  if (cfg.symlog_obs) { symlog_fwd(s, inf_obs, inf_obs_sym, (long long)rows * OF); x = inf_obs_sym; }
  const float* enc = nullptr;
  if (!image) {
    mlp_fwd(enc_mlp, inf_enc_act, x, OF, rows);
    enc = inf_enc_act.act.back();
  } else {
    const float* cur_in = x;
    for (int l = 0; l < 4; ++l) {  // same conv stack as the training encoder, on its (larger) activation buffers
      ConvLayer& c = enc_conv[l];
      const int ho = c.hin / 2;
      const long long r = (long long)rows * ho * ho;
      im2col_k4s2(s, cur_in, cols, rows, c.hin, c.hin, c.cin);
      mm(cols, 16 * c.cin, false, c.w, c.cout, false, c.pre, c.cout, (int)r, c.cout, 16 * c.cin);
      ln_silu_fwd(s, c.pre, c.cout, c.g, c.b, c.act, c.cout, c.mean, c.rstd, 1, r, c.cout);
      cur_in = c.act;
    }
    enc = enc_conv[3].act;
  }
  mm(enc, E, false, Wp.w, D, false, inf_p_pre, D, rows, D, E);
  mm(inf_hz, W, false, Wp.w + (size_t)E * D, D, false, inf_p_pre, D, rows, D, G, nullptr, 1.f);
  ln_silu_fwd(s, inf_p_pre, D, Wp.g, Wp.b, inf_p_act, D, nullptr, nullptr, 1, rows, D);
  mm(inf_p_act, D, false, post_repr.w, Z, false, inf_logits, Z, rows, Z, D, post_repr.bias);
  repr_fwd(s, inf_logits, Z, inf_probs, Z, inf_u_z, Zc, inf_z_idx, Zc, inf_hz + G, W, rows, Zc, Zk, 1);
  // actor (actor_network.py:63-118)
  mlp_fwd(actor_mlp, inf_actor_act, inf_hz, W, rows);
  mm(inf_actor_act.act.back(), D, false, actor_out.w, A, false, inf_actor_out, A, rows, A, D, actor_out.bias);
  if (discrete) {
    disc_actor_fwd(s, inf_actor_out, inf_act_noise, inf_actor_probs, inf_a_idx, inf_a, A, rows, A);
  } else {
    mlp_fwd(actor_std_mlp, inf_std_act, inf_hz, W, rows);
    mm(inf_std_act.act.back(), D, false, actor_std_out.w, A, false, inf_actor_s, A, rows, A, D, actor_std_out.bias);
    box_actor_fwd(s, inf_actor_out, inf_actor_s, inf_act_noise, inf_a, A, rows, A);
  }
  return 0;
}

int Engine::train_wm(const dv3_hparams* hp, cudaStream_t s) {
  int rc = observe_forward(s);
  if (rc) return rc;
  rc = wm_loss_backward(s);
  if (rc) return rc;
  return optimizer_step(0, hp->wm_grad_clip, hp->wm_lr, hp->wm_eps, -1.f, s);
}

}  // namespace dv3
