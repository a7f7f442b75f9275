// Engine construction: parameter table (mirrors dreamer_v3_b200/spec.py), memory layout, weight binding,
// MLP / dense / conv building blocks. The hot-path orchestration lives in dv3_observe.cu and dv3_dream.cu.
#include "dv3_engine.h"
#include "dv3_dream_tc.h"

#include <map>

#include <cstdlib>

#include "dv3_scan.h"

#include <cmath>
#include <cstring>

namespace dv3 {

#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));   \
  } while (0)

static long long prod(const std::vector<int64_t>& s) {
  long long p = 1;
  for (auto v : s) p *= v;
  return p;
}

// ------------------------------------------------------------------------------------------ param table
void Engine::build_param_table() {
  pinfo.clear();
  for (int g = 0; g < kGroups; ++g) numel[g] = 0;
  auto add = [&](int group, const std::string& name, std::vector<int64_t> shape) {
    PInfo p{name, group, shape, numel[group]};
    long long n = prod(shape);
    numel[group] += (n + 3) / 4 * 4;  // keep every tensor 16-byte aligned inside the flat buffer
    pinfo.push_back(p);
  };
  auto mlp = [&](int group, const std::string& prefix, int in_dim, int L_) {
    int k = in_dim;
    for (int l = 0; l < L_; ++l) {
      const std::string b = prefix + ".mlp" + std::to_string(l);
      add(group, b + ".w", {k, D});
      add(group, b + ".g", {D});
      add(group, b + ".b", {D});
      k = D;
    }
  };
  auto dense = [&](int group, const std::string& prefix, int in, int out) {
    add(group, prefix + ".w", {in, out});
    add(group, prefix + ".bias", {out});
  };
  const int m = cfg.cnn_mult;
  if (image) {
    int cin = cfg.img_c;
    for (int l = 0; l < 4; ++l) {
      const int cout = m << l;
      const std::string b = "enc.conv" + std::to_string(l);
      add(0, b + ".w", {4, 4, cin, cout});
      add(0, b + ".g", {cout});
      add(0, b + ".b", {cout});
      cin = cout;
    }
  } else {
    mlp(0, "enc", cfg.obs_dim, L);
  }
  wm_enc_end = numel[0];
  mlp(0, "post", E + G, 1);
  dense(0, "post.repr", D, Z);
  mlp(0, "dyn", G, 1);
  dense(0, "dyn.repr", D, Z);
  add(0, "initial_h", {G});
  mlp(0, "seq.pre", Z + A, 1);
  add(0, "seq.gru.kernel", {D, 3 * G});
  add(0, "seq.gru.recurrent", {G, 3 * G});
  add(0, "seq.gru.bias", {2, 3 * G});
  wm_heads_off = numel[0];   // everything from here to the end of group 0 belongs to the reward / continue / decoder heads
  mlp(0, "rew", G + Z, L);
  dense(0, "rew.head", D, DV3_NB);
  mlp(0, "cont", G + Z, L);
  dense(0, "cont.out", D, 1);
  if (image) {
    const int c0 = 8 * m;
    dense(0, "dec.dense", G + Z, 16 * c0);
    int cin = c0;
    for (int l = 0; l < 3; ++l) {
      const int cout = cin / 2;
      const std::string b = "dec.convT" + std::to_string(l);
      add(0, b + ".w", {4, 4, cout, cin});
      add(0, b + ".g", {cout});
      add(0, b + ".b", {cout});
      cin = cout;
    }
    add(0, "dec.out.w", {4, 4, cfg.img_c, cin});
    add(0, "dec.out.bias", {cfg.img_c});
  } else {
    mlp(0, "dec", G + Z, L);
    dense(0, "dec.out", D, cfg.obs_dim);
  }
  mlp(1, "actor", G + Z, L);
  dense(1, "actor.out", D, A);
  if (!discrete) {
    mlp(1, "actor.std", G + Z, L);
    dense(1, "actor.std_out", D, A);
  }
  mlp(2, "critic", G + Z, L);
  dense(2, "critic.head", D, DV3_NB);
  mlp(3, "critic_ema", G + Z, L);
  dense(3, "critic_ema.head", D, DV3_NB);
  // curiosity: N x (MLP(L) -> RepresentationLayer) on [h, z, a] (models/disagree_networks.py:31-40)
  for (int n = 0; n < cfg.num_curiosity_nets; ++n) {
    const std::string b = "disagree.net" + std::to_string(n);
    mlp(kDisagree, b, G + Z + A, L);
    dense(kDisagree, b + ".repr", D, Z);
  }
}

float* Engine::P(const std::string& name) const {
  for (const auto& p : pinfo)
    if (p.name == name) return params[p.group] + p.off;
  fprintf(stderr, "dv3: unknown parameter %s\n", name.c_str());
  abort();
}
float* Engine::dP(const std::string& name) const {
  for (const auto& p : pinfo)
    if (p.name == name) return has_grad(p.group) ? grads[p.group] + p.off : nullptr;
  fprintf(stderr, "dv3: unknown parameter %s\n", name.c_str());
  abort();
}

void Engine::bind_mlp(Mlp& mdl, const std::string& prefix, int in_dim, int L_, bool has_grad) {
  mdl.layers.clear();
  int k = in_dim;
  for (int l = 0; l < L_; ++l) {
    const std::string b = prefix + ".mlp" + std::to_string(l);
    LnLayer ly;
    ly.in = k; ly.out = D;
    ly.w = P(b + ".w"); ly.g = P(b + ".g"); ly.b = P(b + ".b");
    if (has_grad) { ly.dw = dP(b + ".w"); ly.dg = dP(b + ".g"); ly.db = dP(b + ".b"); }
    mdl.layers.push_back(ly);
    k = D;
  }
}
void Engine::bind_dense(DenseLayer& d, const std::string& prefix, int in, int out, bool has_grad) {
  d.in = in; d.out = out;
  d.w = P(prefix + ".w"); d.bias = P(prefix + ".bias");
  if (has_grad) { d.dw = dP(prefix + ".w"); d.dbias = dP(prefix + ".bias"); }
}

void Engine::bind_params() {
  const int m = cfg.cnn_mult;
  if (image) {
    int cin = cfg.img_c, hin = cfg.img_hw;
    for (int l = 0; l < 4; ++l) {
      const std::string b = "enc.conv" + std::to_string(l);
      ConvLayer& c = enc_conv[l];
      c.cin = cin; c.cout = m << l; c.hin = hin;
      c.w = P(b + ".w"); c.g = P(b + ".g"); c.b = P(b + ".b");
      c.dw = dP(b + ".w"); c.dg = dP(b + ".g"); c.db = dP(b + ".b");
      cin = c.cout; hin /= 2;
    }
    const int c0 = 8 * m;
    bind_dense(dec_dense, "dec.dense", G + Z, 16 * c0);
    cin = c0; hin = 4;
    for (int l = 0; l < 3; ++l) {
      const std::string b = "dec.convT" + std::to_string(l);
      ConvLayer& c = dec_convT[l];
      c.cin = cin; c.cout = cin / 2; c.hin = hin;
      c.w = P(b + ".w"); c.g = P(b + ".g"); c.b = P(b + ".b");
      c.dw = dP(b + ".w"); c.dg = dP(b + ".g"); c.db = dP(b + ".b");
      cin = c.cout; hin *= 2;
    }
    dec_out_conv.in = cin; dec_out_conv.out = cfg.img_c;
    dec_out_conv.w = P("dec.out.w"); dec_out_conv.bias = P("dec.out.bias");
    dec_out_conv.dw = dP("dec.out.w"); dec_out_conv.dbias = dP("dec.out.bias");
  } else {
    bind_mlp(enc_mlp, "enc", cfg.obs_dim, L);
    bind_mlp(dec_mlp, "dec", G + Z, L);
    bind_dense(dec_out, "dec.out", D, cfg.obs_dim);
  }
  bind_mlp(post_mlp, "post", E + G, 1);
  bind_dense(post_repr, "post.repr", D, Z);
  bind_mlp(dyn_mlp, "dyn", G, 1);
  bind_dense(dyn_repr, "dyn.repr", D, Z);
  initial_h = P("initial_h"); d_initial_h = dP("initial_h");
  bind_mlp(pre_mlp, "seq.pre", Z + A, 1);
  gru_K = P("seq.gru.kernel"); d_gru_K = dP("seq.gru.kernel");
  gru_U = P("seq.gru.recurrent"); d_gru_U = dP("seq.gru.recurrent");
  gru_b = P("seq.gru.bias"); d_gru_b = dP("seq.gru.bias");
  bind_mlp(rew_mlp, "rew", G + Z, L);
  bind_dense(rew_head, "rew.head", D, DV3_NB);
  bind_mlp(cont_mlp, "cont", G + Z, L);
  bind_dense(cont_out, "cont.out", D, 1);
  bind_mlp(actor_mlp, "actor", G + Z, L);
  bind_dense(actor_out, "actor.out", D, A);
  if (!discrete) {
    bind_mlp(actor_std_mlp, "actor.std", G + Z, L);
    bind_dense(actor_std_out, "actor.std_out", D, A);
  }
  bind_mlp(critic_mlp, "critic", G + Z, L);
  bind_dense(critic_head, "critic.head", D, DV3_NB);
  bind_mlp(ema_mlp, "critic_ema", G + Z, L, false);
  bind_dense(ema_head, "critic_ema.head", D, DV3_NB, false);
  for (int n = 0; n < (int)dis.size(); ++n) {
    const std::string b = "disagree.net" + std::to_string(n);
    bind_mlp(dis[n].mlp, b, G + Z + A, L);
    bind_dense(dis[n].repr, b + ".repr", D, Z);
  }
}

// ------------------------------------------------------------------------------------------ memory
void* Engine::alloc_bytes(const std::string& name, size_t bytes, int dtype, std::vector<int64_t> shape, int kind,
                          int group) {
  arena_off = (arena_off + 255) / 256 * 256;
  void* p = dry ? nullptr : (void*)(arena + arena_off);
  arena_off += bytes;
  if (!dry && !name.empty()) {
    index[name] = (int)tensors.size();
    tensors.push_back(TensorEntry{name, p, dtype, shape, kind, group, bytes});
  }
  return p;
}
float* Engine::af(const std::string& name, std::vector<int64_t> shape, int kind) {
  return (float*)alloc_bytes(name, (size_t)prod(shape) * 4, 0, shape, kind, -1);
}
int32_t* Engine::ai(const std::string& name, std::vector<int64_t> shape, int kind) {
  return (int32_t*)alloc_bytes(name, (size_t)prod(shape) * 4, 1, shape, kind, -1);
}
void Engine::phase_mark(const char* label) {
  if (!phase_times) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(cur, &st);
  if (st != cudaStreamCaptureStatusNone) return;
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, cur);
  phase_marks.emplace_back(label, ev);
}
std::string Engine::phase_collect() {
  std::string out;
  if (phase_marks.empty()) return out;
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<double, int>> acc;
  std::vector<std::string> order;
  for (size_t i = 1; i < phase_marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, phase_marks[i - 1].second, phase_marks[i].second);
    auto& a = acc[phase_marks[i].first];
    if (a.second == 0) order.push_back(phase_marks[i].first);
    a.first += ms; a.second += 1;
  }
  char line[160];
  for (const auto& k : order) {
    snprintf(line, sizeof(line), "%s %.6f %d\n", k.c_str(), acc[k].first, acc[k].second);
    out += line;
  }
  for (auto& m : phase_marks) cudaEventDestroy(m.second);
  phase_marks.clear();
  return out;
}
void Engine::phase_dump() {
  if (!phase_env || phase_marks.empty()) { return; }
  const std::string rep = phase_collect();
  size_t pos = 0;
  while (pos < rep.size()) {
    const size_t nl = rep.find('\n', pos);
    fprintf(stderr, "[dv3 phase] %s\n", rep.substr(pos, nl - pos).c_str());
    pos = nl + 1;
  }
}

void* Engine::twin_alloc(const float* base, size_t elems) {
  if (cfg.compute_mode != 1) return nullptr;
  void* p = alloc_bytes("", elems * 2, 0, {}, 5, -1);
  if (!dry && base != nullptr) twins[base] = p;
  return p;
}
static inline void* bf16_off(void* p, size_t elems) { return p == nullptr ? nullptr : (void*)((unsigned short*)p + elems); }
static inline const void* bf16_off(const void* p, size_t elems) { return p == nullptr ? nullptr : (const void*)((const unsigned short*)p + elems); }

MlpAct Engine::alloc_mlp_act(const std::string& name, long long rows, int L_) {
  MlpAct a;
  a.rows = rows;
  for (int l = 0; l < L_; ++l) {
    const std::string b = name + ".l" + std::to_string(l);
    a.pre.push_back(af(b + ".pre", {rows, D}));
    a.act.push_back(af(b + ".act", {rows, D}));
    a.act16.push_back(rows >= 128 ? twin_alloc(a.act.back(), (size_t)rows * D) : nullptr);
    a.mean.push_back(af("", {rows}));
    a.rstd.push_back(af("", {rows}));
  }
  return a;
}

void Engine::layout() {
  const long long Nl = N, Rl = R, HN = (long long)H * N;
  const int m = cfg.cnn_mult;
  // ---- inputs ----
  if (image) in_obs = af("in.obs", {B, T, cfg.img_hw, cfg.img_hw, cfg.img_c}, 4);
  else in_obs = af("in.obs", {B, T, cfg.obs_dim}, 4);
  in_actions = af("in.actions", {B, T, A}, 4);
  in_is_first = af("in.is_first", {B, T}, 4);
  in_rewards = af("in.rewards", {B, T}, 4);
  in_is_term = af("in.is_terminated", {B, T}, 4);
  in_u_post = af("in.u_post", {T, B, Zc}, 4);
  in_u_dyn = af("in.u_dyn", {H, Nl, Zc}, 4);
  if (discrete) in_act_noise = af("in.act_noise", {H + 1, Nl}, 4);
  else in_act_noise = af("in.act_noise", {H + 1, Nl, A}, 4);
  in_dream_h0 = af("in.dream_h0", {Nl, G}, 4);
  in_dream_z0 = af("in.dream_z0", {Nl, Z}, 4);
  in_dream_actions = af("in.dream_actions", {H + 1, Nl, A}, 4);  // given actions of dream_trajectory_with_burn_in
  // ---- observe ----
  if (cfg.symlog_obs) obs_sym = af("wm.obs_symlog", {Nl, OF});
  else obs_sym = in_obs;
  if (image) {
    for (int l = 0; l < 4; ++l) {
      const long long rows = Nl * (cfg.img_hw >> (l + 1)) * (cfg.img_hw >> (l + 1));
      const int cout = m << l;
      const std::string b = "wm.enc.conv" + std::to_string(l);
      enc_conv[l].pre = af(b + ".pre", {rows, cout});
      enc_conv[l].act = af(l == 3 ? "wm.enc_out" : b + ".act", l == 3 ? std::vector<int64_t>{Nl, E} : std::vector<int64_t>{rows, cout});
      enc_conv[l].act16 = twin_alloc(enc_conv[l].act, (size_t)rows * cout);
      enc_conv[l].mean = af("", {rows});
      enc_conv[l].rstd = af("", {rows});
    }
    enc_out = enc_conv[3].act;
  } else {
    enc_act = alloc_mlp_act("wm.enc", Nl, L);
    enc_out = dry ? nullptr : enc_act.act.back();
    if (!dry) {
      index["wm.enc_out"] = (int)tensors.size();
      tensors.push_back(TensorEntry{"wm.enc_out", enc_out, 0, {Nl, E}, 5, -1, (size_t)Nl * E * 4});
    }
  }
  a_in = af("wm.a_in", {Nl, A});
  wf = af("wm.first_weight", {Nl});
  actW = af("wm.actW", {Nl, D});
  h0 = af("wm.h0", {G});
  z0 = af("wm.z0", {Z});
  q0_pre = af("", {D}); q0_act = af("", {D}); z0_logits = af("", {Z});
  z0_idx = ai("wm.z0_indices", {Zc});
  Wt_zq = af("", {Z, D}); Wt_post_h = af("", {D, G}); Ut = af("", {3 * G, G}); Kt = af("", {3 * G, D}); Wt_pre_z = af("", {D, Z});
  scan_bar = (unsigned int*)alloc_bytes("", 256, 1, {}, 6, -1);
  scan_stamps = (unsigned long long*)alloc_bytes("debug.scan_stamps", (size_t)(T * 16 + 16 + 4096) * 8, 1, {(long long)(T * 16 + 16 + 4096) * 2}, 6, -1);
  dp_act_all = af("", {Nl, D}); du_all = af("", {Nl, D}); dh_tot = af("", {B, G});
  h_in = af("wm.h_in", {Nl, G});
  z_in = af("wm.z_in", {Nl, Z});
  x_pre = af("wm.x_pre", {Nl, D}); x_mean = af("", {Nl}); x_rstd = af("", {Nl});
  u_act = af("wm.u", {Nl, D});
  gx = af("wm.gates", {Nl, 3 * G});
  gh = af("wm.gh", {Nl, 3 * G});
  if (cfg.compute_mode == 1 && G >= 512) gx_acc = af("", {Nl, 3 * G});
  hz = af("wm.hz", {Nl, G + Z});
  if (Nl >= 128) hz16 = twin_alloc(hz, (size_t)Nl * (G + Z));
  p_pre = af("wm.p_pre", {Nl, D}); p_mean = af("", {Nl}); p_rstd = af("", {Nl});
  p_act = af("wm.p_act", {Nl, D});
  post_logits = af("wm.post_logits", {Nl, Z});
  post_probs = af("wm.z_posterior_probs", {Nl, Zc, Zk});
  z_idx = ai("wm.z_indices", {Nl, Zc});
  q_pre = af("wm.q_pre", {Nl, D}); q_mean = af("", {Nl}); q_rstd = af("", {Nl});
  q_act = af("wm.q_act", {Nl, D});
  prior_logits = af("wm.prior_logits", {Nl, Z});
  prior_probs = af("wm.z_prior_probs", {Nl, Zc, Zk});
  rew_act = alloc_mlp_act("wm.rew", Nl, L);
  cont_act = alloc_mlp_act("wm.cont", Nl, L);
  rew_logits = af("wm.reward_logits", {Nl, DV3_NB});
  rew_E = af("wm.rewards", {Nl});
  cont_logit = af("wm.continue_logits", {Nl});
  obs_loc = af("wm.obs_loc", {Nl, OF});
  if (image) {
    const int c0 = 8 * m;
    dec_dense_out = af("wm.dec.dense_out", {Nl, 16 * c0});
    dec_dense_out16 = twin_alloc(dec_dense_out, (size_t)Nl * 16 * c0);
    int hin = 4, cin = c0;
    for (int l = 0; l < 3; ++l) {
      const long long rows = Nl * (2 * hin) * (2 * hin);
      const int cout = cin / 2;
      const std::string b = "wm.dec.convT" + std::to_string(l);
      dec_convT[l].pre = af(b + ".pre", {rows, cout});
      dec_convT[l].act = af(b + ".act", {rows, cout});
      dec_convT[l].act16 = twin_alloc(dec_convT[l].act, (size_t)rows * cout);
      dec_convT[l].mean = af("", {rows});
      dec_convT[l].rstd = af("", {rows});
      hin *= 2; cin = cout;
    }
    const long long cols_elems = Nl * 4096LL * m > Nl * 1024LL * 16 * cfg.img_c ? Nl * 4096LL * m : Nl * 1024LL * 16 * cfg.img_c;
    cols = af("", {cols_elems});
    cols16 = twin_alloc(cols, (size_t)cols_elems);
    const long long big = Nl * 1024LL * m > Nl * 128LL * m ? Nl * 1024LL * m : Nl * 128LL * m;
    conv_d0 = af("", {big}); conv_d1 = af("", {big}); conv_dpre = af("", {big});
    conv_dpre16 = twin_alloc(conv_dpre, (size_t)big);
  } else {
    dec_act = alloc_mlp_act("wm.dec", Nl, L);
  }
  // ---- wm losses / grads of activations ----
  L_dec = af("wm.L_decoder_B_T", {B, T}); L_rew = af("wm.L_reward_B_T", {B, T}); L_con = af("wm.L_continue_B_T", {B, T});
  L_dyn = af("wm.L_dynamics_B_T", {B, T}); L_rep = af("wm.L_representation_B_T", {B, T}); L_tot = af("wm.L_total_B_T", {B, T});
  wm_scalars = af("wm.scalars", {8});
  rew_k = ai("wm.reward_two_hot_k", {Nl});
  dloc = af("wm.d_obs_loc", {Nl, OF});
  drew_logits = af("wm.d_reward_logits", {Nl, DV3_NB});
  dcont_logit = af("wm.d_continue_logits", {Nl});
  dpost = af("wm.d_post_probs", {Nl, Z});
  dprior = af("wm.d_prior_probs", {Nl, Z});
  d_hz = af("wm.d_hz", {Nl, G + Z});
  dpost_logits = af("wm.d_post_logits", {Nl, Z});
  dp_pre = af("wm.d_p_pre", {Nl, D});
  dgx = af("wm.d_gx", {Nl, 3 * G}); dgh = af("wm.d_gh", {Nl, 3 * G});
  dh_in = af("wm.d_h_in", {Nl, G});
  dx_pre = af("wm.d_x_pre", {Nl, D});
  if (Nl >= 512) {   // bf16 twins of the scan's saved activations / gradients: operands of the N-row weight-gradient contractions
    twin_alloc(h_in, (size_t)Nl * G); twin_alloc(z_in, (size_t)Nl * Z); twin_alloc(u_act, (size_t)Nl * D); twin_alloc(p_act, (size_t)Nl * D);
    twin_alloc(dpost_logits, (size_t)Nl * Z); twin_alloc(dp_pre, (size_t)Nl * D); twin_alloc(dgx, (size_t)Nl * 3 * G);
    twin_alloc(dgh, (size_t)Nl * 3 * G); twin_alloc(dx_pre, (size_t)Nl * D);
  }
  dprior_logits = af("wm.d_prior_logits", {Nl, Z});
  dq_pre = af("wm.d_q_pre", {Nl, D});
  d_enc_out = af("wm.d_enc_out", {Nl, E});
  dh0 = af("wm.d_h0", {G});
  sB_D0 = af("", {B, D}); sB_D1 = af("", {B, D}); sB_Z = af("", {B, Z});
  sN_D0 = af("", {Rl, D}); sN_D1 = af("", {Rl, D});
  sN_D2 = af("", {Rl, D}); sN_D3 = af("", {Rl, D}); sN_D4 = af("", {Rl, D}); sN_D5 = af("", {Rl, D});
  for (float* sc : {sN_D0, sN_D1, sN_D2, sN_D3, sN_D4, sN_D5}) twin_alloc(sc, (size_t)Rl * D);
  // ---- dream ----
  dr_hz = af("dream.hz", {H + 1, Nl, G + Z});
  if (Nl >= 128) dr_hz16 = twin_alloc(dr_hz, (size_t)Rl * (G + Z));
  dr_act = af("dream.actions", {H + 1, Nl, A});
  dr_x_pre = af("", {HN, D}); dr_x_mean = af("", {HN}); dr_x_rstd = af("", {HN});
  dr_u = af("", {HN, D});
  if (Nl >= 128) dr_u16 = twin_alloc(dr_u, (size_t)HN * D);
  dr_gx = af("", {HN, 3 * G}); dr_gh = af("", {HN, 3 * G});
  dr_q_pre = af("", {HN, D}); dr_q_mean = af("", {HN}); dr_q_rstd = af("", {HN}); dr_q_act = af("", {HN, D});
  if (Nl >= 128) dr_q_act16 = twin_alloc(dr_q_act, (size_t)HN * D);
  dr_prior_logits = af("", {HN, Z});
  dr_prior_probs = af("dream.z_prior_probs", {H, Nl, Zc, Zk});
  dr_z_idx = ai("dream.z_indices", {H, Nl, Zc});
  dr_a_idx = ai("dream.action_indices", {H + 1, Nl});
  actor_act = alloc_mlp_act("dream.actor", Rl, L);
  if (!discrete) actor_std_act = alloc_mlp_act("dream.actor_std", Rl, L);
  dr_rew_act = alloc_mlp_act("dream.rew", Rl, L);
  dr_cont_act = alloc_mlp_act("dream.cont", Rl, L);
  critic_act = alloc_mlp_act("dream.critic", Rl, L);
  ema_act = alloc_mlp_act("dream.critic_ema", Rl, L);
  actor_logits = af("dream.actor_out", {H + 1, Nl, A});
  actor_probs = af("dream.actor_probs", {H + 1, Nl, A});
  actor_s = af("dream.actor_std_out", {H + 1, Nl, A});
  dr_rew_logits = af("dream.reward_logits", {Rl, DV3_NB});
  dr_rew_E = af("dream.rewards_symlog", {H + 1, Nl});
  dr_cont_logit = af("dream.continue_logits", {H + 1, Nl});
  critic_logits = af("dream.value_logits", {Rl, DV3_NB});
  v_E = af("dream.values_symlog", {H + 1, Nl});
  ema_logits = af("", {Rl, DV3_NB});
  ema_E = af("dream.values_symlog_ema", {H + 1, Nl});
  ac_r = af("dream.rewards", {H + 1, Nl}); ac_c = af("dream.continues", {H + 1, Nl});
  ac_w = af("dream.loss_weights", {H + 1, Nl}); ac_v = af("dream.values", {H + 1, Nl});
  ac_targets = af("ac.value_targets", {H, Nl});
  ac_targets_all = af("ac.value_targets_all", {(long long)cfg.world_size, H, Nl});
  pct_ema = af("ac.pct_ema", {2}, 6);
  pct_raw = af("ac.pct_raw", {2});
  ac_scalars = af("ac.scalars", {16});
  dcritic_logits = af("ac.d_value_logits", {Rl, DV3_NB});
  dactor_logits = af("ac.d_actor_out", {Rl, A});
  dmu = dactor_logits;
  ds = af("ac.d_actor_std_out", {Rl, A});
  dR = af("ac.d_targets", {H, Nl}); dV = af("ac.d_values", {H + 1, Nl});
  drew_E = af("", {Rl}); dv_E = af("", {Rl});
  d_dr_hz = af("ac.d_dream_hz", {Rl, G + Z});
  dr_dlogits255 = af("", {Rl, DV3_NB});
  L_critic_HB = af("ac.CRITIC_L_total_H_B", {H, Nl}); L_val_HB = af("ac.CRITIC_L_neg_logp_of_value_targets_H_B", {H, Nl});
  L_reg_HB = af("ac.CRITIC_L_slow_critic_regularization_H_B", {H, Nl});
  targets_symlog = af("ac.VALUE_TARGETS_symlog_H_B", {H, Nl});
  L_actor_HB = af("ac.ACTOR_L_total_H_B", {H, Nl}); logp_HB = af("ac.ACTOR_logp_actions_dreamed_H_B", {H, Nl});
  scaled_HB = af("ac.ACTOR_scaled_value_targets_H_B", {H, Nl}); ent_HB = af("ac.ACTOR_action_entropy_H_B", {H, Nl});
  L_reinf_HB = af("ac.ACTOR_L_neglogp_reinforce_term_H_B", {H, Nl}); L_ent_HB = af("ac.ACTOR_L_neg_entropy_term_H_B", {H, Nl});
  // ---- acting step ----
  if (image) inf_obs = af("inf.obs", {B, cfg.img_hw, cfg.img_hw, cfg.img_c}, 4);
  else inf_obs = af("inf.obs", {B, cfg.obs_dim}, 4);
  inf_is_first = af("inf.is_first", {B}, 4);
  inf_prev_h = af("inf.prev_h", {B, G}, 4); inf_prev_z = af("inf.prev_z", {B, Z}, 4); inf_prev_a = af("inf.prev_a", {B, A}, 4);
  inf_u_z = af("inf.u_z", {B, Zc}, 4);
  if (discrete) inf_act_noise = af("inf.act_noise", {B}, 4); else inf_act_noise = af("inf.act_noise", {B, A}, 4);
  inf_h_in = af("", {B, G}); inf_z_in = af("", {B, Z}); inf_a_in = af("", {B, A}); inf_x_pre = af("", {B, D}); inf_u = af("", {B, D});
  inf_gx = af("", {B, 3 * G}); inf_gh = af("", {B, 3 * G});
  inf_hz = af("inf.hz", {B, G + Z});
  inf_obs_sym = af("", {B, OF});
  if (!image) inf_enc_act = alloc_mlp_act("inf.enc", B, L);
  inf_enc = af("", {B, E});
  inf_p_pre = af("", {B, D}); inf_p_act = af("", {B, D}); inf_logits = af("", {B, Z});
  inf_probs = af("inf.z_probs", {B, Zc, Zk});
  inf_z_idx = ai("inf.z_indices", {B, Zc});
  inf_a = af("inf.a", {B, A}); inf_a_idx = ai("inf.action_indices", {B});
  inf_actor_act = alloc_mlp_act("inf.actor", B, L);
  if (!discrete) inf_std_act = alloc_mlp_act("inf.actor_std", B, L);
  inf_actor_out = af("inf.actor_out", {B, A}); inf_actor_probs = af("inf.actor_probs", {B, A}); inf_actor_s = af("inf.actor_std_out", {B, A});
  sN_Z = af("", {Nl, Z}); sN_3G0 = af("", {Nl, 3 * G}); sN_3G1 = af("", {Nl, 3 * G}); sN_A = af("", {Nl, A});
  if (!discrete) {
    dr_dx_pre = af("", {HN, D}); dr_dact = af("", {HN, A});
    if (Nl >= 128) twin_alloc(dr_dx_pre, (size_t)HN * D);
  }
  if (Nl >= 128) { twin_alloc(sN_Z, (size_t)Nl * Z); twin_alloc(sN_3G0, (size_t)Nl * 3 * G); twin_alloc(sN_3G1, (size_t)Nl * 3 * G); }
  // ---- curiosity (models/disagree_networks.py): inputs, per-net activations, statistics, loss / backward scratch ----
  const int Nd = cfg.num_curiosity_nets;
  dis.resize(Nd);
  if (Nd > 0) {
    dis_ld = (G + Z + A + 7) & ~7;
    dis_in = af("", {Rl, dis_ld});
    if (Rl >= 128) twin_alloc(dis_in, (size_t)Rl * dis_ld);
    dis_probs = af("disagree.z_predicted_probs", {Nd, Rl, Zc, Zk});
    for (int n = 0; n < Nd; ++n) {
      dis[n].act = alloc_mlp_act("disagree.net" + std::to_string(n), Rl, L);
      dis[n].logits = af("", {Rl, Z});
      dis[n].probs = dry ? nullptr : dis_probs + (size_t)n * Rl * Z;
    }
    dis_std = af("", {Rl});
    dis_ri = af("dream.rewards_intrinsic", {H + 1, Nl});
    dis_mse = af("", {Nd, HN});
    dis_scalars = af("disagree.scalars", {4});
    dis_dprobs = af("", {HN, Z}); dis_dlogits = af("", {HN, Z});
    dis_s0 = af("", {HN, D}); dis_s1 = af("", {HN, D});
    twin_alloc(dis_s1, (size_t)HN * D);
  }
  noise_step = (unsigned long long*)alloc_bytes("noise_step", 8, 1, {2}, 6, -1);
}

int Engine::init(const dv3_config* c) {
  cfg = *c;
  if (cfg.abi_version != DV3_ABI_VERSION) return fail("dv3_config.abi_version mismatch");
  phase_env = std::getenv("DV3_PHASE_TIMES") != nullptr;  // env: dump at every dv3_sync; otherwise dv3_phase_report switches it
  phase_times = phase_env;
  G = cfg.G; D = cfg.D; L = cfg.L; Zc = cfg.Zc; Zk = cfg.Zk; Z = Zc * Zk; A = cfg.A;
  B = cfg.B; T = cfg.T; H = cfg.H; N = B * T; R = (H + 1) * N;
  discrete = cfg.discrete != 0; image = cfg.image_obs != 0;
  if (cfg.world_size < 1) cfg.world_size = 1;
  inv_world = 1.0f / (float)cfg.world_size;
  if (G <= 0 || D <= 0 || L <= 0 || Zc <= 0 || Zk <= 0 || A <= 0 || B <= 0 || T <= 0 || H <= 0) return fail("non-positive dimension in dv3_config");
  if (Zk > 32) return fail("Zk (classes per categorical) must be <= 32");
  if (D > 1024) return fail("dense width D must be <= 1024");
  if (discrete && A > 32) return fail("Discrete action spaces with more than 32 actions are not supported");
  if (L > 8) return fail("at most 8 MLP layers");
  if (cfg.num_curiosity_nets < 0 || cfg.num_curiosity_nets > 64) return fail("num_curiosity_nets must be in [0, 64]");
  if (cfg.num_curiosity_nets > 0 && cfg.world_size > 1)
    return fail("curiosity (disagree nets) is single-process: the intrinsic rewards are centred over the whole dream batch");
  if (image) {
    if (cfg.img_hw != 64) return fail("image observations must be 64x64");
    if (8 * cfg.cnn_mult > 1024) return fail("cnn multiplier too large");
    E = 16 * 8 * cfg.cnn_mult;
    OF = cfg.img_hw * cfg.img_hw * cfg.img_c;
  } else {
    E = D;
    OF = cfg.obs_dim;
  }
  if (cfg.compute_mode != 0 && cfg.compute_mode != 1) return fail("compute_mode must be 0 (fp32) or 1 (bf16 tensor cores)");
  CK(cudaSetDevice(cfg.device));
  build_param_table();
  {
    // the three gradient buffers share one allocation, so that the actor and the critic gradients (adjacent, 256-byte aligned,
    // zero padding between them) can be all-reduced with ONE collective under data parallelism (dv3_group_buffer group 4)
    size_t off[4] = {0, 0, 0, 0};
    for (int g = 0; g < 3; ++g) off[g + 1] = off[g] + (((size_t)numel[g] * 4 + 255) / 256) * 256;
    char* gb = nullptr;
    CK(cudaMalloc((void**)&gb, off[3])); allocations.push_back(gb);
    CK(cudaMemset(gb, 0, off[3]));
    for (int g = 0; g < 3; ++g) grads[g] = reinterpret_cast<float*>(gb + off[g]);
    grads_ac_numel = (long long)((off[3] - off[1]) / 4);
  }
  for (int g = 0; g < kGroups; ++g) {
    if (numel[g] == 0) continue;
    const size_t bytes = (size_t)numel[g] * 4;
    CK(cudaMalloc((void**)&params[g], bytes)); allocations.push_back(params[g]);
    CK(cudaMemset(params[g], 0, bytes));
    if (g == kDisagree) {   // (groups 0 - 2 share the allocation above)
      CK(cudaMalloc((void**)&grads[g], bytes)); allocations.push_back(grads[g]);
      CK(cudaMemset(grads[g], 0, bytes));
    }
    if (has_grad(g)) {
      CK(cudaMalloc((void**)&adam_m[g], bytes)); allocations.push_back(adam_m[g]);
      CK(cudaMalloc((void**)&adam_v[g], bytes)); allocations.push_back(adam_v[g]);
      CK(cudaMemset(adam_m[g], 0, bytes)); CK(cudaMemset(adam_v[g], 0, bytes));
      char* st = nullptr;
      CK(cudaMalloc((void**)&st, 64)); allocations.push_back(st);
      CK(cudaMemset(st, 0, 64));
      adam[g].sumsq = (double*)st; adam[g].stats = (float*)(st + 16); adam[g].step = (int*)(st + 32);
      const char* gn[kGroups] = {"world_model", "actor", "critic", "", "", "disagree"};
      index[std::string("opt.") + gn[g] + ".stats"] = (int)tensors.size();
      tensors.push_back(TensorEntry{std::string("opt.") + gn[g] + ".stats", adam[g].stats, 0, {4}, 6, g, 16});
      index[std::string("opt.") + gn[g] + ".step"] = (int)tensors.size();
      tensors.push_back(TensorEntry{std::string("opt.") + gn[g] + ".step", adam[g].step, 1, {1}, 6, g, 4});
    }
  }
  for (const auto& p : pinfo) {
    const size_t bytes = (size_t)prod(p.shape) * 4;
    index[p.name] = (int)tensors.size();
    tensors.push_back(TensorEntry{p.name, params[p.group] + p.off, 0, p.shape, 0, p.group, bytes});
    if (has_grad(p.group)) {
      index["grad." + p.name] = (int)tensors.size();
      tensors.push_back(TensorEntry{"grad." + p.name, grads[p.group] + p.off, 0, p.shape, 1, p.group, bytes});
      index["adam_m." + p.name] = (int)tensors.size();
      tensors.push_back(TensorEntry{"adam_m." + p.name, adam_m[p.group] + p.off, 0, p.shape, 2, p.group, bytes});
      index["adam_v." + p.name] = (int)tensors.size();
      tensors.push_back(TensorEntry{"adam_v." + p.name, adam_v[p.group] + p.off, 0, p.shape, 3, p.group, bytes});
    }
  }
  dry = true; arena_off = 0;
  layout();
  arena_size = arena_off + 256;
  CK(cudaMalloc((void**)&arena, arena_size)); allocations.push_back(arena);
  CK(cudaMemset(arena, 0, arena_size));
  dry = false; arena_off = 0;
  layout();
  bind_params();
  if (cfg.compute_mode == 1) build_shadows();
  // LayerNorm gains default to 1 so that an un-initialised engine is still well defined.
  for (const auto& p : pinfo) {
    if (p.name.size() > 2 && p.name.compare(p.name.size() - 2, 2, ".g") == 0)
      fill_f32(nullptr, params[p.group] + p.off, 1.0f, prod(p.shape));
  }
  // persistent scan kernels: rows must fit one 16-row tile, classes must tile a 32-column block
  use_persistent = std::getenv("DV3_NO_PERSISTENT") == nullptr;
  use_branches = std::getenv("DV3_NO_BRANCHES") == nullptr;
  scan_ctas = 0;
  if (B <= 16 && (32 % Zk) == 0) {
    const int max_ctas = scan_fwd_max_ctas(cfg.device);
    const int tg = (3 * G + 31) / 32, td = (D + 31) / 32, tz = (Z + 31) / 32;
    int want = tg + td;
    if (tz > want) want = tz;
    if ((G + 7) / 8 > want) want = (G + 7) / 8;
    if (B > want) want = B;
    if (want < 8) want = 8;
    scan_ctas = want < max_ctas ? want : max_ctas;
    const int max_b = scan_bwd_max_ctas(cfg.device);
    const int tgb = (G + 31) / 32;
    int wb = tgb + td;
    if (tz > wb) wb = tz;
    if (B > wb) wb = B;
    if (wb < 8) wb = 8;
    scan_bwd_ctas = wb < max_b ? wb : max_b;
  }
  dream_tc_ctas = (cfg.compute_mode == 1 && !shadow_host.empty() && dream_tc_supported(N, G, D, Z, Zk, A, L)) ? dream_tc_max_ctas(cfg.device) : 0;
  if (cfg.compute_mode == 1 && !shadow_host.empty() && std::getenv("DV3_NO_CONV_SCRATCH") == nullptr)
    for (auto& p : conv_scratch) {   // (own allocations: the arena was sized by layout())
      void* q = nullptr;
      if (cudaMalloc(&q, kConvScratchElems * 2) != cudaSuccess) { cudaGetLastError(); break; }
      allocations.push_back(q);
      p = (unsigned short*)q;
    }
  scan_tc_ctas = (cfg.compute_mode == 1 && gx_acc != nullptr && B <= 16 && !shadow_host.empty()) ? scan_tc_max_ctas(cfg.device) : 0;
  const float nanv[2] = {NAN, NAN};
  CK(cudaMemcpy(pct_ema, nanv, 8, cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());
  return 0;
}

void Engine::destroy() {
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (slot_wm.exec) cudaGraphExecDestroy(slot_wm.exec);
  if (slot_ac.exec) cudaGraphExecDestroy(slot_ac.exec);
  for (auto& sl : slot_piece) if (sl.exec) cudaGraphExecDestroy(sl.exec);
  for (int d = 0; d < kDepth; ++d) {
    for (int i = 0; i < kAux; ++i) if (aux[d][i]) { cudaStreamDestroy(aux[d][i]); cudaEventDestroy(ev_join[d][i]); }
    if (ev_fork[d]) cudaEventDestroy(ev_fork[d]);
  }
  if (side) { cudaStreamDestroy(side); cudaEventDestroy(ev_side_in); cudaEventDestroy(ev_side_out); }
  if (own_stream) { cudaStreamDestroy(own_stream); cudaEventDestroy(ev_in); cudaEventDestroy(ev_out); }
  for (void* p : allocations) cudaFree(p);
  allocations.clear();
}

// ------------------------------------------------------------------------------------------ bf16 weight shadows
void Engine::build_shadows() {
  shadow_host.clear();
  shadow_tiles = 0;
  size_t bytes = 0;
  for (const auto& p : pinfo) {
    if (p.shape.size() < 2) continue;
    const long long cols = p.shape.back(), rows = prod(p.shape) / cols;
    if (rows * cols < 4096) continue;  // too small to ever reach the tensor-core path
    bytes += (size_t)((rows * cols * 2 + 255) / 256 * 256) + (size_t)((((rows + 7) / 8 * 8) * cols * 2 + 255) / 256 * 256);
  }
  char* base = nullptr;
  if (bytes == 0 || cudaMalloc((void**)&base, bytes) != cudaSuccess) return;
  allocations.push_back(base);
  cudaMemset(base, 0, bytes);   // (the padding columns of the transposed copies are never written)
  size_t off = 0;
  for (const auto& p : pinfo) {
    if (p.shape.size() < 2) continue;
    const long long cols = p.shape.back(), rows = prod(p.shape) / cols;
    if (rows * cols < 4096) continue;
    const size_t sz = (size_t)(rows * cols * 2 + 255) / 256 * 256;
    ShadowDesc d{};
    d.in = params[p.group] + p.off; d.rows = (int)rows; d.cols = (int)cols;
    d.ld_t = (int)((rows + 7) / 8 * 8);
    d.out_t = base + off; off += (size_t)((size_t)d.ld_t * cols * 2 + 255) / 256 * 256;
    d.out_n = base + off; off += sz;
    d.tile_begin = shadow_tiles;
    shadow_tiles += ((int)cols + 31) / 32 * (((int)rows + 31) / 32);
    shadow_host.push_back(d);
  }
  if (cudaMalloc((void**)&shadow_dev, shadow_host.size() * sizeof(ShadowDesc)) != cudaSuccess) { shadow_host.clear(); return; }
  allocations.push_back(shadow_dev);
  cudaMemcpy(shadow_dev, shadow_host.data(), shadow_host.size() * sizeof(ShadowDesc), cudaMemcpyHostToDevice);
}

const void* Engine::shadow_t(const float* w, int* ld) const {
  for (const ShadowDesc& d : shadow_host) {
    if (w < d.in || w >= d.in + (size_t)d.rows * d.cols) continue;
    const size_t off = (size_t)(w - d.in);
    if (off % d.cols != 0) return nullptr;   // only row sub-blocks (K offsets) are addressable in the transposed copy
    *ld = d.ld_t;
    return reinterpret_cast<const unsigned short*>(d.out_t) + off / d.cols;
  }
  return nullptr;
}

const void* Engine::shadow_n(const float* w, int* ld) const {
  for (const ShadowDesc& d : shadow_host) {
    if (w < d.in || w >= d.in + (size_t)d.rows * d.cols) continue;
    const size_t off = (size_t)(w - d.in);
    if (off % d.cols != 0) return nullptr;
    *ld = d.cols;
    return reinterpret_cast<const unsigned short*>(d.out_n) + off;
  }
  return nullptr;
}

void Engine::refresh_shadows(cudaStream_t s) {
  if (cfg.compute_mode != 1 || shadow_host.empty()) return;
  weight_shadow_batched(s, shadow_dev, (int)shadow_host.size(), shadow_tiles);
}

// ------------------------------------------------------------------------------------------ fork / join
void Engine::run_branches(const std::vector<std::function<void()>>& branches) {
  cudaStream_t main = cur;
  if (!use_branches || branches.size() <= 1 || branch_depth >= kDepth) {
    for (auto& b : branches) { cur = main; b(); }
    cur = main;
    return;
  }
  const int d = branch_depth;
  if (aux[d][0] == nullptr) {
    for (int i = 0; i < kAux; ++i) {
      cudaStreamCreateWithFlags(&aux[d][i], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&ev_join[d][i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ev_fork[d], cudaEventDisableTiming);
  }
  ++branch_depth;
  cudaEventRecord(ev_fork[d], main);
  const int n = (int)branches.size();
  for (int i = 1; i < n; ++i) {
    const int a = (i - 1) % kAux;
    if (i - 1 < kAux) cudaStreamWaitEvent(aux[d][a], ev_fork[d], 0);
    cur = aux[d][a];
    branches[i]();
  }
  cur = main;
  branches[0]();
  for (int a = 0; a < kAux && a < n - 1; ++a) {
    cudaEventRecord(ev_join[d][a], aux[d][a]);
    cudaStreamWaitEvent(main, ev_join[d][a], 0);
  }
  --branch_depth;
  cur = main;
}

void Engine::side_run(const std::function<void()>& fn) {
  if (!use_branches) { fn(); return; }
  if (side == nullptr) {
    cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ev_side_in, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ev_side_out, cudaEventDisableTiming);
  }
  cudaStream_t main = cur;
  cudaEventRecord(ev_side_in, main);
  cudaStreamWaitEvent(side, ev_side_in, 0);
  cur = side;
  fn();
  cudaEventRecord(ev_side_out, side);
  cur = main;
}
void Engine::side_join() {
  if (!use_branches || side == nullptr) return;
  cudaStreamWaitEvent(cur, ev_side_out, 0);
}

// ------------------------------------------------------------------------------------------ building blocks
unsigned short* Engine::conv_scratch_for(cudaStream_t s) const {
  for (int d = 0; d < kDepth; ++d)
    for (int i = 0; i < kAux; ++i)
      if (aux[d][i] != nullptr && s == aux[d][i]) return conv_scratch[1 + d * kAux + i];
  if (side != nullptr && s == side) return conv_scratch[1 + kDepth * kAux];
  return conv_scratch[0];
}

void Engine::mm(const float* A_, int lda, bool ta, const float* B_, int ldb, bool tb, float* C, int ldc, int M, int N_,
                int K, const float* bias, float beta, const void* A16, const void* B16) {
  GemmArgs g = mm_args(A_, lda, ta, B_, ldb, tb, C, ldc, M, N_, K, bias, beta, A16, B16);
  // Large contraction that would miss the TMA-fed kernel only because an activation / gradient operand has no bf16 copy: make
  // the copy in this stream's scratch (one extra pass over the operand) and take the TMA kernel.
  unsigned short* sc = cfg.compute_mode == 1 ? conv_scratch_for(cur) : nullptr;
  if (sc != nullptr && (long long)M * N_ * K >= (1LL << 28) && gemm_tc_eligible(g) && !gemv16_eligible(g) && !gemm_uses_tma(g, 1) &&
      ldc % 4 == 0 && (uintptr_t)C % 16 == 0) {
    const size_t half = kConvScratchElems / 2;
    if (!ta && g.Ah == nullptr && g.Bh != nullptr && g.ldbh % 8 == 0 && (uintptr_t)g.Bh % 16 == 0 && M >= 128) {
      const int ld = (K + 7) & ~7;
      if ((size_t)M * ld <= kConvScratchElems) {
        to_bf16(cur, A_, lda, sc, ld, M, K);
        g.Ah = sc; g.ldah = ld;
      }
    } else if (ta && !tb && M >= 64 && K >= 512) {
      const int lda16 = (M + 7) & ~7, ldb16 = (N_ + 7) & ~7;
      const bool fits = (g.Ah != nullptr || (size_t)K * lda16 <= half) && (g.Bmn != nullptr || (size_t)K * ldb16 <= half);
      const bool have_ok = (g.Ah == nullptr || (g.ldah % 8 == 0 && (uintptr_t)g.Ah % 16 == 0)) &&
                           (g.Bmn == nullptr || (g.ldbmn % 8 == 0 && (uintptr_t)g.Bmn % 16 == 0));
      if (fits && have_ok) {
        if (g.Ah == nullptr) { to_bf16(cur, A_, lda, sc, lda16, K, M); g.Ah = sc; g.ldah = lda16; }
        if (g.Bmn == nullptr) { to_bf16(cur, B_, ldb, sc + half, ldb16, K, N_); g.Bmn = sc + half; g.ldbmn = ldb16; }
      }
    }
  }
  gemm_dispatch(cur, g, cfg.compute_mode);
}

bool Engine::mm_tma(const float* A_, int lda, bool ta, const float* B_, int ldb, bool tb, float* C, int ldc, int M, int N_, int K,
                    const void* A16, const void* B16) const {
  return gemm_uses_tma(mm_args(A_, lda, ta, B_, ldb, tb, C, ldc, M, N_, K, nullptr, 0.f, A16, B16), cfg.compute_mode);
}

GemmArgs Engine::mm_args(const float* A_, int lda, bool ta, const float* B_, int ldb, bool tb, float* C, int ldc, int M, int N_,
                         int K, const float* bias, float beta, const void* A16, const void* B16) const {
  GemmArgs g{A_, lda, ta ? 1 : 0, B_, ldb, tb ? 1 : 0, C, ldc, M, N_, K, bias, beta};
  if (A16 != nullptr && cfg.compute_mode == 1 && (!ta || (B16 != nullptr && !tb))) { g.Ah = A16; g.ldah = lda; }
  if (B16 != nullptr && ta && !tb && cfg.compute_mode == 1) { g.Bmn = B16; g.ldbmn = ldb; }
  if (cfg.compute_mode == 1 && !shadow_host.empty() && gemm_tc_eligible(g)) {
    // is B a (sub-block of a) weight matrix with a bf16 shadow?
    for (const ShadowDesc& d : shadow_host) {
      if (B_ < d.in || B_ >= d.in + (size_t)d.rows * d.cols || ldb != d.cols) continue;
      const size_t off = (size_t)(B_ - d.in);
      const int r0 = (int)(off / d.cols), c0 = (int)(off % d.cols);
      if (!tb) {  // op(B)(k, n) = W[r0 + k][c0 + n] -> transposed shadow T[c0 + n][r0 + k]
        g.Bh = reinterpret_cast<const unsigned short*>(d.out_t) + (size_t)c0 * d.ld_t + r0;
        g.ldbh = d.ld_t;
      } else {    // op(B)(k, n) = W[r0 + n][c0 + k] -> plain shadow
        g.Bh = reinterpret_cast<const unsigned short*>(d.out_n) + off;
        g.ldbh = d.cols;
      }
      break;
    }
  }
  return g;
}

void Engine::mlp_fwd(const Mlp& mdl, MlpAct& a, const float* x, int ldx, long long rows, long long row_off,
                     const void* x16, bool skip_last_ln, bool acc_first) {
  const float* in = x;
  const void* in16 = x16;
  int ldin = ldx;
  for (size_t l = 0; l < mdl.layers.size(); ++l) {
    const LnLayer& ly = mdl.layers[l];
    float* pre = a.pre[l] + row_off * ly.out;
    float* act = a.act[l] + row_off * ly.out;
    void* act16 = rows >= 128 ? bf16_off(a.act16[l], (size_t)row_off * ly.out) : nullptr;
    mm(in, ldin, false, ly.w, ly.out, false, pre, ly.out, (int)rows, ly.out, ly.in, nullptr, (acc_first && l == 0) ? 1.f : 0.f, in16);
    if (skip_last_ln && l + 1 == mdl.layers.size()) break;
    ln_silu_fwd(cur, pre, ly.out, ly.g, ly.b, act, ly.out, a.mean[l] + row_off, a.rstd[l] + row_off, 1, rows, ly.out, act16);
    in = act; in16 = act16; ldin = ly.out;
  }
}

void Engine::mlp_bwd(const Mlp& mdl, MlpAct& a, const float* x, int ldx, float* dtop, float* scratch, float* dx,
                     int lddx, bool wgrad, long long rows, long long row_off, const void* x16) {
  float* dcur = dtop;
  float* dpre = scratch;
  void* dpre16 = rows >= 128 ? twin_of(scratch) : nullptr;  // written by the LN backward launch, read by the dX contraction
  for (int l = (int)mdl.layers.size() - 1; l >= 0; --l) {
    const LnLayer& ly = mdl.layers[l];
    const float* pre = a.pre[l] + row_off * ly.out;
    const bool need_dx = l > 0 || dx != nullptr;
    ln_silu_bwd(cur, dcur, ly.out, pre, ly.out, a.mean[l] + row_off, a.rstd[l] + row_off, ly.g, ly.b, dpre, ly.out,
                wgrad ? ly.dg : nullptr, wgrad ? ly.db : nullptr, 1, rows, ly.out, (need_dx || wgrad) ? dpre16 : nullptr);
    const float* in = (l == 0) ? x : a.act[l - 1] + row_off * mdl.layers[l - 1].out;
    const int ldin = (l == 0) ? ldx : mdl.layers[l - 1].out;
    const void* in16 = (l == 0) ? x16 : (rows >= 128 ? bf16_off((const void*)a.act16[l - 1], (size_t)row_off * mdl.layers[l - 1].out) : nullptr);
    // weight gradient X^T dY: both operands have bf16 twins written by their producers -> TMA-fed MN-major contraction
    if (wgrad) mm(in, ldin, true, dpre, ly.out, false, ly.dw, ly.out, ly.in, ly.out, (int)rows, nullptr, 1.f, dpre16 ? in16 : nullptr, dpre16);
    if (l > 0) mm(dpre, ly.out, false, ly.w, ly.out, true, dcur, ly.in, (int)rows, ly.in, ly.out, nullptr, 0.f, dpre16);
    else if (dx != nullptr) mm(dpre, ly.out, false, ly.w, ly.out, true, dx, lddx, (int)rows, ly.in, ly.out, nullptr, 1.f, dpre16);
  }
}

void Engine::dense_bwd(const DenseLayer& d, const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx,
                       float dx_beta, bool wgrad, long long rows) {
  if (wgrad) {
    mm(x, ldx, true, dy, lddy, false, d.dw, d.out, d.in, d.out, (int)rows, nullptr, 1.f);
    if (d.dbias) colsum_acc(cur, dy, lddy, rows, d.out, d.dbias);
  }
  if (dx != nullptr) mm(dy, lddy, false, d.w, d.out, true, dx, lddx, (int)rows, d.in, d.out, nullptr, dx_beta);
}

// ------------------------------------------------------------------------------------------ encoder / decoder
void Engine::encoder_fwd() {
  if (!image) {
    mlp_fwd(enc_mlp, enc_act, obs_sym, OF, N);
    return;
  }
  const float* x = obs_sym;
  for (int l = 0; l < 4; ++l) {
    ConvLayer& c = enc_conv[l];
    const int ho = c.hin / 2;
    const long long rows = (long long)N * ho * ho;
    // (the fp32 patch matrix is only written when the contraction cannot take the TMA path, which reads the bf16 one)
    const bool tma = cols16 != nullptr && mm_tma(cols, 16 * c.cin, false, c.w, c.cout, false, c.pre, c.cout, (int)rows, c.cout, 16 * c.cin, cols16, nullptr);
    im2col_k4s2(cur, x, tma ? nullptr : cols, N, c.hin, c.hin, c.cin, cols16);
    mm(cols, 16 * c.cin, false, c.w, c.cout, false, c.pre, c.cout, (int)rows, c.cout, 16 * c.cin, nullptr, 0.f, cols16);
    ln_silu_fwd(cur, c.pre, c.cout, c.g, c.b, c.act, c.cout, c.mean, c.rstd, 1, rows, c.cout, c.act16);
    x = c.act;
  }
}

void Engine::encoder_bwd(float* d_enc) {
  if (!image) {
    mlp_bwd(enc_mlp, enc_act, obs_sym, OF, d_enc, sN_D1, nullptr, 0, true, N);
    return;
  }
  float* dcur = d_enc;
  float* bufs[2] = {conv_d0, conv_d1};
  int flip = 0;
  for (int l = 3; l >= 0; --l) {
    ConvLayer& c = enc_conv[l];
    const int ho = c.hin / 2;
    const long long rows = (long long)N * ho * ho;
    ln_silu_bwd(cur, dcur, c.cout, c.pre, c.cout, c.mean, c.rstd, c.g, c.b, conv_dpre, c.cout, c.dg, c.db, 1, rows, c.cout, conv_dpre16);
    const float* xin = (l == 0) ? obs_sym : enc_conv[l - 1].act;
    const bool tma = cols16 != nullptr &&
                     mm_tma(cols, 16 * c.cin, true, conv_dpre, c.cout, false, c.dw, c.cout, 16 * c.cin, c.cout, (int)rows, cols16, conv_dpre16);
    im2col_k4s2(cur, xin, tma ? nullptr : cols, N, c.hin, c.hin, c.cin, cols16);
    mm(cols, 16 * c.cin, true, conv_dpre, c.cout, false, c.dw, c.cout, 16 * c.cin, c.cout, (int)rows, nullptr, 1.f, cols16, conv_dpre16);
    if (l > 0) {
      mm(conv_dpre, c.cout, false, c.w, c.cout, true, cols, 16 * c.cin, (int)rows, 16 * c.cin, c.cout, nullptr, 0.f, conv_dpre16);
      float* dnext = bufs[flip]; flip ^= 1;
      col2im_k4s2(cur, cols, dnext, N, c.hin, c.hin, c.cin, nullptr, 0.f);
      dcur = dnext;
    }
  }
}

void Engine::decoder_fwd() {
  if (!image) {
    mlp_fwd(dec_mlp, dec_act, hz, G + Z, N);
    mm(dec_act.act.back(), D, false, dec_out.w, OF, false, obs_loc, OF, N, OF, D, dec_out.bias);
    return;
  }
  const int c0 = 8 * cfg.cnn_mult;
  mm(hz, G + Z, false, dec_dense.w, 16 * c0, false, dec_dense_out, 16 * c0, N, 16 * c0, G + Z, dec_dense.bias);
  if (dec_dense_out16 != nullptr) to_bf16(cur, dec_dense_out, 16 * c0, dec_dense_out16, 16 * c0, N, 16 * c0);
  const float* x = dec_dense_out;
  const void* x16 = dec_dense_out16;
  for (int l = 0; l < 3; ++l) {
    ConvLayer& c = dec_convT[l];
    const long long rows_in = (long long)N * c.hin * c.hin, rows_out = rows_in * 4;
    mm(x, c.cin, false, c.w, c.cin, true, cols, 16 * c.cout, (int)rows_in, 16 * c.cout, c.cin, nullptr, 0.f, x16);
    col2im_k4s2(cur, cols, c.pre, N, 2 * c.hin, 2 * c.hin, c.cout, nullptr, 0.f);
    ln_silu_fwd(cur, c.pre, c.cout, c.g, c.b, c.act, c.cout, c.mean, c.rstd, 1, rows_out, c.cout, c.act16);
    x = c.act; x16 = c.act16;
  }
  const int mi = dec_out_conv.in, ci = dec_out_conv.out;
  mm(x, mi, false, dec_out_conv.w, mi, true, cols, 16 * ci, N * 1024, 16 * ci, mi, nullptr, 0.f, x16);
  col2im_k4s2(cur, cols, obs_loc, N, 64, 64, ci, dec_out_conv.bias, 0.5f);  // conv_transpose_atari.py:121
}

void Engine::decoder_bwd() {
  if (!image) {
    dense_bwd(dec_out, dec_act.act.back(), D, dloc, OF, sN_D0, D, 0.f, true, N);
    mlp_bwd(dec_mlp, dec_act, hz, G + Z, sN_D0, sN_D1, d_hz, G + Z, true, N);
    return;
  }
  const int mi = dec_out_conv.in, ci = dec_out_conv.out;
  colsum_acc(cur, dloc, ci, (long long)N * 4096, ci, dec_out_conv.dbias);
  im2col_k4s2(cur, dloc, cols, N, 64, 64, ci, cols16);
  const float* x3 = dec_convT[2].act;
  mm(cols, 16 * ci, true, x3, mi, false, dec_out_conv.dw, mi, 16 * ci, mi, N * 1024, nullptr, 1.f, cols16, dec_convT[2].act16);
  float* bufs[2] = {conv_d0, conv_d1};
  int flip = 0;
  float* dcur = bufs[flip]; flip ^= 1;
  mm(cols, 16 * ci, false, dec_out_conv.w, mi, false, dcur, mi, N * 1024, mi, 16 * ci, nullptr, 0.f, cols16);
  for (int l = 2; l >= 0; --l) {
    ConvLayer& c = dec_convT[l];
    const long long rows_in = (long long)N * c.hin * c.hin, rows_out = rows_in * 4;
    ln_silu_bwd(cur, dcur, c.cout, c.pre, c.cout, c.mean, c.rstd, c.g, c.b, conv_dpre, c.cout, c.dg, c.db, 1, rows_out, c.cout);
    const float* xin = (l == 0) ? dec_dense_out : dec_convT[l - 1].act;
    const void* xin16 = (l == 0) ? dec_dense_out16 : dec_convT[l - 1].act16;
    float* dnext_q = bufs[flip];
    const bool tma = cols16 != nullptr &&
                     mm_tma(cols, 16 * c.cout, true, xin, c.cin, false, c.dw, c.cin, 16 * c.cout, c.cin, (int)rows_in, cols16, xin16) &&
                     mm_tma(cols, 16 * c.cout, false, c.w, c.cin, false, dnext_q, c.cin, (int)rows_in, c.cin, 16 * c.cout, cols16, nullptr);
    im2col_k4s2(cur, conv_dpre, tma ? nullptr : cols, N, 2 * c.hin, 2 * c.hin, c.cout, cols16);
    mm(cols, 16 * c.cout, true, xin, c.cin, false, c.dw, c.cin, 16 * c.cout, c.cin, (int)rows_in, nullptr, 1.f, cols16, xin16);
    float* dnext = bufs[flip]; flip ^= 1;
    mm(cols, 16 * c.cout, false, c.w, c.cin, false, dnext, c.cin, (int)rows_in, c.cin, 16 * c.cout, nullptr, 0.f, cols16);
    dcur = dnext;
  }
  dense_bwd(dec_dense, hz, G + Z, dcur, dec_dense.out, d_hz, G + Z, 1.f, true, N);
}

int Engine::optimizer_step(int group, float clip, float lr, float eps, float ema_decay, cudaStream_t s) {
  if (!has_grad(group) || numel[group] == 0) return fail("optimizer_step: group must be 0 (world_model), 1 (actor), 2 (critic) or 5 (disagree nets, with curiosity on)");
  cur = s;
  grad_norm(s, grads[group], numel[group], adam[group]);
  float* ema = (group == 2 && ema_decay >= 0.f) ? params[3] : nullptr;   // a decay of 0.0 is a legitimate EMA (ema = w)
  adam_step(s, params[group], grads[group], adam_m[group], adam_v[group], numel[group], clip, lr, eps, adam[group], ema, ema_decay);
  return 0;
}

int Engine::fill_noise(uint64_t seed, uint64_t step, int which, bool device_step, cudaStream_t s) {
  const unsigned long long* ds_ = device_step ? noise_step : nullptr;
  const uint64_t salt = step * 7919ull;
  if (which & 1) philox_fill(s, in_u_post, (long long)T * B * Zc, seed, salt + 1, ds_, 0);
  if (which & 2) {
    philox_fill(s, in_u_dyn, (long long)H * N * Zc, seed, salt + 2, ds_, 0);
    if (discrete) philox_fill(s, in_act_noise, (long long)(H + 1) * N, seed, salt + 3, ds_, 0);
    else philox_fill(s, in_act_noise, (long long)(H + 1) * N * A, seed, salt + 3, ds_, 1);
  }
  return 0;
}

}  // namespace dv3
