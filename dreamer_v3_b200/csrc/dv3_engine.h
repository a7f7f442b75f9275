// Engine: owns parameters / gradients / Adam slots / activations and orchestrates the hot path.
#pragma once
#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dv3.h"
#include "dv3_common.h"

namespace dv3 {

struct TensorEntry {
  std::string name;
  void* ptr;
  int dtype;  // 0 f32, 1 i32
  std::vector<int64_t> shape;
  int kind, group;
  size_t bytes;
};

// Dense(no bias) + LayerNorm(+SiLU) layer weights (components/mlp.py:43-61)
struct LnLayer {
  int in = 0, out = 0;
  float *w = nullptr, *g = nullptr, *b = nullptr;     // params
  float *dw = nullptr, *dg = nullptr, *db = nullptr;  // grads (null for the EMA critic)
};
struct DenseLayer {
  int in = 0, out = 0;
  float *w = nullptr, *bias = nullptr, *dw = nullptr, *dbias = nullptr;
};
struct Mlp {
  std::vector<LnLayer> layers;
};
// activations of an MLP instance over `rows` rows (each [rows, D] contiguous)
struct MlpAct {
  std::vector<float*> pre, act, mean, rstd;
  std::vector<void*> act16;  // bf16 twins of act (tensor-core mode, rows >= 128), written by the same LN+SiLU launch
  long long rows = 0;
};
struct ConvLayer {  // conv / convT (+LN) weights; kernel viewed as a [krows, kcols] row-major matrix
  int cin = 0, cout = 0, hin = 0;  // hin: input spatial side
  float *w = nullptr, *g = nullptr, *b = nullptr, *dw = nullptr, *dg = nullptr, *db = nullptr;
  float *pre = nullptr, *act = nullptr, *mean = nullptr, *rstd = nullptr;  // activations [n*hout*hout, cout]
  void* act16 = nullptr;  // bf16 twin of act (tensor-core mode)
};

struct Engine {
  dv3_config cfg;
  int G, D, L, Z, Zc, Zk, A, B, T, H, N, R, E, OF;  // OF = flattened obs width, R = (H+1)*N
  bool discrete, image;
  float inv_world;
  std::string err;
  cudaStream_t cur = nullptr;

  // ---- memory ----
  std::vector<void*> allocations;
  std::vector<TensorEntry> tensors;
  std::unordered_map<std::string, int> index;
  char* arena = nullptr;
  size_t arena_off = 0, arena_size = 0;
  bool dry = true;

  // optimizer groups: 0 world_model, 1 actor, 2 critic, 3 critic_ema (no gradients), 4 unused (the id of the actor + critic
  // gradient span in dv3_group_buffer), 5 disagree nets (curiosity)
  static constexpr int kGroups = 6;
  static constexpr int kDisagree = 5;
  static bool has_grad(int g) { return g == 0 || g == 1 || g == 2 || g == kDisagree; }
  float* params[kGroups] = {};
  float* grads[kGroups] = {};
  float* adam_m[kGroups] = {};
  float* adam_v[kGroups] = {};
  long long numel[kGroups] = {};
  long long grads_ac_numel = 0;   // floats from grads[1] to the end of grads[2] (one allocation, padding included)
  AdamState adam[kGroups];
  struct PInfo { std::string name; int group; std::vector<int64_t> shape; long long off; };
  std::vector<PInfo> pinfo;

  // ---- weights ----
  Mlp enc_mlp, post_mlp, dyn_mlp, pre_mlp, rew_mlp, cont_mlp, dec_mlp, actor_mlp, actor_std_mlp, critic_mlp, ema_mlp;
  DenseLayer post_repr, dyn_repr, rew_head, cont_out, dec_out, dec_dense, actor_out, actor_std_out, critic_head, ema_head;
  ConvLayer enc_conv[4], dec_convT[3];
  DenseLayer dec_out_conv;  // w = [4,4,img_c,m] viewed [16*img_c, m]; bias [img_c]
  float *initial_h = nullptr, *d_initial_h = nullptr;
  float *gru_K = nullptr, *gru_U = nullptr, *gru_b = nullptr, *d_gru_K = nullptr, *d_gru_U = nullptr, *d_gru_b = nullptr;

  // ---- inputs ----
  float *in_obs, *in_actions, *in_is_first, *in_rewards, *in_is_term, *in_u_post, *in_u_dyn, *in_act_noise;
  float *in_dream_h0, *in_dream_z0, *in_dream_actions;
  int dream_action_mode = 0;  // 0: actor everywhere, 1: a_0 given (in.dream_actions[0]), 2: every action given
  // ---- observe activations ----
  float *obs_sym, *enc_out, *a_in, *wf, *actW, *h0, *z0, *q0_pre, *q0_act, *z0_logits;
  int32_t* z0_idx;
  float *Wt_zq, *Wt_post_h, *Ut, *Kt, *Wt_pre_z, *dp_act_all, *du_all, *dh_tot;
  unsigned long long* scan_stamps = nullptr;
  unsigned int* scan_bar = nullptr;   // [2] grid-barrier counters of the persistent scan kernels (fwd, bwd)
  int scan_bwd_ctas = 0;
  int scan_ctas = 0;          // > 0: persistent cooperative scan kernels are usable with this many CTAs
  int scan_tc_ctas = 0;       // > 0: the weight-streaming tensor-core scan (dv3_scan_tc.cu) is usable with this many CTAs
  int dream_tc_ctas = 0;      // > 0: the persistent imagination-rollout kernel (dv3_dream_tc.cu) is usable with this many CTAs
  bool dream_rollout_tc(cudaStream_t s, int start_from_observe);
  bool dream_rollout_failed = false;
  // runs the H dreamed steps on it; false: not applicable (caller takes the per-kernel path)
  float* gx_acc = nullptr;    // [N, 3G] accumulator of u . K in that kernel
  const void* shadow_t(const float* w, int* ld) const;  // bf16 transposed shadow [cols, rows] of (a row sub-block of) a weight matrix
  const void* shadow_n(const float* w, int* ld) const;  // bf16 plain shadow [rows, cols] of (a row sub-block of) a weight matrix
  bool use_persistent = true;
  MlpAct enc_act;
  float *h_in, *z_in, *x_pre, *x_mean, *x_rstd, *u_act, *gx, *gh, *hz;
  float *p_pre, *p_mean, *p_rstd, *p_act, *post_logits, *post_probs;
  int32_t* z_idx;
  float *q_pre, *q_mean, *q_rstd, *q_act, *prior_logits, *prior_probs;
  MlpAct rew_act, cont_act, dec_act;
  float *rew_logits, *rew_E, *cont_logit, *obs_loc, *dec_dense_out;
  // conv scratch
  float *cols = nullptr, *conv_d0 = nullptr, *conv_d1 = nullptr, *conv_dpre = nullptr;
  void *cols16 = nullptr, *conv_dpre16 = nullptr, *dec_dense_out16 = nullptr;  // bf16 twins (tensor-core mode)
  // ---- wm losses / backward ----
  float *L_dec, *L_rew, *L_con, *L_dyn, *L_rep, *L_tot, *wm_scalars;
  int32_t* rew_k;
  float *dloc, *drew_logits, *dcont_logit, *dpost, *dprior;
  float *d_hz, *dpost_logits, *dp_pre, *dgx, *dgh, *dh_in, *dx_pre, *dprior_logits, *dq_pre, *d_enc_out, *dh0;
  float *sB_D0, *sB_D1, *sB_Z;     // per-step scratch [B, D], [B, Z]
  float *sN_D0, *sN_D1;            // [max(N,R) , max(D, ...)] scratch for MLP backward
  float *sN_D2, *sN_D3, *sN_D4, *sN_D5;  // second / third scratch pairs for chains that run concurrently
  // ---- dream ----
  float *dr_hz, *dr_act;           // [R, G+Z], [R, A]
  float *dr_x_pre, *dr_x_mean, *dr_x_rstd, *dr_u, *dr_gx, *dr_gh, *dr_q_pre, *dr_q_mean, *dr_q_rstd, *dr_q_act;
  float *dr_prior_logits, *dr_prior_probs;
  int32_t *dr_z_idx, *dr_a_idx;
  MlpAct actor_act, actor_std_act, dr_rew_act, dr_cont_act, critic_act, ema_act;
  float *actor_logits, *actor_probs, *actor_s, *dr_rew_logits, *dr_rew_E, *dr_cont_logit, *critic_logits, *v_E,
      *ema_logits, *ema_E;
  float *ac_r, *ac_c, *ac_w, *ac_v, *ac_targets, *ac_targets_all, *pct_ema, *pct_raw, *ac_scalars;
  float *dcritic_logits, *dactor_logits, *dmu, *ds, *dR, *dV, *drew_E, *dv_E, *d_dr_hz, *dr_dlogits255;
  float *L_critic_HB, *L_val_HB, *L_reg_HB, *targets_symlog, *L_actor_HB, *logp_HB, *scaled_HB, *ent_HB,
      *L_reinf_HB, *L_ent_HB;
  float *sN_Z, *sN_3G0, *sN_3G1, *sN_A;  // dream-backward scratch [N, Z], [N, 3G], [N, A]
  float *dr_dx_pre = nullptr, *dr_dact = nullptr;  // Box BPTT: per-step grads w.r.t. the pre-GRU pre-activation [H*N, D] and the actions [H*N, A]
  unsigned long long* noise_step = nullptr;
  float last_gamma = 0.997f, last_lambda = 0.95f;   // lambda of the value targets exposed after dream_forward alone: the last hparams seen

  // ---- curiosity: the Plan2Explore disagree nets (models/disagree_networks.py), cfg.num_curiosity_nets > 0 ----
  struct DisagreeNet { Mlp mlp; DenseLayer repr; MlpAct act; float* logits = nullptr; float* probs = nullptr; };
  std::vector<DisagreeNet> dis;
  int dis_ld = 0;                  // row stride of dis_in ([h, z, a] rows, padded to 8 floats)
  float *dis_in = nullptr, *dis_probs = nullptr, *dis_std = nullptr, *dis_ri = nullptr, *dis_mse = nullptr, *dis_scalars = nullptr;
  float *dis_dprobs = nullptr, *dis_dlogits = nullptr, *dis_s0 = nullptr, *dis_s1 = nullptr;
  void disagree_forward();   // every net on stop_gradient([h, z, a]) of all dreamed rows + the intrinsic rewards
  int disagree_loss_backward(cudaStream_t s);   // disagree_loss + gradients into group 5

  // ---- bf16 weight shadows for the tensor-core path (compute_mode 1) ----
  std::vector<ShadowDesc> shadow_host;
  ShadowDesc* shadow_dev = nullptr;
  int shadow_tiles = 0;
  void build_shadows();
  void refresh_shadows(cudaStream_t s);

  // ---- acting step (forward_inference) ----
  float *inf_obs, *inf_is_first, *inf_prev_h, *inf_prev_z, *inf_prev_a, *inf_u_z, *inf_act_noise;
  float *inf_h_in, *inf_z_in, *inf_a_in, *inf_x_pre, *inf_u, *inf_gx, *inf_gh, *inf_hz, *inf_obs_sym, *inf_enc, *inf_p_pre, *inf_p_act,
      *inf_logits, *inf_probs, *inf_a, *inf_actor_out, *inf_actor_probs, *inf_actor_s;
  int32_t *inf_z_idx, *inf_a_idx;
  MlpAct inf_enc_act, inf_actor_act, inf_std_act;
  unsigned long long inf_calls = 0;
  int forward_inference(int rows, uint64_t noise_seed, cudaStream_t s);
  void dream_action(int i);

  // ---- fork / join of independent op chains onto auxiliary streams (capturable: events only) ----
  static constexpr int kAux = 3;
  static constexpr int kDepth = 2;   // run_branches may nest once (a branch that forks again uses its own stream set)
  cudaStream_t aux[kDepth][kAux] = {};
  cudaEvent_t ev_fork[kDepth] = {}, ev_join[kDepth][kAux] = {};
  // bf16 conversion scratch, one per stream the engine launches on (main, kDepth x kAux branch streams, side): a large contraction
  // whose activation / gradient operand has no bf16 copy gets one made here first and then runs on the TMA-fed kernel instead of
  // the kernel that converts fp32 tiles inside the CTA (2 - 4x slower at these sizes)
  static constexpr size_t kConvScratchElems = 32u << 20;   // bf16 elements per stream (64 MB)
  unsigned short* conv_scratch[2 + kDepth * kAux] = {};
  unsigned short* conv_scratch_for(cudaStream_t s) const;   int branch_depth = 0;
  bool use_branches = true;
  // runs branches[0] on the current stream and the others concurrently on auxiliary streams, then joins them
  void run_branches(const std::vector<std::function<void()>>& branches);
  // a long side chain on its own stream (outside the pool above, so the main chain may still fork internally)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_side_in = nullptr, ev_side_out = nullptr;
  void side_run(const std::function<void()>& fn);  // forks from cur, runs fn on `side`
  void side_join();                                // cur waits for the side chain
  // ---- debug: DV3_PHASE_TIMES=1 records an event on `cur` at every labelled point of the non-graph path and prints
  // the device time between consecutive marks at the next dv3_sync ----
  bool phase_times = false, phase_env = false;
  std::vector<std::pair<std::string, cudaEvent_t>> phase_marks;
  void phase_mark(const char* label);
  void phase_dump();
  std::string phase_collect();  // accumulated per-label device times as text; clears the record

  // ---- graph ----
  struct GraphSlot { cudaGraphExec_t exec = nullptr; long long launches = 0; dv3_hparams hp{}; uint64_t seed = 0; };
  GraphSlot slot_wm, slot_ac, slot_piece[8];
  int dp_piece(const dv3_hparams* hp, int piece, uint64_t seed, int use_graph, cudaStream_t s);
  template <class F> int run_graphed(GraphSlot& slot, const dv3_hparams* hp, uint64_t seed, cudaStream_t s, F body);
  int train_wm_graph(const dv3_hparams* hp, uint64_t seed, cudaStream_t s);
  int train_ac_graph(const dv3_hparams* hp, uint64_t seed, cudaStream_t s);
  cudaGraphExec_t graph_exec = nullptr;
  long long graph_launches = 0;
  uint64_t graph_seed = 0;
  dv3_hparams graph_hp{};
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;

  // ---- methods ----
  int init(const dv3_config* c);
  void destroy();
  void layout();
  void* alloc_bytes(const std::string& name, size_t bytes, int dtype, std::vector<int64_t> shape, int kind, int group);
  float* af(const std::string& name, std::vector<int64_t> shape, int kind = 5);
  int32_t* ai(const std::string& name, std::vector<int64_t> shape, int kind = 5);
  void build_param_table();
  void bind_params();
  float* P(const std::string& name) const;
  float* dP(const std::string& name) const;
  void bind_mlp(Mlp& m, const std::string& prefix, int in_dim, int L_, bool has_grad = true);
  void bind_dense(DenseLayer& d, const std::string& prefix, int in, int out, bool has_grad = true);
  MlpAct alloc_mlp_act(const std::string& name, long long rows, int L_);

  // A16: bf16 twin of A (same element layout), valid only when the kernel that produced A wrote it in the same launch
  void mm(const float* A_, int lda, bool ta, const float* B_, int ldb, bool tb, float* C, int ldc, int M, int N_, int K,
          const float* bias = nullptr, float beta = 0.f, const void* A16 = nullptr, const void* B16 = nullptr);
  GemmArgs mm_args(const float* A_, int lda, bool ta, const float* B_, int ldb, bool tb, float* C, int ldc, int M, int N_, int K,
                   const float* bias, float beta, const void* A16, const void* B16) const;
  // would this mm() run on the TMA-fed kernel, i.e. read only the bf16 copies of its operands?
  bool mm_tma(const float* A_, int lda, bool ta, const float* B_, int ldb, bool tb, float* C, int ldc, int M, int N_, int K,
              const void* A16, const void* B16) const;
  // skip_last_ln: leave the last layer at its pre-activation (the caller's fused kernel applies LN+SiLU)
  // acc_first: the first layer's pre-activation rows are known to be zero and its contraction adds to them (a split-K
  // contraction then needs no zero-fill node in front of it)
  void mlp_fwd(const Mlp& m, MlpAct& a, const float* x, int ldx, long long rows, long long row_off = 0,
               const void* x16 = nullptr, bool skip_last_ln = false, bool acc_first = false);
  // ---- bf16 twins of activation buffers that feed large contractions (tensor-core mode only) ----
  std::unordered_map<const float*, void*> twins;
  void* twin_alloc(const float* base, size_t elems);
  void* twin_of(const float* base) const { auto it = twins.find(base); return it == twins.end() ? nullptr : it->second; }
  void *hz16 = nullptr, *dr_hz16 = nullptr, *dr_u16 = nullptr, *dr_q_act16 = nullptr;
  void mlp_bwd(const Mlp& m, MlpAct& a, const float* x, int ldx, float* dtop, float* scratch, float* dx, int lddx,
               bool wgrad, long long rows, long long row_off = 0, const void* x16 = nullptr);
  void dense_bwd(const DenseLayer& d, const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx,
                 float dx_beta, bool wgrad, long long rows);

  void encoder_fwd();
  void encoder_bwd(float* d_enc);
  void decoder_fwd();
  void decoder_bwd();

  int initial_state(cudaStream_t s);
  int observe_forward(cudaStream_t s);
  int wm_loss_backward(cudaStream_t s, int phase = 7);
  // cuts of the flat world-model gradient buffer: [0, wm_enc_end) encoder | [wm_enc_end, wm_heads_off) recurrent part |
  // [wm_heads_off, numel) reward / continue / decoder heads
  long long wm_enc_end = 0, wm_heads_off = 0;
  int optimizer_step(int group, float clip, float lr, float eps, float ema_decay, cudaStream_t s);
  int dream_forward(float gamma, int start_from_observe, cudaStream_t s);
  void actor_fwd(long long row_off);
  // actor of dream step i fused with the sequence model's input layer of step i + 1 (dv3_actor_step.cu); with_pre = false for the
  // last dreamed state, whose action is not fed back. false: not applicable (caller takes the per-op path)
  bool actor_step_applicable(int i) const;
  void actor_step_mlps(long long row_off);
  int actor_step_tail(int i, bool with_pre);
  int ac_loss_backward(const dv3_hparams* hp, int phase, cudaStream_t s);
  void actor_backward(const dv3_hparams* hp, const AcArgs& a);
  int fill_noise(uint64_t seed, uint64_t step, int which, bool device_step, cudaStream_t s);
  int train_wm(const dv3_hparams* hp, cudaStream_t s);
  int train_ac(const dv3_hparams* hp, cudaStream_t s);
  int train_update(const dv3_hparams* hp, uint64_t seed, int use_graph, cudaStream_t s);
  int fail(const std::string& m) { err = m; return -1; }
};

}  // namespace dv3
