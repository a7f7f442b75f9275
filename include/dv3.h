/*
 * dv3.h -- C ABI of libdv3.so, the B200-native engine for the DreamerV3 training hot path.
 *
 * The reference (sven1977/dreamer_v3) has no FFI/plugin interface: its boundary is the Python
 * API. Each entry point below replaces one reference function; the thin
 * Python mirror in dreamer_v3_b200/ binds these with ctypes and re-exposes the reference's
 * WorldModel / DreamerModel / train_*_one_step surface.
 *
 * Conventions
 *   - plain C types only; every call returns 0 on success, <0 on error (dv3_last_error()).
 *   - all tensors are fp32 row-major unless noted; "rows BxT" are B-major (row = b*T + t),
 *     dream tensors are time-major (row = i*N + n, N = B*T), as in the reference.
 *   - the engine owns parameters, gradients, Adam slots, activations and I/O buffers; they are
 *     looked up by name (dv3_tensor_info) so the caller can wrap them zero-copy (DLPack /
 *     __cuda_array_interface__) or fill them with dv3_write_tensor / read with dv3_read_tensor.
 *   - sampling noise is an explicit input ("in.u_post", "in.u_dyn", "in.act_noise"), or is
 *     generated on the device by dv3_fill_noise (counter-based Philox).
 *   - every call enqueues on the caller's cudaStream_t (passed as void*); no hidden syncs
 *     except dv3_read_tensor / dv3_write_tensor with a host pointer and sync=1.
 *   - a handle is not thread-safe; one handle per device / rank.
 */
#ifndef DV3_H_
#define DV3_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DV3_ABI_VERSION 1
#define DV3_NUM_BUCKETS 255

typedef struct dv3_engine dv3_engine; /* opaque */

/* Mirrors utils/model_dimensions.py (sizes) + run_experiment.py:131-148 (B, T, H). */
typedef struct dv3_config {
  int32_t abi_version;   /* DV3_ABI_VERSION */
  int32_t G;             /* GRU units            get_gru_units            */
  int32_t D;             /* dense hidden units   get_dense_hidden_units   */
  int32_t L;             /* MLP layers           get_num_dense_layers     */
  int32_t cnn_mult;      /* CNN multiplier       get_cnn_multiplier       */
  int32_t Zc, Zk;        /* categoricals x classes (Zk <= 32)             */
  int32_t A;             /* action dim (Discrete.n or prod(Box.shape))    */
  int32_t discrete;      /* 1 = Discrete, 0 = Box                          */
  int32_t image_obs;     /* 1 = 64x64xC images (CNNAtari/ConvTransposeAtari), 0 = vector */
  int32_t obs_dim;       /* vector observation width                       */
  int32_t img_hw, img_c; /* image side (64) and channels                   */
  int32_t B, T, H;       /* LOCAL replay batch rows, batch length, dream horizon */
  int32_t symlog_obs;    /* world_model.py:206-207                         */
  int32_t compute_mode;  /* 0 = fp32 SIMT everywhere; 1 = bf16 tcgen05 tensor cores for large contractions */
  int32_t world_size;    /* data-parallel ranks (losses are normalised by world_size*B) */
  int32_t rank;
  int32_t device;        /* CUDA device ordinal */
  int32_t num_curiosity_nets; /* 0 = use_curiosity False; else the N of DisagreeNetworks (models/disagree_networks.py:23-40,
                                 utils/model_dimensions.py:69-78); single-process only (world_size == 1) */
  float intrinsic_rewards_scale; /* DreamerModel(intrinsic_rewards_scale=0.1) (dreamer_model.py:33) */
  int32_t reserved[6];
} dv3_config;

typedef struct dv3_tensor_desc {
  const char* name;  /* owned by the engine */
  void* ptr;         /* device pointer */
  int32_t dtype;     /* 0 = f32, 1 = i32 */
  int32_t ndim;
  int64_t shape[6];
  int32_t kind;      /* 0 param, 1 grad, 2 adam_m, 3 adam_v, 4 input, 5 output/activation, 6 scalar state */
  int32_t group;     /* 0 world_model, 1 actor, 2 critic, 3 critic_ema, 5 disagree nets, -1 n/a */
} dv3_tensor_desc;

/* Hyper-parameters of the two update steps (run_experiment.py:145-148, 245-252; world_model.py:124;
 * actor_network.py:60; critic_network.py:23,58). dv3_default_hparams fills the reference values. */
typedef struct dv3_hparams {
  float gamma, lambda_, entropy_scale, return_normalization_decay;
  float wm_grad_clip, actor_grad_clip, critic_grad_clip;
  float wm_lr, wm_eps, ac_lr, ac_eps;
  float critic_ema_decay;
  int32_t train_actor;
  float critic_lr, critic_eps;  /* the critic's own Adam (critic_network.py:58); 0 = use ac_lr / ac_eps */
  /* curiosity (num_curiosity_nets > 0): disagree_grad_clip (run_experiment.py:252), the disagree nets' Adam
   * (disagree_networks.py:43) */
  float disagree_grad_clip, disagree_lr, disagree_eps;
  int32_t reserved[2];
} dv3_hparams;

/* ---- lifetime ----------------------------------------------------------------------------- */
int dv3_abi_version(void);
int dv3_create(const dv3_config* cfg, dv3_engine** out); /* replaces WorldModel.__init__ (world_model.py:25-124) + DreamerModel.__init__ (dreamer_model.py:23-67) */
int dv3_destroy(dv3_engine* e);
const char* dv3_last_error(const dv3_engine* e); /* e may be NULL: last create() error */
void dv3_default_hparams(dv3_hparams* hp);
int dv3_sync(dv3_engine* e, void* stream);

/* ---- tensor table ------------------------------------------------------------------------- */
int dv3_num_tensors(const dv3_engine* e);
int dv3_tensor_info(const dv3_engine* e, int index, dv3_tensor_desc* out);
int dv3_find_tensor(const dv3_engine* e, const char* name, dv3_tensor_desc* out);
/* copy nbytes between a host (or device) buffer and a named engine tensor; host_is_device=1 for D2D */
int dv3_write_tensor(dv3_engine* e, const char* name, const void* src, size_t nbytes, int src_is_device, void* stream, int sync);
int dv3_read_tensor(dv3_engine* e, const char* name, void* dst, size_t nbytes, int dst_is_device, void* stream, int sync);
/* flat per-group views for the data-parallel all-reduce: kind in {0 param,1 grad,2 m,3 v}; group 4 with kind 1 = the actor and
 * critic gradient buffers as ONE span (adjacent in memory, zero padding between them): one collective for both; group 5 = the
 * disagree nets (curiosity); groups 6 / 7 / 8 with kind 1 = the world-model gradient buffer as three buckets in the order the
 * backward pass finalises them: 6 = the head networks (reward, continue, decoder: the tail of the buffer, final after
 * dv3_dp_piece 5), 7 = the recurrent part (posterior, prior, sequence model; after piece 6), 8 = the encoder (after piece 7) */
int dv3_group_buffer(const dv3_engine* e, int group, int kind, void** ptr, int64_t* numel);

/* ---- noise ------------------------------------------------------------------------------- */
/* Fills in.u_post / in.u_dyn / in.act_noise on the device (uniform [0,1) or N(0,1) for Box actions). */
int dv3_fill_noise(dv3_engine* e, uint64_t seed, uint64_t step, int which /*1=u_post,2=dream,3=both*/, void* stream);

/* ---- world model -------------------------------------------------------------------------- */
/* WorldModel.get_initial_state (world_model.py:127-134): fills out.h0 [G], out.z0 [Zc*Zk]. */
int dv3_initial_state(dv3_engine* e, void* stream);
/* WorldModel.forward_train (world_model.py:183-326). Reads in.obs [B,T,...], in.actions [B,T,A],
 * in.is_first [B,T], in.u_post [T,B,Zc]; fills the wm.* outputs (see DESIGN.md tensor table). */
int dv3_observe_forward(dv3_engine* e, void* stream);
/* world_model_prediction_losses + world_model_dynamics_and_representation_loss + the loss
 * reduction of train_world_model_one_step (world_model_losses.py:15-151, train_one_step.py:48-77),
 * then BPTT through forward_train into the world_model gradient buffer.
 * Reads in.rewards [B,T], in.is_terminated [B,T]. */
int dv3_wm_loss_backward(dv3_engine* e, void* stream);
/* tf.clip_by_global_norm + Keras Adam (+ critic EMA for group 2) on a group's flat buffers
 * (train_one_step.py:80-86, 214-250; group 5: the disagree nets, :226-246). Call after the gradient all-reduce when world_size > 1. */
int dv3_optimizer_step(dv3_engine* e, int group, float clip, float lr, float eps, float ema_decay, void* stream); /* ema_decay < 0: no EMA update (any value in [0, 1] is a decay, 0 included) */
/* One call = train_world_model_one_step (train_one_step.py:17-126) for world_size == 1. noise_seed != 0: in.u_post is
 * drawn on the device first; use_graph: the call is captured once into a CUDA graph and replayed. */
int dv3_train_world_model_step(dv3_engine* e, const dv3_hparams* hp, uint64_t noise_seed, int use_graph, void* stream);

/* Acting step, DreamerModel.forward_inference = WorldModel.forward_inference + actor (world_model.py:141-180,
 * 329-342; dreamer_model.py:113-128; called once per env step by env_runner_v2.py:274-282) for `rows` <= B parallel
 * envs. Reads inf.obs, inf.is_first, inf.prev_h, inf.prev_z, inf.prev_a, inf.u_z [rows,Zc], inf.act_noise; writes inf.h,
 * inf.z (one-hot), inf.a (and inf.z_indices, inf.z_probs). noise_seed != 0 draws the noise on the device. */
int dv3_forward_inference(dv3_engine* e, int rows, uint64_t noise_seed, void* stream);

/* ---- dreamer / actor / critic --------------------------------------------------------------- */
/* DreamerModel.dream_trajectory (dreamer_model.py:147-303). start states: start_from_observe=1 uses
 * wm.hz of the last dv3_observe_forward (as run_experiment.py:458-460 does); otherwise in.dream_h0
 * [N,G] / in.dream_z0 [N,Zc*Zk]. Reads in.is_terminated, in.u_dyn [H,N,Zc], in.act_noise. */
int dv3_dream_forward(dv3_engine* e, float gamma, int start_from_observe, void* stream);
/* The prior rollout of DreamerModel.dream_trajectory_with_burn_in (dreamer_model.py:337-372; evaluation path,
 * run_experiment.py:629-644) on the same kernels: action_mode 1 = a_0 is in.dream_actions[0] (the last burn-in action)
 * and the actor samples a_1..a_H; 2 = every action is read from in.dream_actions [H+1,N,A] (use_sampled_actions_in_dream /
 * use_random_actions_in_dream); 0 = dv3_dream_forward. Rows are independent: a caller with fewer than N start states
 * fills the first rows and ignores the rest. The burn-in steps themselves are dv3_forward_inference calls. */
int dv3_dream_forward_actions(dv3_engine* e, float gamma, int start_from_observe, int action_mode, void* stream);
/* compute_value_targets + critic_loss + actor_loss (critic_loss.py:15-139, actor_loss.py:12-138) and
 * their gradients into the actor / critic gradient buffers (REINFORCE for Discrete, BPTT through the
 * dream for Box). phase 1 = value targets only (so the caller can all-gather them across ranks for
 * the percentile), phase 2 = the rest, 3 = both. */
int dv3_ac_loss_backward(dv3_engine* e, const dv3_hparams* hp, int phase, void* stream);
/* Curiosity (use_curiosity=True; models/disagree_networks.py:47-95, losses/disagree_loss.py:11-41, train_one_step.py:180-246).
 * With num_curiosity_nets > 0, dv3_dream_forward also runs every disagree net on stop_gradient([h, z, a]) of all (H+1)*N dreamed
 * rows (disagree.z_predicted_probs [nets, (H+1)N, Zc*Zk]) and fills dream.rewards_intrinsic [H+1, N] (row 0 unused:
 * rewards_intrinsic_t1_to_H_B is rows 1..H), which dv3_ac_loss_backward adds to the dreamed rewards inside the lambda-returns
 * (critic_loss.py:118-120). dv3_disagree_loss_backward: disagree_loss (disagree.scalars[0] = DISAGREE_L_total, [1] = mean
 * intrinsic reward) and its gradients into group 5; dv3_train_actor_critic_step / dv3_train_update do all of it and step group 5
 * with hp->disagree_grad_clip / disagree_lr / disagree_eps. */
int dv3_disagree_loss_backward(dv3_engine* e, void* stream);
/* CriticNetwork.init_ema / update_ema (critic_network.py:84-100). */
int dv3_critic_init_ema(dv3_engine* e, void* stream);
int dv3_critic_update_ema(dv3_engine* e, float decay, void* stream);
/* One call = train_actor_and_critic_one_step (train_one_step.py:130-252) for world_size == 1 (dreams from the states of
 * the last observe pass; same noise_seed / use_graph meaning as above). */
int dv3_train_actor_critic_step(dv3_engine* e, const dv3_hparams* hp, uint64_t noise_seed, int use_graph, void* stream);

/* ---- whole update captured as CUDA graphs ---------------------------------------------------- */
/* Captures {wm step, ac step} once into a CUDA graph on `stream` and replays it: one launch per
 * update instead of thousands. noise_seed != 0: noise is regenerated on the device each replay. */
int dv3_train_update(dv3_engine* e, const dv3_hparams* hp, uint64_t noise_seed, int use_graph, void* stream);
/* Data-parallel update (world_size > 1): the same work cut at the collectives into five graph-replayed pieces;
 * piece 0: observe + WM losses/BPTT [all-reduce WM grads] 1: WM clip+Adam 2: dream + value targets [all-gather targets]
 * 3: critic/actor losses + gradients [all-reduce critic, actor grads] 4: actor/critic clip+Adam (+EMA).
 * Pieces 5 + 6 + 7 = piece 0 cut at the points where a gradient bucket becomes final, so that its all-reduce runs under the
 * rest of the backward pass (SURVEY 8e "bucketed + overlapped"):
 * 5: observe + WM losses + head backward + reverse-time scan [start all-reduce of group 6 asynchronously]
 * 6: recurrent weight gradients [start all-reduce of group 7 asynchronously] 7: encoder backward [all-reduce group 8, wait]. */
int dv3_dp_piece(dv3_engine* e, const dv3_hparams* hp, int piece, uint64_t noise_seed, int use_graph, void* stream);
/* ---- replay sampler: the producer of the [B,T,...] sample dict ------------------------- */
/* Row packing of EpisodeReplayBuffer.sample (utils/episode_replay_buffer.py:98-211) over an HBM-resident paged episode
 * store. obs_store [slots, obs_elems] uint8 (obs_is_u8; normalize_u8: x/128-1 as env_runner_v2.py:53-54) or float32;
 * act_store [slots, A]; rew_store [slots]; plan int32 [n,4] = {obs_slot, act_slot, rew_slot (-1 => 0.0), flags
 * (1 is_first | 2 is_last | 4 is_terminated)} for the n = B*T output timesteps, built on the host from the reference's
 * index arithmetic (:121-158: reward shift, last-action repeat, episode-boundary packing). All pointers are device
 * pointers; outputs are float32 [n, ...] (typically the engine's in.obs / in.actions / in.rewards / in.is_first /
 * in.is_terminated tensors). */
int dv3_replay_gather(const void* obs_store, int obs_is_u8, int normalize_u8, int64_t obs_elems, const float* act_store, int A,
                      const float* rew_store, const int32_t* plan, int64_t n, float* obs_out, float* act_out, float* rew_out,
                      float* is_first, float* is_last, float* is_terminated, void* stream);
/* Host side of the same call: builds `plan` [B*T, 4] from the episode tables with the reference's sampling loop
 * (:113-189). chunk_* [n_chunks]: the arrival-ordered index space (`_indices` of the reference, run-length encoded:
 * chunk_cum[c] = first flat index of chunk c, chunk_ep / chunk_ts0 = its episode (row of the ep_* tables) and first
 * timestep); ep_len / ep_term / ep_page_off [episodes], pages: the page lists, `page` timesteps per page. `draws`: the
 * caller's np.random.randint(0, n_indices, size=B*10) blocks, consumed one per segment as the reference does. Returns
 * the number of draws consumed, or -(k) when more than the k-1 given draws are needed (append a block and call again). */
int64_t dv3_replay_plan(int B, int T, int page, const int64_t* chunk_cum, const int32_t* chunk_ep, const int32_t* chunk_ts0, int64_t n_chunks,
                        const int32_t* ep_len, const uint8_t* ep_term, const int64_t* ep_page_off, const int32_t* pages,
                        const int64_t* draws, int64_t n_draws, int32_t* plan);
/* Host-side position of the acting step's device noise stream (the number of dv3_forward_inference calls that drew noise), so
 * that a checkpoint can carry it and a resumed run (run_experiment.py:224-233) draws what the uninterrupted one would have. */
int64_t dv3_get_inference_calls(const dv3_engine* e);
int dv3_set_inference_calls(dv3_engine* e, int64_t calls);
int64_t dv3_kernel_launch_count(const dv3_engine* e); /* kernels of this library launched (or replayed) so far */
/* Measurement aid of bench.py (no reference counterpart): enable = 1 / 0 switches the recording of CUDA events at the
 * labelled phase boundaries of the non-graph path (encoder | observe scan | heads | losses | head backward | backward
 * scan | weight gradients | ...); enable = -1 leaves the switch alone. Synchronises the device and writes the device time
 * accumulated per label since the last report as "label ms count" lines into out (capacity cap, NUL-terminated), then
 * clears the record. Returns the number of bytes written, <0 on error. */
int dv3_phase_report(dv3_engine* e, int enable, char* out, size_t cap);
/* Host-only self-check of the stream-K schedules of the tensor-core observe-scan kernels (world_model.py:244-273 on S ... XL):
 * walks every CTA's segment / staging tables with the code the kernels run. 0 = valid, 1 = these sizes are not handled by those
 * kernels (other scan paths are used), < 0 = defect. Needs no GPU. */
int dv3_scan_tc_schedule_check(int B, int T, int G, int D, int Z, int ctas);

/* ---- isolated ops (parity tests call the same kernels the fused path uses) ---------------------- */
int dv3_op_symlog(const float* x, float* y, int64_t n, void* stream);                 /* utils/symlog.py:9-12 */
int dv3_op_inverse_symlog(const float* y, float* x, int64_t n, void* stream);         /* utils/symlog.py:15-28 */
int dv3_op_two_hot(const float* value, int64_t n, int32_t* k, int32_t* kp1, float* w_k, float* w_kp1,
                   float* dense_or_null, void* stream);                                   /* utils/two_hot.py:9-78 (255 buckets, [-20,20]) */
int dv3_op_sample_categorical(const float* probs, const float* u, int64_t rows, int K, int32_t* idx, void* stream);
int dv3_op_value_targets(const float* r, const float* c, const float* v, int H, int64_t N, float gamma,
                         float lambda_, float* targets, void* stream);                 /* critic_loss.py:92-139 */
int dv3_op_percentiles(const float* x, int64_t n, float* p5_p95, void* stream);        /* tfp.stats.percentile nearest */
int dv3_op_gemm(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C, int ldc,
                int M, int N, int K, const float* bias, float beta, int mode, void* stream);
/* tensor-core GEMM with a pre-converted bf16 K-major B operand (Bh[n*ldbh + k]), as the engine uses for weights */
int dv3_op_gemm_shadow(const float* A, int lda, int ta, const void* Bh, int ldbh, float* C, int ldc, int M, int N, int K,
                       const float* bias, float beta, void* stream);
/* backward of y = silu(LayerNorm(x)) given the saved row statistics; dgamma / dbeta are accumulated (nullable) */
int dv3_op_ln_silu_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                       float* dx, float* dgamma, float* dbeta, int64_t rows, int D, void* stream);
/* fp32 -> bf16 copy of a 2-D block, and the TMA-fed tensor-core GEMM on two bf16 K-major operands
 * (C[M,N] = Ah[M,K] * Bh[N,K]^T): the contraction the engine runs when the producer of A wrote a bf16 copy */
int dv3_op_to_bf16(const float* src, int lds, void* dst, int ldd, int64_t rows, int cols, void* stream);
int dv3_op_gemm_bf16(const void* Ah, int ldah, const void* Bh, int ldbh, float* C, int ldc, int M, int N, int K,
                     const float* bias, float beta, void* stream);
/* the same contraction on the small-shape kernel (mma.sync tiles, no TMA / TMEM set-up) the engine uses for the per-step products
 * of the XS dream rollout (models/dreamer_model.py:189-208), where the launch-to-result latency is the cost; -1: not eligible
 * (M in [64, 4096], M N K <= 1024 x 1024 x 1280, K % 8 == 0) */
int dv3_op_gemm_bf16_small(const void* Ah, int ldah, const void* Bh, int ldbh, float* C, int ldc, int M, int N, int K,
                           const float* bias, float beta, void* stream);
/* weight-gradient contraction C[M,N] (+)= Xh[K,M]^T * dYh[K,N] on bf16 copies as stored (both operands MN-major) */
int dv3_op_gemm_bf16_wgrad(const void* Xh, int ldx, const void* dYh, int ldy, float* C, int ldc, int M, int N, int K, float beta,
                           void* stream);
int dv3_op_ln_silu(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DV3_H_ */
