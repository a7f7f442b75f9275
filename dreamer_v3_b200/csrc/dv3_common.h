// Internal declarations shared by the kernel translation units and the engine.
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define DV3_NB 255            // reward / value buckets
#define DV3_BUCKET_LO (-20.0f)
#define DV3_BUCKET_HI (20.0f)
#define DV3_LN_EPS 1e-3f      // Keras LayerNormalization default

namespace dv3 {

// Number of kernels this library has launched (dv3_kernel_launch_count).
extern long long g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N])
//   op(A)(m,k) = ta ? A[k*lda + m] : A[m*lda + k]
//   op(B)(k,n) = tb ? B[n*ldb + k] : B[k*ldb + n]
struct GemmArgs {
  const float* A; int lda; int ta;
  const float* B; int ldb; int tb;
  float* C; int ldc;
  int M, N, K;
  const float* bias;  // nullable, added once
  float beta;         // 0 -> overwrite, 1 -> accumulate into C
  // optional bf16 shadow of op(B), K-major: Bh[n*ldbh + k] (tensor-core path only; refreshed by the engine after
  // each optimizer step). When set the tensor-core kernel reads it instead of converting B.
  const void* Bh = nullptr; int ldbh = 0;
  // optional bf16 copy of A (row-major [M, K], written by A's producer kernel). With both Ah and Bh set the
  // contraction runs on the TMA-fed kernel (dv3_gemm_tma.cu).
  const void* Ah = nullptr; int ldah = 0;
  // optional bf16 copy of B exactly as given with tb = 0 ([K, N] row-major). Together with ta = 1 and Ah (A as given,
  // [K, M] row-major) this is the weight-gradient case X^T dY: both operands "MN-major", fed to the tensor core by TMA
  // without any transposition.
  const void* Bmn = nullptr; int ldbmn = 0;
};

// ---- programmatic dependent launch ----------------------------------------------------------------
// Kernels launched through launch_pdl may start (and run their prologue) while the previous kernel of the stream is
// still running; they call pdl_wait() (dv3_device.cuh) before their first access to global memory, which blocks until
// the previous grid has completed and its writes are visible. Inside a captured graph this becomes a programmatic edge.
// Opt-in with DV3_PDL=1 (no measurable gain inside the replayed graphs on B200); otherwise plain stream order.
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- gemm (dv3_gemm.cu) ----------------------------------------------------------------------
void gemm_f32(cudaStream_t s, const GemmArgs& g);
// mode 0: always fp32 SIMT; mode 1: bf16 tcgen05 tensor cores for large contractions (dv3_gemm_tc.cu)
void gemm_dispatch(cudaStream_t s, const GemmArgs& g, int mode);
bool gemm_uses_tma(const GemmArgs& g, int mode);   // would gemm_dispatch take the TMA-fed kernel (bf16 operand copies only)?
bool gemm_tc_eligible(const GemmArgs& g);
// weight-streaming kernels for <= 16 rows against a large weight matrix (dv3_gemv.cu), fp32 in every mode
bool gemv16_eligible(const GemmArgs& g);
void gemv16(cudaStream_t s, const GemmArgs& g);
void gemm_tc(cudaStream_t s, const GemmArgs& g);  // bf16 tcgen05 / TMEM path
bool gemm_tma_eligible(const GemmArgs& g);
// small-shape bf16 contraction on mma.sync (dv3_gemm_small.cu): both bf16 operand copies present, <= 1.4 GFLOP
bool gemm_small_eligible(const GemmArgs& g);
bool gemm_small_preferred(const GemmArgs& g);   // eligible and measured faster than the TMA kernel as a graph node
bool gemm_small(cudaStream_t s, const GemmArgs& g);
bool gemm_tma(cudaStream_t s, const GemmArgs& g);  // both operands bf16 K-major in global memory: TMA + tcgen05
void to_bf16(cudaStream_t s, const float* src, int lds, void* dst, int ldd, long long rows, int cols);
// bf16 shadows of weight matrices: out_t[c*rows + r] = bf16(in[r*cols + c]) (transposed), out_n = bf16(in) (plain)
void weight_shadow_bf16(cudaStream_t s, const float* in, void* out_t, void* out_n, int rows, int cols);
// ld_t: row stride of the transposed copy (rows rounded up to 8 elements = 16 bytes, so that every row is a legal TMA row
// even for K = Z + A input widths such as 1025)
struct ShadowDesc { const float* in; void* out_t; void* out_n; int rows, cols, tile_begin, ld_t; };
void weight_shadow_batched(cudaStream_t s, const ShadowDesc* descs_dev, int n, int total_tiles);

// ---- elementwise / normalisation / recurrent cells (dv3_elem.cu) ---------------------------------
void symlog_fwd(cudaStream_t s, const float* x, float* y, long long n);
void inverse_symlog_fwd(cudaStream_t s, const float* y, float* x, long long n);
void fill_f32(cudaStream_t s, float* p, float v, long long n);
void zero_2d(cudaStream_t s, float* p, int rows, int cols, int ld);
void copy_2d(cudaStream_t s, const float* src, int lds, float* dst, int ldd, int rows, int cols);
void add_2d(cudaStream_t s, const float* src, int lds, float* dst, int ldd, int rows, int cols);  // dst += src
void tanh_fwd(cudaStream_t s, const float* x, float* y, int n);

// which LayerNorm + SiLU kernels the two calls below launch: false = the exact-sigmoid warp-per-row kernels (fp32 parity mode),
// true = ex2.approx sigmoid + the rows-per-warp kernels for narrow rows (tensor-core mode; the default of the isolated ops)
void set_ln_fast(bool on);
// y = silu(LayerNorm(x)); saves mean/rstd per row. Rows addressed as x + r*ldx.
void ln_silu_fwd(cudaStream_t s, const float* x, int ldx, const float* gamma, const float* beta, float* y, int ldy,
                 float* mean, float* rstd, int sstride, long long rows, int D, void* y16 = nullptr);  // y16: bf16 twin of y, same ld
// dx = d(silu(LN(x)))/dx * dy; dgamma/dbeta accumulated with atomics when non-null.
void ln_silu_bwd(cudaStream_t s, const float* dy, int lddy, const float* x, int ldx, const float* mean,
                 const float* rstd, const float* gamma, const float* beta, float* dx, int lddx, float* dgamma,
                 float* dbeta, int sstride, long long rows, int D, void* dx16 = nullptr);

// Keras GRU (reset_after=True, gates z,r,h). gx/gh hold x*K+b0 and h*U+b1 ([M,3G]); on return gx holds
// (z, r, c) for the backward pass. h_out may alias nothing.
void gru_fwd(cudaStream_t s, float* gx, int ldgx, const float* gh, int ldgh, const float* h_prev, int ldhp,
             float* h_out, int ldho, int M, int G, void* h16 = nullptr);  // h16: bf16 twin of h_out, same ld
// gates = (z,r,c) as left by gru_fwd, gh as in forward. Produces dgx, dgh ([M,3G]) and dh_prev (=dh*z,
// overwritten or accumulated).
void gru_bwd(cudaStream_t s, const float* dh, int lddh, const float* gates, int ldg, const float* gh, int ldgh,
             const float* h_prev, int ldhp, float* dgx, int lddgx, float* dgh, int lddgh, float* dh_prev, int lddhp,
             int accumulate_dh_prev, int M, int G, void* dgx16 = nullptr, void* dgh16 = nullptr);  // bf16 twins, same ld

// RepresentationLayer (components/representation_layer.py:73-113): logits [M, Zc*Zk] -> unimixed probs,
// optional sample (u [M,Zc] row stride ldu) -> idx [M,Zc] + one-hot z (row stride ldz). mode: 0 = probs only,
// 1 = sample with u, 2 = argmax (get_initial_state).
void repr_fwd(cudaStream_t s, const float* logits, int ldl, float* probs, int ldp, const float* u, int ldu,
              int32_t* idx, int ldi, float* z, int ldz, int M, int Zc, int Zk, int mode, void* z16 = nullptr);
// dlogits = J_softmax^T (0.99 * (dz + dprobs)); dz / dprobs nullable.
void repr_bwd(cudaStream_t s, const float* logits, int ldl, const float* dz, int lddz, const float* dprobs, int lddp,
              float* dlogits, int lddl, int M, int Zc, int Zk, void* dl16 = nullptr);

// observe-scan helpers
// Builds the masked recurrent inputs of step t (world_model.py:246-253).
void scan_mask_inputs(cudaStream_t s, const float* is_first, int ldf, const float* h_prev, int ldhp,
                      const float* z_prev, int ldzp, const float* h0, const float* z0, const float* actW, int ldaw,
                      float* h_in, int ldhi, float* z_in, int ldzi, float* x_pre, int ldx, int B, int G, int Z, int D);
// d_out += (1-f) * d_in (rows): passes gradients to the previous step.
void scan_mask_grads(cudaStream_t s, const float* is_first, int ldf, const float* d_in, int ldin, float* d_out,
                     int ldout, int B, int W);

void colsum_acc(cudaStream_t s, const float* x, int ld, long long rows, int cols, float* out);
void scan_prepare(cudaStream_t s, const float* actions, const float* is_first, float* a_in, float* wf, int B, int T, int A);
void tanh_bwd_acc(cudaStream_t s, const float* dh0, const float* h0, float* dinit, int n);
void infer_mask(cudaStream_t s, const float* f, const float* ph, const float* pz, const float* pa, const float* h0, const float* z0,
                float* h, float* z, float* a, int rows, int G, int Z, int A);

// ---- conv helpers (dv3_conv.cu): 4x4 stride-2 "same" ------------------------------------------------
// cols[(n,oy,ox), (ky,kx,c)] = img[n, 2oy-1+ky, 2ox-1+kx, c] (0 outside); img NHWC [n,H,W,C], out H/2 x W/2
// cols16: bf16 twin of the patch matrix; cols may be null when only the twin is wanted
void im2col_k4s2(cudaStream_t s, const float* img, float* cols, int n, int H, int W, int C, void* cols16 = nullptr);
// img[n,y,x,c] (+)= sum over the (<=4) patches covering (y,x) of cols[(n,oy,ox),(ky,kx,c)] + bias[c] + add_const
void col2im_k4s2(cudaStream_t s, const float* cols, float* img, int n, int H, int W, int C, const float* bias,
                 float add_const);

// ---- losses (dv3_loss.cu) -----------------------------------------------------------------------------
void two_hot_op(cudaStream_t s, const float* v, long long n, int32_t* k, int32_t* kp1, float* wk, float* wkp1,
                float* dense);
// E[bucket] under softmax(logits) per row (reward_predictor_layer.py:84-99)
void bucket_expectation(cudaStream_t s, const float* logits, int ldl, float* out, long long rows);
// dlogits = dE * p * (bucket - E)   (backward of bucket_expectation), overwrite
void bucket_expectation_bwd(cudaStream_t s, const float* logits, int ldl, const float* dE, float* dlogits, int lddl,
                            long long rows);
void sample_categorical_op(cudaStream_t s, const float* probs, const float* u, long long rows, int K, int32_t* idx);

struct WmLossArgs {
  int B, T, Z, Zc, Zk, obs_flat;
  float inv_count;  // 1 / (world_size * B * T)
  const float* obs;            // [N, obs_flat] (symlog'ed if enabled)
  const float* loc;            // [N, obs_flat]
  const float* rew_logits;     // [N, 255]
  const float* rewards;        // [N]
  const float* cont_logit;     // [N]
  const float* is_terminated;  // [N]
  const float* post;           // [N, Z]
  const float* prior;          // [N, Z]
  // outputs
  float* L_dec; float* L_rew; float* L_con; float* L_dyn; float* L_rep; float* L_tot;  // [N]
  float* scalars;              // [8]: means of dec, rew, con, pred, dyn, rep, total
  float* dloc; float* drew_logits; float* dcont_logit; float* dpost; float* dprior;
  int32_t* rew_k;              // [N] two-hot lower bucket index of symlog(reward)
};
void wm_losses(cudaStream_t s, const WmLossArgs& a);

struct AcArgs {
  int H, N;                // rows per step
  int A, discrete;
  float gamma, lambda_, entropy_scale, decay;
  float inv_count;         // 1 / (world_size * H * N)
  const float* rew_E;      // [(H+1)N] symlog-space expected reward
  const float* cont_logit; // [(H+1)N]
  const float* v_E;        // [(H+1)N] symlog-space expected value (fast critic)
  const float* ema_E;      // [(H+1)N]
  const float* is_terminated;  // [N]
  float* r; float* c; float* w; float* v;  // [(H+1)N] real-space outputs
  float* targets;          // [H*N]
  const float* r_intr;     // nullable [(H+1)N]: intrinsic (curiosity) rewards, rows of t >= 1 added to r inside the lambda-returns only
};
void ac_value_targets(cudaStream_t s, const AcArgs& a);
void value_targets_raw(cudaStream_t s, const float* r, const float* c, const float* v, int H, int N, float gamma,
                       float lambda_, float* targets);
// p5/p95 (tfp nearest) of x[n] via radix select; updates ema[0..1] (NaN = unset) in place; writes raw to out[0..1]
void percentile_ema(cudaStream_t s, const float* x, long long n, float decay, float* ema, float* raw_out,
                    uint32_t* scratch);

struct AcLossArgs {
  AcArgs base;
  const float* ema;             // [2] pct5, pct95 EMA (device)
  const float* critic_logits;   // [(H+1)N, 255]
  float* dcritic_logits;        // [(H+1)N, 255] (rows of step H zeroed)
  // discrete actor
  const float* actor_logits;    // [(H+1)N, A] raw (pre-softmax)
  const int32_t* action_idx;    // [(H+1)N]
  float* dactor_logits;         // [(H+1)N, A]
  // box actor
  const float* actor_mu; const float* actor_s;  // raw outputs [(H+1)N, A]
  const float* actions;                         // [(H+1)N, A]
  float* dmu; float* ds;                        // entropy-term gradients (overwritten); BPTT adds to them later
  float* dR; float* dV;                         // dL_actor/dtargets [H*N], dL_actor/dv [(H+1)N] (Box only)
  // per-element outputs [H*N]
  float* L_critic_HB; float* L_val_HB; float* L_reg_HB; float* targets_symlog;
  float* L_actor_HB; float* logp_HB; float* scaled_HB; float* ent_HB; float* L_reinf_HB; float* L_ent_HB;
  float* scalars;               // [16]
};
void ac_losses(cudaStream_t s, const AcLossArgs& a);
// Backward of the lambda-return recursion + inverse_symlog for the Box actor: from dR [H*N], dV [(H+1)N] to
// d rew_E, d v_E (symlog space) [(H+1)N].
void ac_returns_bwd(cudaStream_t s, const AcArgs& a, const float* dR, const float* dV, float* drew_E, float* dv_E);
// Box actor sample a = tanh(mu) + std*eps, std = 0.9*sigmoid(s+2)+0.1 (actor_network.py:102-114)
void box_actor_fwd(cudaStream_t s, const float* mu, const float* sraw, const float* eps, float* a, int lda,
                   long long rows, int A);
// dmu += da*(1-tanh^2), ds += da*eps*0.9*sig*(1-sig)
void box_actor_bwd(cudaStream_t s, const float* mu, const float* sraw, const float* eps, const float* da, int ldda,
                   float* dmu, float* ds, long long rows, int A);
// Discrete actor: logits -> unimix probs, sample with u -> idx, one-hot a (row stride lda)
void disc_actor_fwd(cudaStream_t s, const float* logits, const float* u, float* probs, int32_t* idx, float* a,
                    int lda, long long rows, int A);

// Fused actor tail + sequence-model input layer of one dreamed step (dv3_actor_step.cu): LN+SiLU of the actor's last hidden layer,
// its D x A output layer, the action sample, x_pre += a . W_pre[Z:], u = SiLU(LN(x_pre)). All pointers address the first row.
constexpr int kActorStepMaxA = 8;
struct ActorStepNet {   // last hidden layer of the actor (net 0) / of its std net (net 1, Box spaces)
  const float* pre; float* act; void* act16; float* mean; float* rstd;
  const float* g; const float* b;           // LayerNorm gain / offset
  const float* out_w; const float* out_b;   // [D, A], [A]
};
struct ActorStepArgs {
  long long rows; int D, A, discrete, do_pre;
  ActorStepNet net[2];
  const float* noise;                        // uniforms [rows] (discrete) or standard normals [rows, A] (Box)
  float* logits; float* s_raw; float* probs; int32_t* a_idx; float* act;   // [rows, A] ([rows] for a_idx)
  const float* W_pre_a; const float* pre_g; const float* pre_b;            // rows Z.. of the pre-GRU layer [A, D], its LayerNorm
  float* x_pre; float* x_mean; float* x_rstd; float* u; void* u16;         // x_pre holds z . W_pre[:Z] on entry
};
bool actor_step_supported(int D, int A);
cudaError_t actor_step(cudaStream_t s, const ActorStepArgs& a);

// ---- curiosity (dv3_disagree.cu): models/disagree_networks.py:47-76, losses/disagree_loss.py:11-41 ------------------------
// out[r] = mean over the Z probabilities of the population stddev over the nets; probs [nets][rows, Z] (net_stride elements apart)
void disagree_std_rows(cudaStream_t s, const float* probs, size_t net_stride, int nets, long long rows, int Z, float* out);
// ri = (std_rows - mean(std_rows)) * scale; scalars[1] = mean of ri[first_row:]
void disagree_intrinsic(cudaStream_t s, const float* std_rows, long long rows, long long first_row, float scale, float* ri,
                        float* scalars);
// one net: mse[r] = |p[r] - z_next[r]|^2, dprobs = 2 (p - z_next) inv_count
void disagree_loss_rows(cudaStream_t s, const float* probs, const float* z_next, int ldz, long long rows, int Z, float inv_count,
                        float* dprobs, float* mse);
// scalars[0] = sum(mse[0:n]) / rows
void disagree_loss_total(cudaStream_t s, const float* mse, long long n, long long rows, float* scalars);

// ---- optimizer (dv3_optim.cu) ---------------------------------------------------------------------------
// stats[0] = sum of squares (as double bits in 2 floats? no: separate double*), see implementation.
struct AdamState {
  double* sumsq;     // device scalar
  float* stats;      // [4]: global norm, maxabs grad, maxabs clipped grad, scale
  int* step;         // device step counter (1-based after increment)
};
void grad_norm(cudaStream_t s, const float* g, long long n, AdamState st);
void adam_step(cudaStream_t s, float* w, const float* g, float* m, float* v, long long n, float clip, float lr,
               float eps, AdamState st, float* ema_w, float ema_decay);
void ema_update(cudaStream_t s, float* ema, const float* w, long long n, float decay);

// ---- noise (dv3_optim.cu) ----------------------------------------------------------------------------------
// counter = (element/4, salt + *dev_step) ; dev_step may be null
void philox_fill(cudaStream_t s, float* out, long long n, uint64_t seed, uint64_t salt, const unsigned long long* dev_step, int normal);
void bump_counter(cudaStream_t s, unsigned long long* ctr);

// ---- replay sampler (dv3_replay.cu) -------------------------------------------------------------------------
int replay_gather(cudaStream_t s, const void* obs_store, int obs_is_u8, int normalize_u8, long long obs_elems, const float* act_store, int A,
                  const float* rew_store, const int32_t* plan, long long n, float* obs_out, float* act_out, float* rew_out,
                  float* is_first, float* is_last, float* is_term);
long long replay_plan(int B, int T, int page, const long long* chunk_cum, const int* chunk_ep, const int* chunk_ts0, long long n_chunks,
                      const int* ep_len, const unsigned char* ep_term, const long long* ep_page_off, const int* pages,
                      const long long* draws, long long n_draws, int32_t* plan);

}  // namespace dv3
