// Arguments of the persistent imagination-rollout kernel (dv3_dream_tc.cu).
#pragma once
#include "dv3_common.h"

namespace dv3 {

constexpr int kDreamTcMaxA = 8;   // action dimension handled in registers by the fused actor head
constexpr int kDreamTcMaxL = 2;   // MLP layers of the actor (XS: 1, S: 2)

struct DreamTcMlp {               // activations of one actor net over all (H+1)*N dream rows, per hidden layer
  float* pre[kDreamTcMaxL]; float* act[kDreamTcMaxL]; float* mean[kDreamTcMaxL]; float* rstd[kDreamTcMaxL];
  void* act16[kDreamTcMaxL];      // bf16 twins of act (may be null)
  const float* g[kDreamTcMaxL]; const float* b[kDreamTcMaxL];                    // LayerNorm gain / offset
  const void* wt[kDreamTcMaxL]; int ld_wt[kDreamTcMaxL];                         // bf16 transposed shadows [D, in]
};

struct DreamTcArgs {
  int N, H, G, D, Z, Zc, Zk, A, L, discrete;
  DreamTcMlp actor, actor_std;
  const float* actor_out_w; const float* actor_out_b;     // [D, A], [A]
  const float* std_out_w; const float* std_out_b;
  const float* W_pre_a;                                   // rows Z.. of W_pre: [A, D] fp32
  const float* pre_g; const float* pre_b; const float* gru_b; const float* dyn_g; const float* dyn_b; const float* zp_bias;
  const void* Wt_pre_z; int ld_pre_z;                     // bf16 transposed shadows [N_out, K]
  const void* Wt_U; int ld_U; const void* Wt_K; int ld_K; const void* Wt_dyn; int ld_dyn; const void* Wt_zp; int ld_zp;
  const float* act_noise; const float* u_dyn;             // [(H+1) N (, A)], [H, N, Zc]
  float* hz; void* hz16;                                  // [(H+1) N, G + Z] and its bf16 twin (TMA source of stage 0)
  float* actor_logits; float* actor_s; float* actor_probs; int32_t* a_idx; float* dr_act;
  float* x_pre; float* x_mean; float* x_rstd; float* u; void* u16;
  float* gx; float* gh;                                   // gx receives the gates (z, r, c) like gru_fwd leaves them
  float* zpart; float* gx_raw;                            // per-step scratch [N, D], [N, 3G]: z . W_pre[:Z] and u . K + b0
  float* q_pre; float* q_mean; float* q_rstd; float* q_act; void* q_act16;
  float* prior_logits; float* prior_probs; int32_t* z_idx;
  unsigned int* bar;
  unsigned long long* prof;   // optional per-CTA cycle totals [ctas][16] (profiling), may be null
};

// Row-resident rollout (dv3_dream_rr.cu): same arguments plus the plain bf16 shadows the one-hot row gathers read and the class
// indices of the start states. XS layer sizes, one actor layer, dreams that start from the engine's own observe pass.
struct DreamRrExtra {
  const void* Wn_pre; int ld_n_pre;          // plain shadow of W_pre [Z + A, D]
  const void* Wn_act[2]; int ld_n_act[2];    // plain shadows of the first layer of the actor / its std net [G + Z, D]
  const int32_t* z0_idx;                     // [N, Zc]
};
bool dream_rr_supported(int N, int G, int D, int Z, int Zc, int Zk, int A, int L);
cudaError_t launch_dream_rr(cudaStream_t s, const DreamTcArgs& a, const DreamRrExtra& x);

bool dream_tc_supported(int N, int G, int D, int Z, int Zk, int A, int L);
int dream_tc_max_ctas(int device);
cudaError_t launch_dream_tc(cudaStream_t s, const DreamTcArgs& a, int ctas);

}  // namespace dv3
