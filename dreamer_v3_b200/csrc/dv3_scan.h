// Parameters of the persistent observe-scan kernels (dv3_scan.cu).
#pragma once
#include "dv3_common.h"

namespace dv3 {

struct ScanFwdParams {
  int B, T, G, D, Z, Zc, Zk, E;
  // weights
  const float* W_pre;    // [Z + A, D] (rows [0, Z) are gathered)
  const float* pre_g; const float* pre_b;
  const float* gru_K;    // [D, 3G]
  const float* gru_U;    // [G, 3G]
  const float* gru_b;    // [2, 3G]
  const float* W_post;   // [E + G, D]
  const float* post_g; const float* post_b;
  const float* W_zq;     // [D, Z]
  const float* b_zq;     // [Z]
  // inputs
  const float* is_first; // [B, T]
  const float* actW;     // [B, T, D]  (1 - f) a_{t-1} . W_pre[Z:]
  const float* u_post;   // [T, B, Zc]
  const float* h0;       // [G]
  const int32_t* z0_idx; // [Zc]
  // state / saved activations (all [B, T, ...] B-major)
  float* h_in; float* z_in; float* x_pre; float* x_mean; float* x_rstd; float* u_act;
  float* gates;          // [B, T, 3G] (z, r, c)
  float* gh;             // [B, T, 3G]
  float* hz;             // [B, T, G + Z]
  float* p_pre;          // [B, T, D], pre-loaded with enc . W_post[:E]
  float* p_mean; float* p_rstd; float* p_act;
  float* post_logits; float* post_probs; int32_t* z_idx;
  int cache_weights;           // set by the launcher: weight tiles resident in shared memory
  unsigned int* bar;           // grid-barrier counter (zeroed by the launcher)
  unsigned long long* stamps;  // optional [T*5+1] globaltimer stamps of CTA 0 (debug / profiling), may be null
};

struct ScanBwdParams {
  int B, T, G, D, Z, Zc, Zk;
  // transposed weight copies (refreshed once per update)
  const float* Wt_zq;      // [Z, D]
  const float* Wt_post_h;  // [D, G]
  const float* Ut;         // [3G, G]
  const float* Kt;         // [3G, D]
  const float* Wt_pre_z;   // [D, Z]
  // untransposed fp32 weights (cluster kernel: converts the rows it owns to bf16 itself)
  const float* W_zq;       // [D, Z]
  const float* W_post_h;   // rows E.. of W_post: [G, D]
  const float* gru_U;      // [G, 3G]
  const float* gru_K;      // [D, 3G]
  const float* W_pre;      // [Z + A, D]
  const float* post_g; const float* post_b; const float* pre_g; const float* pre_b;
  // saved by the forward pass (B-major [B, T, ...])
  const float* is_first; const float* post_logits; const float* dpost;
  const float* p_pre; const float* p_mean; const float* p_rstd;
  const float* gates; const float* gh; const float* h_in;
  const float* x_pre; const float* x_mean; const float* x_rstd;
  // gradients
  float* d_hz;       // [B, T, G+Z] in: heads + prior contributions; the kernel adds the recurrent ones
  float* dlogits;    // [B, T, Z]  posterior logits
  float* dp_act;     // [B, T, D]  grad w.r.t. SiLU(LN(p_pre))
  float* dp_pre;     // [B, T, D]
  float* dh_tot;     // [B, G] scratch
  float* dgx; float* dgh;  // [B, T, 3G]
  float* dh_in;      // [B, T, G]
  float* du;         // [B, T, D]  grad w.r.t. SiLU(LN(x_pre))
  float* dx_pre;     // [B, T, D]
  unsigned int* bar; // grid-barrier counter (zeroed by the launcher)
  int cache_weights; // set by the launcher: every CTA keeps its weight slice of each stage in shared memory
  int cwx;           // set by the launcher (two-stage kernel): GRU units per X task (8, 16 or 32)
  unsigned long long* stamps;  // optional [T*8+1] globaltimer stamps of CTA 0 (profiling), may be null
};

size_t scan_fwd_smem_bytes();
int scan_bwd_max_ctas(int device);
cudaError_t launch_observe_scan_bwd(cudaStream_t s, const ScanBwdParams& p, int ctas);
void transpose_f32(cudaStream_t s, const float* in, float* out, int rows, int cols);
int scan_fwd_max_ctas(int device);  // 0 when the kernel cannot be made co-resident
cudaError_t launch_observe_scan_fwd(cudaStream_t s, const ScanFwdParams& p, int ctas);
// single 16-CTA cluster version (dv3_scan_cluster.cu): weights resident in the cluster's shared memory, DSMEM exchange
bool scan_cluster_supported(const ScanFwdParams& p);
// reverse-time BPTT on the same cluster (tensor-core mode only): bf16 weight rows resident, operands exchanged through DSMEM
bool scan_bwd_cluster_supported(const ScanBwdParams& p);
cudaError_t launch_observe_scan_bwd_cluster(cudaStream_t s, const ScanBwdParams& p);
// bf16: tensor-core variant (bf16 weights resident on chip incl. the gathered W_pre rows, mma.sync contractions)
cudaError_t launch_observe_scan_cluster(cudaStream_t s, const ScanFwdParams& p, bool bf16);

// weight-streaming tensor-core version (dv3_scan_tc.cu; S ... XL in compute_mode 1): bf16 K-major weight shadows [N_out, K]
struct ScanTcWeights {
  const void* Kt; int ld_Kt;                // gru kernel^T            [3G, D]
  const void* Ut; int ld_Ut;                // gru recurrent kernel^T  [3G, G]
  const void* Wpost_h_t; int ld_Wpost_h_t;  // h-part of W_post^T      [D, G] (row stride E + G)
  const void* Wzq_t; int ld_Wzq_t;          // posterior representation layer^T [Z, D]
};
// plain bf16 shadows ([out, in] row-major) for the transposed contractions of the backward scan
struct ScanTcWeightsBwd {
  const void* Wzq; int ld_Wzq;          // [D, Z]
  const void* Wpost_h; int ld_Wpost_h;  // rows E.. of W_post: [G, D]
  const void* U; int ld_U;              // [G, 3G]
  const void* K; int ld_K;              // [D, 3G]
  const void* Wpre_z; int ld_Wpre_z;    // rows 0..Z of W_pre: [Z, D]
};
bool scan_tc_supported(int B, int G, int D, int Z, int Zk);
int scan_tc_max_ctas(int device);
int scan_tc_schedule_check(int B, int T, int G, int D, int Z, int ctas);   // host-only walk of the kernels' stream-K schedules
cudaError_t launch_observe_scan_tc(cudaStream_t s, const ScanFwdParams& p, const ScanTcWeights& w, float* gx_acc, int ctas);
// q.dp_act and q.du must be zero on entry (accumulators); q.d_hz holds the head / prior gradients
cudaError_t launch_observe_scan_tc_bwd(cudaStream_t s, const ScanBwdParams& q, const ScanTcWeightsBwd& w, int ctas);

}  // namespace dv3
