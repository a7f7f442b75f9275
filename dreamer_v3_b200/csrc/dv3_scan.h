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
  float* gx_raw;         // [B, 3G] scratch
  float* gates;          // [B, T, 3G] (z, r, c)
  float* gh;             // [B, T, 3G]
  float* hz;             // [B, T, G + Z]
  float* p_pre;          // [B, T, D], pre-loaded with enc . W_post[:E]
  float* p_mean; float* p_rstd; float* p_act;
  float* post_logits; float* post_probs; int32_t* z_idx;
};

size_t scan_fwd_smem_bytes();
int scan_fwd_max_ctas(int device);  // 0 when the kernel cannot be made co-resident
cudaError_t launch_observe_scan_fwd(cudaStream_t s, const ScanFwdParams& p, int ctas);

}  // namespace dv3
