// extern "C" surface of libdv3.so (include/dv3.h).
#include <cstring>
#include <new>

#include "dv3_engine.h"
#include "dv3_scan.h"

using dv3::Engine;

struct dv3_engine {
  Engine e;
};

static std::string g_create_error;

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int check_async(Engine& e) {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return e.fail(std::string("CUDA error: ") + cudaGetErrorString(err));
  return 0;
}

extern "C" {

int dv3_abi_version(void) { return DV3_ABI_VERSION; }

int dv3_create(const dv3_config* cfg, dv3_engine** out) {
  if (cfg == nullptr || out == nullptr) { g_create_error = "null argument"; return -1; }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    g_create_error = "no CUDA device: libdv3 has no CPU path";
    return -2;
  }
  dv3_engine* h = new (std::nothrow) dv3_engine();
  if (!h) { g_create_error = "out of host memory"; return -1; }
  const int rc = h->e.init(cfg);
  if (rc != 0) {
    g_create_error = h->e.err;
    h->e.destroy();
    delete h;
    return rc;
  }
  *out = h;
  return 0;
}

int dv3_destroy(dv3_engine* h) {
  if (!h) return 0;
  h->e.destroy();
  delete h;
  return 0;
}

const char* dv3_last_error(const dv3_engine* h) { return h ? h->e.err.c_str() : g_create_error.c_str(); }

void dv3_default_hparams(dv3_hparams* hp) {
  std::memset(hp, 0, sizeof(*hp));
  hp->gamma = 0.997f; hp->lambda_ = 0.95f; hp->entropy_scale = 3e-4f; hp->return_normalization_decay = 0.99f;
  hp->wm_grad_clip = 1000.f; hp->actor_grad_clip = 100.f; hp->critic_grad_clip = 100.f;
  hp->wm_lr = 1e-4f; hp->wm_eps = 1e-8f; hp->ac_lr = 3e-5f; hp->ac_eps = 1e-5f;
  hp->critic_ema_decay = 0.98f; hp->train_actor = 1;
  hp->disagree_grad_clip = 100.f; hp->disagree_lr = 1e-4f; hp->disagree_eps = 1e-5f;   // run_experiment.py:252, disagree_networks.py:43
}

int dv3_sync(dv3_engine* h, void* stream) {
  cudaError_t err = cudaStreamSynchronize(S(stream));
  if (err != cudaSuccess) return h->e.fail(std::string("cudaStreamSynchronize: ") + cudaGetErrorString(err));
  h->e.phase_dump();
  return 0;
}

int dv3_phase_report(dv3_engine* h, int enable, char* out, size_t cap) {
  const std::string rep = h->e.phase_collect();
  if (enable >= 0) h->e.phase_times = enable != 0;
  if (out == nullptr || cap == 0) return 0;
  const size_t n = rep.size() < cap - 1 ? rep.size() : cap - 1;
  std::memcpy(out, rep.data(), n);
  out[n] = 0;
  return (int)n;
}

int dv3_scan_tc_schedule_check(int B, int T, int G, int D, int Z, int ctas) { return dv3::scan_tc_schedule_check(B, T, G, D, Z, ctas); }

int dv3_num_tensors(const dv3_engine* h) { return (int)h->e.tensors.size(); }

static void fill_desc(const dv3::TensorEntry& t, dv3_tensor_desc* out) {
  out->name = t.name.c_str(); out->ptr = t.ptr; out->dtype = t.dtype; out->ndim = (int)t.shape.size();
  for (int i = 0; i < 6; ++i) out->shape[i] = i < (int)t.shape.size() ? t.shape[i] : 1;
  out->kind = t.kind; out->group = t.group;
}
int dv3_tensor_info(const dv3_engine* h, int index, dv3_tensor_desc* out) {
  if (index < 0 || index >= (int)h->e.tensors.size()) return -1;
  fill_desc(h->e.tensors[index], out);
  return 0;
}
int dv3_find_tensor(const dv3_engine* h, const char* name, dv3_tensor_desc* out) {
  auto it = h->e.index.find(name);
  if (it == h->e.index.end()) return -1;
  fill_desc(h->e.tensors[it->second], out);
  return 0;
}
int dv3_write_tensor(dv3_engine* h, const char* name, const void* src, size_t nbytes, int src_is_device, void* stream, int sync) {
  auto it = h->e.index.find(name);
  if (it == h->e.index.end()) return h->e.fail(std::string("unknown tensor ") + name);
  const auto& t = h->e.tensors[it->second];
  if (nbytes != t.bytes) return h->e.fail(std::string("size mismatch writing ") + name + ": got " + std::to_string(nbytes) + ", expected " + std::to_string(t.bytes));
  cudaError_t err = cudaMemcpyAsync(t.ptr, src, nbytes, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, S(stream));
  if (err == cudaSuccess && sync) err = cudaStreamSynchronize(S(stream));
  if (err != cudaSuccess) return h->e.fail(std::string("copy failed: ") + cudaGetErrorString(err));
  return 0;
}
int dv3_read_tensor(dv3_engine* h, const char* name, void* dst, size_t nbytes, int dst_is_device, void* stream, int sync) {
  auto it = h->e.index.find(name);
  if (it == h->e.index.end()) return h->e.fail(std::string("unknown tensor ") + name);
  const auto& t = h->e.tensors[it->second];
  if (nbytes != t.bytes) return h->e.fail(std::string("size mismatch reading ") + name);
  cudaError_t err = cudaMemcpyAsync(dst, t.ptr, nbytes, dst_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, S(stream));
  if (err == cudaSuccess && sync) err = cudaStreamSynchronize(S(stream));
  if (err != cudaSuccess) return h->e.fail(std::string("copy failed: ") + cudaGetErrorString(err));
  return 0;
}
int dv3_group_buffer(const dv3_engine* h, int group, int kind, void** ptr, int64_t* numel) {
  const Engine& e = h->e;
  if (group == 4 && kind == 1) { *ptr = e.grads[1]; *numel = e.grads_ac_numel; return 0; }   // actor + critic gradients, one span
  if (group >= 6 && group <= 8 && kind == 1) {   // the world-model gradient buffer in three buckets (see dv3.h)
    const long long lo = group == 6 ? e.wm_heads_off : group == 7 ? e.wm_enc_end : 0;
    const long long hi = group == 6 ? e.numel[0] : group == 7 ? e.wm_heads_off : e.wm_enc_end;
    *ptr = e.grads[0] + lo; *numel = hi - lo;
    return 0;
  }
  if (group == Engine::kDisagree && e.numel[group] == 0) return -1;   // curiosity is off
  if (group < 0 || (group > 3 && group != Engine::kDisagree) || kind < 0 || kind > 3 || (group == 3 && kind != 0)) return -1;
  float* p = kind == 0 ? e.params[group] : kind == 1 ? e.grads[group] : kind == 2 ? e.adam_m[group] : e.adam_v[group];
  *ptr = p; *numel = e.numel[group];
  return 0;
}

int dv3_fill_noise(dv3_engine* h, uint64_t seed, uint64_t step, int which, void* stream) {
  h->e.fill_noise(seed, step, which, false, S(stream));
  return check_async(h->e);
}
int dv3_initial_state(dv3_engine* h, void* stream) { h->e.initial_state(S(stream)); return check_async(h->e); }
int dv3_observe_forward(dv3_engine* h, void* stream) { int rc = h->e.observe_forward(S(stream)); return rc ? rc : check_async(h->e); }
int dv3_wm_loss_backward(dv3_engine* h, void* stream) { int rc = h->e.wm_loss_backward(S(stream)); return rc ? rc : check_async(h->e); }
int dv3_optimizer_step(dv3_engine* h, int group, float clip, float lr, float eps, float ema_decay, void* stream) {
  int rc = h->e.optimizer_step(group, clip, lr, eps, ema_decay, S(stream));
  return rc ? rc : check_async(h->e);
}
int dv3_train_world_model_step(dv3_engine* h, const dv3_hparams* hp, uint64_t noise_seed, int use_graph, void* stream) {
  int rc = 0;
  if (use_graph) rc = h->e.train_wm_graph(hp, noise_seed, S(stream));
  else {
    if (noise_seed != 0) { dv3::bump_counter(S(stream), h->e.noise_step); h->e.fill_noise(noise_seed, 0, 1, true, S(stream)); }
    rc = h->e.train_wm(hp, S(stream));
  }
  return rc ? rc : check_async(h->e);
}
int dv3_forward_inference(dv3_engine* h, int rows, uint64_t noise_seed, void* stream) {
  int rc = h->e.forward_inference(rows, noise_seed, S(stream));
  return rc ? rc : check_async(h->e);
}
int dv3_dream_forward(dv3_engine* h, float gamma, int start_from_observe, void* stream) {
  int rc = h->e.dream_forward(gamma, start_from_observe, S(stream));
  return rc ? rc : check_async(h->e);
}
int dv3_dream_forward_actions(dv3_engine* h, float gamma, int start_from_observe, int action_mode, void* stream) {
  if (action_mode < 0 || action_mode > 2) return h->e.fail("dream_forward_actions: action_mode must be 0, 1 or 2");
  h->e.dream_action_mode = action_mode;
  int rc = h->e.dream_forward(gamma, start_from_observe, S(stream));
  h->e.dream_action_mode = 0;
  return rc ? rc : check_async(h->e);
}
int dv3_ac_loss_backward(dv3_engine* h, const dv3_hparams* hp, int phase, void* stream) {
  int rc = h->e.ac_loss_backward(hp, phase, S(stream));
  return rc ? rc : check_async(h->e);
}
int64_t dv3_get_inference_calls(const dv3_engine* h) { return (int64_t)h->e.inf_calls; }
int dv3_set_inference_calls(dv3_engine* h, int64_t calls) {
  if (calls < 0) return h->e.fail("dv3_set_inference_calls: negative count");
  h->e.inf_calls = (unsigned long long)calls;
  return 0;
}
int dv3_disagree_loss_backward(dv3_engine* h, void* stream) {
  int rc = h->e.disagree_loss_backward(S(stream));
  return rc ? rc : check_async(h->e);
}
int dv3_critic_init_ema(dv3_engine* h, void* stream) {
  cudaMemcpyAsync(h->e.params[3], h->e.params[2], (size_t)h->e.numel[2] * 4, cudaMemcpyDeviceToDevice, S(stream));
  return check_async(h->e);
}
int dv3_critic_update_ema(dv3_engine* h, float decay, void* stream) {
  dv3::ema_update(S(stream), h->e.params[3], h->e.params[2], h->e.numel[2], decay);
  return check_async(h->e);
}
int dv3_train_actor_critic_step(dv3_engine* h, const dv3_hparams* hp, uint64_t noise_seed, int use_graph, void* stream) {
  int rc = 0;
  if (use_graph) rc = h->e.train_ac_graph(hp, noise_seed, S(stream));
  else {
    if (noise_seed != 0) h->e.fill_noise(noise_seed, 0, 2, true, S(stream));
    rc = h->e.train_ac(hp, S(stream));
  }
  return rc ? rc : check_async(h->e);
}
int dv3_dp_piece(dv3_engine* h, const dv3_hparams* hp, int piece, uint64_t noise_seed, int use_graph, void* stream) {
  int rc = h->e.dp_piece(hp, piece, noise_seed, use_graph, S(stream));
  return rc ? rc : check_async(h->e);
}
int dv3_train_update(dv3_engine* h, const dv3_hparams* hp, uint64_t noise_seed, int use_graph, void* stream) {
  int rc = h->e.train_update(hp, noise_seed, use_graph, S(stream));
  return rc ? rc : check_async(h->e);
}
int64_t dv3_kernel_launch_count(const dv3_engine*) { return dv3::g_launches; }

// ---- replay sampler ----
int dv3_replay_gather(const void* obs_store, int obs_is_u8, int normalize_u8, int64_t obs_elems, const float* act_store, int A,
                      const float* rew_store, const int32_t* plan, int64_t n, float* obs_out, float* act_out, float* rew_out,
                      float* is_first, float* is_last, float* is_terminated, void* stream) {
  if (!obs_store || !act_store || !rew_store || !plan || !obs_out || !act_out || !rew_out || !is_first || !is_last || !is_terminated) return -1;
  if (obs_elems < 1 || A < 1 || n < 0) return -1;
  dv3::replay_gather(S(stream), obs_store, obs_is_u8, normalize_u8, obs_elems, act_store, A, rew_store, plan, n, obs_out, act_out, rew_out,
                     is_first, is_last, is_terminated);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int64_t dv3_replay_plan(int B, int T, int page, const int64_t* chunk_cum, const int32_t* chunk_ep, const int32_t* chunk_ts0, int64_t n_chunks,
                        const int32_t* ep_len, const uint8_t* ep_term, const int64_t* ep_page_off, const int32_t* pages,
                        const int64_t* draws, int64_t n_draws, int32_t* plan) {
  if (B < 1 || T < 1 || page < 1 || n_chunks < 1 || !chunk_cum || !chunk_ep || !chunk_ts0 || !ep_len || !ep_term || !ep_page_off || !pages ||
      !draws || !plan) return INT64_MIN;
  return dv3::replay_plan(B, T, page, (const long long*)chunk_cum, chunk_ep, chunk_ts0, n_chunks, ep_len, ep_term, (const long long*)ep_page_off,
                          pages, (const long long*)draws, n_draws, plan);
}

// ---- isolated ops ---------------------------------------------------------------------------------------------
static int op_done() { return cudaGetLastError() == cudaSuccess ? 0 : -1; }
int dv3_op_symlog(const float* x, float* y, int64_t n, void* stream) { dv3::symlog_fwd(S(stream), x, y, n); return op_done(); }
int dv3_op_inverse_symlog(const float* y, float* x, int64_t n, void* stream) { dv3::inverse_symlog_fwd(S(stream), y, x, n); return op_done(); }
int dv3_op_two_hot(const float* value, int64_t n, int32_t* k, int32_t* kp1, float* w_k, float* w_kp1, float* dense, void* stream) {
  dv3::two_hot_op(S(stream), value, n, k, kp1, w_k, w_kp1, dense);
  return op_done();
}
int dv3_op_sample_categorical(const float* probs, const float* u, int64_t rows, int K, int32_t* idx, void* stream) {
  if (K > 32 || K < 1) return -1;
  dv3::sample_categorical_op(S(stream), probs, u, rows, K, idx);
  return op_done();
}
int dv3_op_value_targets(const float* r, const float* c, const float* v, int H, int64_t N, float gamma, float lambda_, float* targets, void* stream) {
  dv3::value_targets_raw(S(stream), r, c, v, H, (int)N, gamma, lambda_, targets);
  return op_done();
}
int dv3_op_percentiles(const float* x, int64_t n, float* p5_p95, void* stream) {
  float* ema = nullptr;
  if (cudaMalloc((void**)&ema, 8) != cudaSuccess) return -1;
  const float nanv[2] = {NAN, NAN};
  cudaMemcpyAsync(ema, nanv, 8, cudaMemcpyHostToDevice, S(stream));
  dv3::percentile_ema(S(stream), x, n, 0.99f, ema, p5_p95, nullptr);
  cudaStreamSynchronize(S(stream));
  cudaFree(ema);
  return op_done();
}
int dv3_op_gemm(const float* A, int lda, int ta, const float* B, int ldb, int tb, float* C, int ldc, int M, int N, int K,
                const float* bias, float beta, int mode, void* stream) {
  dv3::GemmArgs g{A, lda, ta, B, ldb, tb, C, ldc, M, N, K, bias, beta};
  dv3::gemm_dispatch(S(stream), g, mode);
  return op_done();
}
int dv3_op_gemm_shadow(const float* A, int lda, int ta, const void* Bh, int ldbh, float* C, int ldc, int M, int N, int K,
                       const float* bias, float beta, void* stream) {
  dv3::GemmArgs g{A, lda, ta, nullptr, 0, 1, C, ldc, M, N, K, bias, beta};
  g.Bh = Bh; g.ldbh = ldbh;
  if ((ldbh % 8) != 0 || ((uintptr_t)Bh % 16) != 0) return -1;
  dv3::gemm_tc(S(stream), g);
  return op_done();
}
int dv3_op_ln_silu_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                       float* dx, float* dgamma, float* dbeta, int64_t rows, int D, void* stream) {
  if (D > 1024) return -1;
  dv3::set_ln_fast(true);
  dv3::ln_silu_bwd(S(stream), dy, D, x, D, mean, rstd, gamma, beta, dx, D, dgamma, dbeta, 1, rows, D);
  return op_done();
}
int dv3_op_to_bf16(const float* src, int lds, void* dst, int ldd, int64_t rows, int cols, void* stream) {
  dv3::to_bf16(S(stream), src, lds, dst, ldd, rows, cols);
  return op_done();
}
int dv3_op_gemm_bf16(const void* Ah, int ldah, const void* Bh, int ldbh, float* C, int ldc, int M, int N, int K,
                     const float* bias, float beta, void* stream) {
  dv3::GemmArgs g{nullptr, 0, 0, nullptr, 0, 1, C, ldc, M, N, K, bias, beta};
  g.Bh = Bh; g.ldbh = ldbh; g.Ah = Ah; g.ldah = ldah;
  if (!dv3::gemm_tma_eligible(g)) return -1;
  if (!dv3::gemm_tma(S(stream), g)) return -2;
  return op_done();
}
int dv3_op_gemm_bf16_small(const void* Ah, int ldah, const void* Bh, int ldbh, float* C, int ldc, int M, int N, int K,
                           const float* bias, float beta, void* stream) {
  dv3::GemmArgs g{nullptr, 0, 0, nullptr, 0, 1, C, ldc, M, N, K, bias, beta};
  g.Bh = Bh; g.ldbh = ldbh; g.Ah = Ah; g.ldah = ldah;
  if (!dv3::gemm_small_eligible(g)) return -1;
  if (!dv3::gemm_small(S(stream), g)) return -2;
  return op_done();
}
int dv3_op_gemm_bf16_wgrad(const void* Xh, int ldx, const void* dYh, int ldy, float* C, int ldc, int M, int N, int K, float beta,
                           void* stream) {
  dv3::GemmArgs g{nullptr, 0, 1, nullptr, 0, 0, C, ldc, M, N, K, nullptr, beta};
  g.Ah = Xh; g.ldah = ldx; g.Bmn = dYh; g.ldbmn = ldy;
  if (!dv3::gemm_tma_eligible(g)) return -1;
  if (!dv3::gemm_tma(S(stream), g)) return -2;
  return op_done();
}
int dv3_op_ln_silu(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int D, void* stream) {
  if (D > 1024) return -1;
  dv3::set_ln_fast(true);
  dv3::ln_silu_fwd(S(stream), x, D, gamma, beta, y, D, nullptr, nullptr, 1, rows, D);
  return op_done();
}

}  // extern "C"
