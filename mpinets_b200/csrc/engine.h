// engine.h -- context object behind the C ABI (include/mpinets_b200.h) and the kernel launcher prototypes.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/mpinets_b200.h"

namespace mpn {

void set_error(const char* fmt, ...);

#define MPN_CHECK_CUDA(expr)                                                                 \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      mpn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MPN_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

#define MPN_REQUIRE(cond, ...)           \
  do {                                   \
    if (!(cond)) {                       \
      mpn::set_error(__VA_ARGS__);       \
      return MPN_ERR_INVALID;            \
    }                                    \
  } while (0)

// network dimensions (model.py:47-66,365-393)
constexpr int SA1_NPOINT = 512, SA2_NPOINT = 128, NSAMPLE = 128;
constexpr float SA1_RADIUS = 0.05f, SA2_RADIUS = 0.3f;
constexpr int ENC_DIM = 2048, QF_DIM = 64;

struct Linear {      // y = W x + b ; W [out][in] row-major fp32 (device)
  int in = 0, out = 0;
  float* w = nullptr;     // [out][in]
  float* wt = nullptr;    // [in][out] transposed copy (SIMT kernels read k-major)
  float* b = nullptr;
  __nv_bfloat16* w_bf16 = nullptr;  // [out][in_pad] K-major, zero padded to in_pad (multiple of 16)
  int in_pad = 0;
};

struct ParamInfo { std::string name; int64_t offset; int64_t numel; };

struct Weights {
  float* params = nullptr;        // flat fp32 master copy of every tensor, state-dict order, each tensor 4-float aligned
  int64_t n_params = 0;           // floats in `params` (padding included)
  std::vector<ParamInfo> info;
  Linear sa[3][3];
  Linear fc[3];           // fc_layer.0 / .3 / .6
  float* gn_w[2] = {nullptr, nullptr};
  float* gn_b[2] = {nullptr, nullptr};
  Linear fe[5];           // feature_encoder.0/2/4/6/8
  Linear dec[4];          // decoder.0/2/4/6
  bool finalized = false;
};

struct Workspace {
  int capacity = 0;  // problems
  float *xyz1 = nullptr, *feat1 = nullptr;  // [B][512][3], [B][512][64]
  float *xyz2 = nullptr, *feat2 = nullptr;  // [B][128][3], [B][128][256]
  float* feat3 = nullptr;                   // [B][1024]
  float *fc_a = nullptr, *fc_b = nullptr;   // [B][4096] ping-pong
  float* cat = nullptr;                     // [B][2112]  (encoder out | q features)
  float *h_a = nullptr, *h_b = nullptr;     // [B][512] small head activations
  float* dq = nullptr;                      // [B][7]
  float* qn = nullptr;                      // [B][7] normalised joint state of the rollout
  float* qu = nullptr;                      // [B][7] unnormalised
  float* frames = nullptr;                  // [B][11][12]
  float* eef = nullptr;                     // [B][12]
  int32_t* done = nullptr;                  // [B]
  int32_t* first_step = nullptr;            // [B]
  uint8_t* flags = nullptr;                 // [B]
  int32_t* live = nullptr;                  // [1] problems still running (early-exit polling)
  int32_t* live_host = nullptr;             // pinned copy
  void* tc_scratch = nullptr;               // tensor-core path scratch (bf16 hand-off tensors)
  size_t tc_scratch_bytes = 0;
  void* x3_scratch = nullptr;               // split-bf16 mode scratch (sa_x3.cu)
  __nv_bfloat16* head_op = nullptr;         // [B][2 x 2112] operand rows of decoder.0's tensor-core GEMM (bf16 or [hi | lo])
};

// buffers of the training step (train.cu): everything the backward pass needs from the forward pass, for B samples,
// plus the per-chunk compacted-row scratch of the set-abstraction backward
struct TrainWs {
  int capacity = 0, chunk = 0, n_points = 0;
  int32_t *fps_idx = nullptr;                     // [B][512] scratch
  int32_t *ball1 = nullptr, *ball2 = nullptr;     // [B][512][128], [B][128][128]
  uint8_t *arg1 = nullptr, *arg2 = nullptr, *arg3 = nullptr;   // pooled-row index per (group, channel)
  float *z1 = nullptr, *a1 = nullptr, *z2 = nullptr, *a2 = nullptr;   // FC head: pre-GroupNorm / post-LeakyReLU
  float *st1 = nullptr, *st2 = nullptr;           // GroupNorm (mean, rstd) [B][16][2]
  float *f[4] = {nullptr, nullptr, nullptr, nullptr};   // feature_encoder activations [B][32|64|128|128]
  float *d[3] = {nullptr, nullptr, nullptr};      // decoder activations [B][512|256|128]
  float *yhat = nullptr, *gy = nullptr;           // [B][7]
  float *ga = nullptr, *gb = nullptr;             // gradient ping-pong [B][4096]
  float *gcat = nullptr;                          // [B][2112]
  float *gfeat3 = nullptr, *gfeat2 = nullptr, *gfeat1 = nullptr;   // [B][1024], [B][128][256], [B][512][64]
  // per chunk
  float *X = nullptr, *H1 = nullptr, *H2 = nullptr;
  int32_t* src = nullptr; uint8_t* slot = nullptr;
  // bf16 mode: the active rows of a chunk compacted over its groups (row -> group, group -> first row, per-group row masks / counts,
  // rows[0] = active rows, rows[1] = padded to a multiple of 256)
  int32_t *row_grp = nullptr, *grp_off = nullptr, *grp_cnt = nullptr; uint4* grp_mask = nullptr; int* rows = nullptr;
  float* partial = nullptr; size_t partial_floats = 0;   // split-reduction partials of the weight-gradient kernels
  __nv_bfloat16* tcw = nullptr; float* b2dup = nullptr;  // bf16 weight tiles of the tensor-core backward (rebuilt per step)
  // bf16 mode, dense layers on the TMA GEMM: operand copies of one layer at a time (activation / gradient, their transposes, W, W^T)
  __nv_bfloat16 *dx = nullptr, *dgy = nullptr, *dgyT = nullptr, *daT = nullptr, *dw = nullptr, *dwT = nullptr;
  float* zeros = nullptr;                                 // [4096] zero bias of the gradient GEMMs
  float* adam_m = nullptr; float* adam_v = nullptr; float* norm = nullptr;
  // bf16 mode: hidden activations of the group-all level saved by the forward GEMMs, [B*128][512] each (null in the fp32 mode)
  const __nv_bfloat16 *sa3_h1 = nullptr, *sa3_h2 = nullptr;
  const __nv_bfloat16* sa3_a3 = nullptr;   // ... and its operand rows [B*128][272] = [256 SA2 features | x y z | 0-pad] (SA2's output rows)
};

}  // namespace mpn

struct mpn_ctx {
  int device = 0;
  mpn_config cfg{};
  int sm_count = 148;
  // robot tables (device)
  bool tables_set = false;
  float limits_host[14] = {0};
  float* limits = nullptr;       // [7][2]
  int P = 0; float* link_points = nullptr; int32_t* link_ids = nullptr;
  float* link_table4 = nullptr;  // [P] float4 (x, y, z, link id bits): one gather per robot row
  float* robot_sel4 = nullptr;   // [P] float4: this step's permuted subset of link_table4 (shared by the whole batch)
  float* robot_sel_steps = nullptr; size_t robot_sel_steps_cap = 0;   // [steps][n] float4 slabs of a whole rollout
  int n_base_points = 0;         // leading link-0 rows of the table (FrankaSampler(with_base_link=False) skips them)
  float* loss_partial = nullptr; size_t loss_partial_cap = 0;   // per-CTA partial sums of the loss kernels
  int Pe = 0; float* ee_points = nullptr;
  int S = 0; float* sph_c = nullptr; float* sph_r = nullptr; int32_t* sph_l = nullptr;
  float prismatic = 0.025f;
  mpn::Weights w;
  mpn::Workspace ws;
  mpn::TrainWs tw;
  // early-exit rollouts: two ping-pong sets of per-problem state for the compacted still-running problems (engine.cu: LiveSet)
  void* live_sets = nullptr;
  int64_t launches = 0;
  // stage profiler
  bool prof = false;
  struct StageRec { int stage; cudaEvent_t a, b; };
  std::vector<StageRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
};

namespace mpn {

// records a CUDA-event pair around a stage on the launching stream when profiling is enabled
struct StageTimer {
  mpn_ctx* c; cudaStream_t s; int idx = -1;
  StageTimer(mpn_ctx* c_, cudaStream_t s_, int stage) : c(c_), s(s_) {
    if (!c->prof) return;
    auto get = [&]() { cudaEvent_t e; if (c->prof_pool.empty()) cudaEventCreate(&e); else { e = c->prof_pool.back(); c->prof_pool.pop_back(); } return e; };
    mpn_ctx::StageRec r{stage, get(), get()};
    cudaEventRecord(r.a, s);
    c->prof_recs.push_back(r);
    idx = (int)c->prof_recs.size() - 1;
  }
  ~StageTimer() { if (idx >= 0) cudaEventRecord(c->prof_recs[idx].b, s); }
};

// ---- geometry.cu
int launch_fk(mpn_ctx* c, cudaStream_t s, const float* q, int B, float* frames, float* eef);
int launch_sample_robot(mpn_ctx* c, cudaStream_t s, const float* frames, int B, int n, uint32_t step, float* cloud, int rows);
int launch_robot_subsets(mpn_ctx* c, cudaStream_t s, int n, uint32_t step0, int count);
int launch_sample_robot_slab(mpn_ctx* c, cudaStream_t s, const float* frames, int B, int n, int slab, float* cloud, int rows);
int pack_link_table(mpn_ctx* c);
int launch_spheres(mpn_ctx* c, cudaStream_t s, const float* frames, int B, float* centers);
int launch_normalize(mpn_ctx* c, cudaStream_t s, const float* in, int n, float* out, bool unnormalize);
int launch_sdf_points(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* pts, int N, int which, float* sdf);
int launch_build_cloud(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* frames, const float* target,
                       uint32_t problem0, float* cloud, const float* obs_points = nullptr, const int32_t* obs_count = nullptr,
                       int obs_max = 0, const uint32_t* problem_ids = nullptr);
// ---- ingest.cu
int launch_augment_joints(mpn_ctx* c, cudaStream_t s, const float* q, int B, float scale, const uint32_t* sample_ids, uint32_t epoch,
                          float* q_out, float* qn_out);
int launch_clean_point_cloud(mpn_ctx* c, cudaStream_t s, const float* xyz, const float* rgba, int N, int n_out, uint32_t cloud_id,
                             float* out_xyz, float* out_rgba, int32_t* kept, int32_t* scratch);
int launch_sample_end_effector(mpn_ctx* c, cudaStream_t s, const float* poses, int B, int n, uint32_t problem0, float* out);
int launch_sweep(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* traj, int T, int traj_stride_t,
                 int t0, int accumulate, uint8_t* flags, int32_t* first_step, const float* frames_in = nullptr);
int launch_evaluate(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* traj, int T1, const int32_t* num_poses,
                    const float* target, const mpn_scene& tv, int V1, int V2, const mpn_scene& nv, int N1, int N2, float* out);
int launch_sparc(mpn_ctx* c, cudaStream_t s, int B, int n_max, const float* movement, const int32_t* num, float fs, int padlevel, float fc,
                 float amp_th, float* out);
int launch_render_depth(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* camera, int per_problem_camera, int W, int H,
                        float sx, float sy, float tnear, float tfar, float* points, int32_t* counts);
// ---- loss.cu
int launch_collision_loss(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, int N, const float* points, float margin, float* loss,
                          float* grad_points);
int launch_point_match_loss(mpn_ctx* c, cudaStream_t s, size_t n, const float* a, const float* b, float* loss, float* grad_a);
int launch_bc_collision_losses(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* input_norm, const float* target_norm,
                               int n_points, float margin, float w_collision, float w_bc, float* losses, float* grad_input);
// ---- pointnet.cu
int launch_fps(mpn_ctx* c, cudaStream_t s, const float* xyz, int B, int N, int stride, int npoint, int32_t* idx, float* new_xyz);
int launch_ball_query(mpn_ctx* c, cudaStream_t s, float radius, int nsample, const float* xyz, int B, int N, int stride,
                      const float* new_xyz, int npoint, int32_t* idx);
int launch_gather(mpn_ctx* c, cudaStream_t s, const float* feat, int B, int C, int N, const int32_t* idx, int m, float* out);
int launch_group(mpn_ctx* c, cudaStream_t s, const float* feat, int B, int C, int N, const int32_t* idx, int m, int ns, float* out);
// ---- sa_simt.cu : fused ball query + group + 3-layer shared MLP + max, fp32
int launch_sa_simt(mpn_ctx* c, cudaStream_t s, int module, const float* xyz, int stride, const float* feats, int feat_stride,
                   int B, int N, const float* new_xyz, float* new_feats, int32_t* ball_idx, uint8_t* arg_out = nullptr);
// ---- linear.cu
int launch_linear(mpn_ctx* c, cudaStream_t s, const Linear& L, const float* x, int ldx, int M, float* y, int ldy, int act);
// Y[M][N] = act(X[M][K] * Wm[N][K]^T + bias) (* f'(mask) when mask != null: mask_mode 1 LeakyReLU(0.01), 2 ReLU); Y may alias mask
int launch_linear_ex(mpn_ctx* c, cudaStream_t s, const float* X, int ldx, const float* Wm, int ldw, const float* bias, int64_t M, int N,
                     int K, float* Y, int ldy, int act, const float* mask = nullptr, int ldmask = 0, int mask_mode = 0);
int launch_groupnorm_lrelu(mpn_ctx* c, cudaStream_t s, float* x, int M, int C, int groups, const float* gamma, const float* beta);
// training forward: out-of-place, keeps the pre-norm input and the per-(row, group) statistics
int launch_groupnorm_lrelu_train(mpn_ctx* c, cudaStream_t s, const float* z, int M, int C, int groups, const float* gamma,
                                 const float* beta, float* out, float* stats);
int launch_groupnorm_lrelu_bf16(mpn_ctx* c, cudaStream_t s, float* x, int M, int C, int groups, const float* gamma, const float* beta,
                                __nv_bfloat16* out);
// ---- heads.cu
int launch_feature_encoder(mpn_ctx* c, cudaStream_t s, const float* qn, int B, float* cat, int ldcat, int operand_mode, __nv_bfloat16* operand);
int launch_decoder_tail(mpn_ctx* c, cudaStream_t s, const float* h0, int B, float* dq);
int launch_linear_skinny(mpn_ctx* c, cudaStream_t s, const void* X, int ldx, int x_mode, const Linear& L, int M, float* Y, int ldy, int act);
constexpr int SKINNY_MAX_ROWS = 16;   // batches up to this size take the fp32 weight-streaming dense layers
// ---- gemm_tc.cu (epi: 0 relu->bf16, 1 fp32, 2 relu + max over each 128-row tile -> bf16)
int launch_gemm_tc(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int K, const float* bias,
                   int M, int N, void* C, int ldc, uint8_t* arg_out = nullptr);
int launch_step_update(mpn_ctx* c, cudaStream_t s, int B, const float* dq, float* qn, float* qu, const float* target,
                       int32_t* done, int early_exit, float* traj_out, int traj_stride, float* frames, float* eef,
                       float* metrics, int step);
int launch_count_live(mpn_ctx* c, cudaStream_t s, int B, const int32_t* done, int32_t* live);
int launch_fill_traj_tail(mpn_ctx* c, cudaStream_t s, int B, const float* qu, float* traj, int traj_stride, int from, int T);
int launch_finalize_metrics(mpn_ctx* c, cudaStream_t s, int B, const float* eef, const float* target, const uint8_t* flags,
                            const int32_t* first_step, const int32_t* done, int T, float* metrics);
// ---- train.cu : training step (model.py:185-240) in fp32 -- forward with saved state, losses, backward, Adam
int train_step_grads(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, int N, const float* cloud, const float* q_norm,
                     const float* supervision, int n_loss_points, float margin, float w_collision, float w_bc, float* losses,
                     float* y_hat, float* grads, int precision = MPN_PREC_FP32);
int adam_step(mpn_ctx* c, cudaStream_t s, const float* grads, float lr, float beta1, float beta2, float eps, float clip_norm,
              int step, float* grad_norm_out);
int refresh_transposes(mpn_ctx* c, cudaStream_t s);
// ---- train_tc.cu : tcgen05 GEMMs over compacted rows (bf16 operands)
int launch_rows_gemm_tc(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias,
                        const __nv_bfloat16* mask, long long M, int N, __nv_bfloat16* C, float* pool_out = nullptr,
                        uint8_t* pool_arg = nullptr, int paired = 0, const int* rows_dev = nullptr, int rows_shift = 0);
int launch_wgrad_tc(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* dY, const __nv_bfloat16* X, long long R, float* partial,
                    size_t partial_floats, int* n_ctas, int swap_lbo_sbo = 0, const int* rows_dev = nullptr, int rows_shift = 0);
int launch_wgrad_tc2d(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* dY, int ldy, int y_cols, const __nv_bfloat16* X, int ldx, int x_cols,
                      long long R, float* partial, size_t partial_floats, int* n_ctas);
void free_train_ws(mpn_ctx* c);

}  // namespace mpn
