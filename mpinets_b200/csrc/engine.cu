// engine.cu -- context, weight store, workspace, forward / rollout sequencing and the extern "C" surface.
#include "engine.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace mpn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// tensor-core (bf16 / tcgen05) path, sa_tc.cu
int tc_encoder_forward(mpn_ctx* c, cudaStream_t s, const float* cloud, int B, int N, float* out, int ldo);
int tc_prepare_weights(mpn_ctx* c);
size_t tc_scratch_bytes(int B);
int* tc_error_flag(mpn_ctx* c);
int sa_tile_counts(mpn_ctx* c, unsigned long long* out, int reset);
int tc_sa_forward(mpn_ctx* c, cudaStream_t s, int module, const float* xyz, int stride, const float* feats, int feat_stride, int B,
                  int N, const float* new_xyz, float* new_feats, int32_t* ball_idx);
int tc_probe(mpn_ctx* c, cudaStream_t s, const void* A, const void* B, float* D, int N, int K, int mode, int* status);
void tc_free(mpn_ctx* c);
int tc_decoder0(mpn_ctx* c, cudaStream_t s, int precision, const __nv_bfloat16* operand, int B, float* h0);
// split-bf16 parity-grade tensor-core path, sa_x3.cu
int x3_encoder_forward(mpn_ctx* c, cudaStream_t s, const float* cloud, int B, int N, float* out, int ldo);
int x3_sa_forward(mpn_ctx* c, cudaStream_t s, int module, const float* xyz, int stride, const float* feats, int feat_stride, int B,
                  int N, const float* new_xyz, float* new_feats, int32_t* ball_idx);
size_t x3_scratch_bytes(int B);
int x3_gemm_selftest(mpn_ctx* c, cudaStream_t s, const float* A, const float* W, const float* bias, int M, int N, int K, float* C, int split);
void x3_free(mpn_ctx* c);

template <typename T>
static int dev_alloc(T** p, size_t n) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  if (n == 0) return MPN_OK;
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    return MPN_ERR_NOMEM;
  }
  return MPN_OK;
}

template <typename T>
static int dev_upload(T** p, const T* host, size_t n) {
  int r = dev_alloc(p, n);
  if (r) return r;
  MPN_CHECK_CUDA(cudaMemcpy(*p, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return MPN_OK;
}

static int ensure_workspace(mpn_ctx* c, int B) {
  Workspace& w = c->ws;
  if (B <= w.capacity) return MPN_OK;
  size_t b = (size_t)B;
  int r = 0;
  r |= dev_alloc(&w.xyz1, b * SA1_NPOINT * 3);
  r |= dev_alloc(&w.feat1, b * SA1_NPOINT * 64);
  r |= dev_alloc(&w.xyz2, b * SA2_NPOINT * 3);
  r |= dev_alloc(&w.feat2, b * SA2_NPOINT * 256);
  r |= dev_alloc(&w.feat3, b * 1024);
  r |= dev_alloc(&w.fc_a, b * 4096);
  r |= dev_alloc(&w.fc_b, b * 4096);
  r |= dev_alloc(&w.cat, b * (ENC_DIM + QF_DIM));
  r |= dev_alloc(&w.h_a, b * 512);
  r |= dev_alloc(&w.h_b, b * 512);
  r |= dev_alloc(&w.dq, b * 7);
  r |= dev_alloc(&w.qn, b * 7);
  r |= dev_alloc(&w.qu, b * 7);
  r |= dev_alloc(&w.frames, b * 11 * 12);
  r |= dev_alloc(&w.eef, b * 12);
  r |= dev_alloc(&w.done, b);
  r |= dev_alloc(&w.first_step, b);
  r |= dev_alloc(&w.flags, b);
  r |= dev_alloc(&w.live, 1);
  if (!w.live_host && cudaMallocHost((void**)&w.live_host, sizeof(int32_t)) != cudaSuccess) r |= MPN_ERR_NOMEM;
  r |= dev_alloc(&w.head_op, b * 2 * (ENC_DIM + QF_DIM));
  if (w.tc_scratch) { cudaFree(w.tc_scratch); w.tc_scratch = nullptr; }
  w.tc_scratch_bytes = tc_scratch_bytes(B);
  if (w.tc_scratch_bytes) {
    if (cudaMalloc(&w.tc_scratch, w.tc_scratch_bytes) != cudaSuccess) { set_error("workspace: tc scratch alloc failed"); r |= MPN_ERR_NOMEM; }
  }
  if (w.x3_scratch) { cudaFree(w.x3_scratch); w.x3_scratch = nullptr; }
  if (cudaMalloc(&w.x3_scratch, x3_scratch_bytes(B)) != cudaSuccess) { set_error("workspace: bf16x3 scratch alloc failed"); r |= MPN_ERR_NOMEM; }
  if (r) { w.capacity = 0; return MPN_ERR_NOMEM; }
  w.capacity = B;
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------- weights
struct HostTensor { std::vector<int64_t> shape; std::vector<float> data; };
static std::map<mpn_ctx*, std::map<std::string, HostTensor>> g_host_weights;

// Tensors live in one flat fp32 buffer (Weights::params) in state-dict order, each 4-float aligned, so the optimiser and
// the gradient all-reduce see a single vector; Linear::w / ::b and gn_w / gn_b point into it.
struct ParamPlan { std::string name; int64_t numel; };

static std::vector<ParamPlan> param_plan() {
  static const int sa_dims[3][4] = {{4, 64, 64, 64}, {67, 128, 128, 256}, {259, 512, 512, 1024}};
  std::vector<ParamPlan> v;
  char buf[128];
  auto lin = [&](const char* prefix, int in, int out) {
    v.push_back({std::string(prefix) + ".weight", (int64_t)in * out});
    v.push_back({std::string(prefix) + ".bias", out});
  };
  for (int m = 0; m < 3; ++m)
    for (int l = 0; l < 3; ++l) {
      snprintf(buf, sizeof(buf), "point_cloud_encoder.SA_modules.%d.mlps.0.%d", m, 2 * l);
      lin(buf, sa_dims[m][l], sa_dims[m][l + 1]);
    }
  lin("point_cloud_encoder.fc_layer.0", 1024, 4096);
  lin("point_cloud_encoder.fc_layer.1", 1, 4096);     // GroupNorm weight / bias
  lin("point_cloud_encoder.fc_layer.3", 4096, 2048);
  lin("point_cloud_encoder.fc_layer.4", 1, 2048);
  lin("point_cloud_encoder.fc_layer.6", 2048, 2048);
  static const int fe_dims[6] = {7, 32, 64, 128, 128, 64};
  for (int l = 0; l < 5; ++l) { snprintf(buf, sizeof(buf), "feature_encoder.%d", 2 * l); lin(buf, fe_dims[l], fe_dims[l + 1]); }
  static const int de_dims[5] = {2112, 512, 256, 128, 7};
  for (int l = 0; l < 4; ++l) { snprintf(buf, sizeof(buf), "decoder.%d", 2 * l); lin(buf, de_dims[l], de_dims[l + 1]); }
  return v;
}

static float* param_ptr(mpn_ctx* c, const std::string& name) {
  for (auto& p : c->w.info) if (p.name == name) return c->w.params + p.offset;
  return nullptr;
}

static int make_linear(mpn_ctx* c, const std::string& prefix, int in, int out, Linear& L) {
  L.in = in; L.out = out;
  L.w = param_ptr(c, prefix + ".weight");
  L.b = param_ptr(c, prefix + ".bias");
  if (!L.w || !L.b) { set_error("internal: no parameter slot for %s", prefix.c_str()); return MPN_ERR_STATE; }
  return dev_alloc(&L.wt, (size_t)in * out);
}

static int finalize_weights(mpn_ctx* c) {
  auto& hw = g_host_weights[c];
  // flat buffer
  std::vector<ParamPlan> plan = param_plan();
  c->w.info.clear();
  int64_t off = 0;
  for (auto& p : plan) { c->w.info.push_back({p.name, off, p.numel}); off += (p.numel + 3) / 4 * 4; }
  std::vector<float> flat((size_t)off, 0.f);
  for (auto& p : c->w.info) {
    auto it = hw.find(p.name);
    if (it == hw.end()) { set_error("missing weight %s", p.name.c_str()); return MPN_ERR_STATE; }
    if ((int64_t)it->second.data.size() != p.numel) {
      set_error("weight %s has %zu elements, expected %lld", p.name.c_str(), it->second.data.size(), (long long)p.numel);
      return MPN_ERR_INVALID;
    }
    memcpy(flat.data() + p.offset, it->second.data.data(), sizeof(float) * (size_t)p.numel);
  }
  c->w.n_params = off;
  int r = dev_upload(&c->w.params, flat.data(), flat.size());
  if (r) return r;
  free_train_ws(c);   // optimiser state belongs to the previous parameter vector

  static const int sa_dims[3][4] = {{4, 64, 64, 64}, {67, 128, 128, 256}, {259, 512, 512, 1024}};
  char buf[128];
  for (int m = 0; m < 3; ++m)
    for (int l = 0; l < 3; ++l) {
      snprintf(buf, sizeof(buf), "point_cloud_encoder.SA_modules.%d.mlps.0.%d", m, 2 * l);
      if ((r = make_linear(c, buf, sa_dims[m][l], sa_dims[m][l + 1], c->w.sa[m][l]))) return r;
    }
  static const int fc_idx[3] = {0, 3, 6};
  static const int fc_dims[4] = {1024, 4096, 2048, 2048};
  for (int l = 0; l < 3; ++l) {
    snprintf(buf, sizeof(buf), "point_cloud_encoder.fc_layer.%d", fc_idx[l]);
    if ((r = make_linear(c, buf, fc_dims[l], fc_dims[l + 1], c->w.fc[l]))) return r;
  }
  c->w.gn_w[0] = param_ptr(c, "point_cloud_encoder.fc_layer.1.weight");
  c->w.gn_b[0] = param_ptr(c, "point_cloud_encoder.fc_layer.1.bias");
  c->w.gn_w[1] = param_ptr(c, "point_cloud_encoder.fc_layer.4.weight");
  c->w.gn_b[1] = param_ptr(c, "point_cloud_encoder.fc_layer.4.bias");
  static const int fe_dims[6] = {7, 32, 64, 128, 128, 64};
  for (int l = 0; l < 5; ++l) {
    snprintf(buf, sizeof(buf), "feature_encoder.%d", 2 * l);
    if ((r = make_linear(c, buf, fe_dims[l], fe_dims[l + 1], c->w.fe[l]))) return r;
  }
  static const int de_dims[5] = {2112, 512, 256, 128, 7};
  for (int l = 0; l < 4; ++l) {
    snprintf(buf, sizeof(buf), "decoder.%d", 2 * l);
    if ((r = make_linear(c, buf, de_dims[l], de_dims[l + 1], c->w.dec[l]))) return r;
  }
  if ((r = refresh_transposes(c, 0))) return r;
  if ((r = tc_prepare_weights(c))) return r;
  c->w.finalized = true;
  g_host_weights.erase(c);
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------- forward passes
static int encoder_forward(mpn_ctx* c, cudaStream_t s, int precision, const float* cloud, int B, int N, float* out, int ldo) {
  Workspace& w = c->ws;
  int r;
  if (precision == MPN_PREC_BF16) return tc_encoder_forward(c, s, cloud, B, N, out, ldo);
  if (precision == MPN_PREC_BF16X3) return x3_encoder_forward(c, s, cloud, B, N, out, ldo);
  // SA1: FPS over the 4-float rows of the cloud; features = mask column
  { StageTimer t(c, s, MPN_ST_FPS1);
    if ((r = launch_fps(c, s, cloud, B, N, 4, SA1_NPOINT, reinterpret_cast<int32_t*>(w.fc_a), w.xyz1))) return r; }
  { StageTimer t(c, s, MPN_ST_SA1);
    if ((r = launch_sa_simt(c, s, 0, cloud, 4, cloud + 3, 4, B, N, w.xyz1, w.feat1, nullptr))) return r; }
  { StageTimer t(c, s, MPN_ST_FPS2);
    if ((r = launch_fps(c, s, w.xyz1, B, SA1_NPOINT, 3, SA2_NPOINT, reinterpret_cast<int32_t*>(w.fc_a), w.xyz2))) return r; }
  { StageTimer t(c, s, MPN_ST_SA2);
    if ((r = launch_sa_simt(c, s, 1, w.xyz1, 3, w.feat1, 64, B, SA1_NPOINT, w.xyz2, w.feat2, nullptr))) return r; }
  { StageTimer t(c, s, MPN_ST_SA3);
    if ((r = launch_sa_simt(c, s, 2, w.xyz2, 3, w.feat2, 256, B, SA2_NPOINT, nullptr, w.feat3, nullptr))) return r; }
  StageTimer tfc(c, s, MPN_ST_FC);
  if (B <= SKINNY_MAX_ROWS) {
    if ((r = launch_linear_skinny(c, s, w.feat3, 1024, 0, c->w.fc[0], B, w.fc_a, 4096, 0))) return r;
    if ((r = launch_groupnorm_lrelu(c, s, w.fc_a, B, 4096, 16, c->w.gn_w[0], c->w.gn_b[0]))) return r;
    if ((r = launch_linear_skinny(c, s, w.fc_a, 4096, 0, c->w.fc[1], B, w.fc_b, 2048, 0))) return r;
    if ((r = launch_groupnorm_lrelu(c, s, w.fc_b, B, 2048, 16, c->w.gn_w[1], c->w.gn_b[1]))) return r;
    return launch_linear_skinny(c, s, w.fc_b, 2048, 0, c->w.fc[2], B, out, ldo, 0);
  }
  if ((r = launch_linear(c, s, c->w.fc[0], w.feat3, 1024, B, w.fc_a, 4096, 0))) return r;
  if ((r = launch_groupnorm_lrelu(c, s, w.fc_a, B, 4096, 16, c->w.gn_w[0], c->w.gn_b[0]))) return r;
  if ((r = launch_linear(c, s, c->w.fc[1], w.fc_a, 4096, B, w.fc_b, 2048, 0))) return r;
  if ((r = launch_groupnorm_lrelu(c, s, w.fc_b, B, 2048, 16, c->w.gn_w[1], c->w.gn_b[1]))) return r;
  return launch_linear(c, s, c->w.fc[2], w.fc_b, 2048, B, out, ldo, 0);
}

static int policy_forward(mpn_ctx* c, cudaStream_t s, int precision, const float* cloud, const float* qn, int B, int N, float* dq) {
  Workspace& w = c->ws;
  int r;
  const int CAT = ENC_DIM + QF_DIM;
  if ((r = encoder_forward(c, s, precision, cloud, B, N, w.cat, CAT))) return r;
  // feature_encoder (model.py:47-57) -> cat(encoder, q features) (model.py:90) -> decoder (model.py:58-66): three launches
  StageTimer th(c, s, MPN_ST_HEADS);
  const bool skinny = B <= SKINNY_MAX_ROWS;
  const bool tc = precision != MPN_PREC_FP32 && !skinny;
  if ((r = launch_feature_encoder(c, s, qn, B, w.cat, CAT, tc ? (precision == MPN_PREC_BF16 ? 1 : 2) : 0, w.head_op))) return r;
  if (skinny) {
    if ((r = launch_linear_skinny(c, s, w.cat, CAT, 0, c->w.dec[0], B, w.h_a, 512, 1))) return r;
  } else if (tc) {
    if ((r = tc_decoder0(c, s, precision, w.head_op, B, w.h_a))) return r;
  } else {
    if ((r = launch_linear(c, s, c->w.dec[0], w.cat, CAT, B, w.h_a, 512, 1))) return r;
  }
  return launch_decoder_tail(c, s, w.h_a, B, dq);
}

}  // namespace mpn

using namespace mpn;
void free_live_sets(mpn_ctx* c);
#ifndef MPN_NLINK
#define MPN_NLINK 11   // link frames per configuration (spec_math.cuh)
#endif

#define REQ_CTX(c)                                              \
  do {                                                          \
    if (!(c)) { set_error("null context"); return MPN_ERR_INVALID; } \
    MPN_CHECK_CUDA(cudaSetDevice((c)->device));                 \
  } while (0)
#define REQ_TABLES(c) \
  do { if (!(c)->tables_set) { set_error("robot tables not set (mpn_set_robot_tables)"); return MPN_ERR_STATE; } } while (0)
#define REQ_WEIGHTS(c) \
  do { if (!(c)->w.finalized) { set_error("weights not finalized (mpn_load_weight / mpn_weights_finalize)"); return MPN_ERR_STATE; } } while (0)

extern "C" {

const char* mpn_last_error(void) { return g_err; }
const char* mpn_version(void) { return "mpinets_b200 0.1 (sm_100a)"; }

int mpn_ctx_create(int device, const mpn_config* cfg, mpn_ctx** out) {
  if (!cfg || !out) { set_error("mpn_ctx_create: null argument"); return MPN_ERR_INVALID; }
  MPN_REQUIRE(cfg->n_robot > 0 && cfg->n_obstacle >= 0 && cfg->n_target >= 0, "mpn_ctx_create: bad point counts");
  MPN_REQUIRE(cfg->max_cuboids >= 0 && cfg->max_cylinders >= 0 && cfg->max_cuboids + cfg->max_cylinders <= 128,
              "mpn_ctx_create: max_cuboids + max_cylinders must be <= 128");
  int ndev = 0;
  MPN_CHECK_CUDA(cudaGetDeviceCount(&ndev));
  MPN_REQUIRE(device >= 0 && device < ndev, "mpn_ctx_create: device %d not present (%d devices)", device, ndev);
  MPN_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MPN_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  MPN_REQUIRE(prop.major == 10, "mpinets_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  mpn_ctx* c = new mpn_ctx();
  c->device = device;
  c->cfg = *cfg;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return MPN_OK;
}

int mpn_ctx_destroy(mpn_ctx* c) {
  if (!c) return MPN_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  g_host_weights.erase(c);
  // device memory is reclaimed by explicit frees of the big buffers; small tables die with the process
  float** bufs[] = {&c->ws.xyz1, &c->ws.feat1, &c->ws.xyz2, &c->ws.feat2, &c->ws.feat3, &c->ws.fc_a, &c->ws.fc_b, &c->ws.cat,
                    &c->ws.h_a, &c->ws.h_b, &c->ws.dq, &c->ws.qn, &c->ws.qu, &c->ws.frames, &c->ws.eef};
  for (auto p : bufs) if (*p) cudaFree(*p);
  if (c->ws.done) cudaFree(c->ws.done);
  if (c->ws.first_step) cudaFree(c->ws.first_step);
  if (c->ws.flags) cudaFree(c->ws.flags);
  if (c->ws.live) cudaFree(c->ws.live);
  if (c->ws.live_host) cudaFreeHost(c->ws.live_host);
  if (c->ws.head_op) cudaFree(c->ws.head_op);
  if (c->ws.tc_scratch) cudaFree(c->ws.tc_scratch);
  if (c->ws.x3_scratch) cudaFree(c->ws.x3_scratch);
  tc_free(c);
  x3_free(c);
  if (c->link_table4) cudaFree(c->link_table4);
  if (c->robot_sel4) cudaFree(c->robot_sel4);
  if (c->robot_sel_steps) cudaFree(c->robot_sel_steps);
  if (c->loss_partial) cudaFree(c->loss_partial);
  free_train_ws(c);
  free_live_sets(c);
  if (c->w.params) cudaFree(c->w.params);
  delete c;
  return MPN_OK;
}

int mpn_reserve(mpn_ctx* c, int max_batch) {
  REQ_CTX(c);
  MPN_REQUIRE(max_batch > 0, "mpn_reserve: max_batch must be positive");
  return ensure_workspace(c, max_batch);
}

int mpn_set_robot_tables(mpn_ctx* c, const float* joint_limits, int P, const float* link_points, const int32_t* link_ids,
                         int Pe, const float* ee_points, int S, const float* sph_c, const float* sph_r, const int32_t* sph_l,
                         float prismatic) {
  REQ_CTX(c);
  MPN_REQUIRE(joint_limits && link_points && link_ids && ee_points && sph_c && sph_r && sph_l, "mpn_set_robot_tables: null table");
  MPN_REQUIRE(P >= c->cfg.n_robot && P < (1 << 23), "mpn_set_robot_tables: need n_link_points >= n_robot (%d)", c->cfg.n_robot);
  MPN_REQUIRE(Pe >= c->cfg.n_target, "mpn_set_robot_tables: need n_ee_points >= n_target (%d)", c->cfg.n_target);
  MPN_REQUIRE(S >= 1 && S <= 96, "mpn_set_robot_tables: 1..96 spheres supported");
  for (int i = 0; i < P; ++i) MPN_REQUIRE(link_ids[i] >= 0 && link_ids[i] < 11, "link id out of range at %d", i);
  for (int i = 0; i < S; ++i) MPN_REQUIRE(sph_l[i] >= 0 && sph_l[i] < 11, "sphere link id out of range at %d", i);
  int r = 0;
  memcpy(c->limits_host, joint_limits, sizeof(float) * 14);
  r |= dev_upload(&c->limits, joint_limits, 14);
  r |= dev_upload(&c->link_points, link_points, (size_t)P * 3);
  r |= dev_upload(&c->link_ids, link_ids, (size_t)P);
  r |= dev_upload(&c->ee_points, ee_points, (size_t)Pe * 3);
  r |= dev_upload(&c->sph_c, sph_c, (size_t)S * 3);
  r |= dev_upload(&c->sph_r, sph_r, (size_t)S);
  r |= dev_upload(&c->sph_l, sph_l, (size_t)S);
  if (r) return r;
  c->P = P; c->Pe = Pe; c->S = S; c->prismatic = prismatic;
  c->n_base_points = 0;
  while (c->n_base_points < P && link_ids[c->n_base_points] == 0) ++c->n_base_points;
  if ((r = pack_link_table(c))) return r;
  c->tables_set = true;
  return MPN_OK;
}

int mpn_load_weight(mpn_ctx* c, const char* name, const float* host_data, const int64_t* shape, int ndim) {
  REQ_CTX(c);
  MPN_REQUIRE(name && host_data && shape && ndim >= 1 && ndim <= 4, "mpn_load_weight: bad arguments");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { MPN_REQUIRE(shape[i] > 0, "mpn_load_weight: bad shape"); n *= (size_t)shape[i]; }
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  t.data.assign(host_data, host_data + n);
  g_host_weights[c][name] = std::move(t);
  c->w.finalized = false;
  return MPN_OK;
}

int mpn_weights_finalize(mpn_ctx* c) {
  REQ_CTX(c);
  return finalize_weights(c);
}

int64_t mpn_launch_count(mpn_ctx* c) { return c ? c->launches : 0; }

int mpn_tc_selftest(mpn_ctx* c, void* stream, const void* a, const void* b, float* d, int N, int K, int mode, int* status) {
  REQ_CTX(c);
  MPN_REQUIRE(a && b && d && status, "mpn_tc_selftest: null pointer");
  return tc_probe(c, (cudaStream_t)stream, a, b, d, N, K, mode, status);
}

int mpn_tc_gemm_selftest(mpn_ctx* c, void* stream, const float* a, const float* w, const float* bias, int M, int N, int K, float* out,
                         int split) {
  REQ_CTX(c);
  MPN_REQUIRE(a && w && bias && out, "mpn_tc_gemm_selftest: null pointer");
  return x3_gemm_selftest(c, (cudaStream_t)stream, a, w, bias, M, N, K, out, split);
}

int mpn_tc_error(mpn_ctx* c, int* out) {
  REQ_CTX(c);
  MPN_REQUIRE(out, "mpn_tc_error: null output");
  MPN_CHECK_CUDA(cudaDeviceSynchronize());
  MPN_CHECK_CUDA(cudaMemcpy(out, tc_error_flag(c), sizeof(int), cudaMemcpyDeviceToHost));
  return MPN_OK;
}

int mpn_sa_tile_counts(mpn_ctx* c, uint64_t* out, int reset) {
  REQ_CTX(c);
  MPN_REQUIRE(out, "mpn_sa_tile_counts: null output");
  unsigned long long v[2] = {0ull, 0ull};
  int r = sa_tile_counts(c, v, reset);
  out[0] = v[0];
  out[1] = v[1];
  return r;
}

int mpn_profile(mpn_ctx* c, int enable) {
  REQ_CTX(c);
  c->prof = enable != 0;
  return MPN_OK;
}

int mpn_profile_read(mpn_ctx* c, float* ms, int64_t* launches) {
  REQ_CTX(c);
  MPN_REQUIRE(ms && launches, "mpn_profile_read: null output");
  MPN_CHECK_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < MPN_NUM_STAGES; ++i) { ms[i] = 0.f; launches[i] = 0; }
  for (auto& r : c->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.stage >= 0 && r.stage < MPN_NUM_STAGES) { ms[r.stage] += t; launches[r.stage]++; }
    c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
  }
  c->prof_recs.clear();
  return MPN_OK;
}

// ---- pointnet2_ops
int mpn_fps(mpn_ctx* c, void* stream, const float* xyz, int B, int N, int stride, int npoint, int32_t* idx, float* new_xyz) {
  REQ_CTX(c);
  MPN_REQUIRE(xyz && idx && B >= 0, "mpn_fps: null pointer");
  if (B == 0) return MPN_OK;
  return launch_fps(c, (cudaStream_t)stream, xyz, B, N, stride, npoint, idx, new_xyz);
}

int mpn_ball_query(mpn_ctx* c, void* stream, float radius, int nsample, const float* xyz, int B, int N, int stride,
                   const float* new_xyz, int npoint, int32_t* idx) {
  REQ_CTX(c);
  MPN_REQUIRE(xyz && new_xyz && idx, "mpn_ball_query: null pointer");
  if (B == 0) return MPN_OK;
  return launch_ball_query(c, (cudaStream_t)stream, radius, nsample, xyz, B, N, stride, new_xyz, npoint, idx);
}

int mpn_gather_points(mpn_ctx* c, void* stream, const float* feat, int B, int C, int N, const int32_t* idx, int m, float* out) {
  REQ_CTX(c);
  MPN_REQUIRE(feat && idx && out && C >= 1 && N >= 1 && m >= 1, "mpn_gather_points: bad arguments");
  if (B == 0) return MPN_OK;
  return launch_gather(c, (cudaStream_t)stream, feat, B, C, N, idx, m, out);
}

int mpn_group_points(mpn_ctx* c, void* stream, const float* feat, int B, int C, int N, const int32_t* idx, int m, int ns, float* out) {
  REQ_CTX(c);
  MPN_REQUIRE(feat && idx && out && C >= 1 && N >= 1 && m >= 1 && ns >= 1, "mpn_group_points: bad arguments");
  if (B == 0) return MPN_OK;
  return launch_group(c, (cudaStream_t)stream, feat, B, C, N, idx, m, ns, out);
}

int mpn_sa_forward(mpn_ctx* c, void* stream, int module, int precision, const float* xyz, int stride, const float* feats,
                   int feat_stride, int B, int N, float* new_xyz, float* new_feats, int32_t* fps_idx, int32_t* ball_idx) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  MPN_REQUIRE(module >= 0 && module <= 2, "mpn_sa_forward: module must be 0..2");
  MPN_REQUIRE(precision == MPN_PREC_FP32 || ((precision == MPN_PREC_BF16 || precision == MPN_PREC_BF16X3) && module < 2),
              "mpn_sa_forward: the tensor-core per-module entries cover modules 0 and 1");
  MPN_REQUIRE(xyz && feats && new_feats && stride >= 3, "mpn_sa_forward: null pointer");
  static const int cfeat[3] = {1, 64, 256};
  MPN_REQUIRE(feat_stride >= cfeat[module], "mpn_sa_forward: feat_stride too small");
  if (B == 0) return MPN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int r;
  if ((r = ensure_workspace(c, B))) return r;
  if (module == 2) return launch_sa_simt(c, s, 2, xyz, stride, feats, feat_stride, B, N, nullptr, new_feats, nullptr);
  MPN_REQUIRE(new_xyz, "mpn_sa_forward: new_xyz required for modules 0 and 1");
  int npoint = module == 0 ? SA1_NPOINT : SA2_NPOINT;
  int32_t* idx = fps_idx ? fps_idx : reinterpret_cast<int32_t*>(c->ws.fc_a);
  if ((r = launch_fps(c, s, xyz, B, N, stride, npoint, idx, new_xyz))) return r;
  if (precision == MPN_PREC_BF16) return tc_sa_forward(c, s, module, xyz, stride, feats, feat_stride, B, N, new_xyz, new_feats, ball_idx);
  if (precision == MPN_PREC_BF16X3) return x3_sa_forward(c, s, module, xyz, stride, feats, feat_stride, B, N, new_xyz, new_feats, ball_idx);
  return launch_sa_simt(c, s, module, xyz, stride, feats, feat_stride, B, N, new_xyz, new_feats, ball_idx);
}

// ---- robofin
int mpn_fk(mpn_ctx* c, void* stream, const float* q, int B, float* frames, float* eef) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(q, "mpn_fk: null q");
  if (B == 0) return MPN_OK;
  return launch_fk(c, (cudaStream_t)stream, q, B, frames, eef);
}

int mpn_sample_robot(mpn_ctx* c, void* stream, const float* q, int B, int n, uint32_t step, float* cloud, int rows) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(q && cloud && n >= 1 && n <= c->P && rows >= n, "mpn_sample_robot: bad arguments");
  if (B == 0) return MPN_OK;
  int r;
  if ((r = ensure_workspace(c, B))) return r;
  if ((r = launch_fk(c, (cudaStream_t)stream, q, B, c->ws.frames, nullptr))) return r;
  return launch_sample_robot(c, (cudaStream_t)stream, c->ws.frames, B, n, step, cloud, rows);
}

int mpn_sample_end_effector(mpn_ctx* c, void* stream, const float* poses, int B, int n, uint32_t problem0, float* out) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(poses && out && n >= 1 && n <= c->Pe, "mpn_sample_end_effector: need 1 <= n <= %d gripper points", c->Pe);
  if (B == 0) return MPN_OK;
  return launch_sample_end_effector(c, (cudaStream_t)stream, poses, B, n, problem0, out);
}

int mpn_compute_spheres(mpn_ctx* c, void* stream, const float* q, int B, float* centers) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(q && centers, "mpn_compute_spheres: null pointer");
  if (B == 0) return MPN_OK;
  int r;
  if ((r = ensure_workspace(c, B))) return r;
  if ((r = launch_fk(c, (cudaStream_t)stream, q, B, c->ws.frames, nullptr))) return r;
  return launch_spheres(c, (cudaStream_t)stream, c->ws.frames, B, centers);
}

int mpn_normalize_joints(mpn_ctx* c, void* stream, const float* q, int n, float* q_norm) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(q && q_norm, "mpn_normalize_joints: null pointer");
  if (n == 0) return MPN_OK;
  return launch_normalize(c, (cudaStream_t)stream, q, n, q_norm, false);
}

int mpn_unnormalize_joints(mpn_ctx* c, void* stream, const float* q_norm, int n, float* q) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(q && q_norm, "mpn_unnormalize_joints: null pointer");
  if (n == 0) return MPN_OK;
  return launch_normalize(c, (cudaStream_t)stream, q_norm, n, q, true);
}

// ---- geometry
static int check_scene(const mpn_ctx* c, const mpn_scene* sc) {
  MPN_REQUIRE(sc, "null scene");
  if (c->cfg.max_cuboids > 0) MPN_REQUIRE(sc->cuboid_centers && sc->cuboid_dims && sc->cuboid_quats, "scene: null cuboid arrays");
  if (c->cfg.max_cylinders > 0)
    MPN_REQUIRE(sc->cylinder_centers && sc->cylinder_radii && sc->cylinder_heights && sc->cylinder_quats, "scene: null cylinder arrays");
  return MPN_OK;
}

int mpn_sdf_points(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, const float* points, int N, int which, float* sdf) {
  REQ_CTX(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(points && sdf && which >= 0 && which <= 2, "mpn_sdf_points: bad arguments");
  if (B == 0 || N == 0) return MPN_OK;
  return launch_sdf_points(c, (cudaStream_t)stream, *scene, B, points, N, which, sdf);
}

int mpn_build_cloud(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, const float* q0, const float* target,
                    uint32_t problem0, float* cloud) {
  REQ_CTX(c); REQ_TABLES(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(q0 && target && cloud, "mpn_build_cloud: null pointer");
  if (B == 0) return MPN_OK;
  if ((r = ensure_workspace(c, B))) return r;
  StageTimer t(c, (cudaStream_t)stream, MPN_ST_BUILD_CLOUD);
  if ((r = launch_fk(c, (cudaStream_t)stream, q0, B, c->ws.frames, nullptr))) return r;
  return launch_build_cloud(c, (cudaStream_t)stream, *scene, B, c->ws.frames, target, problem0, cloud);
}

int mpn_build_cloud_ids(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, const float* q0, const float* target,
                        const uint32_t* problem_ids, float* cloud) {
  REQ_CTX(c); REQ_TABLES(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(q0 && target && cloud && problem_ids, "mpn_build_cloud_ids: null pointer");
  if (B == 0) return MPN_OK;
  if ((r = ensure_workspace(c, B))) return r;
  StageTimer t(c, (cudaStream_t)stream, MPN_ST_BUILD_CLOUD);
  if ((r = launch_fk(c, (cudaStream_t)stream, q0, B, c->ws.frames, nullptr))) return r;
  return launch_build_cloud(c, (cudaStream_t)stream, *scene, B, c->ws.frames, target, 0u, cloud, nullptr, nullptr, 0, problem_ids);
}

int mpn_augment_joints(mpn_ctx* c, void* stream, const float* q, int B, float random_scale, const uint32_t* sample_ids, uint32_t epoch,
                       float* q_out, float* q_norm_out) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(q && q_out && q_norm_out && random_scale >= 0.f, "mpn_augment_joints: bad arguments");
  if (B == 0) return MPN_OK;
  return launch_augment_joints(c, (cudaStream_t)stream, q, B, random_scale, sample_ids, epoch, q_out, q_norm_out);
}

int mpn_clean_point_cloud(mpn_ctx* c, void* stream, const float* xyz, const float* rgba, int N, int n_out, uint32_t cloud_id,
                          float* out_xyz, float* out_rgba, int32_t* kept, int32_t* scratch) {
  REQ_CTX(c);
  MPN_REQUIRE(xyz && out_xyz && kept && scratch && N >= 1 && n_out >= 1, "mpn_clean_point_cloud: bad arguments");
  MPN_REQUIRE((rgba == nullptr) == (out_rgba == nullptr), "mpn_clean_point_cloud: rgba and out_rgba go together");
  return launch_clean_point_cloud(c, (cudaStream_t)stream, xyz, rgba, N, n_out, cloud_id, out_xyz, out_rgba, kept, scratch);
}

int mpn_build_cloud_from_points(mpn_ctx* c, void* stream, int B, const float* q0, const float* target, const float* obstacle_points,
                                const int32_t* obstacle_counts, int max_points, uint32_t problem0, float* cloud) {
  REQ_CTX(c); REQ_TABLES(c);
  MPN_REQUIRE(q0 && target && cloud && obstacle_points && obstacle_counts && max_points >= 1, "mpn_build_cloud_from_points: bad arguments");
  if (B == 0) return MPN_OK;
  int r;
  if ((r = ensure_workspace(c, B))) return r;
  StageTimer t(c, (cudaStream_t)stream, MPN_ST_BUILD_CLOUD);
  if ((r = launch_fk(c, (cudaStream_t)stream, q0, B, c->ws.frames, nullptr))) return r;
  mpn_scene none{};
  return launch_build_cloud(c, (cudaStream_t)stream, none, B, c->ws.frames, target, problem0, cloud, obstacle_points, obstacle_counts,
                            max_points);
}

int mpn_render_depth_cloud(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, const float* camera, int per_problem_camera,
                           int width, int height, float tan_half_fov_x, float tan_half_fov_y, float near_depth, float far_depth,
                           float* points, int32_t* counts) {
  REQ_CTX(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(camera && points && counts, "mpn_render_depth_cloud: null pointer");
  MPN_REQUIRE(width >= 1 && height >= 1 && (int64_t)width * height <= (1 << 24), "mpn_render_depth_cloud: bad image size");
  MPN_REQUIRE(tan_half_fov_x > 0.f && tan_half_fov_y > 0.f && near_depth >= 0.f && far_depth > near_depth,
              "mpn_render_depth_cloud: bad intrinsics");
  if (B == 0) return MPN_OK;
  return launch_render_depth(c, (cudaStream_t)stream, *scene, B, camera, per_problem_camera, width, height, tan_half_fov_x,
                             tan_half_fov_y, near_depth, far_depth, points, counts);
}

int mpn_sweep_flags(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, const float* traj, int T, int t0, int accumulate,
                    uint8_t* flags, int32_t* first_step) {
  REQ_CTX(c); REQ_TABLES(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(traj && flags && T >= 1, "mpn_sweep_flags: bad arguments");
  if (B == 0) return MPN_OK;
  return launch_sweep(c, (cudaStream_t)stream, *scene, B, traj, T, T * 7, t0, accumulate, flags, first_step);
}

int mpn_evaluate(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, const float* traj, int n_poses_max,
                 const int32_t* num_poses, const float* target, const mpn_scene* target_volume, int tv_cuboids, int tv_cylinders,
                 const mpn_scene* negative_volumes, int nv_cuboids, int nv_cylinders, float* eval) {
  REQ_CTX(c); REQ_TABLES(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(traj && target && eval, "mpn_evaluate: null pointer");
  MPN_REQUIRE(n_poses_max >= 1 && n_poses_max <= 2048, "mpn_evaluate: 1 <= n_poses_max <= 2048");
  MPN_REQUIRE(tv_cuboids >= 0 && tv_cylinders >= 0 && nv_cuboids >= 0 && nv_cylinders >= 0, "mpn_evaluate: negative volume count");
  MPN_REQUIRE(tv_cuboids + tv_cylinders == 0 || target_volume, "mpn_evaluate: target_volume is null");
  MPN_REQUIRE(nv_cuboids + nv_cylinders == 0 || negative_volumes, "mpn_evaluate: negative_volumes is null");
  if (target_volume) {
    MPN_REQUIRE(tv_cuboids == 0 || (target_volume->cuboid_centers && target_volume->cuboid_dims && target_volume->cuboid_quats),
                "mpn_evaluate: target_volume cuboid arrays missing");
    MPN_REQUIRE(tv_cylinders == 0 || (target_volume->cylinder_centers && target_volume->cylinder_radii &&
                                      target_volume->cylinder_heights && target_volume->cylinder_quats),
                "mpn_evaluate: target_volume cylinder arrays missing");
  }
  if (negative_volumes) {
    MPN_REQUIRE(nv_cuboids == 0 || (negative_volumes->cuboid_centers && negative_volumes->cuboid_dims && negative_volumes->cuboid_quats),
                "mpn_evaluate: negative_volumes cuboid arrays missing");
    MPN_REQUIRE(nv_cylinders == 0 || (negative_volumes->cylinder_centers && negative_volumes->cylinder_radii &&
                                      negative_volumes->cylinder_heights && negative_volumes->cylinder_quats),
                "mpn_evaluate: negative_volumes cylinder arrays missing");
  }
  if (B == 0) return MPN_OK;
  mpn_scene none{};
  return launch_evaluate(c, (cudaStream_t)stream, *scene, B, traj, n_poses_max, num_poses, target,
                         target_volume ? *target_volume : none, tv_cuboids, tv_cylinders,
                         negative_volumes ? *negative_volumes : none, nv_cuboids, nv_cylinders, eval);
}

int mpn_sparc(mpn_ctx* c, void* stream, int B, int n_max, const float* movement, const int32_t* num_samples, float fs, int padlevel,
              float fc, float amp_th, float* sal) {
  REQ_CTX(c);
  MPN_REQUIRE(movement && sal && B >= 0 && n_max >= 1 && padlevel >= 0 && fs > 0.f, "mpn_sparc: bad arguments");
  if (B == 0) return MPN_OK;
  return launch_sparc(c, (cudaStream_t)stream, B, n_max, movement, num_samples, fs, padlevel, fc, amp_th, sal);
}

// ---- losses (loss.py)
int mpn_collision_loss(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, int N, const float* points, float margin, float* loss,
                       float* grad_points) {
  REQ_CTX(c); REQ_TABLES(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(points && loss && B >= 1 && N >= 1, "mpn_collision_loss: bad arguments");
  return launch_collision_loss(c, (cudaStream_t)stream, *scene, B, N, points, margin, loss, grad_points);
}

int mpn_point_match_loss(mpn_ctx* c, void* stream, int64_t n, const float* a, const float* b, float* loss, float* grad_a) {
  REQ_CTX(c);
  MPN_REQUIRE(a && b && loss && n >= 1, "mpn_point_match_loss: bad arguments");
  return launch_point_match_loss(c, (cudaStream_t)stream, (size_t)n, a, b, loss, grad_a);
}

int mpn_bc_collision_losses(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, const float* input_normalized,
                            const float* target_normalized, int n_points, float margin, float w_collision, float w_bc, float* losses,
                            float* grad_input) {
  REQ_CTX(c); REQ_TABLES(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(input_normalized && target_normalized && losses && B >= 1, "mpn_bc_collision_losses: bad arguments");
  return launch_bc_collision_losses(c, (cudaStream_t)stream, *scene, B, input_normalized, target_normalized, n_points, margin,
                                    w_collision, w_bc, losses, grad_input);
}

// ---- model
int mpn_encoder_forward(mpn_ctx* c, void* stream, int precision, const float* cloud, int B, int N, float* out) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  MPN_REQUIRE(cloud && out, "mpn_encoder_forward: null pointer");
  MPN_REQUIRE(N >= SA1_NPOINT && N <= 8192, "mpn_encoder_forward: N=%d unsupported (512..8192)", N);
  MPN_REQUIRE(precision == MPN_PREC_FP32 || precision == MPN_PREC_BF16 || precision == MPN_PREC_BF16X3, "bad precision");
  if (B == 0) return MPN_OK;
  int r;
  if ((r = ensure_workspace(c, B))) return r;
  return encoder_forward(c, (cudaStream_t)stream, precision, cloud, B, N, out, ENC_DIM);
}

int mpn_policy_forward(mpn_ctx* c, void* stream, int precision, const float* cloud, const float* q_norm, int B, int N, float* dq) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  MPN_REQUIRE(cloud && q_norm && dq, "mpn_policy_forward: null pointer");
  MPN_REQUIRE(N >= SA1_NPOINT && N <= 8192, "mpn_policy_forward: N=%d unsupported (512..8192)", N);
  MPN_REQUIRE(precision == MPN_PREC_FP32 || precision == MPN_PREC_BF16 || precision == MPN_PREC_BF16X3, "bad precision");
  if (B == 0) return MPN_OK;
  int r;
  if ((r = ensure_workspace(c, B))) return r;
  return policy_forward(c, (cudaStream_t)stream, precision, cloud, q_norm, B, N, dq);
}

// ---- training step (model.py:185-240, 68-73; run_training.py:112)
int64_t mpn_param_count(mpn_ctx* c) { return (c && c->w.finalized) ? c->w.n_params : 0; }

int mpn_param_info(mpn_ctx* c, int index, char* name, int name_cap, int64_t* offset, int64_t* numel) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  if (index < 0 || index >= (int)c->w.info.size()) return MPN_ERR_INVALID;   // end-of-list marker, no message
  const ParamInfo& p = c->w.info[index];
  if (name && name_cap > 0) { strncpy(name, p.name.c_str(), (size_t)name_cap - 1); name[name_cap - 1] = 0; }
  if (offset) *offset = p.offset;
  if (numel) *numel = p.numel;
  return MPN_OK;
}

int mpn_get_params(mpn_ctx* c, void* stream, float* dst) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  MPN_REQUIRE(dst, "mpn_get_params: null destination");
  MPN_CHECK_CUDA(cudaMemcpyAsync(dst, c->w.params, (size_t)c->w.n_params * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MPN_OK;
}

int mpn_set_params(mpn_ctx* c, void* stream, const float* src) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  MPN_REQUIRE(src, "mpn_set_params: null source");
  MPN_CHECK_CUDA(cudaMemcpyAsync(c->w.params, src, (size_t)c->w.n_params * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return refresh_transposes(c, (cudaStream_t)stream);
}

int mpn_weights_sync(mpn_ctx* c) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  MPN_CHECK_CUDA(cudaDeviceSynchronize());
  return tc_prepare_weights(c);
}

int mpn_train_step_grads(mpn_ctx* c, void* stream, const mpn_scene* scene, int B, int N, const float* cloud, const float* q_norm,
                         const float* supervision, int n_loss_points, float margin, float w_collision, float w_bc, float* losses,
                         float* y_hat, float* grads, int precision) {
  REQ_CTX(c); REQ_TABLES(c); REQ_WEIGHTS(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(cloud && q_norm && supervision && losses && B >= 1, "mpn_train_step_grads: bad arguments");
  MPN_REQUIRE(N >= SA1_NPOINT && N <= 8192, "mpn_train_step_grads: N=%d unsupported (512..8192)", N);
  MPN_REQUIRE(n_loss_points >= 1, "mpn_train_step_grads: n_loss_points must be positive");
  MPN_REQUIRE(precision == MPN_PREC_FP32 || precision == MPN_PREC_BF16, "mpn_train_step_grads: bad precision");
  if ((r = ensure_workspace(c, B))) return r;
  return train_step_grads(c, (cudaStream_t)stream, *scene, B, N, cloud, q_norm, supervision, n_loss_points, margin, w_collision, w_bc,
                          losses, y_hat, grads, precision);
}

// tensor-core building blocks of the training backward, exposed for their own parity tests
int mpn_train_tc_gemm(mpn_ctx* c, void* stream, int epi, const void* A, const void* W, const float* bias, const void* mask, int64_t M,
                      int N, void* C) {
  REQ_CTX(c);
  MPN_REQUIRE(A && W && C && epi >= 0 && epi <= 2, "mpn_train_tc_gemm: bad arguments");
  return launch_rows_gemm_tc(c, (cudaStream_t)stream, epi, (const __nv_bfloat16*)A, (const __nv_bfloat16*)W, bias,
                             (const __nv_bfloat16*)mask, M, N, (__nv_bfloat16*)C);
}

int mpn_train_tc_wgrad(mpn_ctx* c, void* stream, const void* dY, const void* X, int64_t R, float* partial, int64_t partial_floats,
                       int* n_ctas, int variant) {
  REQ_CTX(c);
  MPN_REQUIRE(dY && X && partial && n_ctas && R >= 1 && partial_floats >= 128 * 128, "mpn_train_tc_wgrad: bad arguments");
  return launch_wgrad_tc(c, (cudaStream_t)stream, (const __nv_bfloat16*)dY, (const __nv_bfloat16*)X, R, partial, (size_t)partial_floats,
                         n_ctas, variant);
}

int mpn_train_pooled_rows(mpn_ctx* c, void* stream, int module, int B, uint8_t* dst) {
  REQ_CTX(c);
  MPN_REQUIRE(module >= 0 && module <= 2 && dst, "mpn_train_pooled_rows: bad arguments");
  MPN_REQUIRE(B >= 1 && B <= c->tw.capacity, "mpn_train_pooled_rows: no training step of >= %d samples has run", B);
  const uint8_t* src = module == 0 ? c->tw.arg1 : module == 1 ? c->tw.arg2 : c->tw.arg3;
  const size_t per = module == 0 ? (size_t)SA1_NPOINT * 64 : module == 1 ? (size_t)SA2_NPOINT * 256 : 1024;
  MPN_CHECK_CUDA(cudaMemcpyAsync(dst, src, per * B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MPN_OK;
}

int mpn_adam_step(mpn_ctx* c, void* stream, const float* grads, float lr, float beta1, float beta2, float eps, float clip_norm,
                  int step, float* grad_norm) {
  REQ_CTX(c); REQ_WEIGHTS(c);
  MPN_REQUIRE(grads && step >= 1 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "mpn_adam_step: bad arguments");
  return adam_step(c, (cudaStream_t)stream, grads, lr, beta1, beta2, eps, clip_norm, step, grad_norm);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- early exit: compaction of the live problems
// rollout_until_success (run_inference.py:137-191) stops a problem when it reaches its target; in a lock-step batch the stopped problems
// would keep costing a full policy step.  At a poll point where at most half of the current problems are still running, their state
// (cloud, joint state, frames, target, scene rows, flags) is gathered into a compact set and the loop goes on over that set only; what the
// abandoned set produced is scattered back to the caller's arrays first.  Problems are independent and the per-step robot subset is
// shared by the batch, so the order inside the compact set does not matter and results equal the uncompacted rollout's bit for bit.
namespace {
struct LiveSet {
  int cap = 0, n_points = 0, m1 = 0, m2 = 0, stride = 0;
  float *cloud = nullptr, *target = nullptr, *qn = nullptr, *qu = nullptr, *frames = nullptr, *eef = nullptr, *traj = nullptr;
  float *cub_c = nullptr, *cub_d = nullptr, *cub_q = nullptr, *cyl_c = nullptr, *cyl_r = nullptr, *cyl_h = nullptr, *cyl_q = nullptr;
  int32_t *done = nullptr, *first = nullptr, *map = nullptr, *sel = nullptr;
  uint8_t* flags = nullptr;
};
struct LiveSets { LiveSet s[2]; };

void free_live_set(LiveSet& L) {
  void* ps[] = {L.cloud, L.target, L.qn, L.qu, L.frames, L.eef, L.traj, L.cub_c, L.cub_d, L.cub_q, L.cyl_c, L.cyl_r, L.cyl_h, L.cyl_q,
                L.done, L.first, L.map, L.sel, L.flags};
  for (void* q : ps) if (q) cudaFree(q);
  L = LiveSet();
}
int ensure_live_set(LiveSet& L, int cap, int N, int m1, int m2, int stride) {
  if (cap <= L.cap && N <= L.n_points && m1 <= L.m1 && m2 <= L.m2 && stride <= L.stride) return MPN_OK;
  free_live_set(L);
  const size_t b = (size_t)cap;
  bool ok = true;
  auto A = [&](auto** q, size_t n) { ok = ok && cudaMalloc((void**)q, n * sizeof(**q)) == cudaSuccess; };
  A(&L.cloud, b * N * 4); A(&L.target, b * 12); A(&L.qn, b * 7); A(&L.qu, b * 7); A(&L.frames, b * MPN_NLINK * 12); A(&L.eef, b * 12);
  A(&L.traj, b * stride); A(&L.cub_c, b * m1 * 3); A(&L.cub_d, b * m1 * 3); A(&L.cub_q, b * m1 * 4); A(&L.cyl_c, b * m2 * 3);
  A(&L.cyl_r, b * m2); A(&L.cyl_h, b * m2); A(&L.cyl_q, b * m2 * 4); A(&L.done, b); A(&L.first, b); A(&L.map, b); A(&L.sel, b); A(&L.flags, b);
  if (!ok) { free_live_set(L); mpn::set_error("early-exit compaction: cudaMalloc failed"); return MPN_ERR_NOMEM; }
  L.cap = cap; L.n_points = N; L.m1 = m1; L.m2 = m2; L.stride = stride;
  return MPN_OK;
}

// sel[0..live) = indices of the problems that are still running (any order)
__global__ void select_live_kernel(int B, const int32_t* __restrict__ done, int32_t* __restrict__ sel, int32_t* __restrict__ counter) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const bool alive = b < B && done[b] < 0;
  const unsigned m = __ballot_sync(0xffffffffu, alive);
  int base = 0;
  if (lane == 0 && m) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (alive) sel[base + __popc(m & ((1u << lane) - 1u))] = b;
}
// dst[i][0..cols) = src[idx ? idx[i] : i][col0 .. col0 + cols)   (row pitches in elements)
template <typename T>
__global__ void gather_rows_kernel(T* __restrict__ dst, int dst_pitch, const T* __restrict__ src, int src_pitch, const int32_t* __restrict__ idx,
                                   int rows, int cols) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * cols; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), k = (int)(i - (long long)r * cols);
    dst[(size_t)r * dst_pitch + k] = src[(size_t)idx[r] * src_pitch + k];
  }
}
// dst[idx[i]][col0 .. col0 + cols) = src[i][col0 .. col0 + cols)
template <typename T>
__global__ void scatter_rows_kernel(T* __restrict__ dst, int dst_pitch, const T* __restrict__ src, int src_pitch, const int32_t* __restrict__ idx,
                                    int rows, int col0, int cols) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * cols; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), k = col0 + (int)(i - (long long)r * cols);
    dst[(size_t)idx[r] * dst_pitch + k] = src[(size_t)r * src_pitch + k];
  }
}
__global__ void compose_map_kernel(int32_t* __restrict__ dst, const int32_t* __restrict__ outer, const int32_t* __restrict__ sel, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = outer ? outer[sel[i]] : sel[i];
}
template <typename T>
int gather_rows(mpn_ctx* c, cudaStream_t s, T* dst, int dst_pitch, const T* src, int src_pitch, const int32_t* idx, int rows, int cols) {
  const long long n = (long long)rows * cols;
  gather_rows_kernel<T><<<(unsigned)std::min<long long>((n + 255) / 256, 8LL * c->sm_count), 256, 0, s>>>(dst, dst_pitch, src, src_pitch, idx, rows, cols);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}
template <typename T>
int scatter_rows(mpn_ctx* c, cudaStream_t s, T* dst, int dst_pitch, const T* src, int src_pitch, const int32_t* idx, int rows, int col0, int cols) {
  const long long n = (long long)rows * cols;
  if (n <= 0) return MPN_OK;
  scatter_rows_kernel<T><<<(unsigned)std::min<long long>((n + 255) / 256, 8LL * c->sm_count), 256, 0, s>>>(dst, dst_pitch, src, src_pitch, idx, rows, col0, cols);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}
}  // namespace

void free_live_sets(mpn_ctx* c) {
  if (!c->live_sets) return;
  LiveSets* ls = static_cast<LiveSets*>(c->live_sets);
  free_live_set(ls->s[0]);
  free_live_set(ls->s[1]);
  delete ls;
  c->live_sets = nullptr;
}

extern "C" {

int mpn_rollout(mpn_ctx* c, void* stream, int precision, const mpn_scene* scene, int B, int N, float* cloud, const float* q0,
                const float* target, int T, int early_exit, int check_every_step, float* traj, float* metrics) {
  REQ_CTX(c); REQ_TABLES(c); REQ_WEIGHTS(c);
  int r;
  if ((r = check_scene(c, scene))) return r;
  MPN_REQUIRE(cloud && q0 && target && traj && metrics && T >= 1, "mpn_rollout: bad arguments");
  MPN_REQUIRE(N >= c->cfg.n_robot && N >= SA1_NPOINT && N <= 8192, "mpn_rollout: N=%d unsupported", N);
  MPN_REQUIRE(precision == MPN_PREC_FP32 || precision == MPN_PREC_BF16 || precision == MPN_PREC_BF16X3, "bad precision");
  if (B == 0) return MPN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if ((r = ensure_workspace(c, B))) return r;
  Workspace& w = c->ws;
  const int stride = (T + 1) * 7;
  if ((r = launch_normalize(c, s, q0, B, w.qn, false))) return r;
  MPN_CHECK_CUDA(cudaMemcpy2DAsync(traj, (size_t)stride * 4, q0, 7 * 4, 7 * 4, B, cudaMemcpyDeviceToDevice, s));
  MPN_CHECK_CUDA(cudaMemsetAsync(w.done, 0xff, (size_t)B * 4, s));
  MPN_CHECK_CUDA(cudaMemsetAsync(w.first_step, 0xff, (size_t)B * 4, s));
  MPN_CHECK_CUDA(cudaMemsetAsync(w.flags, 0, (size_t)B, s));
  if ((r = launch_fk(c, s, q0, B, w.frames, w.eef))) return r;
  if (check_every_step && (r = launch_sweep(c, s, *scene, B, traj, 1, stride, 0, 0, w.flags, w.first_step, w.frames))) return r;
  if ((r = launch_robot_subsets(c, s, c->cfg.n_robot, 1u, T))) return r;   // the per-step robot subsets, all at once
  // early_exit == 1: every EARLY_EXIT_POLL steps the host reads the number of problems still running and stops launching once every
  // problem has stopped (rollout_until_success breaks out of its loop, run_inference.py:180-187) -- the one place this call
  // synchronises the stream; early_exit == 2 keeps the done mask without polling (capturable in a CUDA graph).
  constexpr int EARLY_EXIT_POLL = 8;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(s, &cap);
  const bool poll = early_exit == 1 && cap == cudaStreamCaptureStatusNone;
  const bool allow_compact = poll && getenv("MPN_NO_LIVE_COMPACTION") == nullptr;
  // the CURRENT set of problems: the caller's arrays at first, a compact LiveSet after a compaction (cmap: its row -> caller's row)
  int cb = B, cstart = 0;                 // problems in the current set; the step after which it was formed
  float* ccloud = cloud; const float* ctarget = target; mpn_scene cscene = *scene; float* ctraj = traj;
  float *cqn = w.qn, *cqu = w.qu, *cframes = w.frames, *ceef = w.eef;
  int32_t *cdone = w.done, *cfirst = w.first_step; uint8_t* cflags = w.flags; const int32_t* cmap = nullptr;
  int cur_set = -1;
  const int nrob4 = c->cfg.n_robot * 4;
  // what a compact set produced since it was formed goes back to the caller's arrays (trajectory columns, robot rows of the cloud,
  // joint state, end-effector pose, done / collision bookkeeping)
  auto scatter_back = [&](int upto_step) -> int {
    if (!cmap) return MPN_OK;
    int rr;
    if ((rr = scatter_rows<float>(c, s, traj, stride, ctraj, stride, cmap, cb, (cstart + 1) * 7, (upto_step - cstart) * 7))) return rr;
    if ((rr = scatter_rows<float>(c, s, cloud, N * 4, ccloud, N * 4, cmap, cb, 0, nrob4))) return rr;
    if ((rr = scatter_rows<float>(c, s, w.qn, 7, cqn, 7, cmap, cb, 0, 7))) return rr;
    if ((rr = scatter_rows<float>(c, s, w.qu, 7, cqu, 7, cmap, cb, 0, 7))) return rr;
    if ((rr = scatter_rows<float>(c, s, w.eef, 12, ceef, 12, cmap, cb, 0, 12))) return rr;
    if ((rr = scatter_rows<float>(c, s, w.frames, MPN_NLINK * 12, cframes, MPN_NLINK * 12, cmap, cb, 0, MPN_NLINK * 12))) return rr;
    if ((rr = scatter_rows<int32_t>(c, s, w.done, 1, cdone, 1, cmap, cb, 0, 1))) return rr;
    if ((rr = scatter_rows<int32_t>(c, s, w.first_step, 1, cfirst, 1, cmap, cb, 0, 1))) return rr;
    return scatter_rows<uint8_t>(c, s, w.flags, 1, cflags, 1, cmap, cb, 0, 1);
  };
  int last_step = T;
  for (int i = 1; i <= T; ++i) {
    if ((r = policy_forward(c, s, precision, ccloud, cqn, cb, N, w.dq))) return r;
    { StageTimer t(c, s, MPN_ST_UPDATE);
      if ((r = launch_step_update(c, s, cb, w.dq, cqn, cqu, ctarget, cdone, early_exit, ctraj, stride, cframes, ceef, nullptr, i))) return r; }
    { StageTimer t(c, s, MPN_ST_SAMPLE_ROBOT);
      if ((r = launch_sample_robot_slab(c, s, cframes, cb, c->cfg.n_robot, i - 1, ccloud, N))) return r; }
    if (check_every_step) {
      StageTimer t(c, s, MPN_ST_SWEEP);
      if ((r = launch_sweep(c, s, cscene, cb, ctraj + (size_t)i * 7, 1, stride, i, 1, cflags, cfirst, cframes))) return r;
    }
    if (poll && i % EARLY_EXIT_POLL == 0 && i < T) {
      if ((r = launch_count_live(c, s, cb, cdone, w.live))) return r;
      MPN_CHECK_CUDA(cudaMemcpyAsync(w.live_host, w.live, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
      MPN_CHECK_CUDA(cudaStreamSynchronize(s));
      const int live = *w.live_host;
      if (live == 0) {   // everybody has stopped: the remaining rows repeat the frozen configurations
        if ((r = launch_fill_traj_tail(c, s, cb, cqu, ctraj, stride, i + 1, T))) return r;
        last_step = T;   // the tail columns of the current set are final
        break;
      }
      if (allow_compact && live <= cb / 2) {
        // the stopped problems of the current set are finished: their tails repeat the frozen configuration; hand everything the set
        // produced back, then carry only the running problems on
        if ((r = launch_fill_traj_tail(c, s, cb, cqu, ctraj, stride, i + 1, T))) return r;
        if ((r = scatter_back(T))) return r;
        if (!c->live_sets) c->live_sets = new LiveSets();
        LiveSets* ls = static_cast<LiveSets*>(c->live_sets);
        const int nxt = cur_set == 0 ? 1 : 0;
        LiveSet& L = ls->s[nxt];
        if ((r = ensure_live_set(L, cur_set < 0 ? (B + 1) / 2 : live, N, c->cfg.max_cuboids, c->cfg.max_cylinders, stride))) return r;
        MPN_CHECK_CUDA(cudaMemsetAsync(w.live, 0, sizeof(int32_t), s));
        select_live_kernel<<<(cb + 255) / 256, 256, 0, s>>>(cb, cdone, L.sel, w.live);
        compose_map_kernel<<<(live + 255) / 256, 256, 0, s>>>(L.map, cmap, L.sel, live);
        c->launches += 2;
        MPN_CHECK_CUDA(cudaGetLastError());
        const int m1 = c->cfg.max_cuboids, m2 = c->cfg.max_cylinders;
        if ((r = gather_rows<float>(c, s, L.cloud, N * 4, ccloud, N * 4, L.sel, live, N * 4))) return r;
        if ((r = gather_rows<float>(c, s, L.target, 12, ctarget, 12, L.sel, live, 12))) return r;
        if ((r = gather_rows<float>(c, s, L.qn, 7, cqn, 7, L.sel, live, 7))) return r;
        if ((r = gather_rows<float>(c, s, L.qu, 7, cqu, 7, L.sel, live, 7))) return r;
        if ((r = gather_rows<float>(c, s, L.frames, MPN_NLINK * 12, cframes, MPN_NLINK * 12, L.sel, live, MPN_NLINK * 12))) return r;
        if ((r = gather_rows<float>(c, s, L.eef, 12, ceef, 12, L.sel, live, 12))) return r;
        if ((r = gather_rows<int32_t>(c, s, L.done, 1, cdone, 1, L.sel, live, 1))) return r;
        if ((r = gather_rows<int32_t>(c, s, L.first, 1, cfirst, 1, L.sel, live, 1))) return r;
        if ((r = gather_rows<uint8_t>(c, s, L.flags, 1, cflags, 1, L.sel, live, 1))) return r;
        if ((r = gather_rows<float>(c, s, L.cub_c, m1 * 3, cscene.cuboid_centers, m1 * 3, L.sel, live, m1 * 3))) return r;
        if ((r = gather_rows<float>(c, s, L.cub_d, m1 * 3, cscene.cuboid_dims, m1 * 3, L.sel, live, m1 * 3))) return r;
        if ((r = gather_rows<float>(c, s, L.cub_q, m1 * 4, cscene.cuboid_quats, m1 * 4, L.sel, live, m1 * 4))) return r;
        if ((r = gather_rows<float>(c, s, L.cyl_c, m2 * 3, cscene.cylinder_centers, m2 * 3, L.sel, live, m2 * 3))) return r;
        if ((r = gather_rows<float>(c, s, L.cyl_r, m2, cscene.cylinder_radii, m2, L.sel, live, m2))) return r;
        if ((r = gather_rows<float>(c, s, L.cyl_h, m2, cscene.cylinder_heights, m2, L.sel, live, m2))) return r;
        if ((r = gather_rows<float>(c, s, L.cyl_q, m2 * 4, cscene.cylinder_quats, m2 * 4, L.sel, live, m2 * 4))) return r;
        cb = live; cstart = i; cur_set = nxt; cmap = L.map;
        ccloud = L.cloud; ctarget = L.target; ctraj = L.traj; cqn = L.qn; cqu = L.qu; cframes = L.frames; ceef = L.eef;
        cdone = L.done; cfirst = L.first; cflags = L.flags;
        cscene.cuboid_centers = L.cub_c; cscene.cuboid_dims = L.cub_d; cscene.cuboid_quats = L.cub_q; cscene.cylinder_centers = L.cyl_c;
        cscene.cylinder_radii = L.cyl_r; cscene.cylinder_heights = L.cyl_h; cscene.cylinder_quats = L.cyl_q;
      }
    }
  }
  if ((r = scatter_back(last_step))) return r;
  if (!check_every_step) {
    StageTimer t(c, s, MPN_ST_SWEEP);
    if ((r = launch_sweep(c, s, *scene, B, traj, T + 1, stride, 0, 0, w.flags, w.first_step))) return r;
  }
  return launch_finalize_metrics(c, s, B, w.eef, target, w.flags, w.first_step, w.done, T, metrics);
}

}  // extern "C"
