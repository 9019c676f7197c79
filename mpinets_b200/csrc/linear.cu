// linear.cu -- fp32 (parity-mode) dense layers of the encoder head and the policy head, GroupNorm+LeakyReLU, and the
// per-step joint update.  Replaces nn.Linear / nn.GroupNorm(16,.) / nn.LeakyReLU in mpinets/model.py:47-66,385-393 and
// the clamp / unnormalise / success test of model.py:171-174 and run_inference.py:172-187.
#include "engine.h"
#include "spec_math.cuh"

namespace mpn {

// Y[M][N] = act(X[M][K] * W[N][K]^T + b).  64x64x16 tiles, 256 threads, 4x4 micro-tiles, fp32 FMA.
constexpr int GT = 64, GK = 16;

// blockIdx.x = row tile (M can be millions of compacted rows in the training backward), blockIdx.y = column tile.
// mask (optional): the result is multiplied by the activation derivative taken from mask[m][n] (1: LeakyReLU(0.01) of a
// post-activation value, 2: ReLU); Y may alias mask (each element is read, then written, by the same thread).
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                                                     const float* __restrict__ bias, long long M, int N, int K,
                                                     float* Y, int ldy, int act, const float* mask, int ldmask,
                                                     int mask_mode) {
  __shared__ float xs[GK][GT + 4];
  __shared__ float ws[GK][GT + 4];
  const long long m0 = (long long)blockIdx.x * GT;
  const int n0 = blockIdx.y * GT;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = threadIdx.x / 4, lk = (threadIdx.x % 4) * 4;  // loader: row lr (0..63), k offset lk (0,4,8,12)
  for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int k = k0 + lk + u;
      long long m = m0 + lr;
      int n = n0 + lr;
      xs[lk + u][lr] = (m < M && k < K) ? X[(size_t)m * ldx + k] : 0.f;
      ws[lk + u][lr] = (n < N && k < K) ? W[(size_t)n * ldw + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = xs[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = ws[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = bias ? acc[i][j] + bias[n] : acc[i][j];
      if (act == 1) v = v > 0.f ? v : 0.01f * v;
      else if (act == 2) v = fmaxf(v, 0.f);
      if (mask) {
        float a = mask[(size_t)m * ldmask + n];
        if (mask_mode == 1) v = a > 0.f ? v : 0.01f * v;
        else v = a > 0.f ? v : 0.f;
      }
      Y[(size_t)m * ldy + n] = v;
    }
  }
}

int launch_linear_ex(mpn_ctx* c, cudaStream_t s, const float* X, int ldx, const float* Wm, int ldw, const float* bias, int64_t M, int N,
                     int K, float* Y, int ldy, int act, const float* mask, int ldmask, int mask_mode) {
  if (M <= 0) return MPN_OK;
  dim3 grid((unsigned)((M + GT - 1) / GT), (N + GT - 1) / GT);
  linear_kernel<<<grid, 256, 0, s>>>(X, ldx, Wm, ldw, bias, (long long)M, N, K, Y, ldy, act, mask, ldmask, mask_mode);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

int launch_linear(mpn_ctx* c, cudaStream_t s, const Linear& L, const float* x, int ldx, int M, float* y, int ldy, int act) {
  return launch_linear_ex(c, s, x, ldx, L.w, L.in, L.b, M, L.out, L.in, y, ldy, act);
}

// GroupNorm(groups, C) (eps 1e-5, biased variance, affine) + LeakyReLU(0.01), in place.  One warp per (row, group).
__global__ void __launch_bounds__(256) groupnorm_lrelu_kernel(float* __restrict__ x, int M, int C, int groups,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32 = nullptr,
                                                              float* __restrict__ stats = nullptr, int split = 0) {
  int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= M * groups) return;
  int row = wid / groups, g = wid % groups, gs = C / groups;
  float* p = x + (size_t)row * C + (size_t)g * gs;
  float sum = 0.f;
  for (int i = lane; i < gs; i += 32) sum += p[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  float mean = sum / (float)gs;
  float sq = 0.f;
  for (int i = lane; i < gs; i += 32) { float d = p[i] - mean; sq = fmaf(d, d, sq); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  float rstd = 1.0f / sqrtf(sq / (float)gs + 1e-5f);
  if (stats && lane == 0) { stats[2 * (size_t)wid] = mean; stats[2 * (size_t)wid + 1] = rstd; }
  for (int i = lane; i < gs; i += 32) {
    int ch = g * gs + i;
    float v = (p[i] - mean) * rstd * gamma[ch] + beta[ch];
    v = v > 0.f ? v : 0.01f * v;
    if (out_bf16 && split) {   // [hi | lo] row of the split-bf16 GEMM that follows: pitch 2C
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out_bf16[(size_t)row * 2 * C + ch] = h;
      out_bf16[(size_t)row * 2 * C + C + ch] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    else if (out_bf16) out_bf16[(size_t)row * C + ch] = __float2bfloat16_rn(v);
    else if (out_f32) out_f32[(size_t)row * C + ch] = v;
    else p[i] = v;
  }
}

int launch_groupnorm_lrelu(mpn_ctx* c, cudaStream_t s, float* x, int M, int C, int groups, const float* gamma, const float* beta) {
  int warps = M * groups;
  groupnorm_lrelu_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(x, M, C, groups, gamma, beta, nullptr);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

int launch_groupnorm_lrelu_train(mpn_ctx* c, cudaStream_t s, const float* z, int M, int C, int groups, const float* gamma,
                                 const float* beta, float* out, float* stats) {
  int warps = M * groups;
  groupnorm_lrelu_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(const_cast<float*>(z), M, C, groups, gamma, beta, nullptr, out, stats);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// same, reading fp32 and writing the bf16 operand of the next tensor-core GEMM
int launch_groupnorm_lrelu_bf16(mpn_ctx* c, cudaStream_t s, float* x, int M, int C, int groups, const float* gamma, const float* beta,
                                __nv_bfloat16* out) {
  int warps = M * groups;
  groupnorm_lrelu_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(x, M, C, groups, gamma, beta, out);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// same, writing the [hi | lo] split-bf16 operand rows ([M][2C]) of the bf16x3 mode
int launch_groupnorm_lrelu_split(mpn_ctx* c, cudaStream_t s, float* x, int M, int C, int groups, const float* gamma, const float* beta,
                                 __nv_bfloat16* out) {
  int warps = M * groups;
  groupnorm_lrelu_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(x, M, C, groups, gamma, beta, out, nullptr, nullptr, 1);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// rotation angle (degrees) between two 3x4 poses
__device__ __forceinline__ float pose_angle_deg(const float* A, const float* Bp) {
  float tr = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) tr = fmaf(A[i * 4 + j], Bp[i * 4 + j], tr);
  float cs = fminf(1.0f, fmaxf(-1.0f, (tr - 1.0f) * 0.5f));
  return acosf(cs) * 57.29577951308232f;
}

// q = clamp(q + dq, -1, 1) (model.py:171); unnormalise (:172); append to trajectory; FK for the next resample;
// optional early exit of run_inference.py:180-187 as a per-problem done mask (done[b] = step at which it stopped).
__global__ void step_update_kernel(int B, const float* __restrict__ dq, float* __restrict__ qn, float* __restrict__ qu,
                                   const float* __restrict__ lim, const float* __restrict__ target, int32_t* __restrict__ done,
                                   int early_exit, float* __restrict__ traj, int traj_stride, float prismatic,
                                   float* __restrict__ frames, float* __restrict__ eef, int step) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float q[7];
  bool frozen = early_exit && done[b] >= 0;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    float v = qn[7 * b + j];
    if (!frozen) {
      v = fminf(1.0f, fmaxf(-1.0f, fadd(v, dq[7 * b + j])));
      qn[7 * b + j] = v;
    }
    q[j] = spec_unnormalize(v, lim[2 * j], lim[2 * j + 1]);
    qu[7 * b + j] = q[j];
    traj[(size_t)b * traj_stride + (size_t)step * 7 + j] = q[j];
  }
  if (frozen) return;
  float F[MPN_NLINK * 12], E[12];
  spec_fk(q, prismatic, F, E);
  float* fo = frames + (size_t)b * MPN_NLINK * 12;
#pragma unroll
  for (int i = 0; i < MPN_NLINK * 12; ++i) fo[i] = F[i];
#pragma unroll
  for (int i = 0; i < 12; ++i) eef[(size_t)b * 12 + i] = E[i];
  if (early_exit) {
    const float* T = target + (size_t)b * 12;
    float dx = E[3] - T[3], dy = E[7] - T[7], dz = E[11] - T[11];
    float pos = sqrtf(dx * dx + dy * dy + dz * dz);
    if (pos < 0.01f && pose_angle_deg(E, T) < 15.0f) done[b] = step;
  }
}

int launch_step_update(mpn_ctx* c, cudaStream_t s, int B, const float* dq, float* qn, float* qu, const float* target,
                       int32_t* done, int early_exit, float* traj_out, int traj_stride, float* frames, float* eef,
                       float* metrics, int step) {
  (void)metrics;
  step_update_kernel<<<(B + 63) / 64, 64, 0, s>>>(B, dq, qn, qu, c->limits, target, done, early_exit, traj_out, traj_stride,
                                                  c->prismatic, frames, eef, step);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// early-exit rollouts: number of problems that have not stopped yet, and the trajectory tail of a rollout that ended early
__global__ void count_live_kernel(int B, const int32_t* __restrict__ done, int32_t* __restrict__ live) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  int alive = (b < B && done[b] < 0) ? 1 : 0;
  alive = __reduce_add_sync(0xffffffffu, alive);
  if ((threadIdx.x & 31) == 0 && alive) atomicAdd(live, alive);
}
__global__ void fill_traj_tail_kernel(int B, const float* __restrict__ qu, float* __restrict__ traj, int traj_stride, int from, int T) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 7) return;
  const int b = i / 7, j = i % 7;
  const float v = qu[i];
  for (int t = from; t <= T; ++t) traj[(size_t)b * traj_stride + (size_t)t * 7 + j] = v;
}
int launch_count_live(mpn_ctx* c, cudaStream_t s, int B, const int32_t* done, int32_t* live) {
  MPN_CHECK_CUDA(cudaMemsetAsync(live, 0, sizeof(int32_t), s));
  count_live_kernel<<<(B + 255) / 256, 256, 0, s>>>(B, done, live);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}
int launch_fill_traj_tail(mpn_ctx* c, cudaStream_t s, int B, const float* qu, float* traj, int traj_stride, int from, int T) {
  if (from > T) return MPN_OK;
  fill_traj_tail_kernel<<<(B * 7 + 255) / 256, 256, 0, s>>>(B, qu, traj, traj_stride, from, T);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

__global__ void finalize_metrics_kernel(int B, const float* __restrict__ eef, const float* __restrict__ target,
                                        const uint8_t* __restrict__ flags, const int32_t* __restrict__ first_step,
                                        const int32_t* __restrict__ done, int T, float* __restrict__ metrics) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* E = eef + (size_t)b * 12;
  const float* Tg = target + (size_t)b * 12;
  float dx = E[3] - Tg[3], dy = E[7] - Tg[7], dz = E[11] - Tg[11];
  float pos = sqrtf(dx * dx + dy * dy + dz * dz);
  float ang = pose_angle_deg(E, Tg);
  float* m = metrics + (size_t)b * MPN_METRICS_COLS;
  m[MPN_M_COLLISION] = flags[b] ? 1.f : 0.f;
  m[MPN_M_FIRST_COLLISION_STEP] = (float)first_step[b];
  m[MPN_M_STEPS] = done[b] >= 0 ? (float)done[b] : (float)T;
  m[MPN_M_POS_ERR] = pos;
  m[MPN_M_ORI_ERR_DEG] = ang;
  m[MPN_M_REACHED] = (pos < 0.01f && ang < 15.0f) ? 1.f : 0.f;
  m[MPN_M_MIN_SDF_MARGIN] = 0.f;
  m[MPN_M_RESERVED] = 0.f;
}

int launch_finalize_metrics(mpn_ctx* c, cudaStream_t s, int B, const float* eef, const float* target, const uint8_t* flags,
                            const int32_t* first_step, const int32_t* done, int T, float* metrics) {
  finalize_metrics_kernel<<<(B + 127) / 128, 128, 0, s>>>(B, eef, target, flags, first_step, done, T, metrics);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
