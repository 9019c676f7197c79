// loss.cu -- the training losses of mpinets/loss.py as forward + analytic-gradient kernels.
//
//   collision_loss   (loss.py:47-94)   mean over B*N of max(0, margin - sdf(point)), sdf = min over the scene's cuboids and
//                                      cylinders (geometry.py:238-288, 456-507); gradient w.r.t. the points
//   point_match_loss (loss.py:31-44)   mse(mean) + l1(mean); gradient w.r.t. the first cloud
//   CollisionAndBCLossContainer.__call__ (loss.py:111-166) + the weighting of model.py:232-236: both losses from the
//                                      normalised joint vectors through FK and the fixed 1024-point robot cloud, and the
//                                      gradient of their weighted sum w.r.t. input_normalized (FK Jacobian:
//                                      d x / d q_j = z_j x (x - o_j) for the joints above the point's link)
//
// Sums are deterministic: fixed-order warp/block trees into per-CTA partials, then one single-CTA pass over the partials.
// The sdf VALUES use the same spec arithmetic as every other kernel (so the hinge on/off decisions match the oracle
// bit for bit); the gradients are plain fp32.
#include "engine.h"
#include "scene.cuh"
#include "spec_math.cuh"

namespace mpn {

namespace {

__device__ __forceinline__ float sgn(float v) { return v < 0.f ? -1.f : (v > 0.f ? 1.f : 0.f); }

// value exactly as sdf_cuboid (spec_math.cuh); gradient with torch autograd's conventions (norm at the zero vector -> 0,
// inside term -> first arg-max)
__device__ __forceinline__ float sdf_cuboid_grad(const PrimFrame& f, float px, float py, float pz, float* g) {
  float l[3], d[3], m[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    l[i] = fadd(dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], px, py, pz), f.Rt[i]);
    d[i] = fsub(fabsf(l[i]), f.h[i]);
    m[i] = fmaxf(d[i], 0.0f);
  }
  const float outside = fsqrt(ffma(m[2], m[2], ffma(m[1], m[1], fmul(m[0], m[0]))));
  const float mx = fmaxf(d[0], fmaxf(d[1], d[2]));
  float gl[3] = {0.f, 0.f, 0.f};
  if (outside > 0.0f) {
#pragma unroll
    for (int i = 0; i < 3; ++i) gl[i] = (m[i] / outside) * (l[i] < 0.0f ? -1.0f : 1.0f);
  }
  if (mx < 0.0f) {
    const int a = d[0] >= d[1] ? (d[0] >= d[2] ? 0 : 2) : (d[1] >= d[2] ? 1 : 2);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i == a) gl[i] += sgn(l[i]);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) g[c] = f.R[6 + c] * gl[2] + f.R[3 + c] * gl[1] + f.R[c] * gl[0];
  return fadd(outside, fminf(mx, 0.0f));
}

__device__ __forceinline__ float sdf_cylinder_grad(const PrimFrame& f, float px, float py, float pz, float* g) {
  float l[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) l[i] = fadd(dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], px, py, pz), f.Rt[i]);
  const float rho = fsqrt(ffma(l[1], l[1], fmul(l[0], l[0])));
  const float d0 = fsub(fabsf(rho), f.h[0]), d1 = fsub(fabsf(l[2]), f.h[1]);
  const float m0 = fmaxf(d0, 0.0f), m1 = fmaxf(d1, 0.0f);
  const float outside = fsqrt(ffma(m1, m1, fmul(m0, m0)));
  const float mx = fmaxf(d0, d1);
  float g_rho = 0.f, g_z = 0.f;
  if (outside > 0.0f) { g_rho = m0 / outside; g_z = (m1 / outside) * (l[2] < 0.0f ? -1.0f : 1.0f); }
  if (mx < 0.0f) {
    if (d0 >= d1) g_rho += 1.0f; else g_z += sgn(l[2]);
  }
  float gl[3] = {0.f, 0.f, g_z};
  if (rho > 0.0f) { gl[0] = g_rho * (l[0] / rho); gl[1] = g_rho * (l[1] / rho); }
#pragma unroll
  for (int c = 0; c < 3; ++c) g[c] = f.R[6 + c] * gl[2] + f.R[3 + c] * gl[1] + f.R[c] * gl[0];
  return fadd(outside, fminf(mx, 0.0f));
}

// min over the compacted scene (cuboids first; the first minimum wins, like torch.min / torch.minimum) + its gradient
__device__ __forceinline__ float scene_sdf_grad(const PrimFrame* fr, int nc, int ny, float px, float py, float pz, float* g) {
  float best = __int_as_float(0x7f800000), gt[3];
  g[0] = g[1] = g[2] = 0.f;
  for (int m = 0; m < nc; ++m) {
    const float v = sdf_cuboid_grad(fr[m], px, py, pz, gt);
    if (v < best) { best = v; g[0] = gt[0]; g[1] = gt[1]; g[2] = gt[2]; }
  }
  for (int m = nc; m < nc + ny; ++m) {
    const float v = sdf_cylinder_grad(fr[m], px, py, pz, gt);
    if (v < best) { best = v; g[0] = gt[0]; g[1] = gt[1]; g[2] = gt[2]; }
  }
  return best;
}

// fixed-order sum of NV values per thread over a 256-thread block; result valid in thread 0..NV-1 (value index = thread)
template <int NV>
__device__ __forceinline__ float block_sum(float (&v)[NV], float (*red)[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_down_sync(0xffffffffu, v[i], o);
    if (lane == 0) red[warp][i] = v[i];
  }
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < NV)
    for (int w = 0; w < nw; ++w) s += red[w][threadIdx.x];
  return s;
}

__global__ void __launch_bounds__(256) collision_loss_kernel(mpn_scene sc, int M1, int M2, int quirk, const float* __restrict__ pts, int N,
                                                             float margin, float scale, float* __restrict__ partial,
                                                             float* __restrict__ grad) {
  __shared__ PrimFrame fr[MAX_PRIMS];
  __shared__ int counts[2], wcnt[8];
  __shared__ float red[8][1];
  const int b = blockIdx.y;
  stage_scene_compact(sc, b, M1, M2, quirk != 0, fr, counts, wcnt);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float h[1] = {0.f};
  if (n < N) {
    const float* p = pts + ((size_t)b * N + n) * 3;
    float g[3];
    const float v = scene_sdf_grad(fr, counts[0], counts[1], p[0], p[1], p[2], g);
    const float hh = fsub(margin, v);
    const bool on = hh > 0.0f;
    h[0] = on ? hh : 0.f;
    if (grad) {
      float* go = grad + ((size_t)b * N + n) * 3;
      go[0] = on ? -g[0] * scale : 0.f; go[1] = on ? -g[1] * scale : 0.f; go[2] = on ? -g[2] * scale : 0.f;
    }
  }
  const float s = block_sum<1>(h, red);
  if (threadIdx.x == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = s;
}

// out = scaled column sums of partial [rows][cols] (cols <= 3).  mode 0: out[c] = scale0 * sum_c;
// mode 1 (the two losses of the container): out[0] = scale0 * sum_0, out[1] = scale1 * (sum_1 + sum_2);
// mode 2 (mse + l1): out[0] = scale0 * (sum_0 + sum_1)
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partial, int rows, int cols, int mode, float scale0,
                                                               float scale1, float* __restrict__ out) {
  __shared__ float red[32][3];
  float v[3] = {0.f, 0.f, 0.f};
  for (int r = threadIdx.x; r < rows; r += blockDim.x)
    for (int c = 0; c < cols; ++c) v[c] += partial[(size_t)r * cols + c];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_down_sync(0xffffffffu, v[c], o);
    if (lane == 0) red[warp][c] = v[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s[3] = {0.f, 0.f, 0.f};
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w)
      for (int c = 0; c < 3; ++c) s[c] += red[w][c];
    if (mode == 0) { for (int c = 0; c < cols; ++c) out[c] = s[c] * scale0; }
    else if (mode == 1) { out[0] = s[0] * scale0; out[1] = (s[1] + s[2]) * scale1; }
    else out[0] = (s[0] + s[1]) * scale0;
  }
}

__global__ void __launch_bounds__(256) point_match_kernel(const float* __restrict__ a, const float* __restrict__ bb, size_t n, size_t chunk,
                                                          float inv, float* __restrict__ partial, float* __restrict__ grad) {
  __shared__ float red[8][2];
  const size_t lo = (size_t)blockIdx.x * chunk, hi = min(n, lo + chunk);
  float v[2] = {0.f, 0.f};
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float d = a[i] - bb[i];
    v[0] += d * d; v[1] += fabsf(d);
    if (grad) grad[i] = (2.0f * d + sgn(d)) * inv;
  }
  const float s = block_sum<2>(v, red);
  if (threadIdx.x < 2) partial[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
}

// one CTA per problem: FK of the input and the target configuration, the fixed robot subset through both, hinge / mse / l1
// sums and the joint-space gradient of the weighted loss
__global__ void __launch_bounds__(256) bc_losses_kernel(mpn_scene sc, int M1, int M2, int quirk, const float* __restrict__ input_norm,
                                                        const float* __restrict__ target_norm, const float* __restrict__ lim,
                                                        float prismatic, int P, int n0, const float4* __restrict__ table, int n,
                                                        uint32_t seed_lo, uint32_t seed_hi, float margin, float wc_scaled,
                                                        float wp_scaled, float* __restrict__ partial, float* __restrict__ grad_input) {
  __shared__ PrimFrame fr[MAX_PRIMS];
  __shared__ int counts[2], wcnt[8];
  __shared__ float Fi[MPN_NLINK * 12], Ft[MPN_NLINK * 12];
  __shared__ float red[8][10];
  const int b = blockIdx.x;
  stage_scene_compact(sc, b, M1, M2, quirk != 0, fr, counts, wcnt);
  if (threadIdx.x == 0 || threadIdx.x == 32) {
    const float* src = (threadIdx.x == 0 ? input_norm : target_norm) + (size_t)b * 7;
    float q[7], F[MPN_NLINK * 12];
#pragma unroll
    for (int j = 0; j < 7; ++j) q[j] = spec_unnormalize(src[j], lim[2 * j], lim[2 * j + 1]);
    spec_fk(q, prismatic, F, nullptr);
    float* dst = threadIdx.x == 0 ? Fi : Ft;
#pragma unroll
    for (int i = 0; i < MPN_NLINK * 12; ++i) dst[i] = F[i];
  }
  uint32_t key[4];
  philox4x32(0u, 0u, STREAM_FIXED_PERM, 0u, seed_lo, seed_hi, key);
  const uint32_t np = (uint32_t)(P - n0), half = feistel_bits(np) / 2;
  __syncthreads();
  const int nc = counts[0], ny = counts[1];
  float v[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // hinge, sum d^2, sum |d|, 7 joint gradients
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 t = __ldg(table + n0 + feistel_perm((uint32_t)i, np, half, key));
    const int link = __float_as_int(t.w);
    float x[3], y[3], gs[3], g[3];
    m34_apply(Fi + 12 * link, t.x, t.y, t.z, x[0], x[1], x[2]);
    m34_apply(Ft + 12 * link, t.x, t.y, t.z, y[0], y[1], y[2]);
    const float sd = scene_sdf_grad(fr, nc, ny, x[0], x[1], x[2], gs);
    const float hh = fsub(margin, sd);
    const bool on = hh > 0.0f;
    if (on) v[0] += hh;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = x[c] - y[c];
      v[1] += d * d; v[2] += fabsf(d);
      g[c] = wp_scaled * (2.0f * d + sgn(d)) + (on ? -wc_scaled * gs[c] : 0.f);
    }
    const int jm = link < 7 ? link : 7;
#pragma unroll
    for (int j = 1; j <= 7; ++j) {
      if (j <= jm) {
        const float* F = Fi + 12 * j;
        const float rx = x[0] - F[3], ry = x[1] - F[7], rz = x[2] - F[11];
        const float cx = F[6] * rz - F[10] * ry, cy = F[10] * rx - F[2] * rz, cz = F[2] * ry - F[6] * rx;
        v[2 + j] += g[0] * cx + g[1] * cy + g[2] * cz;
      }
    }
  }
  const float s = block_sum<10>(v, red);
  if (threadIdx.x < 3) partial[(size_t)b * 3 + threadIdx.x] = s;
  else if (threadIdx.x < 10 && grad_input) {
    const int j = threadIdx.x - 3;   // q = (qn + 1) / 2 * (hi - lo) + lo
    grad_input[(size_t)b * 7 + j] = s * 0.5f * (lim[2 * j + 1] - lim[2 * j]);
  }
}

int ensure_partials(mpn_ctx* c, size_t floats) {
  if (c->loss_partial_cap >= floats) return MPN_OK;
  if (c->loss_partial) cudaFree(c->loss_partial);
  c->loss_partial = nullptr; c->loss_partial_cap = 0;
  MPN_CHECK_CUDA(cudaMalloc(&c->loss_partial, floats * sizeof(float)));
  c->loss_partial_cap = floats;
  return MPN_OK;
}

}  // namespace

int launch_collision_loss(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, int N, const float* points, float margin, float* loss,
                          float* grad_points) {
  const int nblk = (N + 255) / 256;
  int r;
  if ((r = ensure_partials(c, (size_t)B * nblk))) return r;
  const float scale = 1.0f / ((float)B * (float)N);
  collision_loss_kernel<<<dim3(nblk, B), 256, 0, s>>>(sc, c->cfg.max_cuboids, c->cfg.max_cylinders, c->cfg.quirk_frames, points, N, margin,
                                                     scale, c->loss_partial, grad_points);
  reduce_partials_kernel<<<1, 1024, 0, s>>>(c->loss_partial, B * nblk, 1, 0, scale, 0.f, loss);
  c->launches += 2;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

int launch_point_match_loss(mpn_ctx* c, cudaStream_t s, size_t n, const float* a, const float* b, float* loss, float* grad_a) {
  const int nblk = (int)std::min<size_t>(1024, (n + 4095) / 4096);
  const size_t chunk = ((n + nblk - 1) / nblk + 255) / 256 * 256;
  int r;
  if ((r = ensure_partials(c, (size_t)nblk * 2))) return r;
  const float inv = 1.0f / (float)n;
  point_match_kernel<<<nblk, 256, 0, s>>>(a, b, n, chunk, inv, c->loss_partial, grad_a);
  reduce_partials_kernel<<<1, 1024, 0, s>>>(c->loss_partial, nblk, 2, 2, inv, 0.f, loss);
  c->launches += 2;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

int launch_bc_collision_losses(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* input_norm, const float* target_norm,
                               int n_points, float margin, float w_collision, float w_bc, float* losses, float* grad_input) {
  MPN_REQUIRE(n_points >= 1 && n_points <= c->P - c->n_base_points, "bc_collision_losses: %d fixed points requested, %d non-base link points",
              n_points, c->P - c->n_base_points);
  int r;
  if ((r = ensure_partials(c, (size_t)B * 3))) return r;
  const float sc_c = 1.0f / ((float)B * (float)n_points), sc_p = 1.0f / ((float)B * (float)n_points * 3.0f);
  bc_losses_kernel<<<B, 256, 0, s>>>(sc, c->cfg.max_cuboids, c->cfg.max_cylinders, c->cfg.quirk_frames, input_norm, target_norm, c->limits,
                                     c->prismatic, c->P, c->n_base_points, reinterpret_cast<const float4*>(c->link_table4), n_points,
                                     (uint32_t)c->cfg.seed, (uint32_t)(c->cfg.seed >> 32), margin, w_collision * sc_c, w_bc * sc_p,
                                     c->loss_partial, grad_input);
  reduce_partials_kernel<<<1, 1024, 0, s>>>(c->loss_partial, B, 3, 1, sc_c, sc_p, losses);
  c->launches += 2;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
