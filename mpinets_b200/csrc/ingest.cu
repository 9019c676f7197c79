// ingest.cu -- device-side pieces of the input pipelines either side of the rollout:
//   * augment_joints_kernel   PointCloudBase.get_inputs' training-time joint noise (mpinets/data_loader.py:167-180):
//                             randomized = random_scale * N(0, 1) + q, clamped to FrankaRealRobot.JOINT_LIMITS, normalised.  The
//                             reference draws torch.randn from the global generator; here the normals are Box-Muller transforms of
//                             Philox4x32-10(counter = (sample id, epoch, STREAM_JOINT_NOISE, pair), key = seed), so a sample's noise
//                             does not depend on batch composition, worker or rank.
//   * clean_point_cloud       planning_node.py:187-228: keep the points of a sensed cloud that lie in the task-tabletop box or the
//                             mount-table box (strict float32 comparisons), then a random subset WITHOUT replacement of n_out of them
//                             (np.random.choice(len, NUM_OBSTACLE_POINTS, replace=False) there; a keyed Feistel permutation of the
//                             kept list here).  In-order compaction, so the kept list equals numpy's boolean-mask order.
#include "engine.h"
#include "spec_math.cuh"

namespace mpn {

enum { STREAM_JOINT_NOISE = 6, STREAM_CLEAN_PERM = 7 };

__global__ void augment_joints_kernel(const float* __restrict__ q, int B, float scale, const uint32_t* __restrict__ ids, uint32_t epoch,
                                      const float* __restrict__ lim, uint32_t seed_lo, uint32_t seed_hi, float* __restrict__ q_out,
                                      float* __restrict__ qn_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint32_t id = ids ? ids[b] : (uint32_t)b;
  float z[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t r[4];
    philox4x32(id, epoch, STREAM_JOINT_NOISE, (uint32_t)h, seed_lo, seed_hi, r);
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      const float u1 = fmul((float)((r[2 * p] >> 8) + 1u), 5.9604644775390625e-08f);   // (0, 1], exact
      const float u2 = fmul((float)(r[2 * p + 1] >> 8), 5.9604644775390625e-08f);     // [0, 1), exact
      const float rad = sqrtf(fmul(-2.0f, logf(u1)));
      const float ang = fmul(6.2831853071795864769f, u2);
      z[4 * h + 2 * p] = fmul(rad, cosf(ang));
      z[4 * h + 2 * p + 1] = fmul(rad, sinf(ang));
    }
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const float lo = lim[2 * j], hi = lim[2 * j + 1];
    float v = fadd(fmul(scale, z[j]), q[7 * b + j]);
    v = fminf(fmaxf(v, lo), hi);                       // torch.minimum(torch.maximum(randomized, limits[:, 0]), limits[:, 1])
    q_out[7 * b + j] = v;
    qn_out[7 * b + j] = spec_normalize(v, lo, hi);
  }
}

int launch_augment_joints(mpn_ctx* c, cudaStream_t s, const float* q, int B, float scale, const uint32_t* sample_ids, uint32_t epoch,
                          float* q_out, float* qn_out) {
  augment_joints_kernel<<<(B + 127) / 128, 128, 0, s>>>(q, B, scale, sample_ids, epoch, c->limits, (uint32_t)c->cfg.seed,
                                                        (uint32_t)(c->cfg.seed >> 32), q_out, qn_out);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// planning_node.py:202-223
__device__ __forceinline__ bool in_workspace(float x, float y, float z) {
  const bool task = x > 0.25f && x < 1.35f && y > -0.3f && y < 1.6f && z > -0.05f && z < 0.35f;
  const bool mount = x > -0.35f && x < 0.30f && y > -0.5f && y < 0.5f && z > -0.05f && z < 0.05f;
  return task || mount;
}

constexpr int CLEAN_THREADS = 1024;
// one CTA walks the cloud in index order: list[0 .. kept) = indices of the points inside the workspace, ascending
__global__ void __launch_bounds__(CLEAN_THREADS) clean_compact_kernel(const float* __restrict__ xyz, int N, int32_t* __restrict__ list,
                                                                      int32_t* __restrict__ kept) {
  __shared__ int wcnt[CLEAN_THREADS / 32];
  __shared__ int running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < N; base += CLEAN_THREADS) {
    const int k = base + threadIdx.x;
    bool keep = false;
    if (k < N) keep = in_workspace(xyz[3 * (size_t)k], xyz[3 * (size_t)k + 1], xyz[3 * (size_t)k + 2]);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wcnt[warp] = __popc(m);
    __syncthreads();
    int off = running;
    for (int w = 0; w < warp; ++w) off += wcnt[w];
    if (keep) list[off + __popc(m & ((1u << lane) - 1u))] = k;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < CLEAN_THREADS / 32; ++w) t += wcnt[w];
      running += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) kept[0] = running;
}

__global__ void clean_select_kernel(const float* __restrict__ xyz, const float* __restrict__ rgba, const int32_t* __restrict__ list,
                                    const int32_t* __restrict__ kept, int n_out, uint32_t cloud_id, uint32_t seed_lo, uint32_t seed_hi,
                                    float* __restrict__ out_xyz, float* __restrict__ out_rgba) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int M = kept[0];
  if (j >= n_out || M < n_out) return;   // fewer kept points than requested: np.random.choice raises; the host reports it from kept[0]
  uint32_t key[4];
  philox4x32(0u, cloud_id, STREAM_CLEAN_PERM, 0u, seed_lo, seed_hi, key);
  const uint32_t e = feistel_perm((uint32_t)j, (uint32_t)M, feistel_bits((uint32_t)M) / 2, key);
  const size_t src = (size_t)list[e];
  out_xyz[3 * j] = xyz[3 * src]; out_xyz[3 * j + 1] = xyz[3 * src + 1]; out_xyz[3 * j + 2] = xyz[3 * src + 2];
  if (rgba && out_rgba) {
#pragma unroll
    for (int i = 0; i < 4; ++i) out_rgba[4 * j + i] = rgba[4 * src + i];
  }
}

int launch_clean_point_cloud(mpn_ctx* c, cudaStream_t s, const float* xyz, const float* rgba, int N, int n_out, uint32_t cloud_id,
                             float* out_xyz, float* out_rgba, int32_t* kept, int32_t* scratch) {
  clean_compact_kernel<<<1, CLEAN_THREADS, 0, s>>>(xyz, N, scratch, kept);
  clean_select_kernel<<<(n_out + 255) / 256, 256, 0, s>>>(xyz, rgba, scratch, kept, n_out, cloud_id, (uint32_t)c->cfg.seed,
                                                          (uint32_t)(c->cfg.seed >> 32), out_xyz, out_rgba);
  c->launches += 2;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
