// geometry.cu -- HBM-bound geometry kernels: FK, robot / obstacle / target surface sampling into the segmented
// cloud, point SDF and the link-sphere collision sweep.  Bit-exact against the CPU oracle (spec_math.cuh).
//
// Replaces: robofin FrankaSampler.sample / sample_end_effector / end_effector_pose and
// FrankaCollisionSampler.compute_spheres (call sites mpinets/model.py:250,275,300; run_inference.py:111-116,188),
// mpinets/geometry.py:238-288,456-507 (sdf), :571-608 (construct_mixed_point_cloud), model.py:293-314 (sweep).
#include "engine.h"
#include "spec_math.cuh"
#include "scene.cuh"

namespace mpn {

// ------------------------------------------------------------------------------------------------ FK
__global__ void fk_kernel(const float* __restrict__ q, int B, float prismatic, float* __restrict__ frames,
                          float* __restrict__ eef) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float qq[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) qq[j] = q[7 * b + j];
  float F[MPN_NLINK * 12], E[12];
  spec_fk(qq, prismatic, F, E);
  if (frames) {
    float* o = frames + (size_t)b * MPN_NLINK * 12;
#pragma unroll
    for (int i = 0; i < MPN_NLINK * 12; ++i) o[i] = F[i];
  }
  if (eef) {
#pragma unroll
    for (int i = 0; i < 12; ++i) eef[(size_t)b * 12 + i] = E[i];
  }
}

int launch_fk(mpn_ctx* c, cudaStream_t s, const float* q, int B, float* frames, float* eef) {
  fk_kernel<<<(B + 63) / 64, 64, 0, s>>>(q, B, c->prismatic, frames, eef);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ------------------------------------------------------------------------------------------------ robot rows
// one CTA per problem; frames staged in smem; the canonical table is packed as float4 (x, y, z, link id) so a row costs
// one 16-byte gather (the 64 KB table lives in L1/L2) and one 16-byte row store (a warp writes 512 contiguous bytes);
// 8 rows per thread are processed with all their loads in flight.
__global__ void pack_link_table_kernel(const float* __restrict__ lp, const int32_t* __restrict__ lid, int P, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) out[i] = make_float4(lp[3 * i], lp[3 * i + 1], lp[3 * i + 2], __int_as_float(lid[i]));
}

// The random subset of FrankaSampler.sample is one np.random.choice shared by the whole batch, so the
// permutation is evaluated once per step into `sel` (n float4 rows, 32 KB: L1/L2 resident) instead of once per problem;
// the per-problem kernel is then a pure streaming pass: coalesced 16-byte reads of `sel`, the link frame from shared
// memory, coalesced 16-byte row writes.
__global__ void robot_subset_kernel(int n, int P, const float4* __restrict__ table, uint32_t seed_lo, uint32_t seed_hi, uint32_t step0,
                                    float4* __restrict__ sel) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t step = step0 + blockIdx.y;     // blockIdx.y: consecutive steps of a rollout, one [n] slab each
  uint32_t key[4];
  philox4x32(0u, step, STREAM_ROBOT_PERM, 0u, seed_lo, seed_hi, key);
  sel[(size_t)blockIdx.y * n + j] = __ldg(table + feistel_perm((uint32_t)j, (uint32_t)P, feistel_bits((uint32_t)P) / 2, key));
}

__global__ void __launch_bounds__(256) sample_robot_kernel(const float* __restrict__ frames, int n, const float4* __restrict__ sel,
                                                           float4* __restrict__ cloud, int rows) {
  __shared__ float F[MPN_NLINK * 12];
  int b = blockIdx.x;
  for (int i = threadIdx.x; i < MPN_NLINK * 12; i += blockDim.x) F[i] = frames[(size_t)b * MPN_NLINK * 12 + i];
  __syncthreads();
  float4* out = cloud + (size_t)b * rows;
  constexpr int U = 8;
  for (int j0 = threadIdx.x; j0 < n; j0 += blockDim.x * U) {
    float4 t[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int j = j0 + u * blockDim.x;
      if (j < n) t[u] = __ldg(sel + j);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int j = j0 + u * blockDim.x;
      if (j < n) {
        float4 o;
        m34_apply(F + 12 * __float_as_int(t[u].w), t[u].x, t[u].y, t[u].z, o.x, o.y, o.z);
        o.w = 0.0f;
        out[j] = o;
      }
    }
  }
}

int launch_sample_robot(mpn_ctx* c, cudaStream_t s, const float* frames, int B, int n, uint32_t step, float* cloud, int rows) {
  MPN_REQUIRE(n <= c->P, "sample_robot: %d points requested, the link table has %d", n, c->P);
  robot_subset_kernel<<<dim3((n + 255) / 256, 1), 256, 0, s>>>(n, c->P, reinterpret_cast<const float4*>(c->link_table4), (uint32_t)c->cfg.seed,
                                                              (uint32_t)(c->cfg.seed >> 32), step, reinterpret_cast<float4*>(c->robot_sel4));
  sample_robot_kernel<<<B, 256, 0, s>>>(frames, n, reinterpret_cast<const float4*>(c->robot_sel4), (float4*)cloud, rows);
  c->launches += 2;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// rollout form: the subsets of steps step0 .. step0 + count - 1 in ONE launch (they depend on the step index only), then one
// streaming kernel per step reading its slab
int launch_robot_subsets(mpn_ctx* c, cudaStream_t s, int n, uint32_t step0, int count) {
  MPN_REQUIRE(n <= c->P, "sample_robot: %d points requested, the link table has %d", n, c->P);
  const size_t need = (size_t)count * n;
  if (c->robot_sel_steps_cap < need) {
    if (c->robot_sel_steps) cudaFree(c->robot_sel_steps);
    c->robot_sel_steps = nullptr; c->robot_sel_steps_cap = 0;
    MPN_CHECK_CUDA(cudaMalloc(&c->robot_sel_steps, need * sizeof(float4)));
    c->robot_sel_steps_cap = need;
  }
  robot_subset_kernel<<<dim3((n + 255) / 256, count), 256, 0, s>>>(n, c->P, reinterpret_cast<const float4*>(c->link_table4),
                                                                  (uint32_t)c->cfg.seed, (uint32_t)(c->cfg.seed >> 32), step0,
                                                                  reinterpret_cast<float4*>(c->robot_sel_steps));
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

int launch_sample_robot_slab(mpn_ctx* c, cudaStream_t s, const float* frames, int B, int n, int slab, float* cloud, int rows) {
  sample_robot_kernel<<<B, 256, 0, s>>>(frames, n, reinterpret_cast<const float4*>(c->robot_sel_steps) + (size_t)slab * n, (float4*)cloud, rows);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

int pack_link_table(mpn_ctx* c) {
  if (c->link_table4) cudaFree(c->link_table4);
  MPN_CHECK_CUDA(cudaMalloc(&c->link_table4, (size_t)c->P * sizeof(float4)));
  if (c->robot_sel4) cudaFree(c->robot_sel4);
  MPN_CHECK_CUDA(cudaMalloc(&c->robot_sel4, (size_t)c->P * sizeof(float4)));
  pack_link_table_kernel<<<(c->P + 255) / 256, 256>>>(c->link_points, c->link_ids, c->P, reinterpret_cast<float4*>(c->link_table4));
  MPN_CHECK_CUDA(cudaGetLastError());
  MPN_CHECK_CUDA(cudaDeviceSynchronize());
  return MPN_OK;
}

// ------------------------------------------------------------------------------------------------ spheres
__global__ void spheres_kernel(const float* __restrict__ frames, int B, int S, const float* __restrict__ sc,
                               const int32_t* __restrict__ sl, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * S) return;
  int b = i / S, k = i % S;
  const float* F = frames + ((size_t)b * MPN_NLINK + sl[k]) * 12;
  float A[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) A[j] = F[j];
  float x, y, z;
  m34_apply(A, sc[3 * k], sc[3 * k + 1], sc[3 * k + 2], x, y, z);
  out[3 * (size_t)i] = x; out[3 * (size_t)i + 1] = y; out[3 * (size_t)i + 2] = z;
}

int launch_spheres(mpn_ctx* c, cudaStream_t s, const float* frames, int B, float* centers) {
  int n = B * c->S;
  spheres_kernel<<<(n + 255) / 256, 256, 0, s>>>(frames, B, c->S, c->sph_c, c->sph_l, centers);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

__global__ void normalize_kernel(const float* __restrict__ in, int n, const float* __restrict__ lim, float* __restrict__ out,
                                 int un) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 7) return;
  int j = i % 7;
  float lo = lim[2 * j], hi = lim[2 * j + 1];
  out[i] = un ? spec_unnormalize(in[i], lo, hi) : spec_normalize(in[i], lo, hi);
}

int launch_normalize(mpn_ctx* c, cudaStream_t s, const float* in, int n, float* out, bool un) {
  normalize_kernel<<<(n * 7 + 255) / 256, 256, 0, s>>>(in, n, c->limits, out, un ? 1 : 0);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

__global__ void __launch_bounds__(256) sdf_points_kernel(mpn_scene sc, int M1, int M2, int quirk, const float* __restrict__ pts,
                                                         int N, int which, float* __restrict__ sdf) {
  __shared__ PrimFrame fr[MAX_PRIMS];
  int b = blockIdx.y;
  stage_scene(sc, b, M1, M2, quirk != 0, fr);
  __syncthreads();
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* p = pts + ((size_t)b * N + n) * 3;
  sdf[(size_t)b * N + n] = scene_sdf(fr, 0, which == 2 ? 0 : M1, M1, which == 1 ? M1 : M1 + M2, p[0], p[1], p[2]);
}

int launch_sdf_points(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* pts, int N, int which, float* sdf) {
  dim3 grid((N + 255) / 256, B);
  sdf_points_kernel<<<grid, 256, 0, s>>>(sc, c->cfg.max_cuboids, c->cfg.max_cylinders, c->cfg.quirk_frames, pts, N, which, sdf);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ------------------------------------------------------------------------------------------------ cloud build (t = 0)
__global__ void __launch_bounds__(256) build_cloud_kernel(mpn_scene sc, int M1, int M2, const float* __restrict__ frames,
                                                          const float* __restrict__ target, int Nr, int No, int Nt, int P,
                                                          const float* __restrict__ lp, const int32_t* __restrict__ lid,
                                                          int Pe, const float* __restrict__ ee, uint32_t seed_lo,
                                                          uint32_t seed_hi, uint32_t problem0, float4* __restrict__ cloud,
                                                          const float* __restrict__ obs_points, const int32_t* __restrict__ obs_count,
                                                          int obs_max, const uint32_t* __restrict__ problem_ids) {
  __shared__ float F[MPN_NLINK * 12];
  __shared__ float TG[12];
  __shared__ uint32_t start[MAX_PRIMS + 1];
  __shared__ int16_t pidx[MAX_PRIMS];  // >=0 cuboid index, <0 : -(cyl index)-1
  __shared__ int nvalid;
  int b = blockIdx.x;
  uint32_t problem = problem_ids ? problem_ids[b] : problem0 + (uint32_t)b;   // the RNG counter of this problem's sampling streams
  for (int i = threadIdx.x; i < MPN_NLINK * 12; i += blockDim.x) F[i] = frames[(size_t)b * MPN_NLINK * 12 + i];
  if (threadIdx.x < 12) TG[threadIdx.x] = target[(size_t)b * 12 + threadIdx.x];
  if (threadIdx.x == 0 && !obs_points) {
    // geometry.py:590-599: proportions in double, cuboids then cylinders, zero-volume skipped
    double area[MAX_PRIMS];
    int Pn = 0;
    for (int m = 0; m < M1; ++m) {
      const float* d = sc.cuboid_dims + ((size_t)b * M1 + m) * 3;
      if (is_close0(d[0]) || is_close0(d[1]) || is_close0(d[2])) continue;
      double x = d[0], y = d[1], z = d[2];
      area[Pn] = __dmul_rn(2.0, __dadd_rn(__dadd_rn(__dmul_rn(x, y), __dmul_rn(x, z)), __dmul_rn(y, z)));
      pidx[Pn] = (int16_t)m; ++Pn;
    }
    for (int m = 0; m < M2; ++m) {
      float rf = sc.cylinder_radii[(size_t)b * M2 + m], hf = sc.cylinder_heights[(size_t)b * M2 + m];
      if (is_close0(rf) || is_close0(hf)) continue;
      double r = rf, h = hf;
      double tpr = __dmul_rn(__dmul_rn(2.0, 3.14159265358979323846), r);
      area[Pn] = __dadd_rn(__dmul_rn(tpr, h), __dmul_rn(tpr, r));
      pidx[Pn] = (int16_t)(-m - 1); ++Pn;
    }
    double total = 0.0;
    for (int i = 0; i < Pn; ++i) total = __dadd_rn(total, area[i]);
    start[0] = 0;
    for (int i = 0; i < Pn; ++i) {
      double prop = __ddiv_rn(area[i], total);
      uint32_t ni = (uint32_t)(int)__dmul_rn(prop, (double)No) + 500u;
      start[i + 1] = start[i] + ni;
    }
    nvalid = Pn;
  }
  __syncthreads();
  float4* out = cloud + (size_t)b * (Nr + No + Nt);
  // robot rows (step 0 subset)
  {
    uint32_t key[4];
    philox4x32(0u, 0u, STREAM_ROBOT_PERM, 0u, seed_lo, seed_hi, key);
    uint32_t half = feistel_bits((uint32_t)P) / 2;
    for (int j = threadIdx.x; j < Nr; j += blockDim.x) {
      uint32_t e = feistel_perm((uint32_t)j, (uint32_t)P, half, key);
      float4 o;
      m34_apply(F + 12 * __ldg(lid + e), __ldg(lp + 3 * e), __ldg(lp + 3 * e + 1), __ldg(lp + 3 * e + 2), o.x, o.y, o.z);
      o.w = 0.0f;
      out[j] = o;
    }
  }
  // obstacle rows
  if (obs_points) {
    // run_inference.make_point_cloud_from_problem (run_inference.py:58-90): a subset without replacement of a given cloud
    const int cnt = min(obs_count[b], obs_max);
    uint32_t key[4];
    philox4x32(0u, problem, STREAM_OBS_PERM, 0u, seed_lo, seed_hi, key);
    const uint32_t half = feistel_bits((uint32_t)max(cnt, 1)) / 2;
    const float* src = obs_points + (size_t)b * obs_max * 3;
    for (int j = threadIdx.x; j < No; j += blockDim.x) {
      float4 o = make_float4(0.f, 0.f, 0.f, 1.0f);
      if (cnt > 0) {
        const uint32_t e = feistel_perm((uint32_t)(j % cnt), (uint32_t)cnt, half, key);
        o.x = src[3 * e]; o.y = src[3 * e + 1]; o.z = src[3 * e + 2];
      }
      out[Nr + j] = o;
    }
  } else {
    int Pn = nvalid;
    if (Pn == 0) {
      for (int j = threadIdx.x; j < No; j += blockDim.x) out[Nr + j] = make_float4(0.f, 0.f, 0.f, 1.0f);
    } else {
      uint32_t pool = start[Pn];
      uint32_t key[4];
      philox4x32(0u, problem, STREAM_OBS_PERM, 0u, seed_lo, seed_hi, key);
      uint32_t half = feistel_bits(pool) / 2;
      for (int j = threadIdx.x; j < No; j += blockDim.x) {
        uint32_t e = feistel_perm((uint32_t)j, pool, half, key);
        int i = 0;
        while (e >= start[i + 1]) ++i;
        uint32_t r[4];
        philox4x32(e, problem, STREAM_OBS_SAMPLE, 0u, seed_lo, seed_hi, r);
        float o[3];
        int m = pidx[i];
        if (m >= 0) {
          sample_cuboid(sc.cuboid_centers + ((size_t)b * M1 + m) * 3, sc.cuboid_dims + ((size_t)b * M1 + m) * 3,
                        sc.cuboid_quats + ((size_t)b * M1 + m) * 4, r, o);
        } else {
          m = -m - 1;
          sample_cylinder(sc.cylinder_centers + ((size_t)b * M2 + m) * 3, sc.cylinder_radii[(size_t)b * M2 + m],
                          sc.cylinder_heights[(size_t)b * M2 + m], sc.cylinder_quats + ((size_t)b * M2 + m) * 4, r, o);
        }
        out[Nr + j] = make_float4(o[0], o[1], o[2], 1.0f);
      }
    }
  }
  // target rows
  {
    uint32_t key[4];
    philox4x32(0u, problem, STREAM_TARGET_PERM, 0u, seed_lo, seed_hi, key);
    uint32_t half = feistel_bits((uint32_t)Pe) / 2;
    for (int j = threadIdx.x; j < Nt; j += blockDim.x) {
      uint32_t e = feistel_perm((uint32_t)j, (uint32_t)Pe, half, key);
      float4 o;
      m34_apply(TG, __ldg(ee + 3 * e), __ldg(ee + 3 * e + 1), __ldg(ee + 3 * e + 2), o.x, o.y, o.z);
      o.w = 2.0f;
      out[Nr + No + j] = o;
    }
  }
}

int launch_build_cloud(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* frames, const float* target,
                       uint32_t problem0, float* cloud, const float* obs_points, const int32_t* obs_count, int obs_max,
                       const uint32_t* problem_ids) {
  build_cloud_kernel<<<B, 256, 0, s>>>(sc, c->cfg.max_cuboids, c->cfg.max_cylinders, frames, target, c->cfg.n_robot,
                                       c->cfg.n_obstacle, c->cfg.n_target, c->P, c->link_points, c->link_ids, c->Pe,
                                       c->ee_points, (uint32_t)c->cfg.seed, (uint32_t)(c->cfg.seed >> 32), problem0,
                                       (float4*)cloud, obs_points, obs_count, obs_max, problem_ids);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// FrankaSampler.sample_end_effector (run_inference.py:113-116; data_loader.py:158-161; planning_node.py:71-74): n gripper points
// (hand + fingers, given in the right_gripper frame) transformed by each pose.  The subset is the same keyed Feistel permutation as
// the target rows of build_cloud_kernel, so out[b] equals rows [Nr + No, Nr + No + n) of the cloud built for problem problem0 + b.
__global__ void __launch_bounds__(128) sample_end_effector_kernel(const float* __restrict__ poses, int n, int Pe, const float* __restrict__ ee,
                                                                   uint32_t seed_lo, uint32_t seed_hi, uint32_t problem0,
                                                                   float* __restrict__ out) {
  __shared__ float TG[12];
  __shared__ uint32_t key[4];
  const int b = blockIdx.x;
  if (threadIdx.x < 12) TG[threadIdx.x] = poses[(size_t)b * 12 + threadIdx.x];
  if (threadIdx.x == 0) philox4x32(0u, problem0 + (uint32_t)b, STREAM_TARGET_PERM, 0u, seed_lo, seed_hi, key);
  __syncthreads();
  const uint32_t half = feistel_bits((uint32_t)Pe) / 2;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const uint32_t e = feistel_perm((uint32_t)j, (uint32_t)Pe, half, key);
    float x, y, z;
    m34_apply(TG, __ldg(ee + 3 * e), __ldg(ee + 3 * e + 1), __ldg(ee + 3 * e + 2), x, y, z);
    float* o = out + ((size_t)b * n + j) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
}

int launch_sample_end_effector(mpn_ctx* c, cudaStream_t s, const float* poses, int B, int n, uint32_t problem0, float* out) {
  sample_end_effector_kernel<<<B, 128, 0, s>>>(poses, n, c->Pe, c->ee_points, (uint32_t)c->cfg.seed, (uint32_t)(c->cfg.seed >> 32), problem0, out);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ------------------------------------------------------------------------------------------------ collision sweep
// One CTA per problem.  Primitive inverse frames + sphere table in shared memory; timesteps processed in chunks
// of SWEEP_TCHUNK: one thread per timestep does FK into smem, then all threads sweep (timestep, sphere) pairs.
constexpr int SWEEP_TCHUNK = 16;
constexpr int SWEEP_THREADS = 128;
constexpr int MAX_SPHERES = 96;

__global__ void __launch_bounds__(SWEEP_THREADS) sweep_kernel(mpn_scene sc, int M1, int M2, int quirk,
                                                              const float* __restrict__ traj, int T, int problem_stride,
                                                              int t0, float prismatic, int S, const float* __restrict__ sph_c,
                                                              const float* __restrict__ sph_r, const int32_t* __restrict__ sph_l,
                                                              int accumulate, uint8_t* __restrict__ flags,
                                                              int32_t* __restrict__ first_step, const float* __restrict__ frames_in) {
  __shared__ PrimFrame fr[MAX_PRIMS];
  __shared__ float F[SWEEP_TCHUNK][MPN_NLINK * 12];
  __shared__ float sc_s[MAX_SPHERES * 3];
  __shared__ float sr_s[MAX_SPHERES];
  __shared__ int sl_s[MAX_SPHERES];
  __shared__ int first;
  __shared__ int counts[2], wcnt[SWEEP_THREADS / 32];
  int b = blockIdx.x;
  stage_scene_compact(sc, b, M1, M2, quirk != 0, fr, counts, wcnt);
  const int nc = counts[0], ny = counts[1];
  for (int k = threadIdx.x; k < S; k += blockDim.x) {
    sc_s[3 * k] = sph_c[3 * k]; sc_s[3 * k + 1] = sph_c[3 * k + 1]; sc_s[3 * k + 2] = sph_c[3 * k + 2];
    sr_s[k] = sph_r[k]; sl_s[k] = sph_l[k];
  }
  if (threadIdx.x == 0) first = 0x7fffffff;
  __syncthreads();
  for (int tc = 0; tc < T; tc += SWEEP_TCHUNK) {
    int nt = min(SWEEP_TCHUNK, T - tc);
    if (frames_in) {   // T == 1: link frames of this configuration were already produced by the step update (same spec FK)
      for (int i = threadIdx.x; i < MPN_NLINK * 12; i += blockDim.x) F[0][i] = frames_in[(size_t)b * MPN_NLINK * 12 + i];
    } else if (threadIdx.x < nt) {
      float qq[7];
      const float* qp = traj + (size_t)b * problem_stride + (size_t)(tc + threadIdx.x) * 7;
#pragma unroll
      for (int j = 0; j < 7; ++j) qq[j] = qp[j];
      float Fl[MPN_NLINK * 12];
      spec_fk(qq, prismatic, Fl, nullptr);
#pragma unroll
      for (int i = 0; i < MPN_NLINK * 12; ++i) F[threadIdx.x][i] = Fl[i];
    }
    __syncthreads();
    // (timestep, sphere, primitive slice) items: "min over primitives <= r" is "any primitive <= r", so when a chunk has
    // fewer pairs than threads (the per-step check, T = 1) each pair is split over two halves of the primitive list
    const int nsl = (nt * S * 2 <= (int)blockDim.x) ? 2 : 1;
    for (int p = threadIdx.x; p < nt * S * nsl; p += blockDim.x) {
      const int sl = p / (nt * S), pr = p - sl * nt * S;
      const int t = pr / S, k = pr - t * S;
      float x, y, z;
      m34_apply(F[t] + 12 * sl_s[k], sc_s[3 * k], sc_s[3 * k + 1], sc_s[3 * k + 2], x, y, z);
      const float d = scene_sdf_packed(fr, nc * sl / nsl, nc * (sl + 1) / nsl, nc + ny * sl / nsl, nc + ny * (sl + 1) / nsl, x, y, z);
      if (d <= sr_s[k]) atomicMin(&first, tc + t);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int hit = first != 0x7fffffff;
    int fs = hit ? first + t0 : -1;
    if (accumulate) {
      if (hit) flags[b] = 1;
      if (first_step && hit && first_step[b] < 0) first_step[b] = fs;
    } else {
      flags[b] = (uint8_t)hit;
      if (first_step) first_step[b] = fs;
    }
  }
}

int launch_sweep(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* traj, int T, int problem_stride,
                 int t0, int accumulate, uint8_t* flags, int32_t* first_step, const float* frames_in) {
  if (T != 1) frames_in = nullptr;
  sweep_kernel<<<B, SWEEP_THREADS, 0, s>>>(sc, c->cfg.max_cuboids, c->cfg.max_cylinders, c->cfg.quirk_frames, traj, T,
                                           problem_stride, t0, c->prismatic, c->S, c->sph_c, c->sph_r, c->sph_l,
                                           accumulate, flags, first_step, frames_in);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ------------------------------------------------------------------------------------------------ trajectory evaluation
// Evaluator.evaluate_trajectory's device-computable subset (metrics.py:311-322,340-384,411-434,487-523).  One CTA per
// problem; poses in chunks of EVAL_TCHUNK: one thread per pose does FK + the joint-limit test, all threads then sweep
// (pose, sphere) pairs against the scene and (pose, sphere, sphere) pairs against each other.  Step lengths are written
// to shared memory and summed by thread 0 in pose order, which is the order the oracle adds them in.
constexpr int EVAL_TCHUNK = 16;
constexpr int EVAL_THREADS = 128;

// rotation angle (degrees) of A * B^T: atan2(|antisymmetric part| / 2, (trace - 1) / 2)
__device__ __forceinline__ float rel_angle_deg(const float* A, const float* Bp) {
  float R[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[3 * i + j] = ffma(A[4 * i + 2], Bp[4 * j + 2], ffma(A[4 * i + 1], Bp[4 * j + 1], fmul(A[4 * i], Bp[4 * j])));
  float ax = fsub(R[7], R[5]), ay = fsub(R[2], R[6]), az = fsub(R[3], R[1]);
  float s = fmul(0.5f, fsqrt(ffma(az, az, ffma(ay, ay, fmul(ax, ax)))));
  float c = fmul(0.5f, fsub(fadd(fadd(R[0], R[4]), R[8]), 1.0f));
  return fmul(atan2f(s, c), 57.29577951308232f);
}

__device__ __forceinline__ int link_group(int link) { return link < 7 ? link : 7; }

__global__ void __launch_bounds__(EVAL_THREADS)
evaluate_kernel(mpn_scene sc, int M1, int M2, int quirk, const float* __restrict__ traj, int T1,
                const int32_t* __restrict__ num_poses, const float* __restrict__ target, const float* __restrict__ lim,
                float prismatic, int S, const float* __restrict__ sph_c, const float* __restrict__ sph_r,
                const int32_t* __restrict__ sph_l, mpn_scene tv, int V1, int V2, mpn_scene nv, int N1, int N2,
                float* __restrict__ out) {
  extern __shared__ float dyn[];                 // eef poses [T1][12] | step lengths pos/ori/config [3][T1]
  __shared__ PrimFrame fr[MAX_PRIMS];
  __shared__ float F[EVAL_TCHUNK][MPN_NLINK * 12];
  __shared__ float W[EVAL_TCHUNK][MAX_SPHERES * 3];
  __shared__ float sc_s[MAX_SPHERES * 3];
  __shared__ float sr_s[MAX_SPHERES];
  __shared__ int sl_s[MAX_SPHERES];
  __shared__ int first, jl, selfc, depth_bits;
  float* E = dyn;
  float* Lp = dyn + (size_t)T1 * 12;
  float* Lo = Lp + T1;
  float* Lc = Lo + T1;
  const int b = blockIdx.x, tid = threadIdx.x;
  stage_scene(sc, b, M1, M2, quirk != 0, fr);
  for (int k = tid; k < S; k += blockDim.x) {
    sc_s[3 * k] = sph_c[3 * k]; sc_s[3 * k + 1] = sph_c[3 * k + 1]; sc_s[3 * k + 2] = sph_c[3 * k + 2];
    sr_s[k] = sph_r[k]; sl_s[k] = sph_l[k];
  }
  if (tid == 0) { first = 0x7fffffff; jl = 0; selfc = 0; depth_bits = 0; }
  int n = num_poses ? num_poses[b] : T1;
  n = max(1, min(n, T1));
  __syncthreads();
  const float* tb = traj + (size_t)b * T1 * 7;
  for (int tc = 0; tc < n; tc += EVAL_TCHUNK) {
    const int nt = min(EVAL_TCHUNK, n - tc);
    if (tid < nt) {
      float qq[7];
      const float* qp = tb + (size_t)(tc + tid) * 7;
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        qq[j] = qp[j];
        bad |= qq[j] < lim[2 * j] || qq[j] > lim[2 * j + 1];
      }
      if (bad) jl = 1;
      float Fl[MPN_NLINK * 12], El[12];
      spec_fk(qq, prismatic, Fl, El);
#pragma unroll
      for (int i = 0; i < MPN_NLINK * 12; ++i) F[tid][i] = Fl[i];
#pragma unroll
      for (int i = 0; i < 12; ++i) E[(size_t)(tc + tid) * 12 + i] = El[i];
    }
    __syncthreads();
    float pen_max = 0.f;
    for (int p = tid; p < nt * S; p += blockDim.x) {
      int t = p / S, k = p - t * S;
      float x, y, z;
      m34_apply(F[t] + 12 * sl_s[k], sc_s[3 * k], sc_s[3 * k + 1], sc_s[3 * k + 2], x, y, z);
      W[t][3 * k] = x; W[t][3 * k + 1] = y; W[t][3 * k + 2] = z;
      float d = scene_sdf(fr, 0, M1, M1, M1 + M2, x, y, z);
      if (d <= sr_s[k]) atomicMin(&first, tc + t);
      pen_max = fmaxf(pen_max, fsub(sr_s[k], d));
    }
    if (pen_max > 0.f) atomicMax(&depth_bits, __float_as_int(pen_max));   // non-negative floats order like ints
    __syncthreads();
    const int pairs = S * S;
    for (int p = tid; p < nt * pairs; p += blockDim.x) {
      int t = p / pairs, r = p - t * pairs;
      int i = r / S, j = r - i * S;
      if (j <= i) continue;
      int g = link_group(sl_s[i]) - link_group(sl_s[j]);
      if (g < 0) g = -g;
      if (g < 2) continue;
      float dx = fsub(W[t][3 * i], W[t][3 * j]), dy = fsub(W[t][3 * i + 1], W[t][3 * j + 1]), dz = fsub(W[t][3 * i + 2], W[t][3 * j + 2]);
      float dist = fsqrt(ffma(dz, dz, ffma(dy, dy, fmul(dx, dx))));
      if (dist < fadd(sr_s[i], sr_s[j])) selfc = 1;
    }
    __syncthreads();
  }
  for (int t = 1 + tid; t < n; t += blockDim.x) {
    const float* a = E + (size_t)t * 12;
    const float* p = a - 12;
    float dx = fsub(a[3], p[3]), dy = fsub(a[7], p[7]), dz = fsub(a[11], p[11]);
    Lp[t] = fsqrt(ffma(dz, dz, ffma(dy, dy, fmul(dx, dx))));
    Lo[t] = rel_angle_deg(a, p);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) { float dq = fsub(tb[(size_t)t * 7 + j], tb[(size_t)(t - 1) * 7 + j]); acc = ffma(dq, dq, acc); }
    Lc[t] = fsqrt(acc);
  }
  __syncthreads();
  if (tid != 0) return;
  float pos_path = 0.f, ori_path = 0.f, cfg_path = 0.f;
  for (int t = 1; t < n; ++t) { pos_path = fadd(pos_path, Lp[t]); ori_path = fadd(ori_path, Lo[t]); cfg_path = fadd(cfg_path, Lc[t]); }
  const float* last = E + (size_t)(n - 1) * 12;
  const float* Tg = target + (size_t)b * 12;
  float dx = fsub(last[3], Tg[3]), dy = fsub(last[7], Tg[7]), dz = fsub(last[11], Tg[11]);
  float pos_cm = fmul(100.0f, fsqrt(ffma(dz, dz, ffma(dy, dy, fmul(dx, dx)))));
  float ori = rel_angle_deg(last, Tg);
  // metrics.py:497-504: negative volumes containing the target are dropped; final xyz inside the target volume and
  // outside every remaining negative volume.  geometrout primitives -> textbook rotations.
  bool region = true;
  if (V1 + V2 > 0) {
    float best = __int_as_float(0x7f800000);
    for (int m = 0; m < V1 + V2; ++m) {
      PrimFrame f = prim_frame_of(tv, b, V1, V2, m, false);
      if (f.valid == 0.f) continue;
      best = fminf(best, m < V1 ? sdf_cuboid(f, last[3], last[7], last[11]) : sdf_cylinder(f, last[3], last[7], last[11]));
    }
    if (!(best <= 0.f)) region = false;
  }
  for (int m = 0; m < N1 + N2; ++m) {
    PrimFrame f = prim_frame_of(nv, b, N1, N2, m, false);
    if (f.valid == 0.f) continue;
    float at_target = m < N1 ? sdf_cuboid(f, Tg[3], Tg[7], Tg[11]) : sdf_cylinder(f, Tg[3], Tg[7], Tg[11]);
    float at_final = m < N1 ? sdf_cuboid(f, last[3], last[7], last[11]) : sdf_cylinder(f, last[3], last[7], last[11]);
    if (at_target > 0.f && !(at_final > 0.f)) region = false;
  }
  const bool hit = first != 0x7fffffff;
  const bool physical = hit || jl || selfc;
  float* o = out + (size_t)b * MPN_EVAL_COLS;
#pragma unroll
  for (int i = 0; i < MPN_EVAL_COLS; ++i) o[i] = 0.f;
  o[MPN_E_COLLISION] = hit ? 1.f : 0.f;
  o[MPN_E_JOINT_LIMIT_VIOLATION] = jl ? 1.f : 0.f;
  o[MPN_E_SELF_COLLISION] = selfc ? 1.f : 0.f;
  o[MPN_E_PHYSICAL_VIOLATIONS] = physical ? 1.f : 0.f;
  o[MPN_E_POSITION_ERROR_CM] = pos_cm;
  o[MPN_E_ORIENTATION_ERROR_DEG] = ori;
  o[MPN_E_EFF_POSITION_PATH_LENGTH] = pos_path;
  o[MPN_E_EFF_ORIENTATION_PATH_LENGTH_DEG] = ori_path;
  o[MPN_E_CORRECT_FINAL_REGION] = region ? 1.f : 0.f;
  o[MPN_E_SUCCESS] = (pos_cm < 1.0f && region && ori < 15.0f && !physical) ? 1.f : 0.f;
  o[MPN_E_NUM_STEPS] = (float)n;
  o[MPN_E_FIRST_COLLISION_STEP] = hit ? (float)first : -1.f;
  o[MPN_E_CONFIG_PATH_LENGTH] = cfg_path;
  o[MPN_E_MAX_COLLISION_DEPTH] = __int_as_float(depth_bits);
}

int launch_evaluate(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* traj, int T1, const int32_t* num_poses,
                    const float* target, const mpn_scene& tv, int V1, int V2, const mpn_scene& nv, int N1, int N2, float* out) {
  size_t dyn = (size_t)T1 * 15 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    MPN_CHECK_CUDA(cudaFuncSetAttribute(evaluate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 15 * (int)sizeof(float)));
    attr_set = true;
  }
  evaluate_kernel<<<B, EVAL_THREADS, dyn, s>>>(sc, c->cfg.max_cuboids, c->cfg.max_cylinders, c->cfg.quirk_frames, traj, T1,
                                               num_poses, target, c->limits, c->prismatic, c->S, c->sph_c, c->sph_r, c->sph_l,
                                               tv, V1, V2, nv, N1, N2, out);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ------------------------------------------------------------------------------------------------ SPARC smoothness
// third_party/sparc.py:48-140 (Evaluator.calculate_smoothness, metrics.py:387-409) for B speed profiles at once.
// One CTA per profile: zero-padded DFT length nfft = 2^(ceil(log2 n) + padlevel) evaluated directly for the bins with
// f = k fs / nfft <= fc (twiddles from a shared-memory table indexed by k t mod nfft), magnitudes normalised by the
// spectrum's maximum, the amplitude-threshold window [first, last] bin >= amp_th, and the arc length over that window.
constexpr int SPARC_THREADS = 256;

__global__ void __launch_bounds__(SPARC_THREADS)
sparc_kernel(const float* __restrict__ movement, int n_max, const int32_t* __restrict__ num, float fs, int padlevel, float fc, float amp_th,
             int nfft_cap, float* __restrict__ out) {
  extern __shared__ float sp[];   // cos[nfft] | sin[nfft] | mag[nfft] | m[n_max]
  __shared__ float redf[SPARC_THREADS / 32];
  __shared__ int redi[2][SPARC_THREADS / 32];
  __shared__ float s_max;
  __shared__ int s_first, s_last, s_any;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n = num ? num[b] : n_max;
  n = max(0, min(n, n_max));
  const float* mv = movement + (size_t)b * n_max;
  int lg = 0;
  while ((1 << lg) < n) ++lg;
  const int nfft = 1 << (lg + padlevel);
  if (n < 1 || nfft > nfft_cap) { if (tid == 0) out[b] = n < 1 ? 0.f : __int_as_float(0x7fc00000); return; }
  float* ct = sp;
  float* st = ct + nfft_cap;
  float* mag = st + nfft_cap;
  float* m = mag + nfft_cap;
  int nz = 0;
  for (int t = tid; t < n; t += blockDim.x) { const float v = mv[t]; m[t] = v; nz |= fabsf(v) > 1e-8f; }   // np.allclose(movement, 0)
  for (int k = tid; k < nfft; k += blockDim.x) { float sv, cv; sincospif(2.0f * (float)k / (float)nfft, &sv, &cv); ct[k] = cv; st[k] = sv; }
  if (tid == 0) s_any = 0;
  __syncthreads();
  if (nz) atomicOr(&s_any, 1);
  __syncthreads();
  if (!s_any) { if (tid == 0) out[b] = 0.f; return; }
  // bins with f[k] = k * (fs / nfft) <= fc, f = np.arange(0, fs, fs / nfft)
  const float df = fs / (float)nfft;
  int kc = min(nfft - 1, (int)floorf(fc / df));
  while (kc > 0 && (float)kc * df > fc) --kc;
  while (kc + 1 < nfft && (float)(kc + 1) * df <= fc) ++kc;
  // the maximum of the whole spectrum: by symmetry bins 0 .. nfft/2; evaluate max(kc, nfft/2) bins, keep 0 .. kc
  const int kend = max(kc, nfft / 2);
  float lmax = 0.f;
  for (int k = tid; k <= kend; k += blockDim.x) {
    float re = 0.f, im = 0.f;
    int ph = 0;                                  // (k * t) mod nfft, advanced incrementally
    for (int t = 0; t < n; ++t) {
      re = fmaf(m[t], ct[ph], re); im = fmaf(-m[t], st[ph], im);
      ph += k; if (ph >= nfft) ph -= nfft;
    }
    const float a = sqrtf(re * re + im * im);
    if (k <= kc) mag[k] = a;
    lmax = fmaxf(lmax, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0) redf[warp] = lmax;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int w = 0; w < SPARC_THREADS / 32; ++w) v = fmaxf(v, redf[w]); s_max = v; }
  __syncthreads();
  const float inv = 1.0f / s_max;
  int first = 0x7fffffff, last = -1;
  for (int k = tid; k <= kc; k += blockDim.x) {
    const float v = mag[k] * inv;
    mag[k] = v;
    if (v >= amp_th) { first = min(first, k); last = max(last, k); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { first = min(first, __shfl_xor_sync(0xffffffffu, first, o)); last = max(last, __shfl_xor_sync(0xffffffffu, last, o)); }
  if (lane == 0) { redi[0][warp] = first; redi[1][warp] = last; }
  __syncthreads();
  if (tid == 0) {
    int f0 = 0x7fffffff, l0 = -1;
    for (int w = 0; w < SPARC_THREADS / 32; ++w) { f0 = min(f0, redi[0][w]); l0 = max(l0, redi[1][w]); }
    s_first = f0; s_last = l0;
  }
  __syncthreads();
  const int k0 = s_first, k1 = s_last;
  if (k1 <= k0) { if (tid == 0) out[b] = 0.f; return; }   // a one-bin window has no arc
  const float dfn = 1.0f / (float)(k1 - k0);               // diff(f_sel) / (f_sel[-1] - f_sel[0]) on the uniform grid
  float acc = 0.f;
  for (int k = k0 + tid; k < k1; k += blockDim.x) {
    const float dm = mag[k + 1] - mag[k];
    acc += sqrtf(dfn * dfn + dm * dm);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) redf[warp] = acc;
  __syncthreads();
  if (tid == 0) { float v = 0.f; for (int w = 0; w < SPARC_THREADS / 32; ++w) v += redf[w]; out[b] = -v; }
}

int launch_sparc(mpn_ctx* c, cudaStream_t s, int B, int n_max, const float* movement, const int32_t* num, float fs, int padlevel, float fc,
                 float amp_th, float* out) {
  int lg = 0;
  while ((1 << lg) < n_max) ++lg;
  const int nfft_cap = 1 << (lg + padlevel);
  MPN_REQUIRE(nfft_cap <= 16384, "sparc: padded length %d exceeds 16384 (n_max=%d, padlevel=%d)", nfft_cap, n_max, padlevel);
  const size_t smem = ((size_t)3 * nfft_cap + n_max) * sizeof(float);
  static size_t attr = 0;
  if (smem > attr) {
    MPN_CHECK_CUDA(cudaFuncSetAttribute(sparc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  sparc_kernel<<<B, SPARC_THREADS, smem, s>>>(movement, n_max, num, fs, padlevel, fc, amp_th, nfft_cap, out);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------- depth-camera clouds
// Stand-in for run_inference.convert_primitive_problems_to_depth (run_inference.py:194-257: Bullet renders a depth image of
// the primitives from a fixed camera, robot removed, and un-projects it).  Here every pixel's ray is intersected
// analytically with the scene's cuboids and cylinders in their own frames (the frames of the SDF, so a hit has sdf = 0);
// the nearest hit with camera depth in [near, far] becomes one world point.  Camera frame: OpenGL convention (x right, y up,
// looking along -z) -- the one under which the reference's evaluation cameras (run_inference.py:215-243) face their scenes;
// ray direction (u sx, -v sy, -1) with v growing downwards, so the ray parameter IS the depth.  Spec arithmetic throughout (bit-exact vs the oracle).
__device__ __forceinline__ float ray_cuboid(const PrimFrame& f, const float* o, const float* d, float tnear, float tfar) {
  float t0 = tnear, t1 = tfar;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float ol = fadd(dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], o[0], o[1], o[2]), f.Rt[i]);
    const float dl = dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], d[0], d[1], d[2]);
    if (dl == 0.0f) {
      if (fabsf(ol) > f.h[i]) return __int_as_float(0x7f800000);
      continue;
    }
    const float inv = fdiv(1.0f, dl);
    const float ta = fmul(fsub(-f.h[i], ol), inv), tb = fmul(fsub(f.h[i], ol), inv);
    t0 = fmaxf(t0, fminf(ta, tb));
    t1 = fminf(t1, fmaxf(ta, tb));
  }
  return t0 <= t1 ? t0 : __int_as_float(0x7f800000);
}

__device__ __forceinline__ float ray_cylinder(const PrimFrame& f, const float* o, const float* d, float tnear, float tfar) {
  float ol[3], dl[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    ol[i] = fadd(dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], o[0], o[1], o[2]), f.Rt[i]);
    dl[i] = dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], d[0], d[1], d[2]);
  }
  const float r2 = fmul(f.h[0], f.h[0]), hh = f.h[1];
  float best = __int_as_float(0x7f800000);
  const float a = ffma(dl[1], dl[1], fmul(dl[0], dl[0]));
  if (a > 0.0f) {                                     // side wall, entry point
    const float bq = ffma(ol[1], dl[1], fmul(ol[0], dl[0]));
    const float cq = fsub(ffma(ol[1], ol[1], fmul(ol[0], ol[0])), r2);
    const float disc = fsub(fmul(bq, bq), fmul(a, cq));
    if (disc >= 0.0f) {
      const float t = fdiv(fsub(-bq, fsqrt(disc)), a);
      const float z = ffma(t, dl[2], ol[2]);
      if (t >= tnear && t <= tfar && fabsf(z) <= hh) best = t;
    }
  }
  if (dl[2] != 0.0f) {                                // the two caps
    const float inv = fdiv(1.0f, dl[2]);
#pragma unroll
    for (int sgn = 0; sgn < 2; ++sgn) {
      const float t = fmul(fsub(sgn ? -hh : hh, ol[2]), inv);
      const float x = ffma(t, dl[0], ol[0]), y = ffma(t, dl[1], ol[1]);
      if (t >= tnear && t <= tfar && ffma(y, y, fmul(x, x)) <= r2) best = fminf(best, t);
    }
  }
  return best;
}

// CTA per problem; hit pixels are compacted to the front of points[b] in pixel (row-major) order.
__global__ void __launch_bounds__(256) render_depth_kernel(mpn_scene sc, int M1, int M2, int quirk, const float* __restrict__ camera,
                                                           int camera_stride, int W, int H, float du, float dv, float sx, float sy,
                                                           float tnear, float tfar, float* __restrict__ points,
                                                           int32_t* __restrict__ counts) {
  __shared__ PrimFrame fr[MAX_PRIMS];
  __shared__ int cnts[2], wcnt[8], wtot[8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  stage_scene_compact(sc, b, M1, M2, quirk != 0, fr, cnts, wcnt);
  const int nc = cnts[0], ny = cnts[1];
  const float* C = camera + (size_t)b * camera_stride;
  const float o[3] = {C[3], C[7], C[11]};
  float* out = points + (size_t)b * W * H * 3;
  int base = 0;
  for (int p0 = 0; p0 < W * H; p0 += 256) {
    const int pix = p0 + threadIdx.x;
    float t = __int_as_float(0x7f800000), d[3] = {0.f, 0.f, 0.f};
    if (pix < W * H) {
      const int row = pix / W, col = pix - row * W;
      const float u = ffma((float)col + 0.5f, du, -1.0f), v = ffma((float)row + 0.5f, dv, -1.0f);
      const float dcx = fmul(u, sx), dcy = fmul(v, sy);
#pragma unroll
      for (int i = 0; i < 3; ++i) d[i] = fsub(ffma(-C[4 * i + 1], dcy, fmul(C[4 * i], dcx)), C[4 * i + 2]);
      for (int m = 0; m < nc; ++m) t = fminf(t, ray_cuboid(fr[m], o, d, tnear, tfar));
      for (int m = nc; m < nc + ny; ++m) t = fminf(t, ray_cylinder(fr[m], o, d, tnear, tfar));
    }
    const bool hit = t <= tfar;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wtot[warp] = __popc(bal);
    __syncthreads();
    int off = base, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { if (w < warp) off += wtot[w]; tot += wtot[w]; }
    if (hit) {
      float* q = out + (size_t)(off + __popc(bal & ((1u << lane) - 1u))) * 3;
      q[0] = ffma(t, d[0], o[0]); q[1] = ffma(t, d[1], o[1]); q[2] = ffma(t, d[2], o[2]);
    }
    base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = base;
}

int launch_render_depth(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, const float* camera, int per_problem_camera, int W, int H,
                        float sx, float sy, float tnear, float tfar, float* points, int32_t* counts) {
  render_depth_kernel<<<B, 256, 0, s>>>(sc, c->cfg.max_cuboids, c->cfg.max_cylinders, c->cfg.quirk_frames, camera,
                                        per_problem_camera ? 12 : 0, W, H, 2.0f / (float)W, 2.0f / (float)H, sx, sy, tnear, tfar, points, counts);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
