// sa_simt.cu -- fp32 (parity-mode) fused set abstraction: ball query -> group -> 3-layer shared MLP -> max.
//
// Replaces pointnet2_ops QueryAndGroup / GroupAll + Conv2d(1x1)+ReLU x3 + max_pool2d inside PointnetSAModule
// (call site mpinets/model.py:365-383,423-424).  The grouped tensor [B,C,npoint,nsample] is never materialised:
// one CTA owns one centroid (or one 32-row tile of the group-all module), keeps the activations K-major in shared
// memory and only the pooled [C_out] row goes back to HBM.  fp32 FMA, fp32 accumulate: this is the 1e-5 mode;
// the throughput mode is the tcgen05 kernel in sa_tc.cu.
#include "engine.h"
#include "spec_math.cuh"

namespace mpn {

constexpr int SA_THREADS = 256;

// One MLP layer on an R-row tile.  act_in [K][R] (smem, k-major), wt [K][NOUT] (global, k-major), bias [NOUT].
// !LAST: act_out [NOUT][R] = relu(.) ; LAST: omax[c] = max(omax[c], max_rows relu(.))  (owner thread per column).
template <int K, int NOUT, int R, bool LAST>
__device__ __forceinline__ void mlp_layer(const float* __restrict__ act_in, const float* __restrict__ wt,
                                          const float* __restrict__ bias, float* __restrict__ act_out,
                                          float* __restrict__ omax, int* __restrict__ oarg = nullptr, int row_base = 0) {
  constexpr int RG = R / 8;                 // row groups (8 rows per thread)
  constexpr int CG = SA_THREADS / RG;       // column groups
  constexpr int CH = (NOUT < CG * 8) ? NOUT : CG * 8;  // columns per chunk
  constexpr int NC = CH / CG;               // columns per thread per chunk
  static_assert(NC >= 4 && NC % 4 == 0, "column tile");
  static_assert(NOUT % CH == 0, "chunking");
  const int rg = threadIdx.x % RG, cg = threadIdx.x / RG;
  const int r0 = rg * 8;
  for (int chunk = 0; chunk < NOUT / CH; ++chunk) {
    const int c0 = chunk * CH + cg * NC;
    float acc[8][NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      float bj = __ldg(bias + c0 + j);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = bj;
    }
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      float a[8], w[NC];
      const float4* ap = reinterpret_cast<const float4*>(act_in + k * R + r0);
      float4 a0 = ap[0], a1 = ap[1];
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      const float4* wp = reinterpret_cast<const float4*>(wt + (size_t)k * NOUT + c0);
#pragma unroll
      for (int j = 0; j < NC / 4; ++j) {
        float4 v = __ldg(wp + j);
        w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (!LAST) {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        float4 o0 = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
        float4 o1 = make_float4(fmaxf(acc[4][j], 0.f), fmaxf(acc[5][j], 0.f), fmaxf(acc[6][j], 0.f), fmaxf(acc[7][j], 0.f));
        float4* op = reinterpret_cast<float4*>(act_out + (size_t)(c0 + j) * R + r0);
        op[0] = o0; op[1] = o1;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        float m = acc[0][j];
        if (oarg == nullptr) {
#pragma unroll
          for (int i = 1; i < 8; ++i) m = fmaxf(m, acc[i][j]);
          m = fmaxf(m, 0.f);
#pragma unroll
          for (int o = RG / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (rg == 0) omax[c0 + j] = fmaxf(omax[c0 + j], m);
        } else {
          // training forward: also the pooled row (first occurrence of the maximum, like max_pool2d's argmax)
          int mi = 0;
#pragma unroll
          for (int i = 1; i < 8; ++i)
            if (acc[i][j] > m) { m = acc[i][j]; mi = i; }
          mi += r0 + row_base;
#pragma unroll
          for (int o = RG / 2; o > 0; o >>= 1) {
            float om = __shfl_xor_sync(0xffffffffu, m, o);
            int oi = __shfl_xor_sync(0xffffffffu, mi, o);
            if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
          }
          if (rg == 0 && m > omax[c0 + j]) { omax[c0 + j] = m; oarg[c0 + j] = mi; }
        }
      }
    }
  }
}

struct SaWeights {
  const float* wt[3];
  const float* b[3];
};

// Ball query by the whole CTA: warp w scans the contiguous segment [w*seg, (w+1)*seg) in index order and records
// its hits; segments are then concatenated in warp order and truncated to nsample (== first nsample in index order).
template <int NS>
__device__ __forceinline__ void cta_ball_query(const float* __restrict__ p, int N, int stride, float cx, float cy, float cz,
                                               float r2, int* __restrict__ idx_s, int* __restrict__ wl /*[8][NS]*/,
                                               int* __restrict__ wcnt /*[8]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = SA_THREADS / 32;
  const int seg = ((N + NW - 1) / NW + 31) & ~31;
  const int k_begin = warp * seg, k_end = min(N, k_begin + seg);
  int cnt = 0;
  for (int k0 = k_begin; k0 < k_end && cnt < NS; k0 += 32) {
    int k = k0 + lane;
    bool hit = false;
    if (k < k_end) {
      float d2 = dist2(cx, cy, cz, __ldg(p + (size_t)k * stride), __ldg(p + (size_t)k * stride + 1), __ldg(p + (size_t)k * stride + 2));
      hit = d2 < r2;
    }
    unsigned m = __ballot_sync(0xffffffffu, hit);
    int pos = cnt + __popc(m & ((1u << lane) - 1u));
    if (hit && pos < NS) wl[warp * NS + pos] = k;
    cnt += __popc(m);
  }
  if (lane == 0) wcnt[warp] = min(cnt, NS);
  __syncthreads();
  int total = 0, first = 0;
  bool have_first = false;
  int base_of_me = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    int cw = wcnt[w];
    if (w == warp) base_of_me = total;
    if (!have_first && cw > 0) { first = wl[w * NS]; have_first = true; }
    total += cw;
  }
  int mine = wcnt[warp];
  for (int l = lane; l < mine; l += 32)
    if (base_of_me + l < NS) idx_s[base_of_me + l] = wl[warp * NS + l];
  total = min(total, NS);
  for (int l = total + threadIdx.x; l < NS; l += SA_THREADS) idx_s[l] = first;
  __syncthreads();
}

// grid (npoint, B).  CFEAT = input feature channels (CIN = 3 + CFEAT).
template <int CFEAT, int C1, int C2, int C3>
__global__ void __launch_bounds__(SA_THREADS) sa_group_kernel(const float* __restrict__ xyz, int stride,
                                                              const float* __restrict__ feats, int feat_stride, int N,
                                                              const float* __restrict__ new_xyz, int npoint, float r2,
                                                              SaWeights W, float* __restrict__ out,
                                                              int32_t* __restrict__ ball_idx, uint8_t* __restrict__ arg_out) {
  constexpr int R = NSAMPLE, CIN = 3 + CFEAT;
  constexpr int CA = (CIN > C2 ? CIN : C2), CB = C1;
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;               // [CA][R]
  float* bufB = bufA + CA * R;      // [CB][R]
  float* omax = bufB + CB * R;      // [C3]
  int* idx_s = reinterpret_cast<int*>(omax + C3);  // [R]
  int* wl = idx_s + R;                             // [8][R]
  int* wcnt = wl + 8 * R;                          // [8]
  int* oarg = arg_out ? wcnt + 8 : nullptr;        // [C3] (training forward only)
  const int b = blockIdx.y, j = blockIdx.x;
  const float* p = xyz + (size_t)b * N * stride;
  const float* cp = new_xyz + ((size_t)b * npoint + j) * 3;
  const float cx = cp[0], cy = cp[1], cz = cp[2];
  for (int c = threadIdx.x; c < C3; c += SA_THREADS) omax[c] = 0.f;  // post-ReLU values are >= 0
  if (oarg) for (int c = threadIdx.x; c < C3; c += SA_THREADS) oarg[c] = 0;
  cta_ball_query<R>(p, N, stride, cx, cy, cz, r2, idx_s, wl, wcnt);
  if (ball_idx)
    for (int l = threadIdx.x; l < R; l += SA_THREADS) ball_idx[((size_t)b * npoint + j) * R + l] = idx_s[l];
  // group: row = neighbour, act0[k][row] = [dx,dy,dz, feats...]  (QueryAndGroup, use_xyz=True)
  {
    const int row = threadIdx.x % R, part = threadIdx.x / R;  // 2 parts
    const int k = idx_s[row];
    if (part == 0) {
      bufA[0 * R + row] = fsub(__ldg(p + (size_t)k * stride), cx);
      bufA[1 * R + row] = fsub(__ldg(p + (size_t)k * stride + 1), cy);
      bufA[2 * R + row] = fsub(__ldg(p + (size_t)k * stride + 2), cz);
    }
    const float* f = feats + ((size_t)b * N + k) * feat_stride;
    if (CFEAT % 8 == 0) {
      constexpr int HALF = CFEAT / 2;
      for (int c = part * HALF; c < (part + 1) * HALF; c += 4) {
        float4 v = __ldg(reinterpret_cast<const float4*>(f + c));
        bufA[(3 + c) * R + row] = v.x; bufA[(4 + c) * R + row] = v.y; bufA[(5 + c) * R + row] = v.z; bufA[(6 + c) * R + row] = v.w;
      }
    } else {
      for (int c = part; c < CFEAT; c += 2) bufA[(3 + c) * R + row] = __ldg(f + c);
    }
  }
  __syncthreads();
  mlp_layer<CIN, C1, R, false>(bufA, W.wt[0], W.b[0], bufB, nullptr);
  __syncthreads();
  mlp_layer<C1, C2, R, false>(bufB, W.wt[1], W.b[1], bufA, nullptr);
  __syncthreads();
  mlp_layer<C2, C3, R, true>(bufA, W.wt[2], W.b[2], nullptr, omax, oarg, 0);
  __syncthreads();
  float* o = out + ((size_t)b * npoint + j) * C3;
  for (int c = threadIdx.x; c < C3; c += SA_THREADS) o[c] = omax[c];
  if (oarg) for (int c = threadIdx.x; c < C3; c += SA_THREADS) arg_out[((size_t)b * npoint + j) * C3 + c] = (uint8_t)oarg[c];
}

// GroupAll module (model.py:383): rows = all N points (xyz NOT centred), tiles of 32 rows, grid (B).
template <int CFEAT, int C1, int C2, int C3>
__global__ void __launch_bounds__(SA_THREADS) sa_all_kernel(const float* __restrict__ xyz, int stride,
                                                            const float* __restrict__ feats, int feat_stride, int N,
                                                            SaWeights W, float* __restrict__ out,
                                                            uint8_t* __restrict__ arg_out) {
  constexpr int R = 32, CIN = 3 + CFEAT;
  constexpr int CA = (CIN > C2 ? CIN : C2), CB = C1;
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;
  float* bufB = bufA + CA * R;
  float* omax = bufB + CB * R;
  int* oarg = arg_out ? reinterpret_cast<int*>(omax + C3) : nullptr;   // [C3] (training forward only)
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C3; c += SA_THREADS) omax[c] = 0.f;
  if (oarg) for (int c = threadIdx.x; c < C3; c += SA_THREADS) oarg[c] = 0;
  for (int t0 = 0; t0 < N; t0 += R) {
    __syncthreads();
    {
      const int row = threadIdx.x % R, part = threadIdx.x / R;  // 8 parts
      const int k = min(t0 + row, N - 1);                        // tail rows duplicate the last point (max-invariant)
      const float* pk = xyz + ((size_t)b * N + k) * stride;
      if (part == 0) { bufA[0 * R + row] = pk[0]; bufA[1 * R + row] = pk[1]; bufA[2 * R + row] = pk[2]; }
      const float* f = feats + ((size_t)b * N + k) * feat_stride;
      for (int c = part * 4; c < CFEAT; c += 32) {
        float4 v = __ldg(reinterpret_cast<const float4*>(f + c));
        bufA[(3 + c) * R + row] = v.x; bufA[(4 + c) * R + row] = v.y; bufA[(5 + c) * R + row] = v.z; bufA[(6 + c) * R + row] = v.w;
      }
    }
    __syncthreads();
    mlp_layer<CIN, C1, R, false>(bufA, W.wt[0], W.b[0], bufB, nullptr);
    __syncthreads();
    mlp_layer<C1, C2, R, false>(bufB, W.wt[1], W.b[1], bufA, nullptr);
    __syncthreads();
    mlp_layer<C2, C3, R, true>(bufA, W.wt[2], W.b[2], nullptr, omax, oarg, t0);
  }
  __syncthreads();
  float* o = out + (size_t)b * C3;
  for (int c = threadIdx.x; c < C3; c += SA_THREADS) o[c] = omax[c];
  if (oarg) for (int c = threadIdx.x; c < C3; c += SA_THREADS) arg_out[(size_t)b * C3 + c] = (uint8_t)min(oarg[c], N - 1);
}

int launch_sa_simt(mpn_ctx* c, cudaStream_t s, int module, const float* xyz, int stride, const float* feats, int feat_stride,
                   int B, int N, const float* new_xyz, float* new_feats, int32_t* ball_idx, uint8_t* arg_out) {
  SaWeights W;
  for (int l = 0; l < 3; ++l) { W.wt[l] = c->w.sa[module][l].wt; W.b[l] = c->w.sa[module][l].b; }
  if (module == 0) {
    constexpr int CA = 64, CB = 64, C3 = 64, R = NSAMPLE;
    size_t smem = (size_t)(CA * R + CB * R + C3) * 4 + (size_t)(R + 8 * R + 8 + C3) * 4;
    auto k = sa_group_kernel<1, 64, 64, 64>;
    MPN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<dim3(SA1_NPOINT, B), SA_THREADS, smem, s>>>(xyz, stride, feats, feat_stride, N, new_xyz, SA1_NPOINT,
                                                     SA1_RADIUS * SA1_RADIUS, W, new_feats, ball_idx, arg_out);
  } else if (module == 1) {
    constexpr int CA = 128, CB = 128, C3 = 256, R = NSAMPLE;
    size_t smem = (size_t)(CA * R + CB * R + C3) * 4 + (size_t)(R + 8 * R + 8 + C3) * 4;
    auto k = sa_group_kernel<64, 128, 128, 256>;
    MPN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<dim3(SA2_NPOINT, B), SA_THREADS, smem, s>>>(xyz, stride, feats, feat_stride, N, new_xyz, SA2_NPOINT,
                                                     SA2_RADIUS * SA2_RADIUS, W, new_feats, ball_idx, arg_out);
  } else {
    constexpr int CA = 512, CB = 512, C3 = 1024, R = 32;
    size_t smem = (size_t)(CA * R + CB * R + C3 + C3) * 4;
    auto k = sa_all_kernel<256, 512, 512, 1024>;
    MPN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<B, SA_THREADS, smem, s>>>(xyz, stride, feats, feat_stride, N, W, new_feats, arg_out);
  }
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
