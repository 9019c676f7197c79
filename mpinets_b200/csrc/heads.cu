// heads.cu -- the small dense stacks around the point-cloud encoder (mpinets/model.py:47-66,75-91) in three launches instead of nine,
// and the skinny-batch dense layer used when the batch is a handful of problems (the reference's own callers run B = 1:
// run_inference.py:268-303, planning_node.py:78-151).
//
//   feature_encoder_kernel   q [B][7] -> 32 -> 64 -> 128 -> 128 -> 64 (LeakyReLU 0.01 between, none after the last; model.py:47-57), one
//                            warp per problem, activations in shared memory, transposed fp32 weights from L2.  Writes the 64 features
//                            behind the encoder output (the torch.cat of model.py:90) and, for the tensor-core modes, converts the whole
//                            2112-wide row into the operand format of the decoder's first GEMM (bf16 or [hi | lo] split bf16).
//   decoder.0 (2112 -> 512)  the tensor-core row GEMM (gemm_tc.cu, LeakyReLU epilogue) -- 84 % of the head's MACs --, or the skinny
//                            layer below for B <= 16, or linear_kernel in the fp32 SIMT mode.
//   decoder_tail_kernel      512 -> 256 -> 128 -> 7 (model.py:60-66) for 8 problems per CTA, activations in shared memory.
//   linear_skinny_kernel     Y [M][N] = act(X [M][K] W[N][K]^T + b) for M <= 16 in fp32: a warp per two output columns streams the
//                            weight rows once for all M rows (the FC head at B = 1 is pure weight streaming) and
//                            uses every SM, where a 256-row tensor-core tile would run on N / 256 of them.
#include "engine.h"
#include "tc_common.cuh"

namespace mpn {

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : 0.01f * v; }

// operand_mode: 0 none, 1 bf16 row [2112], 2 split row [2 x 2112] (hi | lo)
__global__ void __launch_bounds__(256) feature_encoder_kernel(const float* __restrict__ qn, int B, const float* __restrict__ w0,
                                                              const float* __restrict__ b0, const float* __restrict__ w1,
                                                              const float* __restrict__ b1, const float* __restrict__ w2,
                                                              const float* __restrict__ b2, const float* __restrict__ w3,
                                                              const float* __restrict__ b3, const float* __restrict__ w4,
                                                              const float* __restrict__ b4, float* __restrict__ cat, int ldcat,
                                                              int operand_mode, __nv_bfloat16* __restrict__ operand) {
  __shared__ float buf[8][2][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + warp;
  if (b >= B) return;
  float* x = buf[warp][0];
  float* y = buf[warp][1];
  if (lane < 7) x[lane] = qn[7 * b + lane];
  __syncwarp();
  // wt [in][out]: lane o reads consecutive addresses; accumulation in ascending k (= the fp32 parity mode's order)
  auto layer = [&](const float* __restrict__ wt, const float* __restrict__ bias, int in, int out, bool act) {
    for (int o = lane; o < out; o += 32) {
      float acc = 0.f;
      for (int k = 0; k < in; ++k) acc = fmaf(x[k], __ldg(wt + (size_t)k * out + o), acc);
      acc += bias[o];
      y[o] = act ? lrelu(acc) : acc;
    }
    __syncwarp();
    float* t = x; x = y; y = t;
  };
  layer(w0, b0, 7, 32, true);
  layer(w1, b1, 32, 64, true);
  layer(w2, b2, 64, 128, true);
  layer(w3, b3, 128, 128, true);
  layer(w4, b4, 128, 64, false);
  float* row = cat + (size_t)b * ldcat;
  row[ENC_DIM + lane] = x[lane];
  row[ENC_DIM + 32 + lane] = x[32 + lane];
  if (operand_mode) {
    constexpr int W = ENC_DIM + QF_DIM;
    for (int k = lane; k < W; k += 32) {
      const float v = k < ENC_DIM ? row[k] : x[k - ENC_DIM];
      if (operand_mode == 1) {
        operand[(size_t)b * W + k] = __float2bfloat16_rn(v);
      } else {
        __nv_bfloat16 h, l;
        tc::split_bf16(v, h, l);
        operand[(size_t)b * 2 * W + k] = h;
        operand[(size_t)b * 2 * W + W + k] = l;
      }
    }
  }
}

int launch_feature_encoder(mpn_ctx* c, cudaStream_t s, const float* qn, int B, float* cat, int ldcat, int operand_mode,
                           __nv_bfloat16* operand) {
  const Linear* L = c->w.fe;
  feature_encoder_kernel<<<(B + 7) / 8, 256, 0, s>>>(qn, B, L[0].wt, L[0].b, L[1].wt, L[1].b, L[2].wt, L[2].b, L[3].wt, L[3].b, L[4].wt, L[4].b,
                                                     cat, ldcat, operand_mode, operand);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// h0 [B][512] (decoder.0 output, LeakyReLU applied) -> decoder.2 -> decoder.4 -> decoder.6 -> dq [B][7]
constexpr int TAIL_R = 8;
__global__ void __launch_bounds__(256) decoder_tail_kernel(const float* __restrict__ h0, int B, const float* __restrict__ w1t,
                                                           const float* __restrict__ b1, const float* __restrict__ w2t,
                                                           const float* __restrict__ b2, const float* __restrict__ w3,
                                                           const float* __restrict__ b3, float* __restrict__ dq) {
  __shared__ __align__(16) float a0[TAIL_R][512];
  __shared__ __align__(16) float a1[TAIL_R][256];
  __shared__ __align__(16) float a2[TAIL_R][128];
  const int t = threadIdx.x, r0 = blockIdx.x * TAIL_R;
  const int rows = min(TAIL_R, B - r0);
  for (int i = t; i < TAIL_R * 512 / 4; i += 256) {
    const int r = i / 128, k4 = i % 128;
    reinterpret_cast<float4*>(a0[r])[k4] = r < rows ? __ldg(reinterpret_cast<const float4*>(h0 + (size_t)(r0 + r) * 512) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  {   // 512 -> 256: thread = output column, TAIL_R rows in registers
    float acc[TAIL_R];
#pragma unroll
    for (int r = 0; r < TAIL_R; ++r) acc[r] = 0.f;
    for (int k = 0; k < 512; k += 4) {
      const float w_0 = __ldg(w1t + (size_t)k * 256 + t), w_1 = __ldg(w1t + (size_t)(k + 1) * 256 + t),
                  w_2 = __ldg(w1t + (size_t)(k + 2) * 256 + t), w_3 = __ldg(w1t + (size_t)(k + 3) * 256 + t);
#pragma unroll
      for (int r = 0; r < TAIL_R; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(&a0[r][k]);
        acc[r] = fmaf(a.w, w_3, fmaf(a.z, w_2, fmaf(a.y, w_1, fmaf(a.x, w_0, acc[r]))));
      }
    }
    const float bb = b1[t];
#pragma unroll
    for (int r = 0; r < TAIL_R; ++r) a1[r][t] = lrelu(acc[r] + bb);
  }
  __syncthreads();
  {   // 256 -> 128: thread = (column, row half)
    const int col = t & 127, rh = t >> 7;
    float acc[TAIL_R / 2];
#pragma unroll
    for (int r = 0; r < TAIL_R / 2; ++r) acc[r] = 0.f;
    for (int k = 0; k < 256; k += 4) {
      const float w_0 = __ldg(w2t + (size_t)k * 128 + col), w_1 = __ldg(w2t + (size_t)(k + 1) * 128 + col),
                  w_2 = __ldg(w2t + (size_t)(k + 2) * 128 + col), w_3 = __ldg(w2t + (size_t)(k + 3) * 128 + col);
#pragma unroll
      for (int r = 0; r < TAIL_R / 2; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(&a1[rh * (TAIL_R / 2) + r][k]);
        acc[r] = fmaf(a.w, w_3, fmaf(a.z, w_2, fmaf(a.y, w_1, fmaf(a.x, w_0, acc[r]))));
      }
    }
    const float bb = b2[col];
#pragma unroll
    for (int r = 0; r < TAIL_R / 2; ++r) a2[rh * (TAIL_R / 2) + r][col] = lrelu(acc[r] + bb);
  }
  __syncthreads();
  // 128 -> 7: one warp per row, a lane sums 4 inputs per output, butterfly reduction
  const int warp = t >> 5, lane = t & 31;
  if (warp < rows) {
    const float4 a = *reinterpret_cast<const float4*>(&a2[warp][lane * 4]);
#pragma unroll
    for (int o = 0; o < 7; ++o) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(w3 + o * 128) + lane);
      float v = fmaf(a.w, w.w, fmaf(a.z, w.z, fmaf(a.y, w.y, a.x * w.x)));
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) dq[(size_t)(r0 + warp) * 7 + o] = v + b3[o];
    }
  }
}

int launch_decoder_tail(mpn_ctx* c, cudaStream_t s, const float* h0, int B, float* dq) {
  const Linear* D = c->w.dec;
  decoder_tail_kernel<<<(B + TAIL_R - 1) / TAIL_R, 256, 0, s>>>(h0, B, D[1].wt, D[1].b, D[2].wt, D[2].b, D[3].w, D[3].b, dq);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------- skinny batch (M <= 16)
constexpr int SK_M = 16, SK_NPW = 2;
// act: 0 none, 1 LeakyReLU(0.01).  x_mode: 0 fp32 rows (pitch ldx), 1 bf16 rows, 2 split-bf16 rows ([hi | lo], lo at ldx / 2)
__global__ void __launch_bounds__(256) linear_skinny_kernel(const void* __restrict__ Xv, int ldx, int x_mode, const float* __restrict__ W,
                                                            const float* __restrict__ bias, int M, int N, int K, float* __restrict__ Y, int ldy,
                                                            int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * 8 + warp) * SK_NPW;
  if (n0 >= N) return;
  float acc[SK_M][SK_NPW];
#pragma unroll
  for (int m = 0; m < SK_M; ++m)
#pragma unroll
    for (int j = 0; j < SK_NPW; ++j) acc[m][j] = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    float4 w[SK_NPW];
#pragma unroll
    for (int j = 0; j < SK_NPW; ++j)
      w[j] = n0 + j < N ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + j) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int m = 0; m < SK_M; ++m) {
      if (m >= M) break;
      float4 x;
      if (x_mode == 0) {
        x = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(Xv) + (size_t)m * ldx + k));
      } else {
        const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(Xv) + (size_t)m * ldx + k;
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(xb));
        x = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xFFFF0000u), __uint_as_float(h.y << 16), __uint_as_float(h.y & 0xFFFF0000u));
        if (x_mode == 2) {
          const uint2 l = __ldg(reinterpret_cast<const uint2*>(xb + ldx / 2));
          x.x += __uint_as_float(l.x << 16); x.y += __uint_as_float(l.x & 0xFFFF0000u);
          x.z += __uint_as_float(l.y << 16); x.w += __uint_as_float(l.y & 0xFFFF0000u);
        }
      }
#pragma unroll
      for (int j = 0; j < SK_NPW; ++j) acc[m][j] = fmaf(x.w, w[j].w, fmaf(x.z, w[j].z, fmaf(x.y, w[j].y, fmaf(x.x, w[j].x, acc[m][j]))));
    }
  }
#pragma unroll
  for (int m = 0; m < SK_M; ++m) {
    if (m >= M) break;
#pragma unroll
    for (int j = 0; j < SK_NPW; ++j) {
      float v = acc[m][j];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0 && n0 + j < N) {
        v += bias[n0 + j];
        Y[(size_t)m * ldy + n0 + j] = act == 1 ? lrelu(v) : v;
      }
    }
  }
}

int launch_linear_skinny(mpn_ctx* c, cudaStream_t s, const void* X, int ldx, int x_mode, const Linear& L, int M, float* Y, int ldy, int act) {
  MPN_REQUIRE(M >= 1 && M <= SK_M && L.in % 4 == 0 && ldx % 4 == 0, "skinny dense layer: 1 <= M <= %d, K %% 4 == 0 (M=%d K=%d)", SK_M, M, L.in);
  const int warps = (L.out + SK_NPW - 1) / SK_NPW;
  linear_skinny_kernel<<<(warps + 7) / 8, 256, 0, s>>>(X, ldx, x_mode, L.w, L.b, M, L.out, L.in, Y, ldy, act);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
