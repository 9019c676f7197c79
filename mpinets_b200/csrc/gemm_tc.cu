// gemm_tc.cu -- tcgen05 GEMM for the row-shared layers of the encoder: the group-all module (SA3, model.py:383: rows =
// B*128 points, 259->512->512->1024 + max over each problem's 128 rows) and the FC head (model.py:385-393).
//
//   C[M][N] = epilogue(A[M][K] * W[N][K]^T + bias)     A, W bf16 K-major in HBM, fp32 accumulate in TMEM.
//   EPI_RELU_BF16: relu -> bf16 rows;  EPI_F32: fp32 rows (GroupNorm follows);  EPI_MAXPOOL: relu + max over each 128-row
//   sub-tile (= one problem) -> one bf16 row.
//
// CTA tile 256 x 256: two 128-row MMA sub-tiles that share every weight stage and together fill the 512 TMEM columns.
// Operands arrive through the tensor memory accelerator: per K stage of 64 one elected thread issues two
// cp.async.bulk.tensor loads (A box 64 x 256 rows, W box 64 x 256 rows, 128-byte swizzle) that complete on the stage's
// `full` mbarrier; a second elected thread waits on it, issues the eight tcgen05.mma of the stage and commits to the
// stage's `empty` mbarrier, which the producer waits on before refilling.  No block-wide sync inside the main loop, and
// L2 -> shared memory moves whole 128-byte rows.  (A cp.async.cg ring was the first version: ncu showed one 32-byte
// sector request per 16-byte lane copy -- 27 sectors per LDGSTS instruction, 2x the operand bytes over the crossbar --
// and 13-17 % tensor-pipe activity; profiles/r1_final2_kernels.summary.txt.  The TMA kernel is 2.3x faster.)
#include <cuda.h>

#include "engine.h"
#include "tc_common.cuh"

namespace mpn {
using namespace tc;

// 3: max-pool + winning row per column (training); 4 / 5: the split-bf16 ("bf16x3") mode's hand-off formats -- the fp32 result is
// written as TWO bf16 values hi = bf16(x), lo = bf16(x - hi) (hi at column n, lo at column c_lo_off + n of the same row), which the
// next GEMM consumes as its [hi | lo] A operand
// 6: fp32 rows with LeakyReLU(0.01) (decoder.0 of the policy head, model.py:58-60)
enum { EPI_RELU_BF16 = 0, EPI_F32 = 1, EPI_MAXPOOL = 2, EPI_MAXPOOL_ARG = 3, EPI_RELU_SPLIT = 4, EPI_MAXPOOL_SPLIT = 5, EPI_LRELU_F32 = 6 };
constexpr int G_BN = 256;

__device__ __forceinline__ uint32_t cvt_relu_pack(float first, float second) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(second), "f"(first));
  return d;
}

constexpr int T_BK = 64, T_STAGES = 3;
constexpr int T_A_BYTES = 256 * T_BK * 2, T_W_BYTES = G_BN * T_BK * 2, T_STAGE_BYTES = T_A_BYTES + T_W_BYTES;

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}

// passes = 1: plain bf16 GEMM.  passes = 3: split-bf16 GEMM -- both operands are given as (hi, lo) bf16 pairs (tmA / tmA2, tmW / tmW2)
// and the K loop runs three times into the same fp32 accumulator: A_hi W_hi + A_lo W_hi + A_hi W_lo (the lo*lo term, 2^-16 of the
// product, is dropped), i.e. an fp32-grade product on the bf16 tensor pipe at 3x the MMA count.
template <int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmW2, int passes, int K, const float* __restrict__ bias,
                int M, int N, void* __restrict__ Cout, int ldc, int c_lo_off, int* __restrict__ err, uint8_t* __restrict__ arg_out) {
  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B tiles: 1024-byte aligned stage bases
  __shared__ uint64_t full[T_STAGES], empty[T_STAGES], accum;
  __shared__ uint32_t tmem_slot;
  __shared__ int red[4][G_BN];
  __shared__ __align__(16) float sbias[G_BN];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int m0 = blockIdx.y * 256, n0 = blockIdx.x * G_BN;
  const int nst = (K + T_BK - 1) / T_BK;

  sbias[tid] = n0 + tid < N ? bias[n0 + tid] : 0.f;
  if (tid == 0) {
    for (int s = 0; s < T_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&accum, 1);
    mbar_fence_init();
    if ((smem_u32(smem) & 1023u) != 0) atomicExch(err, 2);   // the swizzle pattern assumes 1024-byte aligned tiles
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  bool ok = true;

  if (warp == 0) {
    // ---- producer: one elected lane streams the K stages
    if (elect_one()) {
      for (int it = 0; it < nst * passes; ++it) {
        const int slot = it % T_STAGES;
        if (it >= T_STAGES) ok = ok && mbar_wait(&empty[slot], ((it / T_STAGES) - 1) & 1);
        const uint32_t sA = smem_u32(smem) + slot * T_STAGE_BYTES;
        const int pass = it / nst, kt = it - pass * nst;
        mbar_arrive_expect_tx(&full[slot], T_STAGE_BYTES);
        tma_load_2d(sA, pass == 1 ? &tmA2 : &tmA, kt * T_BK, m0, &full[slot]);
        tma_load_2d(sA + T_A_BYTES, pass == 2 ? &tmW2 : &tmW, kt * T_BK, n0, &full[slot]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer
    if (elect_one()) {
      const uint64_t dA0 = make_smem_desc(smem_u32(smem), 16, 1024, LAYOUT_SW128);
      const uint64_t dW0 = make_smem_desc(smem_u32(smem) + T_A_BYTES, 16, 1024, LAYOUT_SW128);
      constexpr uint32_t id = make_idesc_bf16(128, G_BN);
      constexpr uint32_t SUB1 = 128 * 128 / 16;   // rows 128..255 of the A tile, in 16-byte units
      for (int it = 0; it < nst * passes; ++it) {
        const int slot = it % T_STAGES;
        ok = ok && mbar_wait(&full[slot], (it / T_STAGES) & 1);
        tc_fence_after();
        const uint32_t soff = (uint32_t)slot * (T_STAGE_BYTES / 16);
        const int kt = it % nst;
        const int ksteps = min(T_BK / 16, (K - kt * T_BK + 15) / 16);   // the K tail beyond the tensor is zero-filled by TMA
        for (int ks = 0; ks < ksteps; ++ks) mma_bf16_ss_off(tmem, dA0, soff + ks * 2, dW0, soff + ks * 2, id, (it | ks) != 0);
        for (int ks = 0; ks < ksteps; ++ks) mma_bf16_ss_off(tmem + 256, dA0, soff + SUB1 + ks * 2, dW0, soff + ks * 2, id, (it | ks) != 0);
        mma_commit(&empty[slot]);
      }
      mma_commit(&accum);
    }
    __syncwarp();
  }
  ok = mbar_wait(&accum, 0) && ok;
  tc_fence_after();
  if (!ok && (tid & 31) == 0) atomicExch(err, 1);

  // ---- epilogue: warp q = warp & 3 owns TMEM lanes 32q.., column half h = warp >> 2.  A thread holds one output ROW, so
  // storing straight from registers would touch 32 different rows per instruction (32 half-used sectors, ~8k cycles per
  // tile).  Rows therefore go through the (now idle) operand ring, 16-byte chunks XOR-swizzled by row so that both the
  // row-per-thread writes and the row-per-warp reads are conflict-free, and leave as full 512-byte row segments.
  const int q = warp & 3, h = warp >> 2, row = q * 32 + (tid & 31);
  constexpr int ROW_CHUNKS = (EPI == EPI_F32 || EPI == EPI_LRELU_F32 || EPI == EPI_RELU_SPLIT) ? 64 : 32;   // 16-byte chunks per 256-column output row
  constexpr bool POOL = EPI == EPI_MAXPOOL || EPI == EPI_MAXPOOL_ARG || EPI == EPI_MAXPOOL_SPLIT;
#pragma unroll 1
  for (int sub = 0; sub < 2; ++sub) {
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + sub * 256 + h * 128;
    const int m = m0 + sub * 128 + row;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tl + c0, v);
      tmem_ld_wait();
      const int nl = h * 128 + c0;
      const float4* bt = reinterpret_cast<const float4*>(sbias + nl);
      if (EPI == EPI_RELU_BF16) {
        uint8_t* rowp = smem + (size_t)row * (ROW_CHUNKS * 16);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 b0 = bt[j / 4], b1 = bt[j / 4 + 1];
          const int chunk = (nl + j) >> 3;
          *reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(cvt_relu_pack(__uint_as_float(v[j]) + b0.x, __uint_as_float(v[j + 1]) + b0.y),
                         cvt_relu_pack(__uint_as_float(v[j + 2]) + b0.z, __uint_as_float(v[j + 3]) + b0.w),
                         cvt_relu_pack(__uint_as_float(v[j + 4]) + b1.x, __uint_as_float(v[j + 5]) + b1.y),
                         cvt_relu_pack(__uint_as_float(v[j + 6]) + b1.z, __uint_as_float(v[j + 7]) + b1.w));
        }
      } else if (EPI == EPI_RELU_SPLIT) {   // chunks 0..31 = hi values of the 256 columns, 32..63 = lo values
        uint8_t* rowp = smem + (size_t)row * (ROW_CHUNKS * 16);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 b0 = bt[j / 4], b1 = bt[j / 4 + 1];
          const int chunk = (nl + j) >> 3;
          uint4 hi, lo;
          split_relu_pack(__uint_as_float(v[j]) + b0.x, __uint_as_float(v[j + 1]) + b0.y, hi.x, lo.x);
          split_relu_pack(__uint_as_float(v[j + 2]) + b0.z, __uint_as_float(v[j + 3]) + b0.w, hi.y, lo.y);
          split_relu_pack(__uint_as_float(v[j + 4]) + b1.x, __uint_as_float(v[j + 5]) + b1.y, hi.z, lo.z);
          split_relu_pack(__uint_as_float(v[j + 6]) + b1.z, __uint_as_float(v[j + 7]) + b1.w, hi.w, lo.w);
          *reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4)) = hi;
          *reinterpret_cast<uint4*>(rowp + (((32 + chunk) ^ (row & 7)) << 4)) = lo;
        }
      } else if (EPI == EPI_F32 || EPI == EPI_LRELU_F32) {
        uint8_t* rowp = smem + (size_t)row * (ROW_CHUNKS * 16);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = bt[j / 4];
          const int chunk = (nl + j) >> 2;
          float4 o = make_float4(__uint_as_float(v[j]) + bb.x, __uint_as_float(v[j + 1]) + bb.y, __uint_as_float(v[j + 2]) + bb.z,
                                 __uint_as_float(v[j + 3]) + bb.w);
          if (EPI == EPI_LRELU_F32) {
            o.x = o.x > 0.f ? o.x : 0.01f * o.x; o.y = o.y > 0.f ? o.y : 0.01f * o.y;
            o.z = o.z > 0.f ? o.z : 0.01f * o.z; o.w = o.w > 0.f ? o.w : 0.01f * o.w;
          }
          *reinterpret_cast<float4*>(rowp + ((chunk ^ (row & 7)) << 4)) = o;
        }
      } else {   // POOL
        int keep = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          int bits = (int)v[j];
          bits = m < M ? (bits >= 0 ? bits : (int)(0x80000000u - (uint32_t)bits)) : (int)0x80000000;
          const int mx = __reduce_max_sync(0xffffffffu, bits);
          keep = (tid & 31) == j ? mx : keep;
        }
        red[q][nl + (tid & 31)] = keep;
      }
    }
    __syncthreads();
    if (POOL) {
      const int prob = blockIdx.y * 2 + sub;
      if (prob * 128 < M) {
        const int mi = max(max(red[0][tid], red[1][tid]), max(red[2][tid], red[3][tid]));
        const int bits = mi >= 0 ? mi : (int)(0x80000000u - (uint32_t)mi);
        const float pooled = fmaxf(__int_as_float(bits) + sbias[tid], 0.f);
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)prob * ldc + n0;
        const __nv_bfloat16 phi = __float2bfloat16_rn(pooled);
        o[tid] = phi;
        if (EPI == EPI_MAXPOOL_SPLIT) o[c_lo_off + tid] = __float2bfloat16_rn(pooled - __bfloat162float(phi));
        if (EPI == EPI_MAXPOOL_ARG) { red[0][tid] = mi; red[1][tid] = pooled > 0.f; red[2][tid] = 255; }
      }
      if (EPI == EPI_MAXPOOL_ARG) {
        // training forward: first row attaining the maximum of each live column (the accumulator is still in TMEM);
        // columns pooled to 0 carry no gradient and report row 0
        __syncthreads();
        if (prob * 128 < M) {
#pragma unroll 1
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tl + c0, v);
            tmem_ld_wait();
            const int nl = h * 128 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              int bits = (int)v[j];
              bits = bits >= 0 ? bits : (int)(0x80000000u - (uint32_t)bits);
              if (m < M && red[1][nl + j] && bits == red[0][nl + j]) atomicMin(&red[2][nl + j], row);
            }
          }
        }
        __syncthreads();
        if (prob * 128 < M) arg_out[(size_t)prob * N + n0 + tid] = (uint8_t)(red[2][tid] > 127 ? 0 : red[2][tid]);
      }
    } else {
      // warp w streams rows w, w + 8, ... of this 128-row sub-tile: one row = ROW_CHUNKS x 16 B, a lane per chunk
      const int lane = tid & 31;
      constexpr int ESZ = (EPI == EPI_F32 || EPI == EPI_LRELU_F32) ? 4 : 2;
      const int valid_chunks = min(256, N - n0) * ESZ / 16;   // N tail of the last column tile (N % 8 == 0)
      for (int rr = warp; rr < 128; rr += 8) {
        const int mr = m0 + sub * 128 + rr;
        if (mr >= M) break;
        const uint8_t* rowp = smem + (size_t)rr * (ROW_CHUNKS * 16);
        uint8_t* o = reinterpret_cast<uint8_t*>(Cout) + ((size_t)mr * ldc + n0) * ESZ;
#pragma unroll
        for (int cc = 0; cc < ROW_CHUNKS; cc += 32) {
          const int chunk = cc + lane;
          const uint4 val = *reinterpret_cast<const uint4*>(rowp + ((chunk ^ (rr & 7)) << 4));
          if (EPI == EPI_RELU_SPLIT) {   // chunks 32..63 are the lo half of the row
            const int cidx = chunk & 31;
            if (cidx < valid_chunks) *reinterpret_cast<uint4*>(o + (chunk >= 32 ? (size_t)c_lo_off * 2 : 0) + cidx * 16) = val;
          } else if (chunk < valid_chunks) {
            *reinterpret_cast<uint4*>(o + chunk * 16) = val;
          }
        }
      }
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// [rows][K] bf16 K-major matrix with row pitch `ld` elements -> boxes of 64 (K) x 256 (rows), 128-byte swizzle, zero OOB fill
static int make_tmap(CUtensorMap* tm, const __nv_bfloat16* base, int rows, int K, int ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  MPN_REQUIRE(fn, "gemm_tc: cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)T_BK, 256u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MPN_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%d", (int)r, rows, K, ld);
  return MPN_OK;
}

int* tc_error_flag(mpn_ctx* c);

// split = 0: C = epi(A W^T + bias) with bf16 operands A [M][K] (pitch lda), W [N][K] (pitch ldw).
// split = 1: the bf16x3 product; A and W hold [hi | lo] halves per row: A_lo = A + a_lo_off, W_lo = W + w_lo_off (elements).
int launch_gemm_tc_ex(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, int lda, int a_lo_off, const __nv_bfloat16* W, int ldw,
                      int w_lo_off, int K, const float* bias, int M, int N, void* C, int ldc, int c_lo_off, int split, uint8_t* arg_out) {
  const bool pool = epi == EPI_MAXPOOL || epi == EPI_MAXPOOL_ARG || epi == EPI_MAXPOOL_SPLIT;
  MPN_REQUIRE(K % 16 == 0 && N % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm_tc: K %% 16, N %% 8, lda %% 8, ldw %% 8 required (K=%d N=%d lda=%d)",
              K, N, lda);
  MPN_REQUIRE(!pool || (M % 128 == 0 && N % G_BN == 0), "gemm_tc: max-pool epilogue needs M %% 128 == 0 and N %% 256 == 0");
  MPN_REQUIRE(epi != EPI_MAXPOOL_ARG || arg_out, "gemm_tc: the winning-row epilogue needs an output buffer");
  MPN_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && a_lo_off % 8 == 0 && w_lo_off % 8 == 0,
              "gemm_tc: operands must be 16-byte aligned");
  CUtensorMap tmA, tmW, tmA2, tmW2;
  int r;
  if ((r = make_tmap(&tmA, A, M, K, lda))) return r;
  if ((r = make_tmap(&tmW, W, N, K, ldw))) return r;
  if ((r = make_tmap(&tmA2, A + (split ? a_lo_off : 0), M, K, lda))) return r;
  if ((r = make_tmap(&tmW2, W + (split ? w_lo_off : 0), N, K, ldw))) return r;
  const int passes = split ? 3 : 1;
  dim3 grid_t((N + G_BN - 1) / G_BN, (M + 255) / 256);
  const size_t smem_t = (size_t)T_STAGES * T_STAGE_BYTES + 1024;
  int* errf = tc_error_flag(c);
#define GEMM_TMA(E)                                                                                                       \
  do {                                                                                                                    \
    MPN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));   \
    gemm_tma_kernel<E><<<grid_t, 256, smem_t, s>>>(tmA, tmW, tmA2, tmW2, passes, K, bias, M, N, C, ldc, c_lo_off, errf, arg_out);  \
  } while (0)
  if (epi == EPI_RELU_BF16) GEMM_TMA(EPI_RELU_BF16);
  else if (epi == EPI_F32) GEMM_TMA(EPI_F32);
  else if (epi == EPI_MAXPOOL_ARG) GEMM_TMA(EPI_MAXPOOL_ARG);
  else if (epi == EPI_RELU_SPLIT) GEMM_TMA(EPI_RELU_SPLIT);
  else if (epi == EPI_MAXPOOL_SPLIT) GEMM_TMA(EPI_MAXPOOL_SPLIT);
  else if (epi == EPI_LRELU_F32) GEMM_TMA(EPI_LRELU_F32);
  else GEMM_TMA(EPI_MAXPOOL);
#undef GEMM_TMA
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

int launch_gemm_tc(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int K, const float* bias,
                   int M, int N, void* C, int ldc, uint8_t* arg_out) {
  return launch_gemm_tc_ex(c, s, epi, A, lda, 0, W, K, 0, K, bias, M, N, C, ldc, 0, 0, arg_out);
}

}  // namespace mpn
