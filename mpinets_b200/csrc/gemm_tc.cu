// gemm_tc.cu -- tcgen05 GEMM for the row-shared layers of the encoder: the group-all module (SA3, model.py:383: rows =
// B*128 points, 259->512->512->1024 + max over each problem's 128 rows) and the FC head (model.py:385-393).
//
//   C[M][N] = epilogue(A[M][K] * W[N][K]^T + bias)     A, W bf16 K-major in HBM, fp32 accumulate in TMEM.
//   CTA tile 256 x 256 (two 128-row MMA sub-tiles that share every weight stage), K staged 64 at a time through a 3-deep
//   cp.async ring into the UMMA interleaved (8x16B core matrix) layout; one thread issues tcgen05.mma, completion per stage is tracked with tcgen05.commit -> mbarrier so a
//   ring slot is only refilled after the MMAs that read it have retired; all 8 warps drain the 128x256 accumulator.
//   EPI_RELU_BF16: relu -> bf16 rows;  EPI_F32: fp32 rows (GroupNorm follows);  EPI_MAXPOOL: relu + max over the tile's
//   128 rows (= one problem) -> one bf16 row.
#include <cuda.h>

#include "engine.h"
#include "tc_common.cuh"

namespace mpn {
using namespace tc;

enum { EPI_RELU_BF16 = 0, EPI_F32 = 1, EPI_MAXPOOL = 2 };
constexpr int G_BN = 256;   // CTA tile = G_BM x 256 with G_BM = 128 or 256 (one or two 128-row MMA tiles sharing every weight stage)
// K staging is a template parameter pair: (G_BK, G_STAGES) = (64, 3) or (32, 6) -- the same 192 KB ring, but the deeper
// ring keeps five stages of loads in flight behind the one being multiplied instead of two.

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t cvt_relu_pack(float first, float second) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(second), "f"(first));
  return d;
}

template <int EPI, int G_BM, int G_BK, int G_STAGES>
__global__ void __launch_bounds__(256, G_BM == 128 ? 2 : 1)
gemm_tc_kernel(const __nv_bfloat16* __restrict__ A, int lda, const __nv_bfloat16* __restrict__ W, int K, const float* __restrict__ bias,
               int M, int N, void* __restrict__ Cout, int ldc, int* __restrict__ err, long long* __restrict__ tl) {
  constexpr int G_KC = G_BK / 8, G_KS = G_BK / 16;   // 16-byte chunks / MMA K-steps per stage row
  constexpr int G_A_BYTES = G_BM * G_BK * 2, G_W_BYTES = G_BN * G_BK * 2, G_STAGE_BYTES = G_A_BYTES + G_W_BYTES;
  static_assert(G_KC == 4 || G_KC == 8, "stage K must be 32 or 64");
  extern __shared__ __align__(1024) uint8_t smem[];   // (the no-swizzle operand layout only needs 16-byte alignment)
  __shared__ uint64_t done[G_STAGES];
  __shared__ uint32_t tmem_slot;
  __shared__ int red[4][G_BN];
  __shared__ __align__(16) float sbias[G_BN];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // provably warp-uniform
  const int m0 = blockIdx.y * G_BM, n0 = blockIdx.x * G_BN;
  const int nst = (K + G_BK - 1) / G_BK;
  long long tl_prev = clock64();
  const bool tl_on = tl != nullptr && M > 65536 && blockIdx.x == 0 && blockIdx.y == 64 && threadIdx.x == 0;   // a mid-grid tile of the big row GEMMs
#define GT_MARK(i) do { if (tl_on) { long long _n = clock64(); tl[32 + (i)] += _n - tl_prev; tl_prev = _n; } } while (0)

  sbias[tid] = bias[n0 + tid];   // 256 threads == G_BN columns
  if (tid == 0) {
    for (int s = 0; s < G_STAGES; ++s) mbar_init(&done[s], 1);
    mbar_fence_init();
  }
  constexpr int G_SUBS = G_BM / 128;
  if (warp == 0) tmem_alloc(&tmem_slot, 256 * G_SUBS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint64_t dA0 = make_smem_desc(smem_u32(smem), 128, G_KC * 128, LAYOUT_NONE);                 // stage 0 descriptors; later stages /
  const uint64_t dW0 = make_smem_desc(smem_u32(smem) + G_A_BYTES, 128, G_KC * 128, LAYOUT_NONE);     // K-steps only add to the address field

  // stage loader: A rows m0.., W rows n0.., K chunk [st*64, st*64+64) -> interleaved layout with 8 chunks per row.
  // Lane mapping: each quarter-warp (the unit the 16-byte shared-memory write is served in) covers 8 consecutive rows of
  // one chunk column -> 8 distinct 16-byte bank groups (a row-major lane order put all 8 lanes on the same group: an
  // 8-way conflict on every LDGSTS write); the warp as a whole reads 8 rows x 64 contiguous bytes, full 32-byte sectors.
  auto load_stage = [&](int st) {
    uint8_t* sA = smem + (size_t)(st % G_STAGES) * G_STAGE_BYTES;
    uint8_t* sW = sA + G_A_BYTES;
    const int k0 = st * G_BK;
    const int kc_n = min(G_KC, (K - k0) / 8);
#pragma unroll
    for (int i = 0; i < (G_BM * G_KC) / 256; ++i) {
      const int c = tid + i * 256;
      const int r = G_KC == 8 ? (((c >> 6) << 3) | (c & 7)) : (((c >> 5) << 3) | (c & 7));
      const int kc = G_KC == 8 ? ((((c >> 5) & 1) << 2) | ((c >> 3) & 3)) : ((c >> 3) & 3);
      if (kc < kc_n) {
        int m = m0 + r;
        const __nv_bfloat16* src = A + (size_t)min(m, M - 1) * lda + k0 + kc * 8;
        cp_async16(smem_u32(sA + kmajor_chunk_off(r, kc, G_KC)), src, m < M ? 16u : 0u);
      }
    }
#pragma unroll
    for (int i = 0; i < (G_BN * G_KC) / 256; ++i) {
      const int c = tid + i * 256;
      const int r = G_KC == 8 ? (((c >> 6) << 3) | (c & 7)) : (((c >> 5) << 3) | (c & 7));
      const int kc = G_KC == 8 ? ((((c >> 5) & 1) << 2) | ((c >> 3) & 3)) : ((c >> 3) & 3);
      if (kc < kc_n) cp_async16(smem_u32(sW + kmajor_chunk_off(r, kc, G_KC)), W + (size_t)(n0 + r) * K + k0 + kc * 8, 16u);
    }
  };

  for (int s = 0; s < G_STAGES - 1; ++s) {
    if (s < nst) load_stage(s);
    cp_async_commit();
  }
  bool ok = true;
  GT_MARK(0);   // prologue: barriers, TMEM alloc, first loads issued
  for (int it = 0; it < nst; ++it) {
    cp_async_wait<G_STAGES - 2>();
    GT_MARK(1);   // waiting for this stage's loads
    fence_proxy_async_smem();
    __syncthreads();
    GT_MARK(2);   // fence + block sync
    if (warp == 0) {
      tc_fence_after();
      uint32_t el;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el));
      if (el) {
      const uint32_t soff = (uint32_t)(it % G_STAGES) * (G_STAGE_BYTES / 16);
      const int ksteps = min(G_KS, (K - it * G_BK) / 16);
      constexpr uint32_t id = make_idesc_bf16(128, G_BN);
      constexpr uint32_t SUB1 = (128 / 8) * G_KC * 128 / 16;   // rows 128..255 of the A stage, in 16-byte units
      if (ksteps == G_KS) {
#pragma unroll
        for (int ks = 0; ks < G_KS; ++ks) mma_bf16_ss_off(tmem, dA0, soff + ks * 16, dW0, soff + ks * 16, id, (it | ks) != 0);
#pragma unroll
        for (int ks = 0; ks < G_KS; ++ks)
          if (G_SUBS == 2) mma_bf16_ss_off(tmem + 256, dA0, soff + SUB1 + ks * 16, dW0, soff + ks * 16, id, (it | ks) != 0);
      } else {
        for (int ks = 0; ks < ksteps; ++ks) mma_bf16_ss_off(tmem, dA0, soff + ks * 16, dW0, soff + ks * 16, id, (it | ks) != 0);
        for (int ks = 0; ks < ksteps; ++ks)
          if (G_SUBS == 2) mma_bf16_ss_off(tmem + 256, dA0, soff + SUB1 + ks * 16, dW0, soff + ks * 16, id, (it | ks) != 0);
      }
      mma_commit(&done[it % G_STAGES]);
      }
      __syncwarp();
    }
    GT_MARK(3);   // MMA issue (thread 0 is the warp-0 lane that may be elected)
    const int nxt = it + G_STAGES - 1;
    if (nxt < nst) {
      if (it >= 1) ok = ok && mbar_wait(&done[(it - 1) % G_STAGES], ((it - 1) / G_STAGES) & 1);   // slot of stage it-1 is free
      GT_MARK(4);   // waiting for the previous stage's MMAs
      load_stage(nxt);
      GT_MARK(5);   // issuing the next loads
    }
    cp_async_commit();
  }
  ok = ok && mbar_wait(&done[(nst - 1) % G_STAGES], ((nst - 1) / G_STAGES) & 1);
  tc_fence_after();
  GT_MARK(6);   // drain: last MMAs
  if (!ok && tid == 0) atomicExch(err, 1);

  // ---- epilogue: warp (q = warp & 3) owns lanes 32q..32q+31, column half h = warp >> 2.  The bias tile sits in shared
  // memory (staged at kernel start) and is read with 16-byte broadcast loads.
  const int q = warp & 3, h = warp >> 2, row = q * 32 + (tid & 31);
#pragma unroll 1
  for (int sub = 0; sub < G_SUBS; ++sub) {
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + sub * 256 + h * 128;
    const int m = m0 + sub * 128 + row;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tl + c0, v);
      tmem_ld_wait();
      const int nl = h * 128 + c0;          // column inside the CTA tile
      const float4* bt = reinterpret_cast<const float4*>(sbias + nl);
      if (EPI == EPI_RELU_BF16) {
        if (m < M) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)m * ldc + n0 + nl;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 b0 = bt[j / 4], b1 = bt[j / 4 + 1];
            *reinterpret_cast<uint4*>(o + j) =
                make_uint4(cvt_relu_pack(__uint_as_float(v[j]) + b0.x, __uint_as_float(v[j + 1]) + b0.y),
                           cvt_relu_pack(__uint_as_float(v[j + 2]) + b0.z, __uint_as_float(v[j + 3]) + b0.w),
                           cvt_relu_pack(__uint_as_float(v[j + 4]) + b1.x, __uint_as_float(v[j + 5]) + b1.y),
                           cvt_relu_pack(__uint_as_float(v[j + 6]) + b1.z, __uint_as_float(v[j + 7]) + b1.w));
          }
        }
      } else if (EPI == EPI_F32) {
        if (m < M) {
          float* o = reinterpret_cast<float*>(Cout) + (size_t)m * ldc + n0 + nl;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = bt[j / 4];
            *reinterpret_cast<float4*>(o + j) = make_float4(__uint_as_float(v[j]) + bb.x, __uint_as_float(v[j + 1]) + bb.y,
                                                            __uint_as_float(v[j + 2]) + bb.z, __uint_as_float(v[j + 3]) + bb.w);
          }
        }
      } else {
        // max over the tile's 128 rows: relu(max_r acc + bias) == max_r relu(acc + bias); the raw accumulators are
        // reduced as order-preserving integers (sign-magnitude floats -> two's complement order)
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
    if (EPI == EPI_MAXPOOL) {
      __syncthreads();
      const int prob = blockIdx.y * G_SUBS + sub;   // one pooled row per 128-row sub-tile (= one problem)
      if (prob * 128 < M) {
        const int mi = max(max(red[0][tid], red[1][tid]), max(red[2][tid], red[3][tid]));
        const int bits = mi >= 0 ? mi : (int)(0x80000000u - (uint32_t)mi);
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)prob * ldc + n0;
        o[tid] = __float2bfloat16_rn(fmaxf(__int_as_float(bits) + sbias[tid], 0.f));
      }
      __syncthreads();
    }
  }
  GT_MARK(7);   // epilogue
  if (tl_on) tl[32 + 8] += 1;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256 * G_SUBS);
#undef GT_MARK
}

// ------------------------------------------------------------------------------------------------ TMA-fed variant
// Same tile (256 x 256, two 128-row MMA sub-tiles), but the operands arrive through the tensor memory accelerator:
// per K stage of 64 one elected thread issues two cp.async.bulk.tensor loads (A box 64 x 256 rows, W box 64 x 256 rows,
// 128-byte swizzle) that complete on the stage's `full` mbarrier; a second elected thread waits on it, issues the eight
// tcgen05.mma of the stage and commits to the stage's `empty` mbarrier, which the producer waits on before refilling.
// No block-wide sync inside the main loop, and L2 -> shared memory moves whole 128-byte rows: cp.async.cg issued one
// 32-byte sector request per 16-byte lane copy (ncu: 27 sectors per LDGSTS instruction, 2x the operand bytes over the
// crossbar), which is what bounded the cp.async variant above.
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

template <int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int K, const float* __restrict__ bias,
                int M, int N, void* __restrict__ Cout, int ldc, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];   // SWIZZLE_128B tiles: 1024-byte aligned stage bases
  __shared__ uint64_t full[T_STAGES], empty[T_STAGES], accum;
  __shared__ uint32_t tmem_slot;
  __shared__ int red[4][G_BN];
  __shared__ __align__(16) float sbias[G_BN];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int m0 = blockIdx.y * 256, n0 = blockIdx.x * G_BN;
  const int nst = (K + T_BK - 1) / T_BK;

  sbias[tid] = bias[n0 + tid];
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
      for (int it = 0; it < nst; ++it) {
        const int slot = it % T_STAGES;
        if (it >= T_STAGES) ok = ok && mbar_wait(&empty[slot], ((it / T_STAGES) - 1) & 1);
        const uint32_t sA = smem_u32(smem) + slot * T_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[slot], T_STAGE_BYTES);
        tma_load_2d(sA, &tmA, it * T_BK, m0, &full[slot]);
        tma_load_2d(sA + T_A_BYTES, &tmW, it * T_BK, n0, &full[slot]);
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
      for (int it = 0; it < nst; ++it) {
        const int slot = it % T_STAGES;
        ok = ok && mbar_wait(&full[slot], (it / T_STAGES) & 1);
        tc_fence_after();
        const uint32_t soff = (uint32_t)slot * (T_STAGE_BYTES / 16);
        const int ksteps = min(T_BK / 16, (K - it * T_BK + 15) / 16);   // the K tail beyond the tensor is zero-filled by TMA
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
  constexpr int ROW_CHUNKS = EPI == EPI_F32 ? 64 : 32;   // 16-byte chunks per 256-column output row
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
      } else if (EPI == EPI_F32) {
        uint8_t* rowp = smem + (size_t)row * (ROW_CHUNKS * 16);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = bt[j / 4];
          const int chunk = (nl + j) >> 2;
          *reinterpret_cast<float4*>(rowp + ((chunk ^ (row & 7)) << 4)) =
              make_float4(__uint_as_float(v[j]) + bb.x, __uint_as_float(v[j + 1]) + bb.y, __uint_as_float(v[j + 2]) + bb.z,
                          __uint_as_float(v[j + 3]) + bb.w);
        }
      } else {
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
    if (EPI == EPI_MAXPOOL) {
      const int prob = blockIdx.y * 2 + sub;
      if (prob * 128 < M) {
        const int mi = max(max(red[0][tid], red[1][tid]), max(red[2][tid], red[3][tid]));
        const int bits = mi >= 0 ? mi : (int)(0x80000000u - (uint32_t)mi);
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(Cout) + (size_t)prob * ldc + n0;
        o[tid] = __float2bfloat16_rn(fmaxf(__int_as_float(bits) + sbias[tid], 0.f));
      }
    } else {
      // warp w streams rows w, w + 8, ... of this 128-row sub-tile: one row = ROW_CHUNKS x 16 B, a lane per chunk
      const int lane = tid & 31;
      constexpr int ESZ = EPI == EPI_F32 ? 4 : 2;
      for (int rr = warp; rr < 128; rr += 8) {
        const int mr = m0 + sub * 128 + rr;
        if (mr >= M) break;
        const uint8_t* rowp = smem + (size_t)rr * (ROW_CHUNKS * 16);
        uint8_t* o = reinterpret_cast<uint8_t*>(Cout) + ((size_t)mr * ldc + n0) * ESZ;
#pragma unroll
        for (int cc = 0; cc < ROW_CHUNKS; cc += 32) {
          const int chunk = cc + lane;
          *reinterpret_cast<uint4*>(o + chunk * 16) = *reinterpret_cast<const uint4*>(rowp + ((chunk ^ (rr & 7)) << 4));
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
long long* tc_timeline(mpn_ctx* c);

int launch_gemm_tc(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int K, const float* bias,
                   int M, int N, void* C, int ldc) {
  MPN_REQUIRE(K % 16 == 0 && N % G_BN == 0 && lda % 8 == 0, "gemm_tc: K %% 16, N %% 256, lda %% 8 required (K=%d N=%d lda=%d)", K, N, lda);
  MPN_REQUIRE(epi != EPI_MAXPOOL || M % 128 == 0, "gemm_tc: max-pool epilogue needs M %% 128 == 0");
  // cp.async variants, kept for A/B runs (MPN_GEMM_TILE=128: 128 x 256 tiles, 4 stages of K=32, two CTAs per SM;
  // =256: 256 x 256, 6 stages of K=32; =25664: 256 x 256, 3 stages of K=64).  All measured slower than the TMA kernel:
  // they are bound by the sector-per-lane crossbar traffic of LDGSTS (the 128-row tile, which needs 1.5x the operand
  // bytes per flop, is 1.4x slower than the 256-row one).
  static const bool use_cp_async = getenv("MPN_GEMM_TILE") != nullptr;   // default: the TMA-fed kernel
  if (!use_cp_async) {
    MPN_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, "gemm_tc: operands must be 16-byte aligned");
    CUtensorMap tmA, tmW;
    int r;
    if ((r = make_tmap(&tmA, A, M, K, lda))) return r;
    if ((r = make_tmap(&tmW, W, N, K, K))) return r;
    dim3 grid_t(N / G_BN, (M + 255) / 256);
    const size_t smem_t = (size_t)T_STAGES * T_STAGE_BYTES + 1024;
    int* errf = tc_error_flag(c);
#define GEMM_TMA(E)                                                                                                       \
  do {                                                                                                                    \
    MPN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tma_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));   \
    gemm_tma_kernel<E><<<grid_t, 256, smem_t, s>>>(tmA, tmW, K, bias, M, N, C, ldc, errf);                                \
  } while (0)
    if (epi == EPI_RELU_BF16) GEMM_TMA(EPI_RELU_BF16);
    else if (epi == EPI_F32) GEMM_TMA(EPI_F32);
    else GEMM_TMA(EPI_MAXPOOL);
#undef GEMM_TMA
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    return MPN_OK;
  }
  static const int tile = atoi(getenv("MPN_GEMM_TILE"));
  const int bm = tile == 128 ? 128 : 256;
  dim3 grid(N / G_BN, (M + bm - 1) / bm);
  int* err = tc_error_flag(c);
#define GEMM_LAUNCH_T(E, BM, BK, ST)                                                                                            \
  do {                                                                                                                          \
    const size_t smem = (size_t)ST * (BM + G_BN) * BK * 2 + 1024;                                                               \
    MPN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<E, BM, BK, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gemm_tc_kernel<E, BM, BK, ST><<<grid, 256, smem, s>>>(A, lda, W, K, bias, M, N, C, ldc, err, tc_timeline(c));               \
  } while (0)
#define GEMM_LAUNCH(E)                                \
  do {                                                \
    if (tile == 128) GEMM_LAUNCH_T(E, 128, 32, 4);    \
    else if (tile == 25664) GEMM_LAUNCH_T(E, 256, 64, 3); \
    else GEMM_LAUNCH_T(E, 256, 32, 6);                \
  } while (0)
  if (epi == EPI_RELU_BF16) GEMM_LAUNCH(EPI_RELU_BF16);
  else if (epi == EPI_F32) GEMM_LAUNCH(EPI_F32);
  else GEMM_LAUNCH(EPI_MAXPOOL);
#undef GEMM_LAUNCH_T
#undef GEMM_LAUNCH
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
