// train_tc.cu -- tcgen05 kernels of the training step's set-abstraction backward (bf16 operands, fp32 accumulate in TMEM).
//
// The backward of SA1 / SA2 (train.cu) is a chain of GEMMs over COMPACTED ROWS: millions of rows, 128 columns
// (SA1's 64-wide rows are viewed in pairs as 128-wide rows against block-diagonal weights, so one tile shape serves
// both levels).  At 128 x 128 per row tile these GEMMs are HBM-bound (64 flop per byte), so the kernels are built to keep
// many row tiles in flight per SM rather than to maximise MMA issue:
//
//   rows_gemm_tc_kernel   C[M][N] = epi(A[M][128] W[N][128]^T + bias)   N in {64, 128, 256}
//       W resident in shared memory (no-swizzle K-major core matrices); each warpgroup streams its own 128-row tiles:
//       coalesced 16-byte loads -> core-matrix layout -> 8 tcgen05.mma (K = 128) -> tcgen05.ld epilogue
//       (ReLU / ReLU-derivative mask / plain) -> bf16 rows.  Layers 1-2 recomputed for the active rows, the data
//       gradients dZ1 = (dZ2 W2) * relu'(H1) and dX = dZ1 W1.
//   wgrad_tc_kernel       D[128][128] = sum_r dY[r][:]^T X[r][:]
//       both operands are read in their natural row-major layout and handed to the tensor core as MN-major operands
//       (the reduction index r is the MMA's K): 64-row stages, double buffered, one TMEM accumulator per CTA for its
//       whole row range, fp32 partial tile per CTA, ordered reduction afterwards.
#include <algorithm>

#include "engine.h"
#include "tc_common.cuh"

namespace mpn {
using namespace tc;

int* tc_error_flag(mpn_ctx* c);

namespace {

constexpr int TK = 128, TKC = TK / 8;   // K of every row GEMM, 16-byte chunks per row

__device__ __forceinline__ void wgroup_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// idesc with selectable operand majors (bit 15: A MN-major, bit 16: B MN-major)
__device__ __host__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return make_idesc_bf16(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

enum { EPI_RELU = 0, EPI_PLAIN = 1, EPI_MASK = 2, EPI_POOL = 3 };

template <int N>
struct RowsCfg {
  static constexpr int NWG = N == 256 ? 2 : 4;                 // warpgroups per CTA (N TMEM columns each)
  static constexpr size_t w_bytes = (size_t)N * TK * 2;
  static constexpr size_t a_bytes = (size_t)128 * TK * 2;      // one 128-row tile
  static constexpr size_t red_bytes = (size_t)NWG * 4 * N * 4;   // EPI_POOL: per-warp column maxima (keys)
  static constexpr size_t smem = w_bytes + NWG * a_bytes + N * 4 + 64 + red_bytes;
};

template <int N, int EPI>
__global__ void __launch_bounds__(128 * RowsCfg<N>::NWG, 1)
rows_gemm_tc_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ W, const float* __restrict__ bias,
                    const __nv_bfloat16* mask, long long ntiles, __nv_bfloat16* C, float* __restrict__ pool_out,
                    uint8_t* __restrict__ pool_arg, int paired, int* __restrict__ err, const int* __restrict__ rows_dev, int rows_shift) {
  using Cfg = RowsCfg<N>;
  if (rows_dev) ntiles = (long long)((*rows_dev >> rows_shift) / 128);   // row count decided on the device (compacted active rows, train.cu)
  constexpr int NWG = Cfg::NWG;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + Cfg::w_bytes;
  float* sbias = reinterpret_cast<float*>(sA + NWG * Cfg::a_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + N);
  int* redk = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 64);
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  const int wg = warp >> 2, wq = warp & 3, t = tid & 127;

  // weights: [N][128] bf16 K-major -> core-matrix layout
  for (int i = tid; i < N * TKC; i += blockDim.x) {
    const int r = i / TKC, kc = i - r * TKC;
    *reinterpret_cast<uint4*>(sW + kmajor_chunk_off(r, kc, TKC)) = __ldg(reinterpret_cast<const uint4*>(W + (size_t)r * TK + kc * 8));
  }
  for (int i = tid; i < N; i += blockDim.x) sbias[i] = bias ? bias[i] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < NWG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot + (uint32_t)(wg * N);
  uint8_t* Xs = sA + (size_t)wg * Cfg::a_bytes;
  const uint64_t dA = make_smem_desc(smem_u32(Xs), 128, TKC * 128, LAYOUT_NONE);
  const uint64_t dW = make_smem_desc(smem_u32(sW), 128, TKC * 128, LAYOUT_NONE);
  constexpr uint32_t id = idesc_bf16(128, N, 0, 0);
  uint32_t phase = 0;
  bool ok = true;

  for (long long tile = (long long)blockIdx.x * NWG + wg; tile < ntiles; tile += (long long)gridDim.x * NWG) {
    // ---- A tile: 128 rows x 16 chunks; a warp moves 8 rows x 4 chunks per step (64-byte row segments in, 128-byte runs out)
    const uint4* src = reinterpret_cast<const uint4*>(A + (size_t)tile * 128 * TK);
    uint4 v[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int r = wq * 32 + (it >> 2) * 8 + (lane & 7), kc = (it & 3) * 4 + (lane >> 3);
      v[it] = __ldg(src + (size_t)r * TKC + kc);
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
      const int r = wq * 32 + (it >> 2) * 8 + (lane & 7), kc = (it & 3) * 4 + (lane >> 3);
      *reinterpret_cast<uint4*>(Xs + kmajor_chunk_off(r, kc, TKC)) = v[it];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    wgroup_sync(wg);
    if (wq == 0) {
      if (elect_one()) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < TK / 16; ++ks) mma_bf16_ss_off(tmem, dA, ks * 16, dW, ks * 16, id, ks != 0);
        mma_commit(&bars[wg]);
      }
      __syncwarp();
    }
    ok = mbar_wait(&bars[wg], phase) && ok;
    phase ^= 1u;
    tc_fence_after();
    if (EPI == EPI_POOL) {
      // ---- ReLU + max-pool over the tile's rows with the winning row (max_pool2d's argmax, first occurrence): the maximum
      // is taken on order-preserving integer keys whose low 7 bits carry (127 - row), so one redux.sync per column gives
      // value (to 2^-16 relative) and row at once.  Unpaired: tile = one group of 128 neighbour rows.  Paired (SA1): a tile
      // row holds neighbours 2t (columns 0..63) and 2t + 1 (columns 64..127) and the tile spans two groups of 64 paired rows.
      int* rk = redk + (size_t)(wg * 4 + wq) * N;
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t a[32];
        tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + c0, a);
        tmem_ld_wait();
        const int rowcode = paired ? 2 * (t & 63) + (c0 >= 64 ? 1 : 0) : t;
        const uint32_t low = (uint32_t)(127 - rowcode);
        int keep = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int bits = __float_as_int(__uint_as_float(a[j]) + sbias[c0 + j]);
          const int key = bits > 0 ? (int)(((uint32_t)bits & 0xFFFFFF80u) | low) : 0;
          const int mx = __reduce_max_sync(0xffffffffu, key);
          keep = lane == j ? mx : keep;
        }
        rk[c0 + lane] = keep;
      }
      tc_fence_before();
      wgroup_sync(wg);
      const int* r4 = redk + (size_t)(wg * 4) * N;
      if (!paired) {
        for (int cc = t; cc < N; cc += 128) {
          const int k = max(max(r4[cc], r4[N + cc]), max(r4[2 * N + cc], r4[3 * N + cc]));
          pool_out[(size_t)tile * N + cc] = __int_as_float(k & (int)0xFFFFFF80);
          pool_arg[(size_t)tile * N + cc] = (uint8_t)(127 - (k & 127));
        }
      } else {
        const int gh = t >> 6, ch = t & 63;
        const int* ra = r4 + (size_t)(2 * gh) * N;
        const int k = max(max(ra[ch], ra[N + ch]), max(ra[ch + 64], ra[N + ch + 64]));
        pool_out[((size_t)tile * 2 + gh) * 64 + ch] = __int_as_float(k & (int)0xFFFFFF80);
        pool_arg[((size_t)tile * 2 + gh) * 64 + ch] = (uint8_t)(127 - (k & 127));
      }
      continue;
    }
    // ---- epilogue: thread = output row
    const size_t row = (size_t)tile * 128 + t;
    __nv_bfloat16* crow = C + row * N;
    const __nv_bfloat16* mrow = EPI == EPI_MASK ? mask + row * N : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t a[32];
      tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + c0, a);
      tmem_ld_wait();
      uint4 o[4];
      uint32_t* ow = reinterpret_cast<uint32_t*>(o);
      if (EPI == EPI_MASK) {
        uint4 m4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) m4[q] = *reinterpret_cast<const uint4*>(mrow + c0 + q * 8);
        const uint32_t* mw = reinterpret_cast<const uint32_t*>(m4);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // bf16 pair of the mask: element > 0  <=>  sign bit clear and not zero
          const uint32_t m = mw[j];
          const float lo = ((m & 0x7FFFu) != 0u && (m & 0x8000u) == 0u) ? __uint_as_float(a[2 * j]) : 0.f;
          const float hi = ((m & 0x7FFF0000u) != 0u && (m & 0x80000000u) == 0u) ? __uint_as_float(a[2 * j + 1]) : 0.f;
          ow[j] = pack_bf16(lo, hi);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float lo = __uint_as_float(a[2 * j]) + sbias[c0 + 2 * j], hi = __uint_as_float(a[2 * j + 1]) + sbias[c0 + 2 * j + 1];
          if (EPI == EPI_RELU) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
          ow[j] = pack_bf16(lo, hi);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(crow + c0 + q * 8) = o[q];
    }
    tc_fence_before();
  }
  if (!ok && lane == 0) atomicExch(err, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------- weight gradient
constexpr int WG_ROWS = 64;                                   // rows (= MMA K extent) per stage
constexpr int WG_SBO = WG_ROWS * 16 + 32;                     // byte stride between 8-column chunks (+32: bank spread)
constexpr int WG_TILE = 16 * WG_SBO;                          // one operand stage: 16 column chunks

// smem offset of the 16-byte chunk (row r of the stage, columns [8 c, 8 c + 8)): MN-major core matrices = 8 rows x 16 B
__device__ __forceinline__ uint32_t mnmajor_chunk_off(int r, int c) { return (uint32_t)(c * WG_SBO + r * 16); }

__global__ void __launch_bounds__(256, 2)
wgrad_tc_kernel(const __nv_bfloat16* __restrict__ dY, const __nv_bfloat16* __restrict__ X, long long R, long long rows_per_cta,
                float* __restrict__ partial, int swap_lbo_sbo, int* __restrict__ err, const int* __restrict__ rows_dev, int rows_shift,
                int ldy, int ldx, int nx_tiles, int x_cols) {
  // blockIdx.y = (ty, tx): the 128 x 128 tile dY[:, 128 ty ..]^T X[:, 128 tx ..] of a wider product (row pitches ldy / ldx elements,
  // X columns >= x_cols read as zero); the plain 128-wide call has ldy = ldx = x_cols = 128 and one tile
  extern __shared__ __align__(1024) uint8_t smem[];           // [2 stages][dY tile | X tile]
  const int ty = blockIdx.y / nx_tiles, tx = blockIdx.y - ty * nx_tiles;
  dY += (size_t)ty * 128;
  X += (size_t)tx * 128;
  const int x_chunks = max(0, min(16, (x_cols - tx * 128) / 8));   // valid 8-column chunks of this X tile
  if (rows_dev) {   // row count decided on the device: the launch's CTAs share the rows evenly, in whole stages
    R = (long long)(*rows_dev >> rows_shift);
    rows_per_cta = ((R + gridDim.x - 1) / gridDim.x + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  }
  __shared__ uint64_t empty[2], accum;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  const long long r_begin = (long long)blockIdx.x * rows_per_cta, r_end = min(R, r_begin + rows_per_cta);
  const int nst = (int)((r_end - r_begin + WG_ROWS - 1) / WG_ROWS);
  if (tid == 0) {
    mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); mbar_init(&accum, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // MN-major, no swizzle: K (= row) groups of 8 are 128 B apart, column chunks WG_SBO apart
  const uint32_t lbo = swap_lbo_sbo ? WG_SBO : 128, sbo = swap_lbo_sbo ? 128 : WG_SBO;
  constexpr uint32_t id = idesc_bf16(128, 128, 1, 1);
  bool ok = true;
  for (int it = 0; it < nst; ++it) {
    const int s = it & 1;
    uint8_t* sY = smem + (size_t)s * 2 * WG_TILE;
    uint8_t* sX = sY + WG_TILE;
    if (it >= 2) { ok = mbar_wait(&empty[s], ((it >> 1) - 1) & 1) && ok; tc_fence_after(); }
    const long long r0 = r_begin + (long long)it * WG_ROWS;
    // 64 rows x 16 chunks per operand; a warp moves 8 rows x 4 chunks per step
    uint4 vy[4], vx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = warp * 8 + (lane & 7), c = u * 4 + (lane >> 3);
      const bool in = r0 + r < r_end;
      vy[u] = in ? __ldg(reinterpret_cast<const uint4*>(dY + (size_t)(r0 + r) * ldy) + c) : make_uint4(0u, 0u, 0u, 0u);
      vx[u] = (in && c < x_chunks) ? __ldg(reinterpret_cast<const uint4*>(X + (size_t)(r0 + r) * ldx) + c) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = warp * 8 + (lane & 7), c = u * 4 + (lane >> 3);
      *reinterpret_cast<uint4*>(sY + mnmajor_chunk_off(r, c)) = vy[u];
      *reinterpret_cast<uint4*>(sX + mnmajor_chunk_off(r, c)) = vx[u];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        const uint64_t dA = make_smem_desc(smem_u32(sY), lbo, sbo, LAYOUT_NONE);
        const uint64_t dB = make_smem_desc(smem_u32(sX), lbo, sbo, LAYOUT_NONE);
#pragma unroll
        for (int ks = 0; ks < WG_ROWS / 16; ++ks) mma_bf16_ss_off(tmem, dA, ks * 16, dB, ks * 16, id, (it | ks) != 0);   // 16 rows = 256 B
        mma_commit(&empty[s]);
        if (it == nst - 1) mma_commit(&accum);
      }
      __syncwarp();
    }
  }
  if (nst > 0) {
    ok = mbar_wait(&accum, 0) && ok;
    tc_fence_after();
  }
  if (!ok && lane == 0) atomicExch(err, 1);
  if (warp < 4) {
    float* prow = partial + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 128 + tid) * 128;
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t a[32];
      if (nst > 0) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, a);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<uint4*>(prow + c0 + j) = make_uint4(a[j], a[j + 1], a[j + 2], a[j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace

// C[M][N] (bf16) = epi(A[M][128] W[N][128]^T + bias); M % 128 == 0; epi: 0 relu, 1 plain, 2 multiply by relu'(mask[M][N]).
// rows_dev != null: M is only the capacity (grid sizing); the kernel processes (*rows_dev >> rows_shift) / 128 tiles
int launch_rows_gemm_tc(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, const __nv_bfloat16* W, const float* bias,
                        const __nv_bfloat16* mask, long long M, int N, __nv_bfloat16* C, float* pool_out, uint8_t* pool_arg,
                        int paired, const int* rows_dev, int rows_shift) {
  MPN_REQUIRE(M % 128 == 0 && (N == 64 || N == 128 || N == 256), "rows_gemm_tc: M %% 128 == 0 and N in {64,128,256} required");
  MPN_REQUIRE(epi != EPI_MASK || mask, "rows_gemm_tc: mask epilogue without a mask");
  MPN_REQUIRE(epi != EPI_POOL || (pool_out && pool_arg && (paired ? N == 128 : N >= 128)), "rows_gemm_tc: bad pool arguments");
  if (M == 0) return MPN_OK;
  const long long ntiles = M / 128;
  int* errf = tc_error_flag(c);
#define ROWS_LAUNCH(NN, EE)                                                                                              \
  do {                                                                                                                   \
    auto k = rows_gemm_tc_kernel<NN, EE>;                                                                                \
    const size_t smem = RowsCfg<NN>::smem;                                                                               \
    MPN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                     \
    const long long want = (ntiles + RowsCfg<NN>::NWG - 1) / RowsCfg<NN>::NWG;                                           \
    const int grid = (int)std::min<long long>(want, c->sm_count);                                                        \
    k<<<grid, 128 * RowsCfg<NN>::NWG, smem, s>>>(A, W, bias, mask, ntiles, C, pool_out, pool_arg, paired, errf, rows_dev, rows_shift); \
  } while (0)
#define ROWS_EPI(NN)                                       \
  do {                                                     \
    if (epi == EPI_RELU) ROWS_LAUNCH(NN, EPI_RELU);        \
    else if (epi == EPI_PLAIN) ROWS_LAUNCH(NN, EPI_PLAIN); \
    else if (epi == EPI_MASK) ROWS_LAUNCH(NN, EPI_MASK);   \
    else ROWS_LAUNCH(NN, EPI_POOL);                        \
  } while (0)
  if (N == 64) ROWS_EPI(64);
  else if (N == 128) ROWS_EPI(128);
  else ROWS_EPI(256);
#undef ROWS_EPI
#undef ROWS_LAUNCH
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// partial[ctas][128][128] (fp32) = per-CTA sums of dY[r][:]^T X[r][:] over the CTA's row range; returns the CTA count
int launch_wgrad_tc(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* dY, const __nv_bfloat16* X, long long R, float* partial,
                    size_t partial_floats, int* n_ctas, int swap_lbo_sbo, const int* rows_dev, int rows_shift) {
  long long ctas = std::min<long long>(2LL * c->sm_count, (R + 4 * WG_ROWS - 1) / (4 * WG_ROWS));
  ctas = std::max(1LL, std::min<long long>(ctas, (long long)(partial_floats / (128 * 128))));
  long long rows_per_cta = ((R + ctas - 1) / ctas + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  ctas = std::max(1LL, (R + rows_per_cta - 1) / rows_per_cta);
  const size_t smem = (size_t)2 * 2 * WG_TILE;
  MPN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_tc_kernel<<<(unsigned)ctas, 256, smem, s>>>(dY, X, R, rows_per_cta, partial, swap_lbo_sbo, tc_error_flag(c), rows_dev, rows_shift,
                                                    128, 128, 1, 128);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  *n_ctas = (int)ctas;
  return MPN_OK;
}

// wide product: partial[ty * nx + tx][cta][128][128] = per-CTA sums of dY[r][128 ty ..]^T X[r][128 tx ..] over the CTA's rows; dY [R][ldy]
// with y_cols (multiple of 128) columns used, X [R][ldx] with x_cols (multiple of 8) columns used.  Returns the row-CTA count.
int launch_wgrad_tc2d(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* dY, int ldy, int y_cols, const __nv_bfloat16* X, int ldx, int x_cols,
                      long long R, float* partial, size_t partial_floats, int* n_ctas) {
  MPN_REQUIRE(y_cols % 128 == 0 && x_cols % 8 == 0 && ldy % 8 == 0 && ldx % 8 == 0 && R >= 1, "wgrad_tc2d: bad shapes");
  const int ny = y_cols / 128, nx = (x_cols + 127) / 128;
  long long ctas = std::max(1LL, std::min<long long>(R / (8 * WG_ROWS), (2LL * c->sm_count + ny * nx - 1) / (ny * nx)));
  ctas = std::max(1LL, std::min<long long>(ctas, (long long)(partial_floats / ((size_t)128 * 128 * ny * nx))));
  long long rows_per_cta = ((R + ctas - 1) / ctas + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  ctas = std::max(1LL, (R + rows_per_cta - 1) / rows_per_cta);
  const size_t smem = (size_t)2 * 2 * WG_TILE;
  MPN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wgrad_tc_kernel<<<dim3((unsigned)ctas, (unsigned)(ny * nx)), 256, smem, s>>>(dY, X, R, rows_per_cta, partial, 0, tc_error_flag(c), nullptr, 0, ldy,
                                                                              ldx, nx, x_cols);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  *n_ctas = (int)ctas;
  return MPN_OK;
}

}  // namespace mpn
