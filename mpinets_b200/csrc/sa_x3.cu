// sa_x3.cu -- the PARITY-GRADE tensor-core mode of the PointNet++ encoder (MPN_PREC_BF16X3): every contraction of
// mpinets/model.py:365-393 (the shared MLPs inside the three PointnetSAModules and the FC head) runs on tcgen05 with
// SPLIT-bf16 operands and fp32 accumulation in tensor memory:
//
//     x = x_hi + x_lo (+ 2^-16 x),  x_hi = bf16(x),  x_lo = bf16(x - x_hi)          (activations AND weights)
//     a * w  ~=  a_hi w_hi + a_lo w_hi + a_hi w_lo                                   (three bf16 MMAs, one fp32 accumulator)
//
// which carries ~16 mantissa bits per operand instead of bf16's 8 -- enough for delta-q within 1e-5 of the fp32 reference
// (north_star's tolerance; measured ~1e-6) at 3x the MMA count of the bf16 throughput mode instead of the ~25x slower fp32
// SIMT path.  Index outputs (FPS, ball query) come from the same bit-exact code as in the other modes.
//
//   sa1x3_tc_kernel   SA1 (model.py:365-373): the structure of sa1t_tc_kernel (sa_tc.cu) -- CTA = problem, cloud + hash grid in
//                     shared memory, activations in TENSOR MEMORY as the A operand -- with hi / lo operand columns per chain
//                     (64 accumulator + 32 + 32 operand columns -> 4 chains) and fp32 max-pooling.
//   sa2x3_tc_kernel   SA2 (model.py:374-382).  hi + lo weights of all three layers (232 KB) do not fit in shared memory, so
//                     layer 1 is evaluated per POINT instead of per (centroid, neighbour) pair: W1 [p_k - c_i ; f_k] + b =
//                     (W1f f_k + W1x p_k + b) - W1x c_i.  The bracket is one small split-bf16 GEMM over the 512 points of a
//                     problem (`pre`, fp32, through gemm_tma_kernel), the centroid term a 128-vector per centroid; the kernel
//                     gathers rows of `pre`, subtracts, ReLUs and splits straight into tensor memory.  Layers 2 and 3 then run
//                     with A from TMEM and hi / lo weights resident in shared memory (192 KB); 2 chains of 256 TMEM columns.
//   group-all SA3 and the FC head use gemm_tma_kernel's three-pass mode (gemm_tc.cu) with [hi | lo] rows as hand-off format.
#include <cstdlib>

#include "engine.h"
#include "spec_math.cuh"
#include "tc_common.cuh"

namespace mpn {
using namespace tc;

int* tc_error_flag(mpn_ctx* c);
int sa_split(const mpn_ctx* c, int B, int max_split);
int sa_pack();
int launch_gemm_tc_ex(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, int lda, int a_lo_off, const __nv_bfloat16* W, int ldw,
                      int w_lo_off, int K, const float* bias, int M, int N, void* C, int ldc, int c_lo_off, int split, uint8_t* arg_out);
int launch_groupnorm_lrelu_split(mpn_ctx* c, cudaStream_t s, float* x, int M, int C, int groups, const float* gamma, const float* beta,
                                 __nv_bfloat16* out);
enum { X3_EPI_F32 = 1, X3_EPI_RELU_SPLIT = 4, X3_EPI_MAXPOOL_SPLIT = 5, X3_EPI_LRELU_F32 = 6 };   // gemm_tc.cu's epilogue ids

// ---------------------------------------------------------------------------------------------- weights
constexpr int A1_K = 80;     // SA1 output rows / per-point GEMM operand: [64 features | x y z | 0 x13], then the same as lo parts
constexpr int A3_KX = 272;   // SA2 output rows / SA3 operand: [256 features | x y z | 0 x13]
struct X3Weights {
  __nv_bfloat16* sa1_l1 = nullptr;               // [64][16]: the single K = 16 step of layer 1 (see pack_sa1x3_kernel)
  __nv_bfloat16* sa1_hi[2] = {nullptr, nullptr}; // layers 2, 3: [64][80], columns 64..66 = the bias as three bf16 parts
  __nv_bfloat16* sa1_lo[2] = {nullptr, nullptr}; // [64][64]
  __nv_bfloat16* sa2_w1p = nullptr;              // [128][2*80] (hi | lo), K order [f0..f63, x, y, z, 0-pad]
  float* sa2_w1x = nullptr;                      // [3][128] fp32: layer 1's dx / dy / dz columns
  __nv_bfloat16* sa2_w2 = nullptr;               // [128][2*128]
  __nv_bfloat16* sa2_w3 = nullptr;               // [256][2*128]
  __nv_bfloat16* sa3[3] = {nullptr, nullptr, nullptr};   // [N][2*Kpad], layer 1 K order [256 features, x, y, z, 0-pad]
  __nv_bfloat16* fc[3] = {nullptr, nullptr, nullptr};    // [out][2*in]
  __nv_bfloat16* dec0 = nullptr;                         // decoder.0 [512][2*2112] (its hi half doubles as the bf16 mode's copy)
  bool ready = false;
};
static std::map<mpn_ctx*, X3Weights> g_x3;

// w [out][in] fp32 -> dst [out][ld]: hi parts at columns [0, kpad), lo parts at [lo_off, lo_off + kpad); rot: K order
// [f..., dx, dy, dz] <- the reference's [dx, dy, dz, f...]
__global__ void pack_split_kernel(const float* __restrict__ w, int out, int in, int kpad, int rot, __nv_bfloat16* __restrict__ dst, int ld,
                                  int lo_off) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out * kpad) return;
  int o = i / kpad, k = i % kpad;
  float v = 0.f;
  if (k < in) {
    int src = rot ? (k < in - 3 ? k + 3 : k - (in - 3)) : k;
    v = w[(size_t)o * in + src];
  }
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  dst[(size_t)o * ld + k] = h;
  dst[(size_t)o * ld + lo_off + k] = l;
}

// SA1 (4 -> 64 -> 64 -> 64).  Layer 1 is ONE K = 16 MMA step; with a = (dx, dy, dz, m) (m = the mask feature, exact in bf16):
//   operand row  [dx_h dy_h dz_h m | dx_l dy_l dz_l 0 | dx_h dy_h dz_h m | 1 1 1 0]
//   weight row   [  W_h (4)        |   W_h (3)     0 |   W_l (4)        | b1 b2 b3 0]     b = b1 + b2 + b3 (three bf16 parts)
// Layers 2, 3: hi tile [64][80] = [W_h (64) | b1 b2 b3 0 x13] (the bias chunk multiplies a constant ones tile), lo tile [64][64].
__global__ void pack_sa1x3_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                  const float* __restrict__ b2, const float* __restrict__ w3, const float* __restrict__ b3,
                                  __nv_bfloat16* __restrict__ l1, __nv_bfloat16* __restrict__ hi2, __nv_bfloat16* __restrict__ lo2,
                                  __nv_bfloat16* __restrict__ hi3, __nv_bfloat16* __restrict__ lo3) {
  const int o = blockIdx.x, k = threadIdx.x;   // 64 blocks x 80 threads
  auto bias_part = [](float b, int part) {
    const __nv_bfloat16 p1 = __float2bfloat16_rn(b);
    const float r1 = b - __bfloat162float(p1);
    const __nv_bfloat16 p2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 p3 = __float2bfloat16_rn(r1 - __bfloat162float(p2));
    return part == 0 ? p1 : (part == 1 ? p2 : p3);
  };
  if (k < 16) {
    __nv_bfloat16 v = __float2bfloat16_rn(0.f), h, l;
    if (k < 4) { split_bf16(w1[o * 4 + k], h, l); v = h; }
    else if (k < 7) { split_bf16(w1[o * 4 + k - 4], h, l); v = h; }
    else if (k >= 8 && k < 12) { split_bf16(w1[o * 4 + k - 8], h, l); v = l; }
    else if (k >= 12 && k < 15) v = bias_part(b1[o], k - 12);
    l1[o * 16 + k] = v;
  }
  for (int layer = 0; layer < 2; ++layer) {
    const float* w = layer ? w3 : w2;
    const float* bb = layer ? b3 : b2;
    __nv_bfloat16* hi = layer ? hi3 : hi2;
    __nv_bfloat16* lo = layer ? lo3 : lo2;
    __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
    if (k < 64) split_bf16(w[o * 64 + k], h, l);
    else if (k < 67) h = bias_part(bb[o], k - 64);
    hi[o * 80 + k] = h;
    if (k < 64) lo[o * 64 + k] = l;
  }
}

__global__ void pack_w1x_kernel(const float* __restrict__ w1, float* __restrict__ dst) {   // w1 [128][67] -> dst [3][128]
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * 128) dst[i] = w1[(size_t)(i % 128) * 67 + i / 128];
}

// (re)pack every split-bf16 operand copy from the fp32 master weights on stream s (buffers allocated on first use)
int x3_pack_weights(mpn_ctx* c, cudaStream_t s) {
  X3Weights& t = g_x3[c];
  auto alloc = [&](__nv_bfloat16** p, size_t n) -> int {
    if (*p) return MPN_OK;
    MPN_CHECK_CUDA(cudaMalloc(p, n * sizeof(__nv_bfloat16)));
    return MPN_OK;
  };
  int r;
  if ((r = alloc(&t.sa1_l1, 64 * 16))) return r;
  for (int l = 0; l < 2; ++l) {
    if ((r = alloc(&t.sa1_hi[l], 64 * 80))) return r;
    if ((r = alloc(&t.sa1_lo[l], 64 * 64))) return r;
  }
  pack_sa1x3_kernel<<<64, 80, 0, s>>>(c->w.sa[0][0].w, c->w.sa[0][0].b, c->w.sa[0][1].w, c->w.sa[0][1].b, c->w.sa[0][2].w, c->w.sa[0][2].b,
                                      t.sa1_l1, t.sa1_hi[0], t.sa1_lo[0], t.sa1_hi[1], t.sa1_lo[1]);
  c->launches++;
  auto pack = [&](const Linear& L, int kpad, int rot, __nv_bfloat16** dst) -> int {
    int rr = alloc(dst, (size_t)L.out * 2 * kpad);
    if (rr) return rr;
    const int n = L.out * kpad;
    pack_split_kernel<<<(n + 255) / 256, 256, 0, s>>>(L.w, L.out, L.in, kpad, rot, *dst, 2 * kpad, kpad);
    c->launches++;
    return MPN_OK;
  };
  if ((r = pack(c->w.sa[1][0], A1_K, 1, &t.sa2_w1p))) return r;
  if (!t.sa2_w1x) MPN_CHECK_CUDA(cudaMalloc(&t.sa2_w1x, 3 * 128 * sizeof(float)));
  pack_w1x_kernel<<<2, 192, 0, s>>>(c->w.sa[1][0].w, t.sa2_w1x);
  c->launches++;
  if ((r = pack(c->w.sa[1][1], 128, 0, &t.sa2_w2))) return r;
  if ((r = pack(c->w.sa[1][2], 128, 0, &t.sa2_w3))) return r;
  if ((r = pack(c->w.sa[2][0], A3_KX, 1, &t.sa3[0]))) return r;
  if ((r = pack(c->w.sa[2][1], 512, 0, &t.sa3[1]))) return r;
  if ((r = pack(c->w.sa[2][2], 512, 0, &t.sa3[2]))) return r;
  for (int l = 0; l < 3; ++l)
    if ((r = pack(c->w.fc[l], c->w.fc[l].in, 0, &t.fc[l]))) return r;
  if ((r = pack(c->w.dec[0], c->w.dec[0].in, 0, &t.dec0))) return r;
  MPN_CHECK_CUDA(cudaGetLastError());
  t.ready = true;
  return MPN_OK;
}

void x3_free(mpn_ctx* c) {
  auto it = g_x3.find(c);
  if (it == g_x3.end()) return;
  X3Weights& t = it->second;
  __nv_bfloat16* ps[] = {t.sa1_l1, t.sa1_hi[0], t.sa1_hi[1], t.sa1_lo[0], t.sa1_lo[1], t.sa2_w1p, t.sa2_w2, t.sa2_w3,
                         t.sa3[0], t.sa3[1], t.sa3[2], t.fc[0], t.fc[1], t.fc[2], t.dec0};
  for (auto p : ps) if (p) cudaFree(p);
  if (t.sa2_w1x) cudaFree(t.sa2_w1x);
  g_x3.erase(it);
}

// scratch of the mode, per problem: SA1 output rows a1 [512][160] bf16 | layer-1 pre-activations pre [512][128] f32 | SA2 output rows
// a3 [128][544] bf16 | SA3 hidden h1, h2 [128][1024] bf16 | pooled f3 [2048] bf16 | FC operands g1 [8192], g2 [4096] bf16
struct X3Scratch { __nv_bfloat16 *a1, *a3, *h1, *h2, *f3, *g1, *g2; float* pre; };
static size_t al256(size_t n) { return (n + 255) / 256 * 256; }
size_t x3_scratch_bytes(int B) {
  const size_t b = (size_t)B;
  return al256(b * SA1_NPOINT * 2 * A1_K * 2) + al256(b * SA1_NPOINT * 128 * 4) + al256(b * SA2_NPOINT * 2 * A3_KX * 2) +
         2 * al256(b * SA2_NPOINT * 1024 * 2) + al256(b * 2048 * 2) + al256(b * 8192 * 2) + al256(b * 4096 * 2) + 1024;
}
static X3Scratch x3_scratch(mpn_ctx* c) {
  const size_t b = (size_t)c->ws.capacity;
  uint8_t* p = reinterpret_cast<uint8_t*>(c->ws.x3_scratch);
  X3Scratch s;
  s.a1 = reinterpret_cast<__nv_bfloat16*>(p); p += al256(b * SA1_NPOINT * 2 * A1_K * 2);
  s.pre = reinterpret_cast<float*>(p);        p += al256(b * SA1_NPOINT * 128 * 4);
  s.a3 = reinterpret_cast<__nv_bfloat16*>(p); p += al256(b * SA2_NPOINT * 2 * A3_KX * 2);
  s.h1 = reinterpret_cast<__nv_bfloat16*>(p); p += al256(b * SA2_NPOINT * 1024 * 2);
  s.h2 = reinterpret_cast<__nv_bfloat16*>(p); p += al256(b * SA2_NPOINT * 1024 * 2);
  s.f3 = reinterpret_cast<__nv_bfloat16*>(p); p += al256(b * 2048 * 2);
  s.g1 = reinterpret_cast<__nv_bfloat16*>(p); p += al256(b * 8192 * 2);
  s.g2 = reinterpret_cast<__nv_bfloat16*>(p);
  return s;
}

// ---------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ void wg_sync_x(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// global [rows][K] bf16 (row pitch `ld`) -> smem interleaved core-matrix layout (8 x 16 B core matrices, K-major)
__device__ __forceinline__ void stage_weight_ld(const __nv_bfloat16* __restrict__ g, int rows, int K, int ld, uint8_t* s) {
  const int KC = K / 8;
  for (int i = threadIdx.x; i < rows * KC; i += blockDim.x) {
    int r = i / KC, kc = i - r * KC;
    uint4 v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * ld + kc * 8));
    *reinterpret_cast<uint4*>(s + kmajor_chunk_off(r, kc, KC)) = v;
  }
}

// uniform hash grid of sa_tc.cu (cell edge slightly above the query radius: all points within r of a centroid lie in the 27 cells
// around the centroid's cell even under fp32 rounding of the cell coordinates)
constexpr int X3_BUCKETS = 4096;
__device__ __forceinline__ int grid_coord_x(float v) { return (int)floorf((v + 8.0f) * (1.0f / 0.0501f)); }
__device__ __forceinline__ uint32_t grid_bucket_x(int ix, int iy, int iz) {
  return ((uint32_t)ix * 73856093u ^ (uint32_t)iy * 19349663u ^ (uint32_t)iz * 83492791u) & (X3_BUCKETS - 1);
}

// halving butterfly over the 8 row classes of an accumulator-fragment load (lane bits 4, 3, 2): v[NV] per lane -> v[NV/8] per lane,
// the max over the warp's 32 rows of entries [base, base + NV/8), base = (NV/8) * (lane >> 2)
template <int NV>
__device__ __forceinline__ void rows_max_butterfly(float* v, int lane) {
#pragma unroll
  for (int w = NV / 2, bit = 16; w >= NV / 8; w >>= 1, bit >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = upper ? v[i] : v[i + w], keep = upper ? v[i + w] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, bit));
    }
  }
}

// ---------------------------------------------------------------------------------------------- SA1, split-bf16
constexpr int S1X_NWG = 4, S1X_COLS = 128;
struct Sa1xSmem {
  static constexpr size_t w_l1 = 0;                                        // [64][16] bf16
  static constexpr size_t w_hi = w_l1 + 64 * 16 * 2;                       // 2 x [64][80]
  static constexpr size_t w_lo = w_hi + 2 * 64 * 80 * 2;                   // 2 x [64][64]
  static constexpr size_t ones = w_lo + 2 * 64 * 64 * 2;                   // [128][16]: columns 0..2 = 1.0
  static constexpr size_t cnt = ones + 128 * 16 * 2;                       // u32 [BUCKETS]
  static constexpr size_t lists = cnt + X3_BUCKETS * 4;                    // [NWG][4][128] u16
  static constexpr size_t cand = lists + (size_t)S1X_NWG * 4 * 128 * 2;    // [NWG][4][256] u16
  static constexpr size_t red = cand + (size_t)S1X_NWG * 4 * 256 * 2;      // [NWG][4][64] f32
  static constexpr size_t cxyz = red + (size_t)S1X_NWG * 4 * 64 * 4;       // f32 [512][3]
  static constexpr size_t bars = cxyz + (size_t)SA1_NPOINT * 3 * 4;
  static constexpr size_t bstart = (bars + 64 + 15) / 16 * 16;             // u16 [BUCKETS + 1]
  static constexpr size_t cloud = (bstart + (X3_BUCKETS + 1) * 2 + 15) / 16 * 16;   // float4 [N]
  __host__ __device__ static size_t sidx(int N) { return cloud + (size_t)N * 16; }    // u16 [N]
  static size_t total(int N) { return sidx(N) + (size_t)N * 2 + 64; }
};

__global__ void __launch_bounds__(128 * S1X_NWG, 1)
sa1x3_tc_kernel(const float* __restrict__ cloud, int N, const float* __restrict__ new_xyz, float r2, const __nv_bfloat16* __restrict__ gl1,
                const __nv_bfloat16* __restrict__ ghi2, const __nv_bfloat16* __restrict__ glo2, const __nv_bfloat16* __restrict__ ghi3,
                const __nv_bfloat16* __restrict__ glo3, __nv_bfloat16* __restrict__ out_rows, float* __restrict__ out_f32,
                int* __restrict__ err, int32_t* __restrict__ ball_idx, int split, int pack) {
  using S = Sa1xSmem;
  constexpr int NS = NSAMPLE, NWG = S1X_NWG;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sL1 = smem + S::w_l1;
  uint8_t* sHi = smem + S::w_hi;
  uint8_t* sLo = smem + S::w_lo;
  uint8_t* sOnes = smem + S::ones;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bars + 8 * NWG);
  uint16_t* bstart = reinterpret_cast<uint16_t*>(smem + S::bstart);
  float4* cl = reinterpret_cast<float4*>(smem + S::cloud);
  uint16_t* sidx = reinterpret_cast<uint16_t*>(smem + S::sidx(N));
  float* cxyz = reinterpret_cast<float*>(smem + S::cxyz);

  const int b = blockIdx.x / split, part = blockIdx.x % split;   // small batches: rounds dealt to `split` CTAs per problem (sa_split)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int g = warp >> 2, wq = warp & 3, t = threadIdx.x & 127, lane = threadIdx.x & 31;
  uint16_t* lists = reinterpret_cast<uint16_t*>(smem + S::lists) + (size_t)g * 4 * 128;
  uint16_t* wcand = reinterpret_cast<uint16_t*>(smem + S::cand) + (size_t)(g * 4 + wq) * 256;
  float* red = reinterpret_cast<float*>(smem + S::red) + g * 256;
  const float4* gcl = reinterpret_cast<const float4*>(cloud) + (size_t)b * N;

  stage_weight_ld(gl1, 64, 16, 16, sL1);
  stage_weight_ld(ghi2, 64, 80, 80, sHi);
  stage_weight_ld(ghi3, 64, 80, 80, sHi + 64 * 80 * 2);
  stage_weight_ld(glo2, 64, 64, 64, sLo);
  stage_weight_ld(glo3, 64, 64, 64, sLo + 64 * 64 * 2);
  for (int i = threadIdx.x; i < 128 * 2; i += blockDim.x)   // ones tile: K columns 0..2 of every row = 1.0
    *reinterpret_cast<uint4*>(sOnes + kmajor_chunk_off(i >> 1, i & 1, 2)) = (i & 1) ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
  for (int i = threadIdx.x; i < SA1_NPOINT * 3; i += blockDim.x) cxyz[i] = __ldg(new_xyz + (size_t)b * SA1_NPOINT * 3 + i);
  // ---- the problem's cloud -> shared memory, and the hash grid over it (counting sort of point indices by bucket)
  {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + S::cnt);
    __shared__ uint32_t wsum[16];
    for (int i = threadIdx.x; i < X3_BUCKETS; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const float4 v = __ldg(gcl + k);
      cl[k] = v;
      atomicAdd(&cnt[grid_bucket_x(grid_coord_x(v.x), grid_coord_x(v.y), grid_coord_x(v.z))], 1u);
    }
    __syncthreads();
    constexpr int PER = X3_BUCKETS / 512;   // blockDim.x = 512
    uint32_t loc[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { loc[i] = cnt[threadIdx.x * PER + i]; sum += loc[i]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t v = lane < 16 ? wsum[lane] : 0u, iv = v;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += y; }
      if (lane < 16) wsum[lane] = iv - v;
    }
    __syncthreads();
    {
      uint32_t run = wsum[threadIdx.x >> 5] + inc - sum;
#pragma unroll
      for (int i = 0; i < PER; ++i) { bstart[threadIdx.x * PER + i] = (uint16_t)run; cnt[threadIdx.x * PER + i] = run; run += loc[i]; }
    }
    if (threadIdx.x == 0) bstart[X3_BUCKETS] = (uint16_t)N;
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const float4 v = cl[k];
      sidx[atomicAdd(&cnt[grid_bucket_x(grid_coord_x(v.x), grid_coord_x(v.y), grid_coord_x(v.z))], 1u)] = (uint16_t)k;
    }
  }
  __shared__ int round_ctr[1 + 8];
  __shared__ int hcnt_s[S1X_NWG * 4];   // distinct hits (<= 128) of the round's four ball queries
  int* next_round = round_ctr;
  int* rsel = round_ctr + 1;
  int* hcnt = hcnt_s + g * 4;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NWG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    *next_round = NWG;
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemD = *tmem_slot + (uint32_t)g * S1X_COLS;        // accumulator: 64 columns
  const uint32_t tmemAh = tmemD + 64, tmemAl = tmemD + 96;            // operand hi / lo: 32 columns (64 bf16 per row) each
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  const uint32_t tlane = tmemD + lane_off;
  const uint64_t dOnes = make_smem_desc(smem_u32(sOnes), 128, 2 * 128, LAYOUT_NONE);
  const uint64_t dL1 = make_smem_desc(smem_u32(sL1), 128, 2 * 128, LAYOUT_NONE);
  const uint64_t dHi2 = make_smem_desc(smem_u32(sHi), 128, 10 * 128, LAYOUT_NONE);
  const uint64_t dHi3 = make_smem_desc(smem_u32(sHi + 64 * 80 * 2), 128, 10 * 128, LAYOUT_NONE);
  const uint64_t dLo2 = make_smem_desc(smem_u32(sLo), 128, 8 * 128, LAYOUT_NONE);
  const uint64_t dLo3 = make_smem_desc(smem_u32(sLo + 64 * 64 * 2), 128, 8 * 128, LAYOUT_NONE);
  uint64_t* bar = &bars[g];
  uint32_t phase = 0;
  bool ok = true;
  constexpr uint32_t IDESC = make_idesc_bf16(128, 64);
  const unsigned lt = (1u << lane) - 1u;

  // complete ball query of centroid jc by this warp -> lists[wq][0..H), H = hcnt[wq] distinct hits: the 27 hash cells around the
  // centroid, candidates tested against the cloud in shared memory; hits ranked by ORIGINAL point index (= the linear scan's
  // first-128 order) when more than 128 were found or the caller wants the index lists, otherwise left in bucket order (the max-pool
  // takes the set).  pointnet2's first-hit padding is never materialised: rows beyond H re-read row H - 1 (a duplicate either way).
  // Linear scan when a neighbourhood holds more than 256 candidates.  Same algorithm as sa1t_tc_kernel (sa_tc.cu).
  const bool need_order = ball_idx != nullptr;
  auto warp_ball_query = [&](int jc) {
    uint16_t* widx = lists + wq * 128;
    const float qx = cxyz[3 * jc], qy = cxyz[3 * jc + 1], qz = cxyz[3 * jc + 2];
    const int ix = grid_coord_x(qx), iy = grid_coord_x(qy), iz = grid_coord_x(qz);
    uint32_t bk = 0x10000u + lane;
    if (lane < 27) bk = grid_bucket_x(ix + (lane % 3) - 1, iy + ((lane / 3) % 3) - 1, iz + (lane / 9) - 1);
    const unsigned peers = __match_any_sync(0xffffffffu, bk);
    const bool leader = lane < 27 && lane == __ffs(peers) - 1;
    int s0 = 0, n0 = 0;
    if (leader) { s0 = bstart[bk]; n0 = bstart[bk + 1] - s0; }
    int incl = n0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int C = __shfl_sync(0xffffffffu, incl, 31);
    if (C <= 256) {
      {
        const int o = incl - n0;
        for (int i = 0; i < n0; ++i) wcand[o + i] = sidx[s0 + i];
      }
      __syncwarp();
      int H = 0;
      for (int c0 = 0; c0 < C; c0 += 32) {
        const int ci = c0 + lane;
        bool hit = false;
        int k = 0;
        if (ci < C) { k = wcand[ci]; const float4 v = cl[k]; hit = dist2(qx, qy, qz, v.x, v.y, v.z) < r2; }
        __syncwarp();   // every lane has read its candidate before any lane compacts in place
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hit) wcand[H + __popc(hm & lt)] = (uint16_t)k;         // in place: write position <= read position
        H += __popc(hm);
        __syncwarp();
      }
      if (need_order || H > NS) {   // rank by original index: the linear scan's order (which 128 survive; ball_idx output)
        for (int h = lane; h < H; h += 32) {
          const int my = wcand[h];
          int rank = 0;
          for (int i = 0; i < H; ++i) rank += wcand[i] < my;
          if (rank < NS) widx[rank] = (uint16_t)my;
        }
      } else {                      // the max-pool only needs the SET of (at most 128) hits
        for (int h = lane; h < H; h += 32) widx[h] = wcand[h];
      }
      if (lane == 0) { hcnt[wq] = max(1, min(H, NS)); if (H == 0) widx[0] = 0; }
    } else {
      int cnt = 0;
      uint16_t first = 0;
      for (int k0 = 0; k0 < N && cnt < NS; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < N) { const float4 v = cl[k]; hit = dist2(qx, qy, qz, v.x, v.y, v.z) < r2; }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm && cnt == 0) first = (uint16_t)(k0 + __ffs(hm) - 1);
        const int pos = cnt + __popc(hm & lt);
        if (hit && pos < NS) widx[pos] = (uint16_t)k;
        cnt += __popc(hm);
      }
      if (lane == 0) { hcnt[wq] = max(1, min(cnt, NS)); if (cnt == 0) widx[0] = first; }
    }
    __syncwarp();
  };
  // layers 2 / 3: bias through an SS MMA against the constant ones tile (b1 + b2 + b3), then A_hi W_hi + A_lo W_hi + A_hi W_lo
  // with A straight from tensor memory
  auto issue = [&](uint64_t dHi, uint64_t dLo) {
    if (wq == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_bf16_ss_off(tmemD, dOnes, 0, dHi, 64, IDESC, 0);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_bf16_ts(tmemD, tmemAl + ks * 8, dHi + (uint64_t)(ks * 16), IDESC, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_bf16_ts(tmemD, tmemAh + ks * 8, dLo + (uint64_t)(ks * 16), IDESC, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_bf16_ts(tmemD, tmemAh + ks * 8, dHi + (uint64_t)(ks * 16), IDESC, 1);
        mma_commit(bar);
      }
      __syncwarp();
    }
  };
  // accumulator (bias inside) -> relu -> (hi, lo) bf16 -> the operand columns of this thread's lane
  auto epilogue_to_tmem = [&]() {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tlane + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) split_relu_pack(__uint_as_float(v[q * 16 + 2 * i]), __uint_as_float(v[q * 16 + 2 * i + 1]), hi[i], lo[i]);
        tmem_st8(tmemAh + lane_off + (c0 >> 1) + q * 8, hi);
        tmem_st8(tmemAl + lane_off + (c0 >> 1) + q * 8, lo);
      }
    }
    tmem_st_wait();
  };

  int ntiles_done = 0;
  for (int round = g; (round * split + part) * 4 < SA1_NPOINT && ok;) {
    const int base = (round * split + part) * 4;
    const int nvalid = min(4, SA1_NPOINT - base);
    if (wq < nvalid) warp_ball_query(base + wq);
    wg_sync_x(g);
    const TilePack tp = pack_round(hcnt, nvalid, pack);   // the round's distinct rows packed into 128-row tiles (tc_common.cuh)
    if (ball_idx)   // pointnet2's output format: first-hit padding
      for (int c = 0; c < nvalid; ++c) ball_idx[((size_t)b * SA1_NPOINT + base + c) * NS + t] = lists[c * 128 + (t < hcnt[c] ? t : 0)];
#pragma unroll 1
    for (int tile = 0; tile < tp.ntiles && ok; ++tile) {
      {
        const int mc = pack_owner(tp, nvalid, tile, wq);   // this warp's quarter of the tile belongs to centroid base + mc
        const int j = base + mc;
        const int k = lists[mc * 128 + min((wq - pack_q0(tp, mc)) * 32 + lane, hcnt[mc] - 1)];
        const float4 p = cl[k];
        const float dx = fsub(p.x, cxyz[3 * j]), dy = fsub(p.y, cxyz[3 * j + 1]), dz = fsub(p.z, cxyz[3 * j + 2]);
        uint32_t h0, l0, h1, l1;
        split_pack(dx, dy, h0, l0);
        split_pack(dz, p.w, h1, l1);   // the mask value (0, 1, 2) is exact in bf16: its lo part is 0
        const uint32_t row[8] = {h0, h1, l0, l1 & 0xFFFFu, h0, h1, 0x3F803F80u, 0x00003F80u};
        tmem_st8(tmemAh + lane_off, row);
        tmem_st_wait();
      }
      tc_fence_before();
      wg_sync_x(g);
      if (wq == 0) {                                   // layer 1: one K = 16 step (bias inside)
        tc_fence_after();
        if (elect_one()) { mma_bf16_ts(tmemD, tmemAh, dL1, IDESC, 0); mma_commit(bar); }
        __syncwarp();
      }
#pragma unroll 1
      for (int layer = 1; layer < 3; ++layer) {
        ok = ok && mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        epilogue_to_tmem();
        tc_fence_before();
        wg_sync_x(g);
        issue(layer == 1 ? dHi2 : dHi3, layer == 1 ? dLo2 : dLo3);
      }
      ok = ok && mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      // fp32 max over each warp's 32 rows: accumulator-fragment loads (a thread holds 4 rows x 16 channels) -> in-thread max ->
      // halving butterfly over the warp's 8 row classes -> lane L holds channels 2L, 2L + 1 of the warp's 32 rows
      {
        float v[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t va[16], vb[16];
          tmem_ld_16x256b_x4(tlane + h * 32, va);
          tmem_ld_16x256b_x4(tlane + (16u << 16) + h * 32, vb);
          tmem_ld_wait();
#pragma unroll
          for (int rep = 0; rep < 4; ++rep) {
            v[(h * 4 + rep) * 2] = fmaxf(fmaxf(__uint_as_float(va[rep * 4]), __uint_as_float(va[rep * 4 + 2])),
                                         fmaxf(__uint_as_float(vb[rep * 4]), __uint_as_float(vb[rep * 4 + 2])));
            v[(h * 4 + rep) * 2 + 1] = fmaxf(fmaxf(__uint_as_float(va[rep * 4 + 1]), __uint_as_float(va[rep * 4 + 3])),
                                             fmaxf(__uint_as_float(vb[rep * 4 + 1]), __uint_as_float(vb[rep * 4 + 3])));
          }
        }
        rows_max_butterfly<16>(v, lane);
        *reinterpret_cast<float2*>(red + wq * 64 + 2 * lane) = make_float2(v[0], v[1]);
      }
      tc_fence_before();
      wg_sync_x(g);
      // the tile's centroids: warp c finishes centroid c -- max over the quarters it owns, ReLU, split, output row
      // [64 features | x y z | 0-pad] as (hi | lo) bf16 pairs
      if (wq < nvalid && pack_tile(tp, wq) == tile) {
        const int q0 = pack_q0(tp, wq), q1 = pack_q1(tp, nvalid, wq), j = base + wq;
        float2 m = *reinterpret_cast<const float2*>(red + q0 * 64 + 2 * lane);
        for (int q = q0 + 1; q < q1; ++q) {
          const float2 o = *reinterpret_cast<const float2*>(red + q * 64 + 2 * lane);
          m.x = fmaxf(m.x, o.x); m.y = fmaxf(m.y, o.y);
        }
        m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f);
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out_rows + ((size_t)b * SA1_NPOINT + j) * (2 * A1_K));
        uint32_t hi, lo;
        split_pack(m.x, m.y, hi, lo);
        o32[lane] = hi;
        o32[A1_K / 2 + lane] = lo;
        if (lane < 8) {
          const float x0 = 2 * lane < 3 ? cxyz[3 * j + 2 * lane] : 0.f, x1 = 2 * lane + 1 < 3 ? cxyz[3 * j + 2 * lane + 1] : 0.f;
          split_pack(x0, x1, hi, lo);
          o32[32 + lane] = hi;
          o32[A1_K / 2 + 32 + lane] = lo;
        }
        if (out_f32) *reinterpret_cast<float2*>(out_f32 + ((size_t)b * SA1_NPOINT + j) * 64 + 2 * lane) = m;
      }
    }
    ntiles_done += tp.ntiles;
    if (t == 0) rsel[g] = atomicAdd(next_round, 1);
    wg_sync_x(g);   // the lists are rewritten by the next round
    round = rsel[g];
  }
  if (t == 0 && ntiles_done) atomicAdd(sa_tile_counter(err, 0), (unsigned long long)ntiles_done);
  if (!ok && (threadIdx.x & 31) == 0) atomicExch(err, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------- SA2, split-bf16
constexpr int S2X_NWG = 2, S2X_THREADS = 128 * S2X_NWG;
struct Sa2xSmem {
  static constexpr size_t w2h = 0;                                      // [128][128] bf16, K-major core matrices
  static constexpr size_t w2l = w2h + 128 * 128 * 2;
  static constexpr size_t w3h = w2l + 128 * 128 * 2;                    // [256][128]
  static constexpr size_t w3l = w3h + 256 * 128 * 2;
  static constexpr size_t b2 = w3l + 256 * 128 * 2;                     // [128] f32
  static constexpr size_t b3 = b2 + 128 * 4;                            // [256] f32
  static constexpr size_t w1x = b3 + 256 * 4;                           // [3][128] f32
  static constexpr size_t pts = w1x + 3 * 128 * 4;                      // x[512] | y[512] | z[512]
  static constexpr size_t lists = pts + 3 * SA1_NPOINT * 4;             // [NWG][2 slots][4][128] u16
  static constexpr size_t u = lists + (size_t)S2X_NWG * 2 * 4 * 128 * 2;   // [NWG][4 quarters][128] f32: W1x c_i of the quarter's centroid
  static constexpr size_t pool = u + (size_t)S2X_NWG * 4 * 128 * 4;     // [NWG][2 tiles][4 warps][128] f32
  static constexpr size_t bars = pool + (size_t)S2X_NWG * 2 * 4 * 128 * 4;
  static constexpr size_t total = bars + 64;
};

__global__ void __launch_bounds__(S2X_THREADS, 1)
sa2x3_tc_kernel(const float* __restrict__ xyz, int stride, const float* __restrict__ pre, const float* __restrict__ new_xyz, float r2,
                const __nv_bfloat16* __restrict__ gw2, const __nv_bfloat16* __restrict__ gw3, const float* __restrict__ gb2,
                const float* __restrict__ gb3, const float* __restrict__ gw1x, __nv_bfloat16* __restrict__ out_rows,
                float* __restrict__ out_f32, int* __restrict__ err, int32_t* __restrict__ ball_idx, int split, int pack) {
  using S = Sa2xSmem;
  constexpr int N = SA1_NPOINT, NCENT = SA2_NPOINT, NWG = S2X_NWG;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sB2 = reinterpret_cast<float*>(smem + S::b2);
  float* sB3 = reinterpret_cast<float*>(smem + S::b3);
  float* sW1x = reinterpret_cast<float*>(smem + S::w1x);
  float* px = reinterpret_cast<float*>(smem + S::pts);
  float* py = px + N;
  float* pz = py + N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bars + 8 * NWG);

  const int b = blockIdx.x / split, part = blockIdx.x % split;   // small batches: rounds dealt to `split` CTAs per problem (sa_split)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int g = warp >> 2, wq = warp & 3, lane = threadIdx.x & 31, t = threadIdx.x & 127;
  uint16_t* lists = reinterpret_cast<uint16_t*>(smem + S::lists) + (size_t)g * 2 * 4 * 128;
  float* sU = reinterpret_cast<float*>(smem + S::u) + g * 4 * 128;
  float* sPool = reinterpret_cast<float*>(smem + S::pool) + (size_t)g * 2 * 4 * 128;

  stage_weight_ld(gw2, 128, 128, 256, smem + S::w2h);
  stage_weight_ld(gw2 + 128, 128, 128, 256, smem + S::w2l);
  stage_weight_ld(gw3, 256, 128, 256, smem + S::w3h);
  stage_weight_ld(gw3 + 128, 256, 128, 256, smem + S::w3l);
  for (int i = threadIdx.x; i < 128; i += S2X_THREADS) sB2[i] = gb2[i];
  for (int i = threadIdx.x; i < 256; i += S2X_THREADS) sB3[i] = gb3[i];
  for (int i = threadIdx.x; i < 3 * 128; i += S2X_THREADS) sW1x[i] = gw1x[i];
  {
    const float* p = xyz + (size_t)b * N * stride;
    for (int k = threadIdx.x; k < N; k += S2X_THREADS) {
      px[k] = __ldg(p + (size_t)k * stride); py[k] = __ldg(p + (size_t)k * stride + 1); pz[k] = __ldg(p + (size_t)k * stride + 2);
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NWG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemD = *tmem_slot + (uint32_t)g * 256;              // accumulator: 128 columns
  const uint32_t tmemAh = tmemD + 128, tmemAl = tmemD + 192;           // operand hi / lo: 64 columns (128 bf16 per row) each
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  const uint32_t tlane = tmemD + lane_off;
  const uint64_t dW2h = make_smem_desc(smem_u32(smem + S::w2h), 128, 16 * 128, LAYOUT_NONE);
  const uint64_t dW2l = make_smem_desc(smem_u32(smem + S::w2l), 128, 16 * 128, LAYOUT_NONE);
  const uint64_t dW3h = make_smem_desc(smem_u32(smem + S::w3h), 128, 16 * 128, LAYOUT_NONE);
  const uint64_t dW3l = make_smem_desc(smem_u32(smem + S::w3l), 128, 16 * 128, LAYOUT_NONE);
  constexpr uint32_t W3_TILE1 = (128 / 8) * 16 * 128 / 16;             // rows 128..255 of W3, in 16-byte units
  constexpr uint32_t ID128 = make_idesc_bf16(128, 128);
  uint64_t* bar = &bars[g];
  uint32_t phase = 0;
  bool ok = true;
  const unsigned lt = (1u << lane) - 1u;
  __shared__ int hcnt_s[S2X_NWG * 2 * 4];   // [NWG][2 slots][4]: distinct hits (<= 128) of a round's ball queries
  int* hcnt = hcnt_s + g * 8;

  // one warp = one centroid: in-order scan of the 512 points, first 128 hits, first-hit padding (pointnet2 semantics)
  auto bq_round = [&](int base, int slot) {
    const int jc = base + wq;
    if (jc < NCENT) {
      const float* cp = new_xyz + ((size_t)b * NCENT + jc) * 3;
      const float qx = cp[0], qy = cp[1], qz = cp[2];
      uint16_t* out = lists + (slot * 4 + wq) * 128;
      int cnt = 0, first = 0;
#pragma unroll 4
      for (int k0 = 0; k0 < N; k0 += 32) {
        const int k = k0 + lane;
        const bool hit = dist2(qx, qy, qz, px[k], py[k], pz[k]) < r2;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (cnt == 0 && hm) first = k0 + __ffs(hm) - 1;
        const int pos = cnt + __popc(hm & lt);
        if (hit && pos < NSAMPLE) out[pos] = (uint16_t)k;
        cnt += __popc(hm);
      }
      for (int l = min(cnt, NSAMPLE) + lane; l < NSAMPLE; l += 32) out[l] = (uint16_t)first;
      if (lane == 0) hcnt[slot * 4 + wq] = max(1, min(cnt, NSAMPLE));
    }
  };
  // A_lo W_h + A_hi W_l + A_hi W_h (small terms first), K = 128 each, into the chain's accumulator
  auto issue = [&](uint64_t dWh, uint64_t dWl, uint32_t woff) {
    if (wq == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_bf16_ts(tmemD, tmemAl + ks * 8, dWh + (uint64_t)(woff + ks * 16), ID128, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_bf16_ts(tmemD, tmemAh + ks * 8, dWl + (uint64_t)(woff + ks * 16), ID128, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_bf16_ts(tmemD, tmemAh + ks * 8, dWh + (uint64_t)(woff + ks * 16), ID128, 1);
        mma_commit(bar);
      }
      __syncwarp();
    }
  };

  // The gather of a tile reads 128 rows x 512 B of `pre` from L2; the first quarter of the NEXT tile's row is fetched into
  // registers under the current tile's layer-3 MMAs, and the other quarters are requested one step ahead of their use.
  // A round's four neighbourhoods are packed into as few 128-row tiles as their distinct rows need (TilePack, tc_common.cuh).
  constexpr int STEP = NWG * 4;
  int r = 0, kpre = 0;
  float4 xpre[8];
  const int base0 = (g * split + part) * 4;
  TilePack tp{0u, 0u, 0}, tpn{0u, 0u, 0};
  int nvalid = 0, nvalid_n = 0, ntiles_done = 0;
  // after the ball queries of round (nb, nslot): counts -> packing, neighbour lists -> ball_idx
  auto open_round = [&](int nb, int nslot, TilePack& p, int& nv) {
    nv = min(4, NCENT - nb);
    p = pack_round(hcnt + nslot * 4, nv, pack);
    if (ball_idx)
      for (int c = 0; c < nv; ++c) ball_idx[((size_t)b * NCENT + nb + c) * NSAMPLE + t] = lists[(nslot * 4 + c) * 128 + t];
  };
  auto prefetch = [&](const TilePack& p, int nv, int nslot, int ntile) {
    const int mc = pack_owner(p, nv, ntile, wq);
    kpre = lists[(nslot * 4 + mc) * 128 + (wq - pack_q0(p, mc)) * 32 + lane];
    const float4* prow = reinterpret_cast<const float4*>(pre + ((size_t)b * N + kpre) * 128);
#pragma unroll
    for (int q = 0; q < 8; ++q) xpre[q] = __ldg(prow + q);
  };
  if (base0 < NCENT) { bq_round(base0, 0); wg_sync_x(g); open_round(base0, 0, tp, nvalid); prefetch(tp, nvalid, 0, 0); }
  for (int base = base0; base < NCENT && ok; base += STEP * split, ++r) {
    const int slot = r & 1;
#pragma unroll 1
    for (int tile = 0; tile < tp.ntiles && ok; ++tile) {
      {   // W1x c of the centroid that owns each quarter of this tile
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* cp = new_xyz + ((size_t)b * NCENT + base + pack_owner(tp, nvalid, tile, q)) * 3;
          sU[q * 128 + t] = fmaf(sW1x[256 + t], cp[2], fmaf(sW1x[128 + t], cp[1], sW1x[t] * cp[0]));
        }
      }
      wg_sync_x(g);
      // ---- layer 1 (per-point pre-activation - centroid term), ReLU, split, straight into the operand columns
      {
        const float4* prow = reinterpret_cast<const float4*>(pre + ((size_t)b * N + kpre) * 128);
        const float* uq = sU + wq * 128;
        float4 xa[8], xb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) xa[q] = xpre[q];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int c0 = ch * 32;
          if (ch < 3) {
#pragma unroll
            for (int q = 0; q < 8; ++q) xb[q] = __ldg(prow + (c0 >> 2) + 8 + q);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 xv = xa[q * 4 + i];
              const float4 uv = *reinterpret_cast<const float4*>(uq + c0 + q * 16 + i * 4);
              split_relu_pack(xv.x - uv.x, xv.y - uv.y, hi[2 * i], lo[2 * i]);
              split_relu_pack(xv.z - uv.z, xv.w - uv.w, hi[2 * i + 1], lo[2 * i + 1]);
            }
            tmem_st8(tmemAh + lane_off + (c0 >> 1) + q * 8, hi);
            tmem_st8(tmemAl + lane_off + (c0 >> 1) + q * 8, lo);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) xa[q] = xb[q];
        }
        tmem_st_wait();
      }
      tc_fence_before();
      wg_sync_x(g);
      issue(dW2h, dW2l, 0);                                      // layer 2
      ok = ok && mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t hi[8], lo[8];
          const float* bb = sB2 + c0 + q * 16;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            split_relu_pack(__uint_as_float(v[q * 16 + 2 * i]) + bb[2 * i], __uint_as_float(v[q * 16 + 2 * i + 1]) + bb[2 * i + 1], hi[i], lo[i]);
          tmem_st8(tmemAh + lane_off + (c0 >> 1) + q * 8, hi);
          tmem_st8(tmemAl + lane_off + (c0 >> 1) + q * 8, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      wg_sync_x(g);
      issue(dW3h, dW3l, 0);                                      // layer 3, channels 0..127
      // under the layer-3 MMAs: the next round's ball query (when this was the round's last tile) and the next tile's rows
      if (tile + 1 < tp.ntiles) {
        prefetch(tp, nvalid, slot, tile + 1);
      } else {
        const int nb = base + STEP * split;
        if (nb < NCENT) { bq_round(nb, slot ^ 1); wg_sync_x(g); open_round(nb, slot ^ 1, tpn, nvalid_n); prefetch(tpn, nvalid_n, slot ^ 1, 0); }
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        ok = ok && mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        // fp32 max over each warp's 32 rows (= TMEM lanes): accumulator-fragment loads, in-thread max over a thread's 4 rows,
        // halving butterfly over the warp's 8 row classes; the quarters of a centroid meet in shared memory
        float* pl = sPool + half * 4 * 128;
        {
          float v[32];
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) {
            uint32_t va[16], vb[16];
            tmem_ld_16x256b_x4(tlane + blk * 32, va);
            tmem_ld_16x256b_x4(tlane + (16u << 16) + blk * 32, vb);
            tmem_ld_wait();
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
              v[(blk * 4 + rep) * 2] = fmaxf(fmaxf(__uint_as_float(va[rep * 4]), __uint_as_float(va[rep * 4 + 2])),
                                             fmaxf(__uint_as_float(vb[rep * 4]), __uint_as_float(vb[rep * 4 + 2])));
              v[(blk * 4 + rep) * 2 + 1] = fmaxf(fmaxf(__uint_as_float(va[rep * 4 + 1]), __uint_as_float(va[rep * 4 + 3])),
                                                 fmaxf(__uint_as_float(vb[rep * 4 + 1]), __uint_as_float(vb[rep * 4 + 3])));
            }
          }
          rows_max_butterfly<32>(v, lane);
          // entries i = 4 * (lane >> 2) + jj: column block i / 8, rep (i % 8) / 2, element i % 2 of the fragment layout
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int i = 4 * (lane >> 2) + jj;
            pl[wq * 128 + 32 * (i >> 3) + 8 * ((i & 7) >> 1) + 2 * (lane & 3) + (i & 1)] = v[jj];
          }
        }
        tc_fence_before();
        wg_sync_x(g);                                            // every lane of the accumulator has been read
        if (half == 0) issue(dW3h, dW3l, W3_TILE1);              // channels 128..255 into the same TMEM columns
        for (int c = 0; c < nvalid; ++c) {
          if (pack_tile(tp, c) != tile) continue;
          const int q0 = pack_q0(tp, c), q1 = pack_q1(tp, nvalid, c);
          float m = pl[q0 * 128 + t];
          for (int q = q0 + 1; q < q1; ++q) m = fmaxf(m, pl[q * 128 + t]);
          m = fmaxf(m + sB3[half * 128 + t], 0.f);
          __nv_bfloat16 h, l;
          split_bf16(m, h, l);
          __nv_bfloat16* o = out_rows + ((size_t)b * NCENT + base + c) * (2 * A3_KX);
          o[half * 128 + t] = h;
          o[A3_KX + half * 128 + t] = l;
          if (out_f32) out_f32[((size_t)b * NCENT + base + c) * 256 + half * 128 + t] = m;
        }
      }
      for (int i = t; i < nvalid * 16; i += 128) {   // [x y z | 0-pad] columns of the tile's centroids
        const int c = i >> 4, d = i & 15;
        if (pack_tile(tp, c) != tile) continue;
        const float v = d < 3 ? new_xyz[((size_t)b * NCENT + base + c) * 3 + d] : 0.f;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        __nv_bfloat16* o = out_rows + ((size_t)b * NCENT + base + c) * (2 * A3_KX);
        o[256 + d] = h;
        o[A3_KX + 256 + d] = l;
      }
    }
    ntiles_done += tp.ntiles;
    tp = tpn;
    nvalid = nvalid_n;
  }
  if (t == 0 && ntiles_done) atomicAdd(sa_tile_counter(err, 1), (unsigned long long)ntiles_done);
  if (!ok && (threadIdx.x & 31) == 0) atomicExch(err, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------- SA2, split-bf16, 8 warps per chain
// sa2x3_tc_kernel's chains were bound by their own SIMT phases: between two MMA batches the 4 warps of a chain gather, subtract, ReLU,
// split and store 128 columns per row (~2000 instructions per thread and tile), so with two chains per SM the tensor pipe idled ~40 %
// of the time (ncu: 51 % tensor-active, 30 % issue-active, 8 resident warps).  Here a chain is EIGHT warps: warp w owns TMEM lane
// quarter w % 4 (the hardware's lane restriction) and column half w / 4, i.e. every SIMT phase is split over twice the threads.
// Eight warps also mean eight ball queries per round, and the two 16-lane halves of a warp's accumulator fragment are pooled
// separately, so a round of EIGHT centroids is packed at a granularity of 16 rows (TilePack8): 0.33 instead of 0.39 tiles per group on
// tabletop scenes.  Same arithmetic and outputs as sa2x3_tc_kernel (kept as the A/B reference: MPN_SA2X3_V1=1).
constexpr int S2H_THREADS = 256 * S2X_NWG;
struct Sa2hSmem {
  static constexpr size_t w2h = 0;                                      // [128][128] bf16, K-major core matrices
  static constexpr size_t w2l = w2h + 128 * 128 * 2;
  static constexpr size_t w3h = w2l + 128 * 128 * 2;                    // [256][128]
  static constexpr size_t w3l = w3h + 256 * 128 * 2;
  static constexpr size_t b2 = w3l + 256 * 128 * 2;                     // [128] f32
  static constexpr size_t b3 = b2 + 128 * 4;                            // [256] f32
  static constexpr size_t w1x = b3 + 256 * 4;                           // [3][128] f32
  static constexpr size_t pts = w1x + 3 * 128 * 4;                      // x[512] | y[512] | z[512]
  static constexpr size_t lists = pts + 3 * SA1_NPOINT * 4;             // [NWG][2 slots][8][128] u16
  static constexpr size_t u = lists + (size_t)S2X_NWG * 2 * 8 * 128 * 2;   // [NWG][8 eighths][128] f32: W1x c_i of the eighth's centroid
  static constexpr size_t pool = u + (size_t)S2X_NWG * 8 * 128 * 4;     // [NWG][8 eighths][128] f32
  static constexpr size_t bars = pool + (size_t)S2X_NWG * 8 * 128 * 4;
  static constexpr size_t total = bars + 64;
};
__device__ __forceinline__ void chain_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

__global__ void __launch_bounds__(S2H_THREADS, 1)
sa2x3h_tc_kernel(const float* __restrict__ xyz, int stride, const float* __restrict__ pre, const float* __restrict__ new_xyz, float r2,
                 const __nv_bfloat16* __restrict__ gw2, const __nv_bfloat16* __restrict__ gw3, const float* __restrict__ gb2,
                 const float* __restrict__ gb3, const float* __restrict__ gw1x, __nv_bfloat16* __restrict__ out_rows,
                 float* __restrict__ out_f32, int* __restrict__ err, int32_t* __restrict__ ball_idx, int split, int pack) {
  using S = Sa2hSmem;
  constexpr int N = SA1_NPOINT, NCENT = SA2_NPOINT, NWG = S2X_NWG;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sB2 = reinterpret_cast<float*>(smem + S::b2);
  float* sB3 = reinterpret_cast<float*>(smem + S::b3);
  float* sW1x = reinterpret_cast<float*>(smem + S::w1x);
  float* px = reinterpret_cast<float*>(smem + S::pts);
  float* py = px + N;
  float* pz = py + N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bars + 8 * NWG);

  const int b = blockIdx.x / split, part = blockIdx.x % split;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int g = warp >> 3, wc = warp & 7, wq = wc & 3, hh = wc >> 2, lane = threadIdx.x & 31, u = threadIdx.x & 255;
  const int me = wq * 2 + (lane >> 4);                  // this thread's row sits in eighth `me` of the tile (TMEM lane 32 wq + lane)
  uint16_t* lists = reinterpret_cast<uint16_t*>(smem + S::lists) + (size_t)g * 2 * 8 * 128;
  float* sU = reinterpret_cast<float*>(smem + S::u) + g * 8 * 128;
  float* pl = reinterpret_cast<float*>(smem + S::pool) + (size_t)g * 8 * 128;

  stage_weight_ld(gw2, 128, 128, 256, smem + S::w2h);
  stage_weight_ld(gw2 + 128, 128, 128, 256, smem + S::w2l);
  stage_weight_ld(gw3, 256, 128, 256, smem + S::w3h);
  stage_weight_ld(gw3 + 128, 256, 128, 256, smem + S::w3l);
  for (int i = threadIdx.x; i < 128; i += S2H_THREADS) sB2[i] = gb2[i];
  for (int i = threadIdx.x; i < 256; i += S2H_THREADS) sB3[i] = gb3[i];
  for (int i = threadIdx.x; i < 3 * 128; i += S2H_THREADS) sW1x[i] = gw1x[i];
  {
    const float* p = xyz + (size_t)b * N * stride;
    for (int k = threadIdx.x; k < N; k += S2H_THREADS) {
      px[k] = __ldg(p + (size_t)k * stride); py[k] = __ldg(p + (size_t)k * stride + 1); pz[k] = __ldg(p + (size_t)k * stride + 2);
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < NWG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemD = *tmem_slot + (uint32_t)g * 256;              // accumulator: 128 columns
  const uint32_t tmemAh = tmemD + 128, tmemAl = tmemD + 192;           // operand hi / lo: 64 columns (128 bf16 per row) each
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  const uint32_t tlane = tmemD + lane_off;
  const uint64_t dW2h = make_smem_desc(smem_u32(smem + S::w2h), 128, 16 * 128, LAYOUT_NONE);
  const uint64_t dW2l = make_smem_desc(smem_u32(smem + S::w2l), 128, 16 * 128, LAYOUT_NONE);
  const uint64_t dW3h = make_smem_desc(smem_u32(smem + S::w3h), 128, 16 * 128, LAYOUT_NONE);
  const uint64_t dW3l = make_smem_desc(smem_u32(smem + S::w3l), 128, 16 * 128, LAYOUT_NONE);
  constexpr uint32_t W3_TILE1 = (128 / 8) * 16 * 128 / 16;             // rows 128..255 of W3, in 16-byte units
  constexpr uint32_t ID128 = make_idesc_bf16(128, 128);
  uint64_t* bar = &bars[g];
  uint32_t phase = 0;
  bool ok = true;
  const unsigned lt = (1u << lane) - 1u;
  __shared__ int hcnt_s[S2X_NWG * 2 * 8];   // [NWG][2 slots][8]: distinct hits (<= 128) of a round's ball queries
  int* hcnt = hcnt_s + g * 16;

  // every warp of the chain: one centroid, in-order scan of the 512 points (pointnet2 semantics, see sa2x3_tc_kernel)
  auto bq_round = [&](int base, int slot) {
    const int jc = base + wc;
    if (jc < NCENT) {
      const float* cp = new_xyz + ((size_t)b * NCENT + jc) * 3;
      const float qx = cp[0], qy = cp[1], qz = cp[2];
      uint16_t* out = lists + (slot * 8 + wc) * 128;
      int cnt = 0, first = 0;
#pragma unroll 4
      for (int k0 = 0; k0 < N; k0 += 32) {
        const int k = k0 + lane;
        const bool hit = dist2(qx, qy, qz, px[k], py[k], pz[k]) < r2;
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (cnt == 0 && hm) first = k0 + __ffs(hm) - 1;
        const int pos = cnt + __popc(hm & lt);
        if (hit && pos < NSAMPLE) out[pos] = (uint16_t)k;
        cnt += __popc(hm);
      }
      for (int l = min(cnt, NSAMPLE) + lane; l < NSAMPLE; l += 32) out[l] = (uint16_t)first;
      if (lane == 0) hcnt[slot * 8 + wc] = max(1, min(cnt, NSAMPLE));
    }
  };
  auto issue = [&](uint64_t dWh, uint64_t dWl, uint32_t woff) {
    if (wc == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_bf16_ts(tmemD, tmemAl + ks * 8, dWh + (uint64_t)(woff + ks * 16), ID128, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_bf16_ts(tmemD, tmemAh + ks * 8, dWl + (uint64_t)(woff + ks * 16), ID128, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) mma_bf16_ts(tmemD, tmemAh + ks * 8, dWh + (uint64_t)(woff + ks * 16), ID128, 1);
        mma_commit(bar);
      }
      __syncwarp();
    }
  };

  constexpr int STEP = NWG * 8;
  int r = 0, kpre = 0;
  float4 xpre[8];                                      // channels [64 hh, 64 hh + 32) of the next tile's row
  const int base0 = (g * split + part) * 8;
  TilePack8 tp{0u, 0u, 0}, tpn{0u, 0u, 0};
  int nvalid = 0, nvalid_n = 0, ntiles_done = 0;
  auto open_round = [&](int nb, int nslot, TilePack8& p, int& nv) {
    nv = min(8, NCENT - nb);
    p = pack_round8(hcnt + nslot * 8, nv, pack);
    if (ball_idx && u < 128)
      for (int c = 0; c < nv; ++c) ball_idx[((size_t)b * NCENT + nb + c) * NSAMPLE + u] = lists[(nslot * 8 + c) * 128 + u];
  };
  auto prefetch = [&](const TilePack8& p, int nv, int nslot, int ntile) {
    const int mc = pack8_owner(p, nv, ntile, me);
    kpre = lists[(nslot * 8 + mc) * 128 + (me - pack8_e0(p, mc)) * 16 + (lane & 15)];
    const float4* prow = reinterpret_cast<const float4*>(pre + ((size_t)b * N + kpre) * 128) + 16 * hh;
#pragma unroll
    for (int q = 0; q < 8; ++q) xpre[q] = __ldg(prow + q);
  };
  if (base0 < NCENT) { bq_round(base0, 0); chain_sync(g); open_round(base0, 0, tp, nvalid); prefetch(tp, nvalid, 0, 0); }
  for (int base = base0; base < NCENT && ok; base += STEP * split, ++r) {
    const int slot = r & 1;
#pragma unroll 1
    for (int tile = 0; tile < tp.ntiles && ok; ++tile) {
      {   // W1x c of the centroid that owns each eighth of this tile: 8 x 128 values over the chain's 256 threads
        const int ch = u & 127;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const int e = (u >> 7) * 4 + qq;
          const float* cp = new_xyz + ((size_t)b * NCENT + base + pack8_owner(tp, nvalid, tile, e)) * 3;
          sU[e * 128 + ch] = fmaf(sW1x[256 + ch], cp[2], fmaf(sW1x[128 + ch], cp[1], sW1x[ch] * cp[0]));
        }
      }
      chain_sync(g);
      // ---- layer 1 (per-point pre-activation - centroid term), ReLU, split: this thread's 64 channels of its row
      {
        const float4* prow = reinterpret_cast<const float4*>(pre + ((size_t)b * N + kpre) * 128) + 16 * hh;
        const float* uq = sU + me * 128 + 64 * hh;
        float4 xb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) xb[q] = __ldg(prow + 8 + q);
#pragma unroll
        for (int part2 = 0; part2 < 2; ++part2) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 xv = part2 ? xb[q * 4 + i] : xpre[q * 4 + i];
              const float4 uv = *reinterpret_cast<const float4*>(uq + part2 * 32 + q * 16 + i * 4);
              split_relu_pack(xv.x - uv.x, xv.y - uv.y, hi[2 * i], lo[2 * i]);
              split_relu_pack(xv.z - uv.z, xv.w - uv.w, hi[2 * i + 1], lo[2 * i + 1]);
            }
            tmem_st8(tmemAh + lane_off + 32 * hh + part2 * 16 + q * 8, hi);
            tmem_st8(tmemAl + lane_off + 32 * hh + part2 * 16 + q * 8, lo);
          }
        }
        tmem_st_wait();
      }
      tc_fence_before();
      chain_sync(g);
      issue(dW2h, dW2l, 0);                                      // layer 2
      ok = ok && mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 64 * hh; c0 < 64 * hh + 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t hi[8], lo[8];
          const float* bb = sB2 + c0 + q * 16;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            split_relu_pack(__uint_as_float(v[q * 16 + 2 * i]) + bb[2 * i], __uint_as_float(v[q * 16 + 2 * i + 1]) + bb[2 * i + 1], hi[i], lo[i]);
          tmem_st8(tmemAh + lane_off + (c0 >> 1) + q * 8, hi);
          tmem_st8(tmemAl + lane_off + (c0 >> 1) + q * 8, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      chain_sync(g);
      issue(dW3h, dW3l, 0);                                      // layer 3, channels 0..127
      // under the layer-3 MMAs: the next round's ball query (when this was the round's last tile) and the next tile's rows
      if (tile + 1 < tp.ntiles) {
        prefetch(tp, nvalid, slot, tile + 1);
      } else {
        const int nb = base + STEP * split;
        if (nb < NCENT) { bq_round(nb, slot ^ 1); chain_sync(g); open_round(nb, slot ^ 1, tpn, nvalid_n); prefetch(tpn, nvalid_n, slot ^ 1, 0); }
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        ok = ok && mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        if (half == 1) chain_sync(g);                            // the first half's maxima have been combined: pl may be rewritten
        // fp32 max over each 16-lane half of the warp's accumulator fragment (= one eighth of the tile) x 64 columns: in-thread max
        // over a thread's 2 rows, halving butterfly over the 8 row classes; the eighths of a centroid meet in shared memory
        {
          float va_[16], vb_[16];
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            uint32_t va[16], vb[16];
            tmem_ld_16x256b_x4(tlane + 64 * hh + blk * 32, va);
            tmem_ld_16x256b_x4(tlane + (16u << 16) + 64 * hh + blk * 32, vb);
            tmem_ld_wait();
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
              va_[(blk * 4 + rep) * 2] = fmaxf(__uint_as_float(va[rep * 4]), __uint_as_float(va[rep * 4 + 2]));
              va_[(blk * 4 + rep) * 2 + 1] = fmaxf(__uint_as_float(va[rep * 4 + 1]), __uint_as_float(va[rep * 4 + 3]));
              vb_[(blk * 4 + rep) * 2] = fmaxf(__uint_as_float(vb[rep * 4]), __uint_as_float(vb[rep * 4 + 2]));
              vb_[(blk * 4 + rep) * 2 + 1] = fmaxf(__uint_as_float(vb[rep * 4 + 1]), __uint_as_float(vb[rep * 4 + 3]));
            }
          }
          rows_max_butterfly<16>(va_, lane);
          rows_max_butterfly<16>(vb_, lane);
          // entries i = 2 * (lane >> 2) + jj: column block i / 8, rep (i % 8) / 2, element i % 2 of the fragment layout
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int i = 2 * (lane >> 2) + jj;
            const int col = 64 * hh + 32 * (i >> 3) + 8 * ((i & 7) >> 1) + 2 * (lane & 3) + (i & 1);
            pl[(2 * wq) * 128 + col] = va_[jj];
            pl[(2 * wq + 1) * 128 + col] = vb_[jj];
          }
        }
        tc_fence_before();
        chain_sync(g);                                           // every lane of the accumulator has been read
        if (half == 0) issue(dW3h, dW3l, W3_TILE1);              // channels 128..255 into the same TMEM columns
        {
          const int ch = u & 127;
          for (int c = u >> 7; c < nvalid; c += 2) {             // centroids dealt to the chain's two thread halves
            if (pack8_tile(tp, c) != tile) continue;
            const int e0 = pack8_e0(tp, c), e1 = pack8_e1(tp, nvalid, c);
            float m = pl[e0 * 128 + ch];
            for (int e = e0 + 1; e < e1; ++e) m = fmaxf(m, pl[e * 128 + ch]);
            m = fmaxf(m + sB3[half * 128 + ch], 0.f);
            __nv_bfloat16 h, l;
            split_bf16(m, h, l);
            __nv_bfloat16* o = out_rows + ((size_t)b * NCENT + base + c) * (2 * A3_KX);
            o[half * 128 + ch] = h;
            o[A3_KX + half * 128 + ch] = l;
            if (out_f32) out_f32[((size_t)b * NCENT + base + c) * 256 + half * 128 + ch] = m;
          }
        }
      }
      for (int i = u; i < nvalid * 16; i += 256) {   // [x y z | 0-pad] columns of the tile's centroids
        const int c = i >> 4, d = i & 15;
        if (pack8_tile(tp, c) != tile) continue;
        const float v = d < 3 ? new_xyz[((size_t)b * NCENT + base + c) * 3 + d] : 0.f;
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        __nv_bfloat16* o = out_rows + ((size_t)b * NCENT + base + c) * (2 * A3_KX);
        o[256 + d] = h;
        o[A3_KX + 256 + d] = l;
      }
    }
    ntiles_done += tp.ntiles;
    tp = tpn;
    nvalid = nvalid_n;
  }
  if (u == 0 && ntiles_done) atomicAdd(sa_tile_counter(err, 1), (unsigned long long)ntiles_done);
  if (!ok && (threadIdx.x & 31) == 0) atomicExch(err, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------- launchers
static int launch_sa1x3(mpn_ctx* c, cudaStream_t s, const float* cloud, int N, const float* new_xyz, int B, __nv_bfloat16* out_rows,
                        float* out_f32, int32_t* ball_idx) {
  X3Weights& w = g_x3[c];
  MPN_REQUIRE(w.ready, "bf16x3 weights not packed");
  MPN_REQUIRE(N < 65536, "bf16x3 SA1: at most 65535 points");
  const size_t smem = Sa1xSmem::total(N);
  MPN_REQUIRE(smem + 1024 <= 227 * 1024, "bf16x3 SA1: %d points do not fit in shared memory (use MPN_PREC_FP32)", N);
  MPN_CHECK_CUDA(cudaFuncSetAttribute(sa1x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int split = sa_split(c, B, 128 / S1X_NWG);
  sa1x3_tc_kernel<<<B * split, 128 * S1X_NWG, smem, s>>>(cloud, N, new_xyz, SA1_RADIUS * SA1_RADIUS, w.sa1_l1, w.sa1_hi[0], w.sa1_lo[0], w.sa1_hi[1],
                                                         w.sa1_lo[1], out_rows, out_f32, tc_error_flag(c), ball_idx, split, sa_pack());
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// per-point layer-1 pre-activations of SA2: pre [B*512][128] f32 = a1 (split rows, K = 80) x W1p^T + b1
static int launch_sa2_pre(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* a1, int B, float* pre) {
  X3Weights& w = g_x3[c];
  return launch_gemm_tc_ex(c, s, X3_EPI_F32, a1, 2 * A1_K, A1_K, w.sa2_w1p, 2 * A1_K, A1_K, A1_K, c->w.sa[1][0].b, B * SA1_NPOINT, 128, pre,
                           128, 0, 1, nullptr);
}

static int launch_sa2x3(mpn_ctx* c, cudaStream_t s, const float* xyz1, const float* pre, const float* xyz2, int B, __nv_bfloat16* out_rows,
                        float* out_f32, int32_t* ball_idx) {
  X3Weights& w = g_x3[c];
  MPN_REQUIRE(w.ready, "bf16x3 weights not packed");
  if (getenv("MPN_SA2X3_V1") == nullptr) {   // default: 8 warps per chain, rounds of 8 centroids (A/B switch read per launch)
    const int split = sa_split(c, B, 16 / S2X_NWG);
    MPN_CHECK_CUDA(cudaFuncSetAttribute(sa2x3h_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sa2hSmem::total));
    sa2x3h_tc_kernel<<<B * split, S2H_THREADS, Sa2hSmem::total, s>>>(xyz1, 3, pre, xyz2, SA2_RADIUS * SA2_RADIUS, w.sa2_w2, w.sa2_w3, c->w.sa[1][1].b,
                                                                     c->w.sa[1][2].b, w.sa2_w1x, out_rows, out_f32, tc_error_flag(c), ball_idx, split, sa_pack());
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    return MPN_OK;
  }
  const int split = sa_split(c, B, 32 / S2X_NWG);
  MPN_CHECK_CUDA(cudaFuncSetAttribute(sa2x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sa2xSmem::total));
  sa2x3_tc_kernel<<<B * split, S2X_THREADS, Sa2xSmem::total, s>>>(xyz1, 3, pre, xyz2, SA2_RADIUS * SA2_RADIUS, w.sa2_w2, w.sa2_w3, c->w.sa[1][1].b,
                                                                  c->w.sa[1][2].b, w.sa2_w1x, out_rows, out_f32, tc_error_flag(c), ball_idx, split, sa_pack());
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// fp32 rows [rows][cols] (pitch src_stride) -> split rows [rows][2*kpad]: hi at [0, kpad), lo at [kpad, 2 kpad), zero padded
__global__ void split_rows_kernel(const float* __restrict__ src, size_t rows, int src_stride, int cols, int kpad, __nv_bfloat16* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kpad) return;
  size_t r = i / kpad;
  int k = (int)(i % kpad);
  __nv_bfloat16 h, l;
  split_bf16(k < cols ? src[r * src_stride + k] : 0.f, h, l);
  dst[r * 2 * kpad + k] = h;
  dst[r * 2 * kpad + kpad + k] = l;
}
__global__ void add_xyz_cols_kernel(const float* __restrict__ xyz, size_t rows, int kpad, int col0, __nv_bfloat16* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 3) return;
  size_t r = i / 3;
  int d = (int)(i % 3);
  __nv_bfloat16 h, l;
  split_bf16(xyz[i], h, l);
  dst[r * 2 * kpad + col0 + d] = h;
  dst[r * 2 * kpad + kpad + col0 + d] = l;
}

__global__ void split_rows_kernel(const float* __restrict__ src, size_t rows, int src_stride, int cols, int kpad, __nv_bfloat16* __restrict__ dst);

// test entry: C [M][N] f32 = A [M][K] f32 x W [N][K]^T f32 + bias through the TMA GEMM, operands rounded to bf16 (split = 0) or
// split into (hi, lo) bf16 pairs (split = 1).  Allocates its temporaries (synchronises): not a product path.
int x3_gemm_selftest(mpn_ctx* c, cudaStream_t s, const float* A, const float* W, const float* bias, int M, int N, int K, float* C, int split) {
  MPN_REQUIRE(K % 16 == 0 && N % 8 == 0 && M >= 1, "gemm selftest: K %% 16 == 0, N %% 8 == 0 required");
  __nv_bfloat16 *a = nullptr, *w = nullptr;
  MPN_CHECK_CUDA(cudaMalloc(&a, (size_t)M * 2 * K * 2));
  MPN_CHECK_CUDA(cudaMalloc(&w, (size_t)N * 2 * K * 2));
  split_rows_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, s>>>(A, (size_t)M, K, K, K, a);
  split_rows_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, s>>>(W, (size_t)N, K, K, K, w);
  c->launches += 2;
  int r = launch_gemm_tc_ex(c, s, X3_EPI_F32, a, 2 * K, K, w, 2 * K, K, K, bias, M, N, C, N, 0, split, nullptr);
  cudaStreamSynchronize(s);
  cudaFree(a);
  cudaFree(w);
  return r;
}

// per-module entry (tests / mpn_sa_forward with MPN_PREC_BF16X3): fp32 in, fp32 out, split-bf16 inside
int x3_sa_forward(mpn_ctx* c, cudaStream_t s, int module, const float* xyz, int stride, const float* feats, int feat_stride, int B, int N,
                  const float* new_xyz, float* new_feats, int32_t* ball_idx) {
  X3Scratch sc = x3_scratch(c);
  int r;
  if (module == 0) {
    MPN_REQUIRE(stride == 4 && feats == xyz + 3 && feat_stride == 4, "bf16x3 SA1 takes the [B][N][4] cloud (features = 4th column)");
    return launch_sa1x3(c, s, xyz, N, new_xyz, B, sc.a1, new_feats, ball_idx);
  }
  MPN_REQUIRE(module == 1 && N == SA1_NPOINT && stride == 3, "bf16x3 per-module entry supports modules 0 and 1 (N = 512, xyz stride 3 for module 1)");
  const size_t rows = (size_t)B * N;
  split_rows_kernel<<<(unsigned)((rows * A1_K + 255) / 256), 256, 0, s>>>(feats, rows, feat_stride, 64, A1_K, sc.a1);
  add_xyz_cols_kernel<<<(unsigned)((rows * 3 + 255) / 256), 256, 0, s>>>(xyz, rows, A1_K, 64, sc.a1);
  c->launches += 2;
  MPN_CHECK_CUDA(cudaGetLastError());
  if ((r = launch_sa2_pre(c, s, sc.a1, B, sc.pre))) return r;
  return launch_sa2x3(c, s, xyz, sc.pre, new_xyz, B, sc.a3, new_feats, ball_idx);
}

int x3_encoder_forward(mpn_ctx* c, cudaStream_t s, const float* cloud, int B, int N, float* out, int ldo) {
  Workspace& w = c->ws;
  X3Weights& xw = g_x3[c];
  MPN_REQUIRE(xw.ready, "bf16x3 weights not packed");
  X3Scratch sc = x3_scratch(c);
  int r;
  { StageTimer t(c, s, MPN_ST_FPS1);
    if ((r = launch_fps(c, s, cloud, B, N, 4, SA1_NPOINT, reinterpret_cast<int32_t*>(w.fc_a), w.xyz1))) return r; }
  { StageTimer t(c, s, MPN_ST_SA1);
    if ((r = launch_sa1x3(c, s, cloud, N, w.xyz1, B, sc.a1, nullptr, nullptr))) return r; }
  { StageTimer t(c, s, MPN_ST_FPS2);
    if ((r = launch_fps(c, s, w.xyz1, B, SA1_NPOINT, 3, SA2_NPOINT, reinterpret_cast<int32_t*>(w.fc_a), w.xyz2))) return r; }
  { StageTimer t(c, s, MPN_ST_SA2);
    if ((r = launch_sa2_pre(c, s, sc.a1, B, sc.pre))) return r;
    if ((r = launch_sa2x3(c, s, w.xyz1, sc.pre, w.xyz2, B, sc.a3, nullptr, nullptr))) return r; }
  const int M3 = B * SA2_NPOINT;
  { StageTimer t(c, s, MPN_ST_SA3);   // group-all module: three row-shared split GEMMs, the last one pools each problem's 128 rows
    if ((r = launch_gemm_tc_ex(c, s, X3_EPI_RELU_SPLIT, sc.a3, 2 * A3_KX, A3_KX, xw.sa3[0], 2 * A3_KX, A3_KX, A3_KX, c->w.sa[2][0].b, M3, 512,
                               sc.h1, 1024, 512, 1, nullptr))) return r;
    if ((r = launch_gemm_tc_ex(c, s, X3_EPI_RELU_SPLIT, sc.h1, 1024, 512, xw.sa3[1], 1024, 512, 512, c->w.sa[2][1].b, M3, 512, sc.h2, 1024, 512, 1,
                               nullptr))) return r;
    if ((r = launch_gemm_tc_ex(c, s, X3_EPI_MAXPOOL_SPLIT, sc.h2, 1024, 512, xw.sa3[2], 1024, 512, 512, c->w.sa[2][2].b, M3, 1024, sc.f3, 2048, 1024,
                               1, nullptr))) return r; }
  StageTimer tfc(c, s, MPN_ST_FC);
  if (B <= SKINNY_MAX_ROWS) {   // a handful of problems: fp32 weight streaming on every SM instead of N / 256 tensor-core tiles
    if ((r = launch_linear_skinny(c, s, sc.f3, 2048, 2, c->w.fc[0], B, w.fc_a, 4096, 0))) return r;
    if ((r = launch_groupnorm_lrelu(c, s, w.fc_a, B, 4096, 16, c->w.gn_w[0], c->w.gn_b[0]))) return r;
    if ((r = launch_linear_skinny(c, s, w.fc_a, 4096, 0, c->w.fc[1], B, w.fc_b, 2048, 0))) return r;
    if ((r = launch_groupnorm_lrelu(c, s, w.fc_b, B, 2048, 16, c->w.gn_w[1], c->w.gn_b[1]))) return r;
    return launch_linear_skinny(c, s, w.fc_b, 2048, 0, c->w.fc[2], B, out, ldo, 0);
  }
  if ((r = launch_gemm_tc_ex(c, s, X3_EPI_F32, sc.f3, 2048, 1024, xw.fc[0], 2048, 1024, 1024, c->w.fc[0].b, B, 4096, w.fc_a, 4096, 0, 1, nullptr))) return r;
  if ((r = launch_groupnorm_lrelu_split(c, s, w.fc_a, B, 4096, 16, c->w.gn_w[0], c->w.gn_b[0], sc.g1))) return r;
  if ((r = launch_gemm_tc_ex(c, s, X3_EPI_F32, sc.g1, 8192, 4096, xw.fc[1], 8192, 4096, 4096, c->w.fc[1].b, B, 2048, w.fc_b, 2048, 0, 1, nullptr))) return r;
  if ((r = launch_groupnorm_lrelu_split(c, s, w.fc_b, B, 2048, 16, c->w.gn_w[1], c->w.gn_b[1], sc.g2))) return r;
  return launch_gemm_tc_ex(c, s, X3_EPI_F32, sc.g2, 4096, 2048, xw.fc[2], 4096, 2048, 2048, c->w.fc[2].b, B, 2048, out, ldo, 0, 1, nullptr);
}

// decoder.0 (2112 -> 512, LeakyReLU; model.py:58-60) of the tensor-core modes: operand rows prepared by feature_encoder_kernel
int tc_decoder0(mpn_ctx* c, cudaStream_t s, int precision, const __nv_bfloat16* operand, int B, float* h0) {
  X3Weights& xw = g_x3[c];
  MPN_REQUIRE(xw.ready, "tensor-core weights not packed");
  constexpr int K = ENC_DIM + QF_DIM;
  const Linear& L = c->w.dec[0];
  if (precision == MPN_PREC_BF16)
    return launch_gemm_tc_ex(c, s, X3_EPI_LRELU_F32, operand, K, 0, xw.dec0, 2 * K, 0, K, L.b, B, 512, h0, 512, 0, 0, nullptr);
  return launch_gemm_tc_ex(c, s, X3_EPI_LRELU_F32, operand, 2 * K, K, xw.dec0, 2 * K, K, K, L.b, B, 512, h0, 512, 0, 1, nullptr);
}

}  // namespace mpn
