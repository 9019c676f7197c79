// sa_tc.cu -- bf16 tcgen05 tensor-core path of the PointNet++ encoder (throughput mode, MPN_PREC_BF16).
//
// Replaces the same reference code as sa_simt.cu (pointnet2_ops QueryAndGroup + Conv2d1x1/ReLU x3 + max_pool2d inside
// PointnetSAModule, mpinets/model.py:365-383) with ONE fused kernel per set-abstraction level (sa1_tc_kernel,
// sa2_tc_kernel); the group-all level and the FC head run on the tcgen05 GEMM of gemm_tc.cu.
//
//   CTA = one problem.  Row warpgroups (128 threads = the 128 neighbour rows = the 128 TMEM lanes) stream centroids
//   independently, so one group's SIMT phases overlap another group's MMAs on the SM's single tensor pipe.
//   per centroid:  ball query (pointnet2 semantics kept bit-exact: first 128 hits in index order, first-hit padding)
//     -> gather the neighbours' rows straight into the UMMA operand layout in shared memory (bf16, K-major 8x16 B core
//        matrices, bias carried by a column of ones)  -> tcgen05.mma (accumulator in TMEM)  -> epilogue: tcgen05.ld,
//        cvt.rn.relu.bf16x2, st.shared over the SAME buffer (the previous operand is dead once its MMA retired)
//     -> layer 2 -> layer 3 -> max over the 128 neighbours -> one pooled bf16 row to HBM.
//   SA2: last layer transposed (TMEM lane = channel) so the max-pool is a per-thread reduction; ball query and feature
//        prefetch of the next centroid are issued while layer 3 runs; MMAs are issued by a dedicated warp per group.
//   SA1: exact hash-grid ball query on a producer warp per group (2-slot list ring, full/empty mbarriers); integer
//        redux.sync max-pool.
//   K order of layer 1 is [features..., dx, dy, dz, 1, 0-pad] (a permutation of the reference's [dx,dy,dz,features]
//   applied to both operand and weight) so feature rows are copied as aligned 16-byte chunks.
#include <cstdlib>

#include "engine.h"
#include "spec_math.cuh"
#include "tc_common.cuh"

namespace mpn {
using namespace tc;

// ---------------------------------------------------------------------------------------------- weight packing
constexpr int SA1_XK = 80, SA1_BIAS_COL = 64;   // SA1 operand rows: [64 values | 1.0 | 0 x15]
struct TcWeights {
  __nv_bfloat16* sa[3][3] = {{nullptr}};  // [Cout][Kpad] K-major, layer-0 K order permuted
  int kpad[3][3] = {{0}};
  __nv_bfloat16* fc[3] = {nullptr};       // [out][in]
  __nv_bfloat16* w2_nofold = nullptr;     // SA2 layer 2 as plain [128][128] (3-warpgroup variant adds the bias in the epilogue)
};
static std::map<mpn_ctx*, TcWeights> g_tc;
int* tc_error_flag(mpn_ctx* c);
int x3_pack_weights(mpn_ctx* c, cudaStream_t s);   // sa_x3.cu: the split-bf16 copies are refreshed together with the bf16 ones

// bias_col >= 0: the layer's bias is folded into the GEMM as K-column `bias_col` (the operand carries a 1.0 there)
__global__ void pack_weight_kernel(const float* __restrict__ w, int out, int in, int kpad, int rot, __nv_bfloat16* __restrict__ dst,
                                   const float* __restrict__ bias = nullptr, int bias_col = -1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out * kpad) return;
  int o = i / kpad, k = i % kpad;
  float v = 0.f;
  if (k < in) {
    int src = rot ? (k < in - 3 ? k + 3 : k - (in - 3)) : k;  // rot: [f..., dx,dy,dz] <- [dx,dy,dz, f...]
    v = w[(size_t)o * in + src];
  } else if (k == bias_col) {
    v = bias[o];
  }
  dst[i] = __float2bfloat16_rn(v);
}

// (re)pack the bf16 operand copies from the fp32 master weights on stream s; buffers are allocated on first use
static int tc_pack_weights(mpn_ctx* c, cudaStream_t s, bool sa_only) {
  TcWeights& t = g_tc[c];
  for (int m = 0; m < 3; ++m)
    for (int l = 0; l < 3; ++l) {
      const Linear& L = c->w.sa[m][l];
      int kpad = m == 0 ? SA1_XK : (L.in + 15) / 16 * 16;   // SA1: every layer is [64][80] with the bias in column 64
      int bias_col = m == 0 ? SA1_BIAS_COL : -1;
      if (m == 1 && l == 0) bias_col = 67;                  // SA2 layer 1: [f0..f63, dx, dy, dz, bias, 0 x12]
      if (m == 1 && l == 1) { kpad = 144; bias_col = 128; } // SA2 layer 2: [128 weights, bias, 0 x15]
      t.kpad[m][l] = kpad;
      if (!t.sa[m][l]) MPN_CHECK_CUDA(cudaMalloc(&t.sa[m][l], (size_t)L.out * kpad * sizeof(__nv_bfloat16)));
      int n = L.out * kpad;
      pack_weight_kernel<<<(n + 255) / 256, 256, 0, s>>>(L.w, L.out, L.in, kpad, (l == 0 && m > 0) ? 1 : 0, t.sa[m][l],
                                                         bias_col >= 0 ? L.b : nullptr, bias_col);
      c->launches++;
      MPN_CHECK_CUDA(cudaGetLastError());
    }
  {
    const Linear& L = c->w.sa[1][1];
    if (!t.w2_nofold) MPN_CHECK_CUDA(cudaMalloc(&t.w2_nofold, (size_t)L.out * L.in * sizeof(__nv_bfloat16)));
    pack_weight_kernel<<<(L.out * L.in + 255) / 256, 256, 0, s>>>(L.w, L.out, L.in, L.in, 0, t.w2_nofold);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  if (sa_only) return MPN_OK;
  for (int l = 0; l < 3; ++l) {
    const Linear& L = c->w.fc[l];
    if (!t.fc[l]) MPN_CHECK_CUDA(cudaMalloc(&t.fc[l], (size_t)L.out * L.in * sizeof(__nv_bfloat16)));
    int n = L.out * L.in;
    pack_weight_kernel<<<(n + 255) / 256, 256, 0, s>>>(L.w, L.out, L.in, L.in, 0, t.fc[l]);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  return x3_pack_weights(c, s);
}

void tc_free(mpn_ctx* c) {
  auto it = g_tc.find(c);
  if (it == g_tc.end()) return;
  TcWeights& t = it->second;
  for (int m = 0; m < 3; ++m)
    for (int l = 0; l < 3; ++l) if (t.sa[m][l]) cudaFree(t.sa[m][l]);
  for (int l = 0; l < 3; ++l) if (t.fc[l]) cudaFree(t.fc[l]);
  if (t.w2_nofold) cudaFree(t.w2_nofold);
  g_tc.erase(it);
}

int tc_prepare_weights(mpn_ctx* c) {
  int r = tc_pack_weights(c, 0, false);
  if (r) return r;
  MPN_CHECK_CUDA(cudaDeviceSynchronize());
  return MPN_OK;
}

// after an optimizer step: refresh every bf16 copy on the training stream (no allocation, no host sync)
int tc_refresh_weights(mpn_ctx* c, cudaStream_t s) { return tc_pack_weights(c, s, false); }

// scratch: feat1 bf16 [B][512][64] | SA3 operand rows bf16 [B*128][272] ([256 feats, x, y, z, 0-pad])
constexpr int A3_K = 272;
static size_t feat1_bytes(int B) { return ((size_t)B * SA1_NPOINT * 64 * sizeof(__nv_bfloat16) + 1023) / 1024 * 1024; }
static size_t a3_bytes(int B) { return ((size_t)B * SA2_NPOINT * A3_K * 2 + 1023) / 1024 * 1024; }
static size_t h_bytes(int B) { return (size_t)B * SA2_NPOINT * 512 * 2; }     // SA3 hidden activations [B*128][512] bf16
// layout: feat1 | a3 | h1 | h2 | feat3 bf16 [B][1024] | g1 bf16 [B][4096] | g2 bf16 [B][2048]
size_t tc_scratch_bytes(int B) {
  return feat1_bytes(B) + a3_bytes(B) + 2 * h_bytes(B) + (size_t)B * (1024 + 4096 + 2048) * 2 + 1024;
}

// ---------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ void wg_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// global [rows][K] bf16 (K-major) -> smem interleaved core-matrix layout
__device__ __forceinline__ void stage_weight(const __nv_bfloat16* __restrict__ g, int rows, int K, uint8_t* s) {
  const int KC = K / 8;
  for (int i = threadIdx.x; i < rows * KC; i += blockDim.x) {
    int r = i / KC, kc = i - r * KC;
    uint4 v = __ldg(reinterpret_cast<const uint4*>(g + (size_t)r * K + kc * 8));
    *reinterpret_cast<uint4*>(s + kmajor_chunk_off(r, kc, KC)) = v;
  }
}

// descriptor of K-step ks (16 elements) of a [rows][K] interleaved tile starting at row r0
__device__ __forceinline__ uint64_t tile_desc(uint32_t base, int K, int r0, int ks) {
  const int KC = K / 8;
  return make_smem_desc(base + (uint32_t)((r0 >> 3) * KC * 128 + ks * 256), 128, KC * 128, LAYOUT_NONE);
}

// ---------------------------------------------------------------------------------------------- shared helpers of the SA kernels
__device__ __forceinline__ uint32_t cvt_relu_bf16x2(float first, float second) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(second), "f"(first));
  return d;
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
// TMEM [128 lanes][NC cols] fp32 (bias already inside the accumulator) -> relu -> bf16 -> chunks 0..NC/8-1 of row `row`
template <int NC, int KCX>
__device__ __forceinline__ void epilogue_pack_relu(uint32_t taddr, uint8_t* X, int row) {
#pragma unroll
  for (int c0 = 0; c0 < NC; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(X + kmajor_chunk_off(row, (c0 >> 3) + q, KCX)) =
          make_uint4(cvt_relu_bf16x2(__uint_as_float(v[q * 8]), __uint_as_float(v[q * 8 + 1])),
                     cvt_relu_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])),
                     cvt_relu_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])),
                     cvt_relu_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])));
  }
}

// ---------------------------------------------------------------------------------------------- SA2, 3 warpgroups
// Variant of sa2_tc_kernel with THREE row warpgroups per CTA (3 centroids in flight instead of 2).  To fit shared memory
// the operand rows shrink to 128 K-columns (32 KB per group): layer 1 keeps its bias in the free column 67, layer 2 adds
// its bias in the epilogue (fp32), and the ball query is done by one warp per centroid over the smem-resident xyz
// (index order falls out of the in-order scan, no merge lists).  TMEM: 128 columns per group, so the transposed last
// layer runs as two sequential channel tiles.
constexpr int SA2W_NWG = 3, SA2W_KC = 16, SA2W_THREADS = 128 * SA2W_NWG;
struct Sa2wSmem {
  static constexpr size_t w1 = 0;                                     // [128][80]
  static constexpr size_t w2 = w1 + 128 * 80 * 2;                     // [128][128]
  static constexpr size_t w3 = w2 + 128 * 128 * 2;                    // [256][128]
  static constexpr size_t x = w3 + 256 * 128 * 2;                     // NWG x [128][128] bf16
  static constexpr size_t b2 = x + (size_t)SA2W_NWG * 128 * 128 * 2;  // [128] float
  static constexpr size_t b3 = b2 + 128 * 4;                          // [256] float
  static constexpr size_t pts = b3 + 256 * 4;                         // x[512] | y[512] | z[512]
  static constexpr size_t lists = pts + 3 * SA1_NPOINT * 4;           // [NWG][2 slots][4][128] u16
  static constexpr size_t bars = lists + (size_t)SA2W_NWG * 2 * 4 * 128 * 2;   // NWG mbarriers + tmem slot
  static constexpr size_t total = bars + 64;
};

template <bool ARG>
__global__ void __launch_bounds__(SA2W_THREADS, 1)
sa2w3_tc_kernel(const float* __restrict__ xyz, int stride, const __nv_bfloat16* __restrict__ feat_bf16, const float* __restrict__ new_xyz,
                float r2, const __nv_bfloat16* __restrict__ gw1, const __nv_bfloat16* __restrict__ gw2, const __nv_bfloat16* __restrict__ gw3,
                const float* __restrict__ gb2, const float* __restrict__ gb3, __nv_bfloat16* __restrict__ out_bf16, int out_stride,
                int* __restrict__ err, int32_t* __restrict__ ball_idx, uint8_t* __restrict__ arg_out, int split, int pack) {
  using S = Sa2wSmem;
  constexpr int N = SA1_NPOINT, NCENT = SA2_NPOINT, KC = SA2W_KC;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW1 = smem + S::w1;
  uint8_t* sW2 = smem + S::w2;
  uint8_t* sW3 = smem + S::w3;
  float* sB2 = reinterpret_cast<float*>(smem + S::b2);
  float* sB3 = reinterpret_cast<float*>(smem + S::b3);
  float* px = reinterpret_cast<float*>(smem + S::pts);
  float* py = px + N;
  float* pz = py + N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bars + 8 * SA2W_NWG);

  const int b = blockIdx.x / split, part = blockIdx.x % split;   // small batches: a problem's centroid rounds are dealt to `split` CTAs
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int g = warp >> 2, wq = warp & 3, lane = threadIdx.x & 31;
  const int t = threadIdx.x & 127;
  uint8_t* X = smem + S::x + (size_t)g * 128 * 128 * 2;
  uint16_t* lists = reinterpret_cast<uint16_t*>(smem + S::lists) + (size_t)g * 2 * 4 * 128;

  stage_weight(gw1, 128, 80, sW1);
  stage_weight(gw2, 128, 128, sW2);
  stage_weight(gw3, 256, 128, sW3);
  for (int i = threadIdx.x; i < 128; i += SA2W_THREADS) sB2[i] = gb2[i];
  for (int i = threadIdx.x; i < 256; i += SA2W_THREADS) sB3[i] = gb3[i];
  {
    const float* p = xyz + (size_t)b * N * stride;
    for (int k = threadIdx.x; k < N; k += SA2W_THREADS) {
      px[k] = __ldg(p + (size_t)k * stride); py[k] = __ldg(p + (size_t)k * stride + 1); pz[k] = __ldg(p + (size_t)k * stride + 2);
    }
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < SA2W_NWG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + (uint32_t)g * 128;
  const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
  const uint64_t dX = make_smem_desc(smem_u32(X), 128, KC * 128, LAYOUT_NONE);
  const uint64_t dW1 = make_smem_desc(smem_u32(sW1), 128, 10 * 128, LAYOUT_NONE);
  const uint64_t dW2 = make_smem_desc(smem_u32(sW2), 128, 16 * 128, LAYOUT_NONE);
  const uint64_t dW3 = make_smem_desc(smem_u32(sW3), 128, 16 * 128, LAYOUT_NONE);
  constexpr uint32_t W3_TILE1 = (128 / 8) * 16 * 128 / 16;
  constexpr uint32_t ID128 = make_idesc_bf16(128, 128);
  uint64_t* bar = &bars[g];
  uint32_t phase = 0;
  bool ok = true;
  const unsigned lt = (1u << lane) - 1u;
  __shared__ int hcnt_s[SA2W_NWG * 2 * 4];   // [NWG][2 slots][4]: distinct hits (<= 128) of a round's ball queries
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
  // A round's four neighbourhoods are packed into as few 128-row tiles as their distinct rows need (TilePack, tc_common.cuh); the rows
  // of the next tile (coordinates + 64 features) are fetched into registers under the current tile's layer-3 MMAs.
  constexpr int STEP = SA2W_NWG * 4;
  float dx, dy, dz;
  uint4 fr[8];
  TilePack tp{0u, 0u, 0}, tpn{0u, 0u, 0};
  int nvalid = 0, nvalid_n = 0, ntiles_done = 0;
  auto open_round = [&](int nb, int nslot, TilePack& p, int& nv) {
    nv = min(4, NCENT - nb);
    p = pack_round(hcnt + nslot * 4, nv, pack);
    if (ball_idx)
      for (int c = 0; c < nv; ++c) ball_idx[((size_t)b * NCENT + nb + c) * NSAMPLE + t] = lists[(nslot * 4 + c) * 128 + t];
  };
  auto prefetch = [&](const TilePack& p, int nv, int nb, int nslot, int ntile) {
    const int mc = pack_owner(p, nv, ntile, wq);
    const float* cp = new_xyz + ((size_t)b * NCENT + nb + mc) * 3;
    const int k = lists[(nslot * 4 + mc) * 128 + (wq - pack_q0(p, mc)) * 32 + lane];
    dx = fsub(px[k], cp[0]); dy = fsub(py[k], cp[1]); dz = fsub(pz[k], cp[2]);
    const uint4* f = reinterpret_cast<const uint4*>(feat_bf16 + ((size_t)b * N + k) * 64);
#pragma unroll
    for (int q = 0; q < 8; ++q) fr[q] = __ldg(f + q);
  };
  auto issue = [&](uint64_t da, uint32_t aoff, uint64_t db, uint32_t boff, int ksteps) {
    if (wq == 0) {
      tc_fence_after();
      if (elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks) mma_bf16_ss_off(tmem, da, aoff + ks * 16, db, boff + ks * 16, ID128, ks > 0);
        mma_commit(bar);
      }
      __syncwarp();
    }
  };

  // centroids of warpgroup g: rounds of 4 consecutive centroids, rounds interleaved over the warpgroups
  int r = 0;
  const int base0 = (g * split + part) * 4;
  if (base0 < NCENT) {
    bq_round(base0, 0);
    wg_sync(g);
    open_round(base0, 0, tp, nvalid);
    prefetch(tp, nvalid, base0, 0, 0);
  }
  for (int base = base0; base < NCENT && ok; base += STEP * split, ++r) {
    const int slot = r & 1;
#pragma unroll 1
    for (int tile = 0; tile < tp.ntiles && ok; ++tile) {
#pragma unroll
      for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(X + kmajor_chunk_off(t, q, KC)) = fr[q];
      *reinterpret_cast<uint4*>(X + kmajor_chunk_off(t, 8, KC)) = make_uint4(pack_bf16(dx, dy), pack_bf16(dz, 1.0f), 0u, 0u);
      *reinterpret_cast<uint4*>(X + kmajor_chunk_off(t, 9, KC)) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
      tc_fence_before();
      wg_sync(g);
      issue(dX, 0, dW1, 0, 5);                                  // layer 1: K = 80 (bias in column 67)
      ok = mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      epilogue_pack_relu<128, KC>(tlane, X, t);
      fence_proxy_async_smem();
      tc_fence_before();
      wg_sync(g);
      issue(dX, 0, dW2, 0, 8);                                  // layer 2: K = 128, bias added below
      ok = ok && mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tlane + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* bb = sB2 + c0 + q * 8;
          *reinterpret_cast<uint4*>(X + kmajor_chunk_off(t, (c0 >> 3) + q, KC)) =
              make_uint4(cvt_relu_bf16x2(__uint_as_float(v[q * 8]) + bb[0], __uint_as_float(v[q * 8 + 1]) + bb[1]),
                         cvt_relu_bf16x2(__uint_as_float(v[q * 8 + 2]) + bb[2], __uint_as_float(v[q * 8 + 3]) + bb[3]),
                         cvt_relu_bf16x2(__uint_as_float(v[q * 8 + 4]) + bb[4], __uint_as_float(v[q * 8 + 5]) + bb[5]),
                         cvt_relu_bf16x2(__uint_as_float(v[q * 8 + 6]) + bb[6], __uint_as_float(v[q * 8 + 7]) + bb[7]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      wg_sync(g);
      issue(dW3, 0, dX, 0, 8);                                  // layer 3, channel tile 0 (transposed)
      // next tile: its ball-query round (if this was the last tile of the round) and its rows, under the MMAs
      if (tile + 1 < tp.ntiles) {
        prefetch(tp, nvalid, base, slot, tile + 1);
      } else {
        const int nb = base + STEP * split;
        if (nb < NCENT) { bq_round(nb, slot ^ 1); wg_sync(g); open_round(nb, slot ^ 1, tpn, nvalid_n); prefetch(tpn, nvalid_n, nb, slot ^ 1, 0); }
      }
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        ok = ok && mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        // transposed accumulator: lane = channel, column = tile row; the max over each quarter's 32 columns
        float mq[4];
#pragma unroll
        for (int q = 0; q < 4; q += 2) {
          uint32_t v[32], u[32];
          tmem_ld32(tlane + q * 32, v);
          tmem_ld32(tlane + q * 32 + 32, u);
          tmem_ld_wait();
          float mv = __uint_as_float(v[0]), mu = __uint_as_float(u[0]);
#pragma unroll
          for (int i = 1; i < 32; ++i) { mv = fmaxf(mv, __uint_as_float(v[i])); mu = fmaxf(mu, __uint_as_float(u[i])); }
          mq[q] = mv;
          mq[q + 1] = mu;
        }
        for (int c = 0; c < nvalid; ++c) {
          if (pack_tile(tp, c) != tile) continue;
          const int q0 = pack_q0(tp, c), q1 = pack_q1(tp, nvalid, c);
          float m = -3.0e38f;
#pragma unroll
          for (int q = 0; q < 4; ++q) m = (q >= q0 && q < q1) ? fmaxf(m, mq[q]) : m;
          out_bf16[((size_t)b * NCENT + base + c) * out_stride + half * 128 + t] = __float2bfloat16_rn(fmaxf(m + sB3[half * 128 + t], 0.f));
          if constexpr (ARG) {   // training forward: the winning row of this channel (first row of the centroid's list attaining the maximum)
            int arg = 0;
            for (int q = q1 - 1; q >= q0; --q) {
              uint32_t v[32];
              tmem_ld32(tlane + q * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 31; i >= 0; --i) arg = __uint_as_float(v[i]) == m ? (q - q0) * 32 + i : arg;
            }
            arg_out[((size_t)b * NCENT + base + c) * 256 + half * 128 + t] = (uint8_t)arg;
          }
        }
        tc_fence_before();
        wg_sync(g);                                             // every lane of the accumulator has been read
        if (half == 0) issue(dW3, W3_TILE1, dX, 0, 8);           // channel tile 1 into the same TMEM columns
      }
      for (int i = t; i < nvalid * 16; i += 128) {   // [x y z | 0-pad] columns of the tile's centroids
        const int c = i >> 4, d = i & 15;
        if (pack_tile(tp, c) != tile) continue;
        const float v = d < 3 ? new_xyz[((size_t)b * NCENT + base + c) * 3 + d] : 0.f;
        out_bf16[((size_t)b * NCENT + base + c) * out_stride + 256 + d] = __float2bfloat16_rn(v);
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

// ---------------------------------------------------------------------------------------------- SA1
constexpr int SA1_BUCKETS = 4096;

// uniform hash grid: cell edge slightly above the query radius so that every point within r of a centroid lies in
// one of the 27 cells around the centroid's cell even under fp32 rounding of the cell coordinates
__device__ __forceinline__ int grid_coord(float v) { return (int)floorf((v + 8.0f) * (1.0f / 0.0501f)); }
__device__ __forceinline__ uint32_t grid_bucket(int ix, int iy, int iz) {
  return ((uint32_t)ix * 73856093u ^ (uint32_t)iy * 19349663u ^ (uint32_t)iz * 83492791u) & (SA1_BUCKETS - 1);
}

// ---------------------------------------------------------------------------------------------- SA1, 6 warpgroups
// Variant of sa1_tc_kernel with SIX row warpgroups per CTA (6 centroids in flight instead of 4).  Shared memory is freed
// by keeping only the hash grid's index structure on chip (bucket starts + point indices); candidate coordinates are
// gathered from the problem's cloud in L2.  Every warp runs the complete ball query of one centroid of its group's next
// round (cells -> candidates -> hits -> rank by point index), so there are no block-level syncs inside the query.
template <int SA1W_NWG>
struct Sa1wSmem {
  static constexpr size_t w = 0;                                           // 3 x [64][80] bf16
  static constexpr size_t x = w + 3 * 64 * SA1_XK * 2;                     // NWG x [128][80] bf16 (aliased by the grid build)
  static constexpr size_t lists = x + (size_t)SA1W_NWG * 128 * SA1_XK * 2; // [NWG][4][128] u16
  static constexpr size_t cand = lists + SA1W_NWG * 4 * 128 * 2;           // [NWG][4][256] u16
  static constexpr size_t red = cand + SA1W_NWG * 4 * 256 * 2;             // [NWG][4][64] int
  static constexpr size_t bars = red + SA1W_NWG * 4 * 64 * 4;              // NWG mbarriers + tmem slot
  static constexpr size_t bstart = (bars + 64 + 15) / 16 * 16;             // u16 [BUCKETS + 1]
  static constexpr size_t sidx = (bstart + (SA1_BUCKETS + 1) * 2 + 15) / 16 * 16;   // u16 [N]
  __host__ __device__ static size_t cxyz(int N) { return (sidx + (size_t)N * 2 + 64 + 15) / 16 * 16; }   // f32 [512][3] centroid coordinates
  static size_t total(int N) { return cxyz(N) + (size_t)SA1_NPOINT * 3 * 4; }
};

template <int SA1W_NWG, bool ARG = false>
__global__ void __launch_bounds__(128 * SA1W_NWG, 1)
sa1w_tc_kernel(const float* __restrict__ cloud, int N, const float* __restrict__ new_xyz, float r2, const __nv_bfloat16* __restrict__ gw1,
                const __nv_bfloat16* __restrict__ gw2, const __nv_bfloat16* __restrict__ gw3, __nv_bfloat16* __restrict__ out_bf16,
                int* __restrict__ err, int32_t* __restrict__ ball_idx, uint8_t* __restrict__ arg_out = nullptr) {
  using S = Sa1wSmem<SA1W_NWG>;
  constexpr int KC = SA1_XK / 8, NS = NSAMPLE;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW1 = smem + S::w;
  uint8_t* sW2 = sW1 + 64 * SA1_XK * 2;
  uint8_t* sW3 = sW2 + 64 * SA1_XK * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bars + 8 * SA1W_NWG);
  uint16_t* bstart = reinterpret_cast<uint16_t*>(smem + S::bstart);
  uint16_t* sidx = reinterpret_cast<uint16_t*>(smem + S::sidx);

  const int b = blockIdx.x;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int g = warp >> 2, wq = warp & 3, t = threadIdx.x & 127, lane = threadIdx.x & 31;
  uint8_t* X = smem + S::x + (size_t)g * 128 * SA1_XK * 2;
  uint16_t* lists = reinterpret_cast<uint16_t*>(smem + S::lists) + (size_t)g * 4 * 128;
  uint16_t* wcand = reinterpret_cast<uint16_t*>(smem + S::cand) + (size_t)(g * 4 + wq) * 256;
  int* red = reinterpret_cast<int*>(smem + S::red) + g * 4 * 64;
  const float4* cl = reinterpret_cast<const float4*>(cloud) + (size_t)b * N;

  stage_weight(gw1, 64, SA1_XK, sW1);
  stage_weight(gw2, 64, SA1_XK, sW2);
  stage_weight(gw3, 64, SA1_XK, sW3);
  float* cxyz = reinterpret_cast<float*>(smem + S::cxyz(N));
  for (int i = threadIdx.x; i < SA1_NPOINT * 3; i += blockDim.x) cxyz[i] = __ldg(new_xyz + (size_t)b * SA1_NPOINT * 3 + i);
  // ---- hash-grid build: counting sort of point indices by bucket (counters alias the operand buffers)
  {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + S::x);
    __shared__ uint32_t wsum[16];
    for (int i = threadIdx.x; i < SA1_BUCKETS; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const float4 v = __ldg(cl + k);
      atomicAdd(&cnt[grid_bucket(grid_coord(v.x), grid_coord(v.y), grid_coord(v.z))], 1u);
    }
    __syncthreads();
    constexpr int PER = SA1_BUCKETS / 512;
    const bool scanner = threadIdx.x < 512;
    uint32_t loc[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { loc[i] = scanner ? cnt[threadIdx.x * PER + i] : 0u; sum += loc[i]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31 && scanner) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t v = lane < 16 ? wsum[lane] : 0u, iv = v;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += y; }
      if (lane < 16) wsum[lane] = iv - v;
    }
    __syncthreads();
    if (scanner) {
      uint32_t run = wsum[threadIdx.x >> 5] + inc - sum;
#pragma unroll
      for (int i = 0; i < PER; ++i) { bstart[threadIdx.x * PER + i] = (uint16_t)run; cnt[threadIdx.x * PER + i] = run; run += loc[i]; }
    }
    if (threadIdx.x == 0) bstart[SA1_BUCKETS] = (uint16_t)N;
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const float4 v = __ldg(cl + k);
      sidx[atomicAdd(&cnt[grid_bucket(grid_coord(v.x), grid_coord(v.y), grid_coord(v.z))], 1u)] = (uint16_t)k;
    }
    __syncthreads();
  }
  *reinterpret_cast<uint4*>(X + kmajor_chunk_off(t, 8, KC)) = make_uint4(0x00003F80u, 0u, 0u, 0u);
  *reinterpret_cast<uint4*>(X + kmajor_chunk_off(t, 9, KC)) = make_uint4(0u, 0u, 0u, 0u);
  __shared__ int round_ctr[1 + 8];
  int* next_round = round_ctr;
  int* rsel = round_ctr + 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < SA1W_NWG; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
    *next_round = SA1W_NWG;
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot + (uint32_t)g * 64;
  const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
  const uint64_t dX = tile_desc(smem_u32(X), SA1_XK, 0, 0), dW1 = tile_desc(smem_u32(sW1), SA1_XK, 0, 0);
  const uint64_t dW2 = tile_desc(smem_u32(sW2), SA1_XK, 0, 0), dW3 = tile_desc(smem_u32(sW3), SA1_XK, 0, 0);
  // layer 1 as ONE K = 16 MMA: its two core matrices along K are chunk 0 (dx dy dz w) and chunk 8 (the ones column that carries
  // the bias) -- a leading-dimension byte offset of 8 chunks instead of 1.  Saves the zero chunk, a second MMA and their
  // shared-memory traffic (the kernel is bound by shared-memory bandwidth: operand reads of the N = 64 MMAs + epilogue stores)
  const uint64_t dX1 = make_smem_desc(smem_u32(X), 8 * 128, KC * 128, LAYOUT_NONE);
  const uint64_t dW1k = make_smem_desc(smem_u32(sW1), 8 * 128, KC * 128, LAYOUT_NONE);
  uint64_t* bar = &bars[g];
  uint32_t phase = 0;
  bool ok = true;
  constexpr uint32_t IDESC = make_idesc_bf16(128, 64);
  const unsigned lt = (1u << lane) - 1u;

  // complete ball query of centroid jc by this warp -> lists[wq][0..127]
  auto warp_ball_query = [&](int jc) {
    uint16_t* widx = lists + wq * 128;
    const float* cpw = cxyz + jc * 3;
    const float qx = cpw[0], qy = cpw[1], qz = cpw[2];
    const int ix = grid_coord(qx), iy = grid_coord(qy), iz = grid_coord(qz);
    uint32_t bk = 0x10000u + lane;
    if (lane < 27) bk = grid_bucket(ix + (lane % 3) - 1, iy + ((lane / 3) % 3) - 1, iz + (lane / 9) - 1);
    const unsigned peers = __match_any_sync(0xffffffffu, bk);
    const bool leader = lane < 27 && lane == __ffs(peers) - 1;
    int s0 = 0, n0 = 0;
    if (leader) { s0 = bstart[bk]; n0 = bstart[bk + 1] - s0; }
    int incl = n0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int C = __shfl_sync(0xffffffffu, incl, 31);
    if (C <= 256) {
      {   // candidates by ORIGINAL point index
        const int o = incl - n0;
        int i = 0;
        for (; i + 4 <= n0; i += 4) {
          const uint16_t a0 = sidx[s0 + i], a1 = sidx[s0 + i + 1], a2 = sidx[s0 + i + 2], a3 = sidx[s0 + i + 3];
          wcand[o + i] = a0; wcand[o + i + 1] = a1; wcand[o + i + 2] = a2; wcand[o + i + 3] = a3;
        }
        for (; i < n0; ++i) wcand[o + i] = sidx[s0 + i];
      }
      __syncwarp();
      int H = 0;
      for (int c0 = 0; c0 < C; c0 += 64) {   // 2 x 32 candidates with their L2 loads in flight together
        const int ci0 = c0 + lane, ci1 = c0 + 32 + lane;
        const int k0 = ci0 < C ? (int)wcand[ci0] : 0, k1 = ci1 < C ? (int)wcand[ci1] : 0;
        const float4 v0 = __ldg(cl + k0);
        float4 v1 = v0;
        if (c0 + 32 < C) v1 = __ldg(cl + k1);
        const bool hit0 = ci0 < C && dist2(qx, qy, qz, v0.x, v0.y, v0.z) < r2;
        const bool hit1 = ci1 < C && dist2(qx, qy, qz, v1.x, v1.y, v1.z) < r2;
        __syncwarp();
        const unsigned hm0 = __ballot_sync(0xffffffffu, hit0), hm1 = __ballot_sync(0xffffffffu, hit1);
        if (hit0) wcand[H + __popc(hm0 & lt)] = (uint16_t)k0;         // in place: write position <= read position of later batches
        H += __popc(hm0);
        if (hit1) wcand[H + __popc(hm1 & lt)] = (uint16_t)k1;
        H += __popc(hm1);
        __syncwarp();
      }
      for (int h = lane; h < H; h += 32) {
        const int my = wcand[h];
        int rank = 0;
        for (int i = 0; i < H; ++i) rank += wcand[i] < my;
        if (rank < NS) widx[rank] = (uint16_t)my;
      }
      __syncwarp();
      const uint16_t first = H > 0 ? widx[0] : (uint16_t)0;
      for (int l = min(H, NS) + lane; l < NS; l += 32) widx[l] = first;
    } else {
      int cnt = 0;
      uint16_t first = 0;
      for (int k0 = 0; k0 < N && cnt < NS; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < N) { const float4 v = __ldg(cl + k); hit = dist2(qx, qy, qz, v.x, v.y, v.z) < r2; }
        const unsigned hm = __ballot_sync(0xffffffffu, hit);
        if (hm && cnt == 0) first = (uint16_t)(k0 + __ffs(hm) - 1);
        const int pos = cnt + __popc(hm & lt);
        if (hit && pos < NS) widx[pos] = (uint16_t)k;
        cnt += __popc(hm);
      }
      for (int l = min(cnt, NS) + lane; l < NS; l += 32) widx[l] = first;
    }
    __syncwarp();
  };
  auto issue = [&](uint64_t dW, bool first_layer) {
    if (wq == 0) {
      tc_fence_after();
      if (elect_one()) {
        if (first_layer) {
          mma_bf16_ss_off(tmem, dX1, 0, dW1k, 0, IDESC, 0);
        } else {
#pragma unroll
          for (int ks = 0; ks < SA1_XK / 16; ++ks) mma_bf16_ss_off(tmem, dX, ks * 16, dW, ks * 16, IDESC, ks > 0);
        }
        mma_commit(bar);
      }
      __syncwarp();
    }
  };

  // rounds of 4 centroids are handed out dynamically (shared counter): ball-query cost varies with the local point density,
  // and with a static stride the groups finish up to several rounds apart and idle at the CTA's final barrier
  for (int round = g; round * 4 < SA1_NPOINT && ok;) {
    const int base = round * 4;
    if (base + wq < SA1_NPOINT) warp_ball_query(base + wq);
    wg_sync(g);
    float4 pn = __ldg(cl + lists[t]);
    const float* cpn = cxyz + base * 3;
    float nx = cpn[0], ny = cpn[1], nz = cpn[2];
#pragma unroll 1
    for (int cc = 0; cc < 4 && ok; ++cc) {
      const int j = base + cc;
      if (j >= SA1_NPOINT) break;
      const float cx = nx, cy = ny, cz = nz;
      const float4 p = pn;
      if (ball_idx) ball_idx[((size_t)b * SA1_NPOINT + j) * NS + t] = lists[cc * 128 + t];
      {
        const float dx = fsub(p.x, cx), dy = fsub(p.y, cy), dz = fsub(p.z, cz);
        *reinterpret_cast<uint4*>(X + kmajor_chunk_off(t, 0, KC)) = make_uint4(pack_bf16(dx, dy), pack_bf16(dz, p.w), 0u, 0u);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      wg_sync(g);
      issue(dW1, true);
      if (cc < 3 && j + 1 < SA1_NPOINT) {   // next centroid's gathered point and coordinates, under the MMAs
        pn = __ldg(cl + lists[(cc + 1) * 128 + t]);
        const float* cq = cxyz + (j + 1) * 3;
        nx = cq[0]; ny = cq[1]; nz = cq[2];
      }
#pragma unroll 1
      for (int layer = 1; layer < 3; ++layer) {
        ok = ok && mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        epilogue_pack_relu<64, KC>(tlane, X, t);
        fence_proxy_async_smem();
        tc_fence_before();
        wg_sync(g);
        issue(layer == 1 ? dW2 : dW3, false);
      }
      ok = ok && mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      // max over the 128 neighbour rows: accumulator-fragment loads (every thread: 4 rows x 16 channels, in two column halves) -> in-thread max ->
      // ReLU + bf16x2 (monotone, so rounding before the max gives the same result) -> 3 halving butterfly steps over the 8
      // row groups of the warp; lane L ends up with channels 2L, 2L+1 of the warp's 32 rows
      {
        uint32_t pk[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t va[16], vb[16];
          tmem_ld_16x256b_x4(tlane + h * 32, va);
          tmem_ld_16x256b_x4(tlane + (16u << 16) + h * 32, vb);
          tmem_ld_wait();
#pragma unroll
          for (int rep = 0; rep < 4; ++rep) {
            const float m0 = fmaxf(fmaxf(__uint_as_float(va[rep * 4]), __uint_as_float(va[rep * 4 + 2])),
                                   fmaxf(__uint_as_float(vb[rep * 4]), __uint_as_float(vb[rep * 4 + 2])));
            const float m1 = fmaxf(fmaxf(__uint_as_float(va[rep * 4 + 1]), __uint_as_float(va[rep * 4 + 3])),
                                   fmaxf(__uint_as_float(vb[rep * 4 + 1]), __uint_as_float(vb[rep * 4 + 3])));
            pk[h * 4 + rep] = cvt_relu_bf16x2(m0, m1);
          }
        }
#pragma unroll
        for (int w = 4; w >= 1; w >>= 1) {
          const bool upper = (lane & (w << 2)) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const uint32_t send = upper ? pk[i] : pk[i + w], keepv = upper ? pk[i + w] : pk[i];
            pk[i] = bf16x2_max(keepv, __shfl_xor_sync(0xffffffffu, send, w << 2));
          }
        }
        red[wq * 32 + lane] = (int)pk[0];
      }
      tc_fence_before();
      wg_sync(g);
      if (t < 32) {
        const uint32_t m = bf16x2_max(bf16x2_max((uint32_t)red[t], (uint32_t)red[32 + t]), bf16x2_max((uint32_t)red[64 + t], (uint32_t)red[96 + t]));
        reinterpret_cast<uint32_t*>(out_bf16 + ((size_t)b * SA1_NPOINT + j) * 64)[t] = m;
        if constexpr (ARG) red[128 + t] = (int)m;
      }
      if constexpr (ARG) {
        // training forward: the winning neighbour row of every channel = the first row whose ReLU'd bf16 output equals the pooled
        // value (channels pooled to 0 carry no gradient: row 0).  The accumulator is still in TMEM; red[160..223] = row per channel
        if (t < 64) red[160 + t] = 255;
        wg_sync(g);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t va[16], vb[16];
          tmem_ld_16x256b_x4(tlane + h * 32, va);
          tmem_ld_16x256b_x4(tlane + (16u << 16) + h * 32, vb);
          tmem_ld_wait();
#pragma unroll
          for (int rep = 0; rep < 4; ++rep) {
            const uint32_t fin = (uint32_t)red[128 + 4 * (h * 4 + rep) + (lane & 3)];
            const int ch0 = 8 * (h * 4 + rep) + 2 * (lane & 3);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const uint32_t* src = rr < 2 ? va : vb;
              const int o = rep * 4 + (rr & 1) * 2;
              const uint32_t pv = cvt_relu_bf16x2(__uint_as_float(src[o]), __uint_as_float(src[o + 1]));
              const int row = wq * 32 + (lane >> 2) + 8 * rr;
              if ((fin & 0xFFFFu) != 0u && (pv & 0xFFFFu) == (fin & 0xFFFFu)) atomicMin(&red[160 + ch0], row);
              if ((fin >> 16) != 0u && (pv >> 16) == (fin >> 16)) atomicMin(&red[160 + ch0 + 1], row);
            }
          }
        }
        tc_fence_before();
        wg_sync(g);
        if (t < 64) {
          const int a = red[160 + t];
          arg_out[((size_t)b * SA1_NPOINT + j) * 64 + t] = (uint8_t)(a > 127 ? 0 : a);
        }
      }
    }
    if (t == 0) rsel[g] = atomicAdd(next_round, 1);
    wg_sync(g);   // the lists are rewritten by the next round
    round = rsel[g];
  }
  if (!ok && (threadIdx.x & 31) == 0) atomicExch(err, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// ---------------------------------------------------------------------------------------------- SA1, activations in TMEM
// sa1w_tc_kernel is bound by shared-memory bandwidth: its N = 64 MMAs read 6 KB of operands per 32 tensor-pipe cycles and
// every layer output goes registers -> st.shared -> (async proxy) -> MMA.  Here the A operand never touches shared memory:
// the gathered rows and every layer's ReLU'd bf16 output are written to TENSOR MEMORY with tcgen05.st (row = lane, two bf16
// per column) and tcgen05.mma reads A from there; only the 2 KB weight slices (and a constant ones tile for the bias MMA)
// come from shared memory.  Per chain: 64 accumulator columns + 32 operand columns -> 5 chains (warpgroups) per CTA.
// The shared memory this frees holds the problem's whole cloud (16 B rows), so the ball query's candidate tests and the
// gathers are LDS instead of L2 round trips.
constexpr int SA1T_NWG = 5, SA1T_COLS = 96;
struct Sa1tSmem {
  static constexpr size_t w = 0;                                           // 3 x [64][80] bf16
  static constexpr size_t ones = w + 3 * 64 * SA1_XK * 2;                  // [128][16] bf16: column 0 = 1.0 (A operand of the bias MMAs)
  static constexpr size_t cnt = ones + 128 * 16 * 2;                       // u32 [BUCKETS] grid-build counters
  static constexpr size_t lists = cnt + SA1_BUCKETS * 4;                   // [NWG][4][128] u16
  static constexpr size_t cand = lists + (size_t)SA1T_NWG * 4 * 128 * 2;   // [NWG][4][256] u16
  static constexpr size_t red = cand + (size_t)SA1T_NWG * 4 * 256 * 2;     // [NWG][512] int
  static constexpr size_t cxyz = red + (size_t)SA1T_NWG * 512 * 4;         // f32 [512][3]
  static constexpr size_t bars = cxyz + (size_t)SA1_NPOINT * 3 * 4;
  static constexpr size_t bstart = (bars + 64 + 15) / 16 * 16;             // u16 [BUCKETS + 1]
  static constexpr size_t cloud = (bstart + (SA1_BUCKETS + 1) * 2 + 15) / 16 * 16;   // float4 [N]
  __host__ __device__ static size_t sidx(int N) { return cloud + (size_t)N * 16; }      // u16 [N]
  static size_t total(int N) { return sidx(N) + (size_t)N * 2 + 64; }
};

template <bool ARG>
__global__ void __launch_bounds__(128 * SA1T_NWG, 1)
sa1t_tc_kernel(const float* __restrict__ cloud, int N, const float* __restrict__ new_xyz, float r2, const __nv_bfloat16* __restrict__ gw1,
               const __nv_bfloat16* __restrict__ gw2, const __nv_bfloat16* __restrict__ gw3, __nv_bfloat16* __restrict__ out_bf16,
               int* __restrict__ err, int32_t* __restrict__ ball_idx, uint8_t* __restrict__ arg_out, int split, int pack) {
  using S = Sa1tSmem;
  constexpr int KC = SA1_XK / 8, NS = NSAMPLE, NWG = SA1T_NWG;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW1 = smem + S::w;
  uint8_t* sW2 = sW1 + 64 * SA1_XK * 2;
  uint8_t* sW3 = sW2 + 64 * SA1_XK * 2;
  uint8_t* sOnes = smem + S::ones;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::bars + 8 * NWG);
  uint16_t* bstart = reinterpret_cast<uint16_t*>(smem + S::bstart);
  float4* cl = reinterpret_cast<float4*>(smem + S::cloud);
  uint16_t* sidx = reinterpret_cast<uint16_t*>(smem + S::sidx(N));
  float* cxyz = reinterpret_cast<float*>(smem + S::cxyz);

  const int b = blockIdx.x / split, part = blockIdx.x % split;   // small batches: a problem's centroid rounds are dealt to `split` CTAs
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int g = warp >> 2, wq = warp & 3, t = threadIdx.x & 127, lane = threadIdx.x & 31;
  uint16_t* lists = reinterpret_cast<uint16_t*>(smem + S::lists) + (size_t)g * 4 * 128;
  uint16_t* wcand = reinterpret_cast<uint16_t*>(smem + S::cand) + (size_t)(g * 4 + wq) * 256;
  int* red = reinterpret_cast<int*>(smem + S::red) + g * 512;   // [0,128) per-warp pooled pairs | [128,256) per-centroid pooled pairs | [256,512) winning rows
  const float4* gcl = reinterpret_cast<const float4*>(cloud) + (size_t)b * N;

  stage_weight(gw1, 64, SA1_XK, sW1);
  stage_weight(gw2, 64, SA1_XK, sW2);
  stage_weight(gw3, 64, SA1_XK, sW3);
  for (int i = threadIdx.x; i < 128 * 2; i += blockDim.x)   // ones tile: chunk 0 of every row = (1, 0, ..), chunk 1 = 0
    *reinterpret_cast<uint4*>(sOnes + kmajor_chunk_off(i >> 1, i & 1, 2)) = make_uint4((i & 1) ? 0u : 0x00003F80u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < SA1_NPOINT * 3; i += blockDim.x) cxyz[i] = __ldg(new_xyz + (size_t)b * SA1_NPOINT * 3 + i);
  // ---- the problem's cloud -> shared memory, and the hash grid over it (counting sort of point indices by bucket)
  {
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + S::cnt);
    __shared__ uint32_t wsum[16];
    for (int i = threadIdx.x; i < SA1_BUCKETS; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const float4 v = __ldg(gcl + k);
      cl[k] = v;
      atomicAdd(&cnt[grid_bucket(grid_coord(v.x), grid_coord(v.y), grid_coord(v.z))], 1u);
    }
    __syncthreads();
    constexpr int PER = SA1_BUCKETS / 512;
    const bool scanner = threadIdx.x < 512;
    uint32_t loc[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { loc[i] = scanner ? cnt[threadIdx.x * PER + i] : 0u; sum += loc[i]; }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31 && scanner) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t v = lane < 16 ? wsum[lane] : 0u, iv = v;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += y; }
      if (lane < 16) wsum[lane] = iv - v;
    }
    __syncthreads();
    if (scanner) {
      uint32_t run = wsum[threadIdx.x >> 5] + inc - sum;
#pragma unroll
      for (int i = 0; i < PER; ++i) { bstart[threadIdx.x * PER + i] = (uint16_t)run; cnt[threadIdx.x * PER + i] = run; run += loc[i]; }
    }
    if (threadIdx.x == 0) bstart[SA1_BUCKETS] = (uint16_t)N;
    __syncthreads();
    for (int k = threadIdx.x; k < N; k += blockDim.x) {
      const float4 v = cl[k];
      sidx[atomicAdd(&cnt[grid_bucket(grid_coord(v.x), grid_coord(v.y), grid_coord(v.z))], 1u)] = (uint16_t)k;
    }
  }
  __shared__ int round_ctr[1 + 8];
  __shared__ int hcnt_s[SA1T_NWG * 4];   // distinct hits (<= 128) of the round's four ball queries
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
  const uint32_t tmemD = *tmem_slot + (uint32_t)g * SA1T_COLS;       // accumulator: 64 columns
  const uint32_t tmemA = tmemD + 64;                                  // operand: 32 columns = 64 bf16 per row
  const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
  const uint32_t tlane = tmemD + lane_off;
  const uint64_t dOnes = make_smem_desc(smem_u32(sOnes), 128, 2 * 128, LAYOUT_NONE);
  const uint64_t dW1 = tile_desc(smem_u32(sW1), SA1_XK, 0, 0), dW2 = tile_desc(smem_u32(sW2), SA1_XK, 0, 0),
                 dW3 = tile_desc(smem_u32(sW3), SA1_XK, 0, 0);
  uint64_t* bar = &bars[g];
  uint32_t phase = 0;
  bool ok = true;
  constexpr uint32_t IDESC = make_idesc_bf16(128, 64);
  const unsigned lt = (1u << lane) - 1u;

  // complete ball query of centroid jc by this warp -> lists[wq][0..H), H = hcnt[wq] distinct hits (same algorithm as sa1w_tc_kernel,
  // cloud in smem).  Hits are ranked by original point index (the linear scan's first-128 order) only when more than 128 were found or
  // the caller wants the index lists / winning rows (training); otherwise the max-pool takes the set in bucket order.  pointnet2's
  // first-hit padding is never materialised: rows beyond H re-read row H - 1 (a duplicate either way).
  const bool need_order = ARG || ball_idx != nullptr;
  auto warp_ball_query = [&](int jc) {
    uint16_t* widx = lists + wq * 128;
    const float qx = cxyz[3 * jc], qy = cxyz[3 * jc + 1], qz = cxyz[3 * jc + 2];
    const int ix = grid_coord(qx), iy = grid_coord(qy), iz = grid_coord(qz);
    uint32_t bk = 0x10000u + lane;
    if (lane < 27) bk = grid_bucket(ix + (lane % 3) - 1, iy + ((lane / 3) % 3) - 1, iz + (lane / 9) - 1);
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
        int i = 0;
        for (; i + 4 <= n0; i += 4) {
          const uint16_t a0 = sidx[s0 + i], a1 = sidx[s0 + i + 1], a2 = sidx[s0 + i + 2], a3 = sidx[s0 + i + 3];
          wcand[o + i] = a0; wcand[o + i + 1] = a1; wcand[o + i + 2] = a2; wcand[o + i + 3] = a3;
        }
        for (; i < n0; ++i) wcand[o + i] = sidx[s0 + i];
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
  // one layer: bias through an SS MMA against the constant ones tile (K = 16: ones column x the weights' bias chunk), then the
  // K-steps of the activations straight from tensor memory
  auto issue = [&](uint64_t dW, int a_col0, int ksteps) {
    if (wq == 0) {
      tc_fence_after();
      if (elect_one()) {
        mma_bf16_ss_off(tmemD, dOnes, 0, dW, 64, IDESC, 0);
        for (int ks = 0; ks < ksteps; ++ks) mma_bf16_ts(tmemD, tmemA + a_col0 + ks * 8, dW + (uint64_t)(ks * 16), IDESC, 1);
        mma_commit(bar);
      }
      __syncwarp();
    }
  };
  // accumulator (bias inside) -> relu -> bf16 -> the operand columns of this thread's lane
  auto epilogue_to_tmem = [&]() {
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tlane + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = cvt_relu_bf16x2(__uint_as_float(v[q * 16 + 2 * i]), __uint_as_float(v[q * 16 + 2 * i + 1]));
        tmem_st8(tmemA + lane_off + (c0 >> 1) + q * 8, pk);
      }
    }
    tmem_st_wait();
  };

  int ntiles_done = 0;
  for (int round = g; (round * split + part) * 4 < SA1_NPOINT && ok;) {
    const int base = (round * split + part) * 4;
    const int nvalid = min(4, SA1_NPOINT - base);
    if (wq < nvalid) warp_ball_query(base + wq);
    wg_sync(g);
    const TilePack tp = pack_round(hcnt, nvalid, pack);   // the round's distinct rows packed into 128-row tiles (tc_common.cuh)
    if (ball_idx)   // pointnet2's output format: first-hit padding
      for (int c = 0; c < nvalid; ++c) ball_idx[((size_t)b * SA1_NPOINT + base + c) * NS + t] = lists[c * 128 + (t < hcnt[c] ? t : 0)];
#pragma unroll 1
    for (int tile = 0; tile < tp.ntiles && ok; ++tile) {
      const int mc = pack_owner(tp, nvalid, tile, wq);   // this warp's quarter of the tile belongs to centroid base + mc
      const int mq0 = pack_q0(tp, mc);
      {
        const int j = base + mc;
        const int k = lists[mc * 128 + min((wq - mq0) * 32 + lane, hcnt[mc] - 1)];
        const float4 p = cl[k];
        const float dx = fsub(p.x, cxyz[3 * j]), dy = fsub(p.y, cxyz[3 * j + 1]), dz = fsub(p.z, cxyz[3 * j + 2]);
        const uint32_t row[8] = {pack_bf16(dx, dy), pack_bf16(dz, p.w), 0u, 0u, 0u, 0u, 0u, 0u};
        tmem_st8(tmemA + lane_off + 24, row);          // K elements 48..63 of the operand columns: (dx dy dz w 0 ...)
        tmem_st_wait();
      }
      tc_fence_before();
      wg_sync(g);
      issue(dW1, 24, 1);                               // layer 1: bias + one K = 16 step against W1's chunks 0-1
#pragma unroll 1
      for (int layer = 1; layer < 3; ++layer) {
        ok = ok && mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        epilogue_to_tmem();
        tc_fence_before();
        wg_sync(g);
        issue(layer == 1 ? dW2 : dW3, 0, 4);
      }
      ok = ok && mbar_wait(bar, phase); phase ^= 1;
      tc_fence_after();
      {
        uint32_t pk[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t va[16], vb[16];
          tmem_ld_16x256b_x4(tlane + h * 32, va);
          tmem_ld_16x256b_x4(tlane + (16u << 16) + h * 32, vb);
          tmem_ld_wait();
#pragma unroll
          for (int rep = 0; rep < 4; ++rep) {
            const float m0 = fmaxf(fmaxf(__uint_as_float(va[rep * 4]), __uint_as_float(va[rep * 4 + 2])),
                                   fmaxf(__uint_as_float(vb[rep * 4]), __uint_as_float(vb[rep * 4 + 2])));
            const float m1 = fmaxf(fmaxf(__uint_as_float(va[rep * 4 + 1]), __uint_as_float(va[rep * 4 + 3])),
                                   fmaxf(__uint_as_float(vb[rep * 4 + 1]), __uint_as_float(vb[rep * 4 + 3])));
            pk[h * 4 + rep] = cvt_relu_bf16x2(m0, m1);
          }
        }
#pragma unroll
        for (int w = 4; w >= 1; w >>= 1) {
          const bool upper = (lane & (w << 2)) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const uint32_t send = upper ? pk[i] : pk[i + w], keepv = upper ? pk[i + w] : pk[i];
            pk[i] = bf16x2_max(keepv, __shfl_xor_sync(0xffffffffu, send, w << 2));
          }
        }
        red[wq * 32 + lane] = (int)pk[0];
      }
      tc_fence_before();
      wg_sync(g);
      {   // the tile's centroids: warp c finishes centroid c (max over the quarters it owns)
        const int c = wq;
        if (c < nvalid && pack_tile(tp, c) == tile) {
          const int q0 = pack_q0(tp, c), q1 = pack_q1(tp, nvalid, c);
          uint32_t m = (uint32_t)red[q0 * 32 + lane];
          for (int q = q0 + 1; q < q1; ++q) m = bf16x2_max(m, (uint32_t)red[q * 32 + lane]);
          reinterpret_cast<uint32_t*>(out_bf16 + ((size_t)b * SA1_NPOINT + base + c) * 64)[lane] = m;
          if constexpr (ARG) { red[128 + c * 32 + lane] = (int)m; red[256 + c * 64 + lane] = 255; red[256 + c * 64 + 32 + lane] = 255; }
        }
      }
      if constexpr (ARG) {
        wg_sync(g);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t va[16], vb[16];
          tmem_ld_16x256b_x4(tlane + h * 32, va);
          tmem_ld_16x256b_x4(tlane + (16u << 16) + h * 32, vb);
          tmem_ld_wait();
#pragma unroll
          for (int rep = 0; rep < 4; ++rep) {
            const uint32_t fin = (uint32_t)red[128 + mc * 32 + 4 * (h * 4 + rep) + (lane & 3)];
            const int ch0 = 8 * (h * 4 + rep) + 2 * (lane & 3);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const uint32_t* src = rr < 2 ? va : vb;
              const int o = rep * 4 + (rr & 1) * 2;
              const uint32_t pv = cvt_relu_bf16x2(__uint_as_float(src[o]), __uint_as_float(src[o + 1]));
              const int row = (wq - mq0) * 32 + (lane >> 2) + 8 * rr;   // row of the centroid's (padded) neighbour list
              if ((fin & 0xFFFFu) != 0u && (pv & 0xFFFFu) == (fin & 0xFFFFu)) atomicMin(&red[256 + mc * 64 + ch0], row);
              if ((fin >> 16) != 0u && (pv >> 16) == (fin >> 16)) atomicMin(&red[256 + mc * 64 + ch0 + 1], row);
            }
          }
        }
        tc_fence_before();
        wg_sync(g);
        for (int i = t; i < nvalid * 64; i += 128) {
          const int c = i >> 6;
          if (pack_tile(tp, c) != tile) continue;
          const int a = red[256 + i];
          arg_out[((size_t)b * SA1_NPOINT + base + c) * 64 + (i & 63)] = (uint8_t)(a > 127 ? 0 : a);
        }
      }
    }
    ntiles_done += tp.ntiles;
    if (t == 0) rsel[g] = atomicAdd(next_round, 1);
    wg_sync(g);   // the lists are rewritten by the next round
    round = rsel[g];
  }
  if (t == 0 && ntiles_done) atomicAdd(sa_tile_counter(err, 0), (unsigned long long)ntiles_done);
  if (!ok && (threadIdx.x & 31) == 0) atomicExch(err, 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(*tmem_slot, 512);
}

// bf16 -> fp32 widening of the pooled SA2 rows for the (still fp32) group-all / FC stages
__global__ void widen_kernel(const __nv_bfloat16* __restrict__ src, int rows, int src_stride, int cols, float* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols) return;
  size_t r = i / cols, c = i % cols;
  dst[i] = __bfloat162float(src[r * src_stride + c]);
}

int* tc_error_flag(mpn_ctx* c) {
  static std::map<mpn_ctx*, int*> flags;
  auto it = flags.find(c);
  if (it != flags.end()) return it->second;
  int* p = nullptr;   // [0] sticky error flag | [2..5] two 64-bit counters: 128-row MMA tiles issued by the SA1 / SA2 kernels (sa_tile_counter)
  cudaMalloc(&p, 32);
  cudaMemset(p, 0, 32);
  flags[c] = p;
  return p;
}
// executed-work accounting of the packed SA kernels (bench.py's roofline: issued MMA flops = tiles x flops per tile)
int sa_tile_counts(mpn_ctx* c, unsigned long long* out, int reset) {
  int* p = tc_error_flag(c);
  MPN_CHECK_CUDA(cudaDeviceSynchronize());
  MPN_CHECK_CUDA(cudaMemcpy(out, p + 2, 16, cudaMemcpyDeviceToHost));
  if (reset) MPN_CHECK_CUDA(cudaMemset(p + 2, 0, 16));
  return MPN_OK;
}

__global__ void narrow_kernel(const float* __restrict__ src, size_t rows, int src_stride, int cols, __nv_bfloat16* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  size_t r = i / cols, cc = i % cols;
  dst[i] = __float2bfloat16_rn(src[r * src_stride + cc]);
}

// A/B switch of the packed row tiles (tc_common.cuh: TilePack), read per launch: MPN_SA_NOPACK=1 gives every centroid its own 128-row tile
int sa_pack() { return getenv("MPN_SA_NOPACK") == nullptr; }

// Small batches leave most SMs without a problem: deal each problem's rounds of 4 centroids to `split` CTAs (each rebuilds the
// problem's shared-memory state) so that about one CTA per SM is in flight.  max_split: rounds per warpgroup of an unsplit CTA.
int sa_split(const mpn_ctx* c, int B, int max_split) {
  int sp = c->sm_count / (B > 0 ? B : 1);
  if (sp > max_split) sp = max_split;
  return sp < 1 ? 1 : sp;
}

template <int MODULE>
static int launch_sa_tc(mpn_ctx* c, cudaStream_t s, const float* xyz, int stride, int N, const __nv_bfloat16* feat, const float* new_xyz,
                        int B, __nv_bfloat16* out, int out_stride, int32_t* ball_idx = nullptr, uint8_t* arg_out = nullptr) {
  TcWeights& tw = g_tc[c];
  if (MODULE == 0) {
    MPN_REQUIRE(stride == 4, "tensor-core SA1 takes the [B][N][4] cloud");
    MPN_REQUIRE(N < 65536, "tensor-core SA1: at most 65535 points");
    const bool sa1_ss = getenv("MPN_SA1_SS") != nullptr;   // A/B switch (read per launch so tests can toggle it): the shared-memory-operand kernel below
    const size_t smem_t = Sa1tSmem::total(N);
    if (!sa1_ss && smem_t + 2048 <= 227 * 1024) {
      auto kern = arg_out ? sa1t_tc_kernel<true> : sa1t_tc_kernel<false>;
      MPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      const int split = sa_split(c, B, 128 / SA1T_NWG);
      kern<<<B * split, 128 * SA1T_NWG, smem_t, s>>>(xyz, N, new_xyz, SA1_RADIUS * SA1_RADIUS, tw.sa[0][0], tw.sa[0][1], tw.sa[0][2], out,
                                                     tc_error_flag(c), ball_idx, arg_out, split, sa_pack());
    } else {
      size_t smem7 = Sa1wSmem<7>::total(N);
      MPN_REQUIRE(smem7 <= 227 * 1024, "tensor-core SA1: %d points do not fit", N);
      auto kern = arg_out ? sa1w_tc_kernel<7, true> : sa1w_tc_kernel<7, false>;
      MPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem7));
      kern<<<B, 128 * 7, smem7, s>>>(xyz, N, new_xyz, SA1_RADIUS * SA1_RADIUS, tw.sa[0][0], tw.sa[0][1], tw.sa[0][2], out, tc_error_flag(c),
                                     ball_idx, arg_out);
    }
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    return MPN_OK;
  }
  MPN_REQUIRE(N == SA1_NPOINT, "tensor-core SA2 expects the 512 SA1 centroids as input points");
  {
    size_t smem3 = Sa2wSmem::total;
    auto kern = arg_out ? sa2w3_tc_kernel<true> : sa2w3_tc_kernel<false>;
    MPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    const int split = sa_split(c, B, 32 / SA2W_NWG);
    kern<<<B * split, SA2W_THREADS, smem3, s>>>(xyz, stride, feat, new_xyz, SA2_RADIUS * SA2_RADIUS, tw.sa[1][0], tw.w2_nofold, tw.sa[1][2],
                                                c->w.sa[1][1].b, c->w.sa[1][2].b, out, out_stride, tc_error_flag(c), ball_idx, arg_out, split, sa_pack());
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    return MPN_OK;
  }
}

// per-module entry (tests / mpn_sa_forward with MPN_PREC_BF16): fp32 in, fp32 out, bf16 inside
int tc_sa_forward(mpn_ctx* c, cudaStream_t s, int module, const float* xyz, int stride, const float* feats, int feat_stride, int B,
                  int N, const float* new_xyz, float* new_feats, int32_t* ball_idx) {
  Workspace& w = c->ws;
  __nv_bfloat16* feat1 = reinterpret_cast<__nv_bfloat16*>(w.tc_scratch);
  __nv_bfloat16* a3 = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w.tc_scratch) + feat1_bytes(w.capacity));
  int r;
  if (module == 0) {
    MPN_REQUIRE(stride == 4 && feats == xyz + 3 && feat_stride == 4, "bf16 SA1 takes the [B][N][4] cloud (features = 4th column)");
    if ((r = launch_sa_tc<0>(c, s, xyz, 4, N, nullptr, new_xyz, B, feat1, 64, ball_idx))) return r;
    size_t n = (size_t)B * SA1_NPOINT * 64;
    widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(feat1, B * SA1_NPOINT, 64, 64, new_feats);
  } else {
    MPN_REQUIRE(module == 1 && N == SA1_NPOINT, "bf16 per-module entry supports modules 0 and 1 (N = 512 for module 1)");
    size_t n = (size_t)B * N * 64;
    narrow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(feats, (size_t)B * N, feat_stride, 64, feat1);
    if ((r = launch_sa_tc<1>(c, s, xyz, stride, N, feat1, new_xyz, B, a3, A3_K, ball_idx))) return r;
    size_t m = (size_t)B * SA2_NPOINT * 256;
    widen_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(a3, B * SA2_NPOINT, A3_K, 256, new_feats);
  }
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// Point-cloud encoder forward (SA1, SA2, group-all SA3) of the bf16 training step through the inference kernels' ARG variants:
// besides the pooled features they record the ball-query indices and the winning neighbour row of every (group, channel) --
// the routing the backward replays.  Outputs in the training step's formats: feat1 f32 [B][512][64], feat2 f32 [B][128][256],
// feat3 f32 [B][1024], ball1 / ball2 i32, arg1 / arg2 / arg3 u8.
int tc_train_forward_sa(mpn_ctx* c, cudaStream_t s, const float* cloud, int B, int N, int32_t* fps_idx, float* xyz1, float* xyz2,
                        float* feat1_f32, float* feat2_f32, float* feat3_f32, int32_t* ball1, int32_t* ball2, uint8_t* arg1,
                        uint8_t* arg2, uint8_t* arg3) {
  Workspace& w = c->ws;
  MPN_REQUIRE(B <= w.capacity, "tc_train_forward_sa: batch %d exceeds the workspace capacity %d", B, w.capacity);
  __nv_bfloat16* feat1 = reinterpret_cast<__nv_bfloat16*>(w.tc_scratch);
  __nv_bfloat16* a3 = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w.tc_scratch) + feat1_bytes(w.capacity));
  uint8_t* base = reinterpret_cast<uint8_t*>(w.tc_scratch) + feat1_bytes(w.capacity) + a3_bytes(w.capacity);
  __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* h2 = reinterpret_cast<__nv_bfloat16*>(base + h_bytes(w.capacity));
  __nv_bfloat16* f3 = reinterpret_cast<__nv_bfloat16*>(base + 2 * h_bytes(w.capacity));
  TcWeights& tw = g_tc[c];
  int r;
  if ((r = launch_fps(c, s, cloud, B, N, 4, SA1_NPOINT, fps_idx, xyz1))) return r;
  if ((r = launch_sa_tc<0>(c, s, cloud, 4, N, nullptr, xyz1, B, feat1, 64, ball1, arg1))) return r;
  if ((r = launch_fps(c, s, xyz1, B, SA1_NPOINT, 3, SA2_NPOINT, fps_idx, xyz2))) return r;
  if ((r = launch_sa_tc<1>(c, s, xyz1, 3, SA1_NPOINT, feat1, xyz2, B, a3, A3_K, ball2, arg2))) return r;
  const int M3 = B * SA2_NPOINT;
  if ((r = launch_gemm_tc(c, s, 0, a3, A3_K, tw.sa[2][0], A3_K, c->w.sa[2][0].b, M3, 512, h1, 512))) return r;
  if ((r = launch_gemm_tc(c, s, 0, h1, 512, tw.sa[2][1], 512, c->w.sa[2][1].b, M3, 512, h2, 512))) return r;
  if ((r = launch_gemm_tc(c, s, 3, h2, 512, tw.sa[2][2], 512, c->w.sa[2][2].b, M3, 1024, f3, 1024, arg3))) return r;
  const size_t n1 = (size_t)B * SA1_NPOINT * 64, n2 = (size_t)B * SA2_NPOINT * 256, n3 = (size_t)B * 1024;
  widen_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, s>>>(feat1, B * SA1_NPOINT, 64, 64, feat1_f32);
  widen_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(a3, B * SA2_NPOINT, A3_K, 256, feat2_f32);
  widen_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, s>>>(f3, B, 1024, 1024, feat3_f32);
  c->launches += 3;
  MPN_CHECK_CUDA(cudaGetLastError());
  c->tw.sa3_h1 = h1;   // the SA3 backward reads them instead of recomputing layers 1-2
  c->tw.sa3_h2 = h2;
  c->tw.sa3_a3 = a3;
  return MPN_OK;
}

int tc_encoder_forward(mpn_ctx* c, cudaStream_t s, const float* cloud, int B, int N, float* out, int ldo) {
  Workspace& w = c->ws;
  int r;
  __nv_bfloat16* feat1 = reinterpret_cast<__nv_bfloat16*>(w.tc_scratch);
  __nv_bfloat16* a3 = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(w.tc_scratch) + feat1_bytes(w.capacity));
  { StageTimer t(c, s, MPN_ST_FPS1);
    if ((r = launch_fps(c, s, cloud, B, N, 4, SA1_NPOINT, reinterpret_cast<int32_t*>(w.fc_a), w.xyz1))) return r; }
  { StageTimer t(c, s, MPN_ST_SA1);
    if ((r = launch_sa_tc<0>(c, s, cloud, 4, N, nullptr, w.xyz1, B, feat1, 64))) return r; }
  { StageTimer t(c, s, MPN_ST_FPS2);
    if ((r = launch_fps(c, s, w.xyz1, B, SA1_NPOINT, 3, SA2_NPOINT, reinterpret_cast<int32_t*>(w.fc_a), w.xyz2))) return r; }
  { StageTimer t(c, s, MPN_ST_SA2);
    if ((r = launch_sa_tc<1>(c, s, w.xyz1, 3, SA1_NPOINT, feat1, w.xyz2, B, a3, A3_K))) return r; }
  TcWeights& tw = g_tc[c];
  uint8_t* base = reinterpret_cast<uint8_t*>(w.tc_scratch) + feat1_bytes(w.capacity) + a3_bytes(w.capacity);
  __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(base);
  __nv_bfloat16* h2 = reinterpret_cast<__nv_bfloat16*>(base + h_bytes(w.capacity));
  __nv_bfloat16* f3 = reinterpret_cast<__nv_bfloat16*>(base + 2 * h_bytes(w.capacity));
  __nv_bfloat16* g1 = f3 + (size_t)w.capacity * 1024;
  __nv_bfloat16* g2 = g1 + (size_t)w.capacity * 4096;
  const int M3 = B * SA2_NPOINT;
  { StageTimer t(c, s, MPN_ST_SA3);   // group-all module as three row-shared GEMMs; the last one pools each problem's 128 rows
    if ((r = launch_gemm_tc(c, s, 0, a3, A3_K, tw.sa[2][0], A3_K, c->w.sa[2][0].b, M3, 512, h1, 512))) return r;
    if ((r = launch_gemm_tc(c, s, 0, h1, 512, tw.sa[2][1], 512, c->w.sa[2][1].b, M3, 512, h2, 512))) return r;
    if ((r = launch_gemm_tc(c, s, 2, h2, 512, tw.sa[2][2], 512, c->w.sa[2][2].b, M3, 1024, f3, 1024))) return r; }
  StageTimer tfc(c, s, MPN_ST_FC);
  if (B <= SKINNY_MAX_ROWS) {   // a handful of problems: fp32 weight streaming on every SM instead of N / 256 tensor-core tiles
    if ((r = launch_linear_skinny(c, s, f3, 1024, 1, c->w.fc[0], B, w.fc_a, 4096, 0))) return r;
    if ((r = launch_groupnorm_lrelu(c, s, w.fc_a, B, 4096, 16, c->w.gn_w[0], c->w.gn_b[0]))) return r;
    if ((r = launch_linear_skinny(c, s, w.fc_a, 4096, 0, c->w.fc[1], B, w.fc_b, 2048, 0))) return r;
    if ((r = launch_groupnorm_lrelu(c, s, w.fc_b, B, 2048, 16, c->w.gn_w[1], c->w.gn_b[1]))) return r;
    return launch_linear_skinny(c, s, w.fc_b, 2048, 0, c->w.fc[2], B, out, ldo, 0);
  }
  if ((r = launch_gemm_tc(c, s, 1, f3, 1024, tw.fc[0], 1024, c->w.fc[0].b, B, 4096, w.fc_a, 4096))) return r;
  if ((r = launch_groupnorm_lrelu_bf16(c, s, w.fc_a, B, 4096, 16, c->w.gn_w[0], c->w.gn_b[0], g1))) return r;
  if ((r = launch_gemm_tc(c, s, 1, g1, 4096, tw.fc[1], 4096, c->w.fc[1].b, B, 2048, w.fc_b, 2048))) return r;
  if ((r = launch_groupnorm_lrelu_bf16(c, s, w.fc_b, B, 2048, 16, c->w.gn_w[1], c->w.gn_b[1], g2))) return r;
  return launch_gemm_tc(c, s, 1, g2, 2048, tw.fc[2], 2048, c->w.fc[2].b, B, 2048, out, ldo);
}

}  // namespace mpn
