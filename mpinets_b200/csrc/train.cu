// train.cu -- the training step of the policy (config 5), fp32: forward with saved state, the two losses, the backward pass
// through the Δq head, the FC head (GroupNorm) and the three set-abstraction levels, gradient clipping and Adam.
//
// Replaces, for one data-parallel rank, TrainingMotionPolicyNetwork.training_step (mpinets/model.py:185-240:
// y_hat = clamp(q + net(xyz, q), -1, 1); losses of loss.py:111-166; weights 1 / 5 of jobconfig.yaml:24-25), the
// torch.autograd graph under it (pointnet2_ops' group_points_grad / max_pool2d backward included),
// configure_optimizers (model.py:68-73: Adam, lr 1e-4) and the Trainer's gradient_clip_val (run_training.py:112).
// The DDP all-reduce stays outside: gradients come back as ONE flat fp32 vector in the layout of the parameter vector
// (engine.cu), which the host reduces with NCCL before calling adam_step.
//
// Set-abstraction backward.  The pooled output of a group depends on ONE neighbour row per output channel, so after the
// max-pool only the rows that won at least one channel ("active" rows: <= min(128, C3) of the 128) carry gradient.  The
// forward kernels record the winning row per (group, channel); the backward
//   1. compacts each group's active rows into SLOTS fixed slots (64 for SA1, 128 for SA2 / SA3)        sa_prepare_kernel
//   2. gathers their operand rows [dx,dy,dz,features] and recomputes layers 1-2 for those rows only    sa_gather_kernel + GEMM
//   3. layer 3: dZ2[slot] = sum over the channels pooled from that slot of g_c W3[c], masked by ReLU;
//      dW3[c] += g_c H2[slot_c] -- weight-stationary, accumulators in registers, W3 in shared memory   sa_l3_bwd_kernel
//   4. layers 2 and 1 are dense GEMMs over the compacted rows (data gradient, weight gradient)          linear_kernel / wgrad_kernel
//   5. the feature part of dX is scatter-added to the previous level's feature gradient                 sa_scatter_add_kernel
// in chunks of samples, so the compacted-row scratch stays a few GB.  Weight-gradient reductions are two-level and
// ordered (partials per CTA, then a fixed-order sum): apart from the scatter-add atomics the step is deterministic.
#include <algorithm>
#include <cstdlib>

#include "engine.h"

namespace mpn {

// ---------------------------------------------------------------------------------------------- weight gradient GEMM
// pW[split][n][k] = sum over the split's rows m of dY[m][n] * X[m][k];  pb[split][n] = sum_m dY[m][n]
constexpr int WT = 64, WM = 16;

__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx,
                                                    long long M, int N, int K, long long rows_per_split,
                                                    float* __restrict__ pW, float* __restrict__ pb) {
  __shared__ __align__(16) float ys[WM][WT + 4];
  __shared__ __align__(16) float xs[WM][WT + 4];
  const int nt = (N + WT - 1) / WT;
  const int n0 = (blockIdx.x % nt) * WT, k0 = (blockIdx.x / nt) * WT;
  const int split = blockIdx.y;
  const long long m_begin = (long long)split * rows_per_split;
  const long long m_end = min(M, m_begin + rows_per_split);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;  // ty -> 4 rows of dW (n), tx -> 4 columns (k)
  const int lm = threadIdx.x / 16, lc = threadIdx.x % 16;  // loader: row lm, columns lc + 16 u
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  const bool do_bias = pb != nullptr && k0 == 0 && threadIdx.x < WT;
  for (long long m0 = m_begin; m0 < m_end; m0 += WM) {
    const long long m = m0 + lm;
    const bool mv = m < m_end;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int cc = lc + 16 * u;
      ys[lm][cc] = (mv && n0 + cc < N) ? dY[(size_t)m * ldy + n0 + cc] : 0.f;
      xs[lm][cc] = (mv && k0 + cc < K) ? X[(size_t)m * ldx + k0 + cc] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WM; ++mm) {
      const float4 a4 = *reinterpret_cast<const float4*>(&ys[mm][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&xs[mm][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (do_bias) {
#pragma unroll
      for (int mm = 0; mm < WM; ++mm) bsum += ys[mm][threadIdx.x];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) pW[((size_t)split * N + n) * K + k] = acc[i][j];
    }
  }
  if (do_bias && n0 + (int)threadIdx.x < N) pb[(size_t)split * N + n0 + threadIdx.x] = bsum;
}

// dst[i] (+)= sum_s p[s][i], s in order
__global__ void reduce_partials_kernel(const float* __restrict__ p, int splits, long long n, float* __restrict__ dst, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += p[(size_t)sp * n + i];
  dst[i] = accumulate ? dst[i] + s : s;
}

static int reduce_partials(mpn_ctx* c, cudaStream_t s, const float* p, int splits, long long n, float* dst, int accumulate) {
  reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, splits, n, dst, accumulate);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// gW[N][K] += dY^T X, gb[N] += column sums of dY (gb may be null)
static int wgrad(mpn_ctx* c, cudaStream_t s, const float* dY, int ldy, const float* X, int ldx, long long M, int N, int K, float* gW,
                 float* gb) {
  if (M <= 0) return MPN_OK;
  TrainWs& t = c->tw;
  const int tiles = ((N + WT - 1) / WT) * ((K + WT - 1) / WT);
  long long splits = (8LL * c->sm_count + tiles - 1) / tiles;   // 40 registers, 17 KB smem: 8 CTAs / SM keep the FMA pipe fed
  splits = std::min(splits, (M + 255) / 256);
  const long long per = (long long)N * K + N;
  splits = std::max(1LL, std::min(splits, (long long)(t.partial_floats / per)));
  MPN_REQUIRE((size_t)per <= t.partial_floats, "wgrad: partial buffer too small");
  long long rps = ((M + splits - 1) / splits + WM - 1) / WM * WM;
  splits = (M + rps - 1) / rps;
  float* pW = t.partial;
  float* pb = gb ? pW + (size_t)splits * N * K : nullptr;
  wgrad_kernel<<<dim3(tiles, (unsigned)splits), 256, 0, s>>>(dY, ldy, X, ldx, M, N, K, rps, pW, pb);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  int r;
  if ((r = reduce_partials(c, s, pW, (int)splits, (long long)N * K, gW, 1))) return r;
  if (gb && (r = reduce_partials(c, s, pb, (int)splits, N, gb, 1))) return r;
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------- GroupNorm + LeakyReLU backward
// y = xhat * gamma + beta, a = lrelu(y); dy = da * lrelu'(y).  Parameter gradients: partial sums over a split of rows.
__global__ void __launch_bounds__(256) gn_param_grad_kernel(const float* __restrict__ z, const float* __restrict__ stats,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ da, int M, int C, int groups, int rows_per_split,
                                                            float* __restrict__ pG, float* __restrict__ pB) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const int split = blockIdx.y, gs = C / groups, grp = ch / gs;
  const int r0 = split * rows_per_split, r1 = min(M, r0 + rows_per_split);
  const float gm = gamma[ch], bt = beta[ch];
  float sg = 0.f, sb = 0.f;
  for (int row = r0; row < r1; ++row) {
    const float mean = stats[2 * ((size_t)row * groups + grp)], rstd = stats[2 * ((size_t)row * groups + grp) + 1];
    const float xh = (z[(size_t)row * C + ch] - mean) * rstd;
    const float y = xh * gm + bt;
    const float dy = da[(size_t)row * C + ch] * (y > 0.f ? 1.0f : 0.01f);
    sg = fmaf(dy, xh, sg);
    sb += dy;
  }
  pG[(size_t)split * C + ch] = sg;
  pB[(size_t)split * C + ch] = sb;
}

// one warp per (row, group): g (in: da, out: dz) = rstd * (gamma dy - mean(gamma dy) - xhat mean(gamma dy xhat))
__global__ void __launch_bounds__(256) gn_bwd_kernel(const float* __restrict__ z, const float* __restrict__ stats,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ g,
                                                     int M, int C, int groups) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= M * groups) return;
  const int row = wid / groups, grp = wid % groups, gs = C / groups;
  const float mean = stats[2 * (size_t)wid], rstd = stats[2 * (size_t)wid + 1];
  const float* zp = z + (size_t)row * C + (size_t)grp * gs;
  float* gp = g + (size_t)row * C + (size_t)grp * gs;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < gs; i += 32) {
    const int ch = grp * gs + i;
    const float xh = (zp[i] - mean) * rstd;
    const float y = xh * gamma[ch] + beta[ch];
    const float gd = gamma[ch] * (gp[i] * (y > 0.f ? 1.0f : 0.01f));
    s1 += gd;
    s2 = fmaf(gd, xh, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const float m1 = s1 / (float)gs, m2 = s2 / (float)gs;
  for (int i = lane; i < gs; i += 32) {
    const int ch = grp * gs + i;
    const float xh = (zp[i] - mean) * rstd;
    const float y = xh * gamma[ch] + beta[ch];
    const float gd = gamma[ch] * (gp[i] * (y > 0.f ? 1.0f : 0.01f));
    gp[i] = rstd * (gd - m1 - xh * m2);
  }
}

static int gn_backward(mpn_ctx* c, cudaStream_t s, const float* z, const float* stats, const float* gamma, const float* beta, float* g,
                       int M, int C, float* g_gamma, float* g_beta) {
  TrainWs& t = c->tw;
  int splits = std::max(1, std::min(32, M / 64));
  int rps = (M + splits - 1) / splits;
  splits = (M + rps - 1) / rps;
  float* pG = t.partial;
  float* pB = pG + (size_t)splits * C;
  gn_param_grad_kernel<<<dim3((C + 255) / 256, splits), 256, 0, s>>>(z, stats, gamma, beta, g, M, C, 16, rps, pG, pB);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  int r;
  if ((r = reduce_partials(c, s, pG, splits, C, g_gamma, 1))) return r;
  if ((r = reduce_partials(c, s, pB, splits, C, g_beta, 1))) return r;
  const int warps = M * 16;
  gn_bwd_kernel<<<(warps * 32 + 255) / 256, 256, 0, s>>>(z, stats, gamma, beta, g, M, C, 16);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------- set-abstraction backward
// One warp per group: masks the pooled-output gradient (ReLU / zero gradient), finds the active rows, assigns slots.
template <int C3, int SLOTS>
__global__ void __launch_bounds__(256) sa_prepare_kernel(const uint8_t* __restrict__ arg, float* __restrict__ g,
                                                         const float* __restrict__ out, const int32_t* __restrict__ ball, int G,
                                                         uint8_t* __restrict__ slot_of_ch, int32_t* __restrict__ src) {
  __shared__ unsigned mask_s[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x * 8 + warp;
  if (grp >= G) return;
  unsigned* mk = mask_s[warp];
  if (lane < 4) mk[lane] = 0u;
  __syncwarp();
  const size_t gc = (size_t)grp * C3;
  for (int ch = lane; ch < C3; ch += 32) {
    const float gv = g[gc + ch];
    const bool act = gv != 0.f && out[gc + ch] > 0.f;
    if (!act) {
      g[gc + ch] = 0.f;
    } else {
      const int r = arg[gc + ch] & 127;
      atomicOr(&mk[r >> 5], 1u << (r & 31));
    }
  }
  __syncwarp();
  const unsigned m0 = mk[0], m1 = mk[1], m2 = mk[2], m3 = mk[3];
  const int b1 = __popc(m0), b2 = b1 + __popc(m1), b3 = b2 + __popc(m2);
  auto slot_of_row = [&](int r) {
    const int wd = r >> 5;
    const unsigned w = wd == 0 ? m0 : wd == 1 ? m1 : wd == 2 ? m2 : m3;
    const int base = wd == 0 ? 0 : wd == 1 ? b1 : wd == 2 ? b2 : b3;
    return base + __popc(w & ((1u << (r & 31)) - 1u));
  };
  for (int sl = lane; sl < SLOTS; sl += 32) src[(size_t)grp * SLOTS + sl] = -1;
  __syncwarp();
#pragma unroll
  for (int wd = 0; wd < 4; ++wd) {
    const unsigned w = wd == 0 ? m0 : wd == 1 ? m1 : wd == 2 ? m2 : m3;
    if ((w >> lane) & 1u) {
      const int r = wd * 32 + lane;
      const int sl = slot_of_row(r);
      if (sl < SLOTS) src[(size_t)grp * SLOTS + sl] = ball ? ball[(size_t)grp * NSAMPLE + r] : r;
    }
  }
  for (int ch = lane; ch < C3; ch += 32) {
    int sl = 255;
    if (g[gc + ch] != 0.f) {
      sl = slot_of_row(arg[gc + ch] & 127);
      if (sl >= SLOTS) { sl = 255; g[gc + ch] = 0.f; }
    }
    slot_of_ch[gc + ch] = (uint8_t)sl;
  }
}

// ---- compacted active rows (bf16 mode).  A group of SA1 has ~7 active rows of its 64 slots and a group of SA2 ~25 of 128 (the rows
// that won at least one channel of the max-pool and carry a non-zero gradient), so instead of a fixed slot range per group the active
// rows of a chunk are laid out back to back: pass 1 counts them per group (and masks the pooled-output gradient), a single-CTA scan
// turns the counts into first-row offsets, pass 2 writes, per row, the source point and the owning group.  The row count stays on the
// device (rows[0]; rows[1] = padded to `pad` with empty rows): the GEMMs behind read it there, no host synchronisation.  Slots inside
// a group keep their ascending-row order, so every sum runs in a fixed order (deterministic like the fixed-slot layout).
template <int C3>
__global__ void __launch_bounds__(256) sa_rows_count_kernel(const uint8_t* __restrict__ arg, float* __restrict__ g,
                                                            const float* __restrict__ out, int G, uint4* __restrict__ masks,
                                                            int32_t* __restrict__ cnt) {
  __shared__ unsigned mask_s[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x * 8 + warp;
  if (grp >= G) return;
  unsigned* mk = mask_s[warp];
  if (lane < 4) mk[lane] = 0u;
  __syncwarp();
  const size_t gc = (size_t)grp * C3;
  for (int ch = lane; ch < C3; ch += 32) {
    const float gv = g[gc + ch];
    const bool act = gv != 0.f && out[gc + ch] > 0.f;
    if (!act) g[gc + ch] = 0.f;
    else { const int r = arg[gc + ch] & 127; atomicOr(&mk[r >> 5], 1u << (r & 31)); }
  }
  __syncwarp();
  if (lane == 0) {
    masks[grp] = make_uint4(mk[0], mk[1], mk[2], mk[3]);
    cnt[grp] = __popc(mk[0]) + __popc(mk[1]) + __popc(mk[2]) + __popc(mk[3]);
  }
}

__global__ void __launch_bounds__(1024) sa_rows_scan_kernel(const int32_t* __restrict__ cnt, int G, int32_t* __restrict__ off,
                                                            int* __restrict__ rows, int32_t* __restrict__ src,
                                                            int32_t* __restrict__ row_grp, int pad) {
  // one CTA, warp w owns the contiguous segment [w seg, (w + 1) seg): coalesced 32-wide loads, a first pass for the warp totals, a
  // second one for the running offsets (the counts of a chunk stay in L2 between the passes)
  __shared__ int wtot[32];
  __shared__ int total_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = ((G + 31) / 32 + 31) / 32 * 32;
  const int s0 = min(G, warp * seg), s1 = min(G, s0 + seg);
  int sum = 0;
#pragma unroll 4
  for (int i = s0 + lane; i < s1; i += 32) sum += cnt[i];
  sum = __reduce_add_sync(0xffffffffu, sum);
  if (lane == 0) wtot[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    const int v = wtot[lane];
    int iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += y; }
    wtot[lane] = iv - v;
    if (lane == 31) total_s = iv;
  }
  __syncthreads();
  int carry = wtot[warp];
#pragma unroll 2
  for (int i0 = s0; i0 < s1; i0 += 32) {
    const int i = i0 + lane;
    const int v = i < s1 ? cnt[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (i < s1) off[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  const int total = total_s, padded = (total + pad - 1) / pad * pad;
  if (tid == 0) { off[G] = total; rows[0] = total; rows[1] = padded; }
  for (int r = total + tid; r < padded; r += 1024) { src[r] = -1; row_grp[r] = -1; }
}

// ---- layer 3 of a grouped level on the tensor cores (bf16 mode, compacted rows).  The gradient of the pooled output reaches, per
// channel, ONE row of its group; written out as a matrix dY3 [rows][C3] (bf16, zero elsewhere) it is an ordinary operand:
//   dW3 = dY3^T H2 (wgrad_tc_kernel),  db3 = column sums,  dZ2 = (dY3 W3) * relu'(H2) (rows_gemm_tc_kernel).
// With the rows compacted dY3 is small (SA1 ~1 M rows x 64, SA2 ~0.8 M rows x 256 per chunk of 256 samples) and the three products cost
// a fraction of the per-group list building of sa_l3_bwd_kernel (which stays for the fp32 mode and the fixed-slot layout).
__global__ void __launch_bounds__(256) zero_rows_kernel(uint4* __restrict__ p, int chunks_per_row, const int* __restrict__ rows_dev) {
  const long long n = (long long)(*rows_dev) * chunks_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}
// dY3[grp_off[grp] + slot_of_ch][ch] = g[grp][ch]; C3 = 64: one matrix [rows][64]; C3 = 256: two matrices [rows][128] (channel halves)
template <int C3>
__global__ void __launch_bounds__(256) sa_dy3_scatter_kernel(const float* __restrict__ g, const uint8_t* __restrict__ slot_of_ch,
                                                             const int32_t* __restrict__ grp_off, long long n, __nv_bfloat16* __restrict__ da,
                                                             __nv_bfloat16* __restrict__ db) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int sl = slot_of_ch[i];
    if (sl == 255) continue;
    const long long grp = i / C3;
    const int ch = (int)(i - grp * C3);
    const size_t row = (size_t)grp_off[grp] + sl;
    const __nv_bfloat16 v = __float2bfloat16_rn(g[i]);
    if (C3 == 64) da[row * 64 + ch] = v;
    else if (ch < 128) da[row * 128 + ch] = v;
    else db[row * 128 + ch - 128] = v;
  }
}
// H2 <- (T1 + T2) * relu'(H2), rows [0, *rows_dev) x 128 bf16
__global__ void __launch_bounds__(256) add_mask_rows_kernel(const uint4* __restrict__ t1, const uint4* __restrict__ t2, uint4* __restrict__ h2,
                                                            const int* __restrict__ rows_dev) {
  const long long n = (long long)(*rows_dev) * 16;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint4 a = t1[i], b = t2[i], m = h2[i];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w}, mw[4] = {m.x, m.y, m.z, m.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = __uint_as_float(aw[k] << 16) + __uint_as_float(bw[k] << 16);
      const float hi = __uint_as_float(aw[k] & 0xFFFF0000u) + __uint_as_float(bw[k] & 0xFFFF0000u);
      const bool mlo = (mw[k] & 0x7FFFu) != 0u && (mw[k] & 0x8000u) == 0u, mhi = (mw[k] & 0x7FFF0000u) != 0u && (mw[k] & 0x80000000u) == 0u;
      __nv_bfloat162 v = __floats2bfloat162_rn(mlo ? lo : 0.f, mhi ? hi : 0.f);
      o[k] = *reinterpret_cast<uint32_t*>(&v);
    }
    h2[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

template <int C3>
__global__ void __launch_bounds__(256) sa_rows_fill_kernel(const uint8_t* __restrict__ arg, const float* __restrict__ g,
                                                           const int32_t* __restrict__ ball, int G, const uint4* __restrict__ masks,
                                                           const int32_t* __restrict__ off, uint8_t* __restrict__ slot_of_ch,
                                                           int32_t* __restrict__ src, int32_t* __restrict__ row_grp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x * 8 + warp;
  if (grp >= G) return;
  const uint4 mk = masks[grp];
  const unsigned m0 = mk.x, m1 = mk.y, m2 = mk.z, m3 = mk.w;
  const int b1 = __popc(m0), b2 = b1 + __popc(m1), b3 = b2 + __popc(m2);
  auto slot_of_row = [&](int r) {
    const int wd = r >> 5;
    const unsigned w = wd == 0 ? m0 : wd == 1 ? m1 : wd == 2 ? m2 : m3;
    const int base = wd == 0 ? 0 : wd == 1 ? b1 : wd == 2 ? b2 : b3;
    return base + __popc(w & ((1u << (r & 31)) - 1u));
  };
  const int o = off[grp];
#pragma unroll
  for (int wd = 0; wd < 4; ++wd) {
    const unsigned w = wd == 0 ? m0 : wd == 1 ? m1 : wd == 2 ? m2 : m3;
    if ((w >> lane) & 1u) {
      const int r = wd * 32 + lane;
      const int row = o + slot_of_row(r);
      src[row] = ball[(size_t)grp * NSAMPLE + r];
      row_grp[row] = grp;
    }
  }
  const size_t gc = (size_t)grp * C3;
  for (int ch = lane; ch < C3; ch += 32)
    slot_of_ch[gc + ch] = g[gc + ch] != 0.f ? (uint8_t)slot_of_row(arg[gc + ch] & 127) : (uint8_t)255;
}

// Group-all level with saved activations (bf16 mode): rows keep their natural order, so slot = pooled row and src = row.
__global__ void __launch_bounds__(256) sa3_identity_prepare_kernel(const uint8_t* __restrict__ arg, float* __restrict__ g,
                                                                   const float* __restrict__ out, int G, uint8_t* __restrict__ slot_of_ch,
                                                                   int32_t* __restrict__ src) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)G * 128) src[i] = (int32_t)(i & 127);
  if (i >= (size_t)G * 1024) return;
  const bool act = g[i] != 0.f && out[i] > 0.f;
  if (!act) g[i] = 0.f;
  slot_of_ch[i] = act ? (uint8_t)(arg[i] & 127) : (uint8_t)255;
}
__global__ void __launch_bounds__(256) widen_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long n, float* __restrict__ dst) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const uint4 v = *reinterpret_cast<const uint4*>(src + i);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float o[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) { o[2 * k] = __uint_as_float(w[k] << 16); o[2 * k + 1] = __uint_as_float(w[k] & 0xFFFF0000u); }
  *reinterpret_cast<float4*>(dst + i) = make_float4(o[0], o[1], o[2], o[3]);
  *reinterpret_cast<float4*>(dst + i + 4) = make_float4(o[4], o[5], o[6], o[7]);
}

__global__ void __launch_bounds__(256) pack_weight_kernel_f32(const float* __restrict__ src, int n, __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}
__global__ void __launch_bounds__(256) narrow_bf16_kernel(const float* __restrict__ src, long long n, __nv_bfloat16* __restrict__ dst) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const float4 a = *reinterpret_cast<const float4*>(src + i), b = *reinterpret_cast<const float4*>(src + i + 4);
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w), p2 = __floats2bfloat162_rn(b.x, b.y),
                 p3 = __floats2bfloat162_rn(b.z, b.w);
  *reinterpret_cast<uint4*>(dst + i) = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                  *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
}
// act[i] = act[i] > 0 ? g[i] : 0   (ReLU derivative applied to a data gradient, in place over the activation)
__global__ void __launch_bounds__(256) relu_mask_kernel(float* __restrict__ act, const float* __restrict__ g, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 a = *reinterpret_cast<float4*>(act + i);
  const float4 v = *reinterpret_cast<const float4*>(g + i);
  a.x = a.x > 0.f ? v.x : 0.f; a.y = a.y > 0.f ? v.y : 0.f; a.z = a.z > 0.f ? v.z : 0.f; a.w = a.w > 0.f ? v.w : 0.f;
  *reinterpret_cast<float4*>(act + i) = a;
}

// X[row][:] = [p[src] - centroid | features[src] | 0-pad]; empty slots -> zero rows
template <int CFEAT, int CINP, bool CENTER>
__global__ void __launch_bounds__(256) sa_gather_kernel(const int32_t* __restrict__ src, long long R, int slots, int npoint,
                                                        const float* __restrict__ xyz, int stride, int N, const float* __restrict__ feats,
                                                        int fstride, const float* __restrict__ new_xyz, float* __restrict__ X) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * CINP) return;
  const long long row = i / CINP;
  const int k = (int)(i - row * CINP);
  const long long grp = row / slots;
  const long long b = grp / npoint;
  const int si = src[row];
  float v = 0.f;
  if (si >= 0 && k < 3 + CFEAT) {
    if (k < 3) {
      v = xyz[((size_t)b * N + si) * stride + k];
      if (CENTER) v -= new_xyz[(size_t)grp * 3 + k];
    } else {
      v = feats[((size_t)b * N + si) * fstride + (k - 3)];
    }
  }
  X[i] = v;
}

// Layer 3 of a grouped level, persistent CTAs (W3 resident in shared memory, dW3 / db3 accumulators in registers):
//   H2 (in: relu activations of the group's slots, out: dZ2);  pW[cta][C3][C2], pb[cta][C3] partial sums.
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The group's activations H2 are staged RAW (bf16 in the tensor-core mode) by 16-byte cp.async; in the bf16 mode they are
// double-buffered, so the next group's 32 KB arrive while this group is processed (ncu before this change: half of the
// kernel's stall samples sat on the scalar 2-byte loads of this stage and on the barrier behind them, at one CTA per SM).
template <int C2, int C3, int SLOTS, typename T>
__global__ void __launch_bounds__(512) sa_l3_bwd_kernel(const float* __restrict__ geff, const uint8_t* __restrict__ slot_of_ch,
                                                        const float* __restrict__ W3, T* __restrict__ H2, int G,
                                                        float* __restrict__ pW, float* __restrict__ pb,
                                                        const int32_t* __restrict__ grp_off = nullptr) {
  // grp_off != null: compacted rows -- group grp owns rows [grp_off[grp], grp_off[grp + 1]) of H2 instead of a fixed SLOTS range
  constexpr int NT = 512, Q = NT / C2, CPT = C3 / Q, PER = SLOTS / 32;
  constexpr int NBUF = sizeof(T) == 2 ? 2 : 1, CHUNKS = SLOTS * C2 * (int)sizeof(T) / 16, PARTS = NT / SLOTS, CPP = C3 / PARTS;
  static_assert(NT % C2 == 0 && C3 % Q == 0 && SLOTS % 32 == 0 && SLOTS <= 128 && C3 <= NT && NT % SLOTS == 0 && C3 % PARTS == 0, "layout");
  extern __shared__ __align__(16) float sm[];
  float* W3s = sm;                                  // [C3][C2]
  T* Hbuf = reinterpret_cast<T*>(W3s + C3 * C2);    // [NBUF][SLOTS][C2] raw
  float* gs = reinterpret_cast<float*>(Hbuf + (size_t)NBUF * SLOTS * C2);   // [C3]
  int* sl = reinterpret_cast<int*>(gs + C3);        // [C3]
  int* cnt = sl + C3;                               // [PARTS][SLOTS]
  int* start = cnt + NT;                            // [SLOTS + 1]
  int* chl = start + SLOTS + 1;                     // [C3]
  const int tid = threadIdx.x, j = tid % C2, q = tid / C2;
  float accW[CPT];
#pragma unroll
  for (int i = 0; i < CPT; ++i) accW[i] = 0.f;
  float accb = 0.f;
  for (int i = tid; i < C3 * C2; i += NT) W3s[i] = W3[i];
  auto stage = [&](int grp, int buf) {
    const size_t row0 = grp_off ? (size_t)grp_off[grp] : (size_t)grp * SLOTS;
    const int chunks = grp_off ? (grp_off[grp + 1] - grp_off[grp]) * C2 * (int)sizeof(T) / 16 : CHUNKS;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(H2 + row0 * C2);
    uint8_t* dst = reinterpret_cast<uint8_t*>(Hbuf + (size_t)buf * SLOTS * C2);
    for (int i = tid; i < chunks; i += NT) cp_async16(dst + (size_t)i * 16, src + (size_t)i * 16);
    cp_async_commit();
  };
  if (NBUF == 2 && (int)blockIdx.x < G) stage(blockIdx.x, 0);
  int it = 0;
  for (int grp = blockIdx.x; grp < G; grp += gridDim.x, ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    const T* Hs = Hbuf + (size_t)buf * SLOTS * C2;
    __syncthreads();
    for (int c = tid; c < C3; c += NT) {
      gs[c] = geff[(size_t)grp * C3 + c];
      sl[c] = slot_of_ch[(size_t)grp * C3 + c];
    }
    T* Hg = H2 + (grp_off ? (size_t)grp_off[grp] : (size_t)grp * SLOTS) * C2;
    const int nrows = grp_off ? grp_off[grp + 1] - grp_off[grp] : SLOTS;
    if (NBUF == 2) {
      const int next = grp + gridDim.x;
      if (next < G) { stage(next, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    } else {
      stage(grp, 0);
      cp_async_wait<0>();
    }
    __syncthreads();
    // channel lists per slot, in channel order (deterministic): thread (slot, part) scans its share of the channels
    {
      const int ls = tid % SLOTS, part = tid / SLOTS;
      int n = 0;
#pragma unroll 8
      for (int c = part * CPP; c < (part + 1) * CPP; ++c) n += (sl[c] == ls) ? 1 : 0;
      cnt[part * SLOTS + ls] = n;
    }
    __syncthreads();
    if (tid < 32) {
      int loc[PER], sum = 0;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        int tot = 0;
#pragma unroll
        for (int pp = 0; pp < PARTS; ++pp) tot += cnt[pp * SLOTS + tid * PER + u];
        loc[u] = tot; sum += tot;
      }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      int ex = incl - sum;
#pragma unroll
      for (int u = 0; u < PER; ++u) { start[tid * PER + u] = ex; ex += loc[u]; }
      if (tid == 31) start[SLOTS] = incl;
    }
    __syncthreads();
    {
      const int ls = tid % SLOTS, part = tid / SLOTS;
      int p = start[ls];
      for (int pp = 0; pp < part; ++pp) p += cnt[pp * SLOTS + ls];
#pragma unroll 8
      for (int c = part * CPP; c < (part + 1) * CPP; ++c)
        if (sl[c] == ls) chl[p++] = c;
    }
    __syncthreads();
    // dZ2[slot] = (sum over the slot's channels of g_c W3[c]) masked by ReLU
    for (int s = q; s < nrows; s += Q) {
      float a = 0.f;
      const int e1 = start[s + 1];
      for (int e = start[s]; e < e1; ++e) {
        const int c = chl[e];
        a = fmaf(gs[c], W3s[c * C2 + j], a);
      }
      stf(Hg + (size_t)s * C2 + j, ldf(Hs + s * C2 + j) > 0.f ? a : 0.f);
    }
    // dW3[c] += g_c H2[slot_c]
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int c = q * CPT + i;
      const int s = sl[c];
      if (s != 255) accW[i] = fmaf(gs[c], ldf(Hs + s * C2 + j), accW[i]);
    }
    if (tid < C3) accb += gs[tid];
  }
#pragma unroll
  for (int i = 0; i < CPT; ++i) pW[(size_t)blockIdx.x * C3 * C2 + (size_t)(q * CPT + i) * C2 + j] = accW[i];
  if (tid < C3) pb[(size_t)blockIdx.x * C3 + tid] = accb;
}

// group-all level (C3 = 1024, C2 = 512): dW3 partials over a split of samples; grid (C3 / 8, splits), thread = column
__global__ void __launch_bounds__(512) sa3_dw3_kernel(const float* __restrict__ geff, const uint8_t* __restrict__ slot_of_ch,
                                                      const float* __restrict__ H2, int B, int per_split, float* __restrict__ pW) {
  constexpr int C2 = 512, C3 = 1024, SLOTS = 128;
  const int c0 = blockIdx.x * 8, split = blockIdx.y, j = threadIdx.x;
  const int b0 = split * per_split, b1 = min(B, b0 + per_split);
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int b = b0; b < b1; ++b) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int s = slot_of_ch[(size_t)b * C3 + c0 + i];
      if (s != 255) acc[i] = fmaf(geff[(size_t)b * C3 + c0 + i], H2[((size_t)b * SLOTS + s) * C2 + j], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) pW[((size_t)split * C3 + c0 + i) * C2 + j] = acc[i];
}

// group-all level: H2 (in: activations, out: dZ2) for one sample per CTA; W3 read from L2
__global__ void __launch_bounds__(512) sa3_dz2_kernel(const float* __restrict__ geff, const uint8_t* __restrict__ slot_of_ch,
                                                      const float* __restrict__ W3, float* __restrict__ H2) {
  constexpr int C2 = 512, C3 = 1024, SLOTS = 128;
  __shared__ float gs[C3];
  __shared__ uint8_t sl[C3];
  __shared__ int cnt[SLOTS], start[SLOTS + 1];
  __shared__ short chl[C3];
  const int tid = threadIdx.x, b = blockIdx.x;
  for (int c = tid; c < C3; c += 512) { gs[c] = geff[(size_t)b * C3 + c]; sl[c] = slot_of_ch[(size_t)b * C3 + c]; }
  __syncthreads();
  if (tid < SLOTS) {
    int n = 0;
    for (int c = 0; c < C3; ++c) n += (sl[c] == tid) ? 1 : 0;
    cnt[tid] = n;
  }
  __syncthreads();
  if (tid < 32) {
    int loc[4], sum = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { loc[u] = cnt[tid * 4 + u]; sum += loc[u]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += v;
    }
    int ex = incl - sum;
#pragma unroll
    for (int u = 0; u < 4; ++u) { start[tid * 4 + u] = ex; ex += loc[u]; }
    if (tid == 31) start[SLOTS] = incl;
  }
  __syncthreads();
  if (tid < SLOTS) {
    int p = start[tid];
    for (int c = 0; c < C3; ++c)
      if (sl[c] == tid) chl[p++] = (short)c;
  }
  __syncthreads();
  float* Hb = H2 + (size_t)b * SLOTS * C2;
  for (int s = 0; s < SLOTS; ++s) {
    float a = 0.f;
    const int e1 = start[s + 1];
    for (int e = start[s]; e < e1; ++e) {
      const int c = chl[e];
      a = fmaf(gs[c], __ldg(W3 + (size_t)c * C2 + tid), a);
    }
    const float h = Hb[(size_t)s * C2 + tid];
    Hb[(size_t)s * C2 + tid] = h > 0.f ? a : 0.f;
  }
}

// column sums of Y[M][N] (ordered): one CTA per 32 columns, 8 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ Y, int ld, int M, int N, float* __restrict__ dst) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  float s = 0.f;
  if (col < N)
    for (int m = rl; m < M; m += 8) s += Y[(size_t)m * ld + col];
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && col < N) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += red[i][threadIdx.x];
    dst[col] += tsum;
  }
}

// column sums over many rows: grid (column blocks of 32, row splits) -> partial[split][N], summed in split order by colsum_finish_kernel
__global__ void __launch_bounds__(256) colsum_split_kernel(const float* __restrict__ Y, int ld, int M, int N, int rows_per_split,
                                                           float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  const int m0 = blockIdx.y * rows_per_split, m1 = min(M, m0 + rows_per_split);
  float s0 = 0.f, s1 = 0.f;
  if (col < N) {
    int m = m0 + rl;
    for (; m + 8 < m1; m += 16) { s0 += Y[(size_t)m * ld + col]; s1 += Y[(size_t)(m + 8) * ld + col]; }
    if (m < m1) s0 += Y[(size_t)m * ld + col];
  }
  red[rl][threadIdx.x & 31] = s0 + s1;
  __syncthreads();
  if (rl == 0 && col < N) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += red[i][threadIdx.x];
    partial[(size_t)blockIdx.y * N + col] = tsum;
  }
}
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* __restrict__ partial, int splits, int N, float* __restrict__ dst) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  float s = 0.f;
  for (int i = 0; i < splits; ++i) s += partial[(size_t)i * N + col];
  dst[col] += s;
}
// dst[col] += sum over the M rows of Y[:, col]; deterministic (fixed split, ordered finish).  Uses the tail of the partial buffer so that
// it can follow a weight-gradient launch whose partials are still being reduced on the stream.
static int colsum_into(mpn_ctx* c, cudaStream_t s, const float* Y, int ld, int M, int N, float* dst) {
  TrainWs& t = c->tw;
  if (M <= 512) {
    colsum_kernel<<<(N + 31) / 32, 256, 0, s>>>(Y, ld, M, N, dst);
    c->launches++;
  } else {
    int splits = std::min(64, (M + 255) / 256);
    const int rps = (M + splits - 1) / splits;
    splits = (M + rps - 1) / rps;
    float* p = t.partial + t.partial_floats - (size_t)64 * 4096;
    MPN_REQUIRE(N <= 4096, "colsum_into: at most 4096 columns");
    colsum_split_kernel<<<dim3((N + 31) / 32, splits), 256, 0, s>>>(Y, ld, M, N, rps, p);
    colsum_finish_kernel<<<(N + 255) / 256, 256, 0, s>>>(p, splits, N, dst);
    c->launches += 2;
  }
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// previous level's feature gradient: dfeat[b][src][f] += dX[row][f]   (pointnet2's group_points_grad)
template <typename T>
__global__ void __launch_bounds__(256) sa_scatter_add_kernel(const T* __restrict__ dX, const int32_t* __restrict__ src, long long R,
                                                             int slots, int npoint, int N, int CF, float* __restrict__ dfeat,
                                                             const int32_t* __restrict__ row_grp = nullptr,
                                                             const int* __restrict__ rows_dev = nullptr) {
  if (rows_dev) R = *rows_dev;   // compacted rows: count on the device, group of a row from row_grp
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * CF; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / CF;
    const int f = (int)(i - row * CF);
    const int si = src[row];
    if (si < 0) continue;
    const long long b = (row_grp ? (long long)row_grp[row] : row / slots) / npoint;
    atomicAdd(dfeat + ((size_t)b * N + si) * CF + f, ldf(dX + i));
  }
}


// ---------------------------------------------------------------------------------------------- bf16 / tensor-core variant
// (MPN_PREC_BF16 training: the compacted-row GEMMs of SA1 / SA2 run on tcgen05, train_tc.cu; accumulators, the sparse layer 3,
//  the pooled-feature gradients and every parameter gradient stay fp32)

// SA2 operand rows, bf16 [R][128] = [dx dy dz | 64 features | 0 x 60 | 1]: the ones column makes column 127 of dZ1^T X the
// bias gradient of layer 1
__global__ void __launch_bounds__(256) sa2_gather_bf16_kernel(const int32_t* __restrict__ src, long long R, int slots, int npoint,
                                                              const float* __restrict__ xyz, int N, const float* __restrict__ feats,
                                                              const float* __restrict__ new_xyz, __nv_bfloat16* __restrict__ X,
                                                              const int32_t* __restrict__ row_grp = nullptr,
                                                              const int* __restrict__ rows_dev = nullptr) {
  if (rows_dev) R = *rows_dev;   // compacted rows (padded count on the device), group of a row from row_grp
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * 16; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i >> 4;
    const int ch = (int)(i & 15);
    const int si = src[row];
    const long long grp = si < 0 ? 0 : (row_grp ? (long long)row_grp[row] : row / slots), b = grp / npoint;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (si >= 0) {
      const float* p = xyz + ((size_t)b * N + si) * 3;
      const float* f = feats + ((size_t)b * N + si) * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = ch * 8 + j;
        if (k < 3) v[j] = p[k] - new_xyz[(size_t)grp * 3 + k];
        else if (k < 67) v[j] = f[k - 3];
        else if (k == 127) v[j] = 1.0f;
      }
    }
    __nv_bfloat162 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(X + (size_t)row * 128 + ch * 8) = *reinterpret_cast<uint4*>(o);
  }
}

// SA1: operand rows X4 fp32 [R][4] = [dx dy dz mask] and layer 1 (4 -> 64, K too small for a tensor-core tile) computed on
// the spot: H1 bf16 [R][64] = relu(W1 x + b1); empty slots give zero rows
__global__ void __launch_bounds__(256) sa1_gather_h1_kernel(const int32_t* __restrict__ src, long long R, int slots, int npoint,
                                                            const float* __restrict__ cloud, int N, const float* __restrict__ new_xyz,
                                                            const float* __restrict__ W1, const float* __restrict__ b1,
                                                            float* __restrict__ X4, __nv_bfloat16* __restrict__ H1,
                                                            const int32_t* __restrict__ row_grp = nullptr,
                                                            const int* __restrict__ rows_dev = nullptr) {
  __shared__ float w[4][72], bs[64];   // w[k][c], rows padded: the 8 channel-chunk threads of a row read distinct banks
  w[threadIdx.x & 3][threadIdx.x >> 2] = W1[threadIdx.x];
  if (threadIdx.x < 64) bs[threadIdx.x] = b1[threadIdx.x];
  __syncthreads();
  if (rows_dev) R = *rows_dev;   // compacted rows (padded count on the device), group of a row from row_grp
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < R * 8; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i >> 3;
    const int ch = (int)(i & 7);
    const int si = src[row];
    const long long grp = si < 0 ? 0 : (row_grp ? (long long)row_grp[row] : row / slots), b = grp / npoint;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    if (si >= 0) {
      const float4 p = *reinterpret_cast<const float4*>(cloud + ((size_t)b * N + si) * 4);
      x0 = p.x - new_xyz[(size_t)grp * 3]; x1 = p.y - new_xyz[(size_t)grp * 3 + 1]; x2 = p.z - new_xyz[(size_t)grp * 3 + 2]; x3 = p.w;
    }
    if (ch == 0 && X4) *reinterpret_cast<float4*>(X4 + (size_t)row * 4) = make_float4(x0, x1, x2, x3);
    float h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      const float z = fmaf(w[3][c], x3, fmaf(w[2][c], x2, fmaf(w[1][c], x1, fmaf(w[0][c], x0, bs[c]))));
      h[j] = si >= 0 ? fmaxf(z, 0.f) : 0.f;
    }
    __nv_bfloat162 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(h[2 * j], h[2 * j + 1]);
    *reinterpret_cast<uint4*>(H1 + (size_t)row * 64 + ch * 8) = *reinterpret_cast<uint4*>(o);
  }
}

// column sums of Y bf16 [R][128] over a CTA's row range -> partial[cta][128]
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ Y, long long R, long long rows_per_cta,
                                                          float* __restrict__ partial, const int* __restrict__ rows_dev = nullptr,
                                                          int rows_shift = 0) {
  __shared__ float red[4][128];
  if (rows_dev) { R = (long long)(*rows_dev >> rows_shift); rows_per_cta = (R + gridDim.x - 1) / gridDim.x; }
  const int cp = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  float s0 = 0.f, s1 = 0.f;
  for (long long r = r0 + rl; r < r1; r += 4) {
    const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(Y + (size_t)r * 128)[cp];
    s0 += __low2float(v); s1 += __high2float(v);
  }
  red[rl][2 * cp] = s0; red[rl][2 * cp + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 128)
    partial[(size_t)blockIdx.x * 128 + threadIdx.x] = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
}

// gW[o][k] += sum_s P[s][o][k] (+ P[s][o + 64][k + 64] for SA1's paired rows); optional bias from column `bias_col`
__global__ void wgrad_extract_kernel(const float* __restrict__ P, int n, int out, int in, int paired, float* __restrict__ gW, int bias_col,
                                     float* __restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < out * in) {
    const int o = i / in, k = i - o * in;
    float s = 0.f;
    for (int sp = 0; sp < n; ++sp) {
      s += P[((size_t)sp * 128 + o) * 128 + k];
      if (paired) s += P[((size_t)sp * 128 + o + 64) * 128 + k + 64];
    }
    gW[i] += s;
  } else if (bias_col >= 0 && i < out * in + out) {
    const int o = i - out * in;
    float s = 0.f;
    for (int sp = 0; sp < n; ++sp) s += P[((size_t)sp * 128 + o) * 128 + bias_col];
    gb[o] += s;
  }
}
// one CTA of 1024 threads = 8 partial lanes x 128 columns; lanes summed in order (deterministic)
__global__ void __launch_bounds__(1024) bias_extract_kernel(const float* __restrict__ P, int n, int out, int paired, float* __restrict__ gb) {
  __shared__ float red[8][128];
  const int o = threadIdx.x & 127, rl = threadIdx.x >> 7;
  float s = 0.f;
  for (int sp = rl; sp < n; sp += 8) s += P[(size_t)sp * 128 + o];
  red[rl][o] = s;
  __syncthreads();
  if (rl == 0 && o < out) {
    float tsum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tsum += red[i][o] + (paired ? red[i][o + 64] : 0.f);
    gb[o] += tsum;
  }
}

// SA1 layer 1 (64 x 4): partial[cta][64][5] = sums over the CTA's rows of dZ1[r][c] * [x0 x1 x2 x3 1]
__global__ void __launch_bounds__(256) sa1_wgrad1_kernel(const __nv_bfloat16* __restrict__ dZ1, const float* __restrict__ X4, long long R,
                                                         long long rows_per_cta, float* __restrict__ partial,
                                                         const int* __restrict__ rows_dev = nullptr) {
  __shared__ float red[4][64][5];
  if (rows_dev) { R = *rows_dev; rows_per_cta = (R + gridDim.x - 1) / gridDim.x; }
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
  for (long long r = r0 + rl; r < r1; r += 4) {
    const float d = __bfloat162float(dZ1[(size_t)r * 64 + c]);
    const float4 x = *reinterpret_cast<const float4*>(X4 + (size_t)r * 4);
    a0 = fmaf(d, x.x, a0); a1 = fmaf(d, x.y, a1); a2 = fmaf(d, x.z, a2); a3 = fmaf(d, x.w, a3); a4 += d;
  }
  red[rl][c][0] = a0; red[rl][c][1] = a1; red[rl][c][2] = a2; red[rl][c][3] = a3; red[rl][c][4] = a4;
  __syncthreads();
  for (int i = threadIdx.x; i < 320; i += 256) {
    const int cc = i / 5, k = i - cc * 5;
    partial[(size_t)blockIdx.x * 320 + i] = red[0][cc][k] + red[1][cc][k] + red[2][cc][k] + red[3][cc][k];
  }
}
__global__ void sa1_w1_extract_kernel(const float* __restrict__ P, int n, float* __restrict__ gW, float* __restrict__ gb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 320) return;
  float s = 0.f;
  for (int sp = 0; sp < n; ++sp) s += P[(size_t)sp * 320 + i];
  const int c = i / 5, k = i - c * 5;
  if (k < 4) gW[c * 4 + k] += s; else gb[c] += s;
}

// dst bf16 [dst_rows][128]: mode 0 zero-padded copy of src[rows][cols] (row pitch ld); mode 1 block-diagonal diag(src, src)
// of a 64 x 64 matrix (SA1's rows are processed in pairs)
__global__ void pack_train_weight_kernel(const float* __restrict__ src, int rows, int cols, int ld, int mode, int dst_rows,
                                         __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dst_rows * 128) return;
  const int r = i >> 7, c = i & 127;
  float v = 0.f;
  if (mode == 0) { if (r < rows && c < cols) v = src[(size_t)r * ld + c]; }
  else if ((r >> 6) == (c >> 6)) v = src[(size_t)(r & 63) * ld + (c & 63)];
  dst[i] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------- small elementwise kernels
__global__ void yhat_kernel(const float* __restrict__ qn, const float* __restrict__ dq, int n, float* __restrict__ yhat) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) yhat[i] = fminf(1.0f, fmaxf(-1.0f, __fadd_rn(qn[i], dq[i])));
}
// torch.clamp backward: the gradient passes where min <= x <= max
__global__ void clamp_bwd_kernel(const float* __restrict__ qn, const float* __restrict__ dq, int n, float* __restrict__ g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = __fadd_rn(qn[i], dq[i]);
  if (!(x >= -1.0f && x <= 1.0f)) g[i] = 0.f;
}
__global__ void transpose_kernel(const float* __restrict__ w, int out, int in, float* __restrict__ wt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out * in) return;
  int o = i / in, k = i - o * in;
  wt[(size_t)k * out + o] = w[i];
}

// ---------------------------------------------------------------------------------------------- optimiser
constexpr int NORM_CTAS = 512;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  __shared__ float red[256];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)NORM_CTAS * 256) s = fmaf(g[i], g[i], s);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(NORM_CTAS) sumsq_final_kernel(const float* __restrict__ partial, float* __restrict__ norm) {
  __shared__ float red[NORM_CTAS];
  red[threadIdx.x] = partial[threadIdx.x];
  __syncthreads();
  for (int o = NORM_CTAS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norm[0] = sqrtf(red[0]);
}
// torch.optim.Adam (no weight decay, no amsgrad) after torch.nn.utils.clip_grad_norm_(max_norm = clip)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float clip, const float* __restrict__ norm) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float scale = 1.0f;
  if (clip > 0.f) scale = fminf(1.0f, clip / (norm[0] + 1e-6f));
  const float gi = g[i] * scale;
  const float mi = m[i] + (1.0f - b1) * (gi - m[i]);
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

int refresh_transposes(mpn_ctx* c, cudaStream_t s) {
  Weights& W = c->w;
  std::vector<Linear*> all;
  for (int m = 0; m < 3; ++m) for (int l = 0; l < 3; ++l) all.push_back(&W.sa[m][l]);
  for (int l = 0; l < 3; ++l) all.push_back(&W.fc[l]);
  for (int l = 0; l < 5; ++l) all.push_back(&W.fe[l]);
  for (int l = 0; l < 4; ++l) all.push_back(&W.dec[l]);
  for (Linear* L : all) {
    const int n = L->in * L->out;
    transpose_kernel<<<(n + 255) / 256, 256, 0, s>>>(L->w, L->out, L->in, L->wt);
    c->launches++;
  }
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------- workspace
template <typename T>
static int talloc(T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (cudaMalloc((void**)p, n * sizeof(T)) != cudaSuccess) {
    set_error("training workspace: cudaMalloc(%zu bytes) failed", n * sizeof(T));
    return MPN_ERR_NOMEM;
  }
  return MPN_OK;
}

void free_train_ws(mpn_ctx* c) {
  TrainWs& t = c->tw;
  void** ptrs[] = {(void**)&t.fps_idx, (void**)&t.ball1, (void**)&t.ball2, (void**)&t.arg1, (void**)&t.arg2, (void**)&t.arg3,
                   (void**)&t.dx, (void**)&t.dgy, (void**)&t.dgyT, (void**)&t.daT, (void**)&t.dw, (void**)&t.dwT, (void**)&t.zeros,
                   (void**)&t.z1, (void**)&t.a1, (void**)&t.z2, (void**)&t.a2, (void**)&t.st1, (void**)&t.st2, (void**)&t.f[0],
                   (void**)&t.f[1], (void**)&t.f[2], (void**)&t.f[3], (void**)&t.d[0], (void**)&t.d[1], (void**)&t.d[2],
                   (void**)&t.yhat, (void**)&t.gy, (void**)&t.ga, (void**)&t.gb, (void**)&t.gcat, (void**)&t.gfeat3, (void**)&t.gfeat2,
                   (void**)&t.gfeat1, (void**)&t.X, (void**)&t.H1, (void**)&t.H2, (void**)&t.src, (void**)&t.slot, (void**)&t.partial,
                   (void**)&t.tcw, (void**)&t.b2dup, (void**)&t.row_grp, (void**)&t.grp_off, (void**)&t.grp_cnt, (void**)&t.grp_mask,
                   (void**)&t.rows,
                   (void**)&t.adam_m, (void**)&t.adam_v, (void**)&t.norm};
  for (auto p : ptrs)
    if (*p) { cudaFree(*p); *p = nullptr; }
  t.capacity = 0; t.chunk = 0; t.partial_floats = 0;
}

// samples per backward chunk of the set-abstraction levels (scratch is sized for every slot of every group of a chunk: ~21 MB per
// sample).  1024 since the rows are compacted: the launches of a chunk work on ~1/6 of that capacity, so larger chunks mean fewer, fuller
// launches (4096 samples: 58.0 k samples/s at 256, 62.4 k at 512, 65.3 k at 1024)
static int train_chunk_size(int B) {
  int chunk = 1024;
  if (const char* e = getenv("MPN_TRAIN_CHUNK")) chunk = std::max(1, atoi(e));
  return std::min(B, chunk);
}

static int ensure_train_ws(mpn_ctx* c, int B, int N) {
  TrainWs& t = c->tw;
  const int chunk = train_chunk_size(B);
  if (B <= t.capacity && chunk <= t.chunk && N <= t.n_points) return MPN_OK;
  float* keep_m = t.adam_m; float* keep_v = t.adam_v; float* keep_n = t.norm;   // optimiser state survives a resize
  t.adam_m = t.adam_v = t.norm = nullptr;
  free_train_ws(c);
  t.adam_m = keep_m; t.adam_v = keep_v; t.norm = keep_n;
  const size_t b = (size_t)B, k = (size_t)chunk;
  int r = 0;
  r |= talloc(&t.fps_idx, b * SA1_NPOINT);
  r |= talloc(&t.ball1, b * SA1_NPOINT * NSAMPLE);
  r |= talloc(&t.ball2, b * SA2_NPOINT * NSAMPLE);
  r |= talloc(&t.arg1, b * SA1_NPOINT * 64);
  r |= talloc(&t.arg2, b * SA2_NPOINT * 256);
  r |= talloc(&t.arg3, b * 1024);
  r |= talloc(&t.z1, b * 4096); r |= talloc(&t.a1, b * 4096);
  r |= talloc(&t.z2, b * 2048); r |= talloc(&t.a2, b * 2048);
  r |= talloc(&t.st1, b * 32); r |= talloc(&t.st2, b * 32);
  r |= talloc(&t.f[0], b * 32); r |= talloc(&t.f[1], b * 64); r |= talloc(&t.f[2], b * 128); r |= talloc(&t.f[3], b * 128);
  r |= talloc(&t.d[0], b * 512); r |= talloc(&t.d[1], b * 256); r |= talloc(&t.d[2], b * 128);
  r |= talloc(&t.yhat, b * 7); r |= talloc(&t.gy, b * 7);
  r |= talloc(&t.ga, b * 4096); r |= talloc(&t.gb, b * 4096);
  r |= talloc(&t.gcat, b * (ENC_DIM + QF_DIM));
  r |= talloc(&t.gfeat3, b * 1024);
  r |= talloc(&t.gfeat2, b * SA2_NPOINT * 256);
  r |= talloc(&t.gfeat1, b * SA1_NPOINT * 64);
  // compacted rows of one chunk: SA1 512 x 64 slots x (4 | 64 | 64), SA2 128 x 128 x (68 | 128 | 128), SA3 128 x (260 | 512 | 512)
  r |= talloc(&t.X, k * SA2_NPOINT * 128 * 68);
  r |= talloc(&t.H1, k * SA2_NPOINT * 128 * 128);
  r |= talloc(&t.H2, k * SA2_NPOINT * 128 * 128);
  r |= talloc(&t.src, k * SA1_NPOINT * 64);
  r |= talloc(&t.slot, k * SA1_NPOINT * 64);
  r |= talloc(&t.row_grp, k * SA1_NPOINT * 64);
  r |= talloc(&t.grp_off, k * SA1_NPOINT + 1);
  r |= talloc(&t.grp_cnt, k * SA1_NPOINT);
  r |= talloc(&t.grp_mask, k * SA1_NPOINT);
  r |= talloc(&t.rows, (size_t)4);
  r |= talloc(&t.tcw, (size_t)8 * 128 * 128 + 512 * 512 + 256 * 512 + 2 * 128 * 128);   // + transposed bf16 tiles of SA3 layers 2 / 1 (data gradients)
  r |= talloc(&t.b2dup, (size_t)256 + 512);                              // + 512 zeros (bias of the data-gradient GEMMs)
  if (!r) cudaMemset(t.b2dup, 0, (256 + 512) * sizeof(float));
  r |= talloc(&t.dx, b * 4096); r |= talloc(&t.dgy, b * 4096); r |= talloc(&t.dgyT, b * 4096); r |= talloc(&t.daT, b * 4096);
  r |= talloc(&t.dw, (size_t)4096 * 2112); r |= talloc(&t.dwT, (size_t)4096 * 2112);
  r |= talloc(&t.zeros, (size_t)4096);
  if (!r) cudaMemset(t.zeros, 0, 4096 * sizeof(float));
  t.partial_floats = (size_t)20 << 20;   // >= the largest single weight tensor (fc_layer.3: 8.4 M) + bias
  r |= talloc(&t.partial, t.partial_floats);
  if (r) { free_train_ws(c); return MPN_ERR_NOMEM; }
  t.capacity = B; t.chunk = chunk; t.n_points = N;
  return MPN_OK;
}

static inline float* gw(const mpn_ctx* c, float* grads, const Linear& L) { return grads + (L.w - c->w.params); }
static inline float* gbias(const mpn_ctx* c, float* grads, const Linear& L) { return grads + (L.b - c->w.params); }

// ---------------------------------------------------------------------------------------------- dense layers on the TMA GEMM (bf16 mode)
// The FC head (1024 -> 4096 -> 2048 -> 2048) and decoder.0 (2112 -> 512) hold 97 % of the dense-stack MACs.  In the bf16 training mode
// their forward, data-gradient and weight-gradient products run on gemm_tma_kernel (bf16 operands, fp32 accumulate / outputs):
//   forward  Y [M][out] = X [M][in] W^T           A = X,      B rows = W [out][in],   K = in
//   dgrad    gX [M][in] = gY [M][out] W            A = gY,     B rows = W^T [in][out], K = out
//   wgrad    gW [out][in] = gY^T X                 A = gY^T,   B rows = X^T [in][M],   K = M  (needs M % 16 == 0)
int launch_gemm_tc_ex(mpn_ctx* c, cudaStream_t s, int epi, const __nv_bfloat16* A, int lda, int a_lo_off, const __nv_bfloat16* W, int ldw,
                      int w_lo_off, int K, const float* bias, int M, int N, void* C, int ldc, int c_lo_off, int split, uint8_t* arg_out);

__global__ void __launch_bounds__(256) narrow_rows_kernel(const float* __restrict__ src, int ld, long long rows, int cols, __nv_bfloat16* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  dst[i] = __float2bfloat16_rn(src[r * ld + (i - r * cols)]);
}
// dst [cols][rows] bf16 = transpose of src [rows][cols] fp32 (row pitch ld): 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) narrow_transpose_kernel(const float* __restrict__ src, int ld, int rows, int cols, __nv_bfloat16* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? src[(size_t)(r0 + i) * ld + c0 + tx] : 0.f;
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) dst[(size_t)(c0 + i) * rows + r0 + tx] = __float2bfloat16_rn(tile[tx][i]);
}
static int narrow_rows(mpn_ctx* c, cudaStream_t s, const float* src, int ld, long long rows, int cols, __nv_bfloat16* dst) {
  narrow_rows_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, s>>>(src, ld, rows, cols, dst);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}
static int narrow_transpose(mpn_ctx* c, cudaStream_t s, const float* src, int ld, int rows, int cols, __nv_bfloat16* dst) {
  narrow_transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, s>>>(src, ld, rows, cols, dst);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}
static bool dense_tc_ok(const Linear& L, int M) { return M % 16 == 0 && M >= 16 && L.in % 16 == 0 && L.out % 16 == 0 && L.in <= 4096 && L.out <= 4096; }

// act: 0 none, 1 LeakyReLU(0.01)
static int dense_forward_tc(mpn_ctx* c, cudaStream_t s, const Linear& L, const float* X, int ldx, int M, float* Y, int ldy, int act) {
  TrainWs& t = c->tw;
  int r;
  if ((r = narrow_rows(c, s, X, ldx, M, L.in, t.dx))) return r;
  if ((r = narrow_rows(c, s, L.w, L.in, L.out, L.in, t.dw))) return r;
  return launch_gemm_tc_ex(c, s, act ? 6 : 1, t.dx, L.in, 0, t.dw, L.in, 0, L.in, L.b, M, L.out, Y, ldy, 0, 0, nullptr);
}

static int dense_backward_tc(mpn_ctx* c, cudaStream_t s, const Linear& L, float* grads, const float* gY, int ldgy, const float* a_in, int lda,
                             int M, float* gX, int ldgx) {
  TrainWs& t = c->tw;
  int r;
  if ((r = narrow_transpose(c, s, gY, ldgy, M, L.out, t.dgyT))) return r;         // [out][M]
  if ((r = narrow_transpose(c, s, a_in, lda, M, L.in, t.daT))) return r;          // [in][M]
  if ((r = launch_gemm_tc_ex(c, s, 1, t.dgyT, M, 0, t.daT, M, 0, M, t.zeros, L.out, L.in, gw(c, grads, L), L.in, 0, 0, nullptr))) return r;
  { int rr = colsum_into(c, s, gY, ldgy, M, L.out, gbias(c, grads, L)); if (rr) return rr; }
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  if (!gX) return MPN_OK;
  if ((r = narrow_rows(c, s, gY, ldgy, M, L.out, t.dgy))) return r;               // [M][out]
  if ((r = narrow_rows(c, s, L.wt, L.out, L.in, L.out, t.dwT))) return r;         // W^T [in][out]
  return launch_gemm_tc_ex(c, s, 1, t.dgy, L.out, 0, t.dwT, L.out, 0, L.out, t.zeros, M, L.in, gX, ldgx, 0, 0, nullptr);
}

// ---------------------------------------------------------------------------------------------- forward (saves state)
int pack_w(mpn_ctx* c, cudaStream_t s, const float* src, int rows, int cols, int ld, int mode, int dst_rows, __nv_bfloat16* dst);
int tc_train_forward_sa(mpn_ctx* c, cudaStream_t s, const float* cloud, int B, int N, int32_t* fps_idx, float* xyz1, float* xyz2,
                        float* feat1_f32, float* feat2_f32, float* feat3_f32, int32_t* ball1, int32_t* ball2, uint8_t* arg1,
                        uint8_t* arg2, uint8_t* arg3);
int tc_refresh_weights(mpn_ctx* c, cudaStream_t s);

// SA1 / SA2 forward of the bf16 training mode: ball query -> gather ALL 128 neighbour rows per group -> the three layers as
// row GEMMs on tcgen05 (the same kernels, hence the same activations, as the backward's recomputation), the last one with
// the max-pool + winning-row epilogue.  Chunked like the backward; rows go through HBM in bf16.
static int train_forward_sa_tc(mpn_ctx* c, cudaStream_t s, const float* cloud, int B, int N) {
  Workspace& w = c->ws;
  TrainWs& t = c->tw;
  const Weights& W = c->w;
  __nv_bfloat16* Xb = reinterpret_cast<__nv_bfloat16*>(t.X);
  __nv_bfloat16* H1b = reinterpret_cast<__nv_bfloat16*>(t.H1);
  __nv_bfloat16* H2b = reinterpret_cast<__nv_bfloat16*>(t.H2);
  __nv_bfloat16* Wa = t.tcw + 4 * 128 * 128;   // forward tiles live in slots 4..7 (the backward re-packs 0..3 per chunk)
  __nv_bfloat16* Wb = t.tcw + 5 * 128 * 128;
  __nv_bfloat16* Wc = t.tcw + 6 * 128 * 128;   // SA2 layer 3: [256][128] = slots 6, 7
  const int chunk = t.chunk;
  int r;
  static const bool rows_fwd = getenv("MPN_TRAIN_ROWS_FWD") != nullptr;   // A/B switch: the row-GEMM forward below
  if (!rows_fwd)
    return tc_train_forward_sa(c, s, cloud, B, N, t.fps_idx, w.xyz1, w.xyz2, w.feat1, w.feat2, w.feat3, t.ball1, t.ball2, t.arg1, t.arg2,
                               t.arg3);
  // ---- SA1
  if ((r = launch_fps(c, s, cloud, B, N, 4, SA1_NPOINT, t.fps_idx, w.xyz1))) return r;
  if ((r = launch_ball_query(c, s, SA1_RADIUS, NSAMPLE, cloud, B, N, 4, w.xyz1, SA1_NPOINT, t.ball1))) return r;
  {
    const Linear* L = W.sa[0];
    if ((r = pack_w(c, s, L[1].w, 64, 64, 64, 1, 128, Wa))) return r;
    if ((r = pack_w(c, s, L[2].w, 64, 64, 64, 1, 128, Wb))) return r;
    for (int h = 0; h < 2; ++h) {
      MPN_CHECK_CUDA(cudaMemcpyAsync(t.b2dup + 64 * h, L[1].b, 64 * 4, cudaMemcpyDeviceToDevice, s));
      MPN_CHECK_CUDA(cudaMemcpyAsync(t.b2dup + 128 + 64 * h, L[2].b, 64 * 4, cudaMemcpyDeviceToDevice, s));
    }
    for (int b0 = 0; b0 < B; b0 += chunk) {
      const int bc = std::min(chunk, B - b0);
      const long long R = (long long)bc * SA1_NPOINT * NSAMPLE;
      sa1_gather_h1_kernel<<<(unsigned)((R * 8 + 255) / 256), 256, 0, s>>>(t.ball1 + (size_t)b0 * SA1_NPOINT * NSAMPLE, R, NSAMPLE, SA1_NPOINT,
                                                                           cloud + (size_t)b0 * N * 4, N, w.xyz1 + (size_t)b0 * SA1_NPOINT * 3,
                                                                           L[0].w, L[0].b, nullptr, H1b);
      c->launches++;
      MPN_CHECK_CUDA(cudaGetLastError());
      if ((r = launch_rows_gemm_tc(c, s, 0, H1b, Wa, t.b2dup, nullptr, R / 2, 128, H2b))) return r;
      if ((r = launch_rows_gemm_tc(c, s, 3, H2b, Wb, t.b2dup + 128, nullptr, R / 2, 128, nullptr, w.feat1 + (size_t)b0 * SA1_NPOINT * 64,
                                   t.arg1 + (size_t)b0 * SA1_NPOINT * 64, 1))) return r;
    }
  }
  // ---- SA2
  if ((r = launch_fps(c, s, w.xyz1, B, SA1_NPOINT, 3, SA2_NPOINT, t.fps_idx, w.xyz2))) return r;
  if ((r = launch_ball_query(c, s, SA2_RADIUS, NSAMPLE, w.xyz1, B, SA1_NPOINT, 3, w.xyz2, SA2_NPOINT, t.ball2))) return r;
  {
    const Linear* L = W.sa[1];
    if ((r = pack_w(c, s, L[0].w, 128, 67, 67, 0, 128, Wa))) return r;
    if ((r = pack_w(c, s, L[1].w, 128, 128, 128, 0, 128, Wb))) return r;
    if ((r = pack_w(c, s, L[2].w, 256, 128, 128, 0, 256, Wc))) return r;
    for (int b0 = 0; b0 < B; b0 += chunk) {
      const int bc = std::min(chunk, B - b0);
      const long long R = (long long)bc * SA2_NPOINT * NSAMPLE;
      sa2_gather_bf16_kernel<<<(unsigned)((R * 16 + 255) / 256), 256, 0, s>>>(t.ball2 + (size_t)b0 * SA2_NPOINT * NSAMPLE, R, NSAMPLE, SA2_NPOINT,
                                                                              w.xyz1 + (size_t)b0 * SA1_NPOINT * 3, SA1_NPOINT,
                                                                              w.feat1 + (size_t)b0 * SA1_NPOINT * 64,
                                                                              w.xyz2 + (size_t)b0 * SA2_NPOINT * 3, Xb);
      c->launches++;
      MPN_CHECK_CUDA(cudaGetLastError());
      if ((r = launch_rows_gemm_tc(c, s, 0, Xb, Wa, L[0].b, nullptr, R, 128, H1b))) return r;
      if ((r = launch_rows_gemm_tc(c, s, 0, H1b, Wb, L[1].b, nullptr, R, 128, H2b))) return r;
      if ((r = launch_rows_gemm_tc(c, s, 3, H2b, Wc, L[2].b, nullptr, R, 256, nullptr, w.feat2 + (size_t)b0 * SA2_NPOINT * 256,
                                   t.arg2 + (size_t)b0 * SA2_NPOINT * 256, 0))) return r;
    }
  }
  return launch_sa_simt(c, s, 2, w.xyz2, 3, w.feat2, 256, B, SA2_NPOINT, nullptr, w.feat3, nullptr, t.arg3);
}

static int train_forward(mpn_ctx* c, cudaStream_t s, const float* cloud, const float* qn, int B, int N, bool tcp) {
  Workspace& w = c->ws;
  TrainWs& t = c->tw;
  const Weights& W = c->w;
  const int CAT = ENC_DIM + QF_DIM;
  int r;
  t.sa3_h1 = t.sa3_h2 = t.sa3_a3 = nullptr;
  if (tcp) {
    if ((r = train_forward_sa_tc(c, s, cloud, B, N))) return r;
  } else {
    if ((r = launch_fps(c, s, cloud, B, N, 4, SA1_NPOINT, t.fps_idx, w.xyz1))) return r;
    if ((r = launch_sa_simt(c, s, 0, cloud, 4, cloud + 3, 4, B, N, w.xyz1, w.feat1, t.ball1, t.arg1))) return r;
    if ((r = launch_fps(c, s, w.xyz1, B, SA1_NPOINT, 3, SA2_NPOINT, t.fps_idx, w.xyz2))) return r;
    if ((r = launch_sa_simt(c, s, 1, w.xyz1, 3, w.feat1, 64, B, SA1_NPOINT, w.xyz2, w.feat2, t.ball2, t.arg2))) return r;
  }
  if (!tcp && (r = launch_sa_simt(c, s, 2, w.xyz2, 3, w.feat2, 256, B, SA2_NPOINT, nullptr, w.feat3, nullptr, t.arg3))) return r;
  const bool dtc = tcp && dense_tc_ok(W.fc[0], B);   // bf16 mode: the big dense layers on the tensor-core GEMM
  auto dense = [&](const Linear& L, const float* X, int ldx, float* Y, int ldy, int act) {
    return dtc ? dense_forward_tc(c, s, L, X, ldx, B, Y, ldy, act) : launch_linear(c, s, L, X, ldx, B, Y, ldy, act);
  };
  if ((r = dense(W.fc[0], w.feat3, 1024, t.z1, 4096, 0))) return r;
  if ((r = launch_groupnorm_lrelu_train(c, s, t.z1, B, 4096, 16, W.gn_w[0], W.gn_b[0], t.a1, t.st1))) return r;
  if ((r = dense(W.fc[1], t.a1, 4096, t.z2, 2048, 0))) return r;
  if ((r = launch_groupnorm_lrelu_train(c, s, t.z2, B, 2048, 16, W.gn_w[1], W.gn_b[1], t.a2, t.st2))) return r;
  if ((r = dense(W.fc[2], t.a2, 2048, w.cat, CAT, 0))) return r;
  if ((r = launch_linear(c, s, W.fe[0], qn, 7, B, t.f[0], 32, 1))) return r;
  if ((r = launch_linear(c, s, W.fe[1], t.f[0], 32, B, t.f[1], 64, 1))) return r;
  if ((r = launch_linear(c, s, W.fe[2], t.f[1], 64, B, t.f[2], 128, 1))) return r;
  if ((r = launch_linear(c, s, W.fe[3], t.f[2], 128, B, t.f[3], 128, 1))) return r;
  if ((r = launch_linear(c, s, W.fe[4], t.f[3], 128, B, w.cat + ENC_DIM, CAT, 0))) return r;
  if ((r = dense(W.dec[0], w.cat, CAT, t.d[0], 512, 1))) return r;
  if ((r = launch_linear(c, s, W.dec[1], t.d[0], 512, B, t.d[1], 256, 1))) return r;
  if ((r = launch_linear(c, s, W.dec[2], t.d[1], 256, B, t.d[2], 128, 1))) return r;
  return launch_linear(c, s, W.dec[3], t.d[2], 128, B, w.dq, 7, 0);
}

// ---------------------------------------------------------------------------------------------- backward

// one dense layer: gW, gb += ; gX[M][in] = (gY W) * f'(a_in)  (mask_mode 0: none)
static int dense_backward(mpn_ctx* c, cudaStream_t s, const Linear& L, float* grads, const float* gY, int ldgy, const float* a_in,
                          int lda, int M, float* gX, int ldgx, int mask_mode) {
  int r;
  if ((r = wgrad(c, s, gY, ldgy, a_in, lda, M, L.out, L.in, gw(c, grads, L), gbias(c, grads, L)))) return r;
  if (!gX) return MPN_OK;
  return launch_linear_ex(c, s, gY, ldgy, L.wt, L.out, nullptr, M, L.in, L.out, gX, ldgx, 0, mask_mode ? a_in : nullptr, lda, mask_mode);
}

template <int C2, int C3, int SLOTS, typename T>
static int launch_sa_l3(mpn_ctx* c, cudaStream_t s, const float* geff, const uint8_t* slot, const Linear& L3, T* H2, int G,
                        float* grads, const int32_t* grp_off = nullptr) {
  TrainWs& t = c->tw;
  auto k = sa_l3_bwd_kernel<C2, C3, SLOTS, T>;
  const size_t smem = (size_t)(C3 * C2 + C3) * 4 + (size_t)(sizeof(T) == 2 ? 2 : 1) * SLOTS * C2 * sizeof(T) + (size_t)(C3 + 512 + SLOTS + 1 + C3) * 4 + 16;
  MPN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = smem > 100 * 1024 ? 1 : 2;
  int grid = std::min(G, c->sm_count * per_sm);
  const size_t per = (size_t)C3 * C2 + C3;
  grid = (int)std::min<size_t>(grid, t.partial_floats / per);
  float* pW = t.partial;
  float* pb = pW + (size_t)grid * C3 * C2;
  k<<<grid, 512, smem, s>>>(geff, slot, L3.w, H2, G, pW, pb, grp_off);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  int r;
  if ((r = reduce_partials(c, s, pW, grid, (long long)C3 * C2, gw(c, grads, L3), 1))) return r;
  return reduce_partials(c, s, pb, grid, C3, gbias(c, grads, L3), 1);
}

// gW[o][k] += sum over the row CTAs of partial[(o / 128) * nx + col / 128][cta][o % 128][col % 128], col = the operand column that holds
// input k: k itself, or (rot > 0) the rotated order [features.. | x y z] of the SA2 output rows: k < 3 -> rot + k, else k - 3
__global__ void wgrad_extract2d_kernel(const float* __restrict__ P, int n, int nx, int out, int in, int rot, float* __restrict__ gW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out * in) return;
  const int o = i / in, k = i - o * in;
  const int col = rot > 0 ? (k < 3 ? rot + k : k - 3) : k;
  const float* p = P + ((size_t)((o >> 7) * nx + (col >> 7)) * n * 128 + (o & 127)) * 128 + (col & 127);
  float sum = 0.f;
  for (int sp = 0; sp < n; ++sp) sum += p[(size_t)sp * 128 * 128];
  gW[i] += sum;
}
// weight gradient of a wide layer on tcgen05: gW [out][in] += dY[R][out]^T X[R][..] (bf16 operands, fp32 accumulation and partials)
static int wgrad_tc2d_into(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* dY, int ldy, int out, const __nv_bfloat16* X, int ldx, int x_cols,
                           long long R, int in, int rot, float* gW) {
  TrainWs& t = c->tw;
  int n = 0, r;
  if ((r = launch_wgrad_tc2d(c, s, dY, ldy, out, X, ldx, x_cols, R, t.partial, t.partial_floats, &n))) return r;
  wgrad_extract2d_kernel<<<(out * in + 255) / 256, 256, 0, s>>>(t.partial, n, (x_cols + 127) / 128, out, in, rot, gW);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// module m backward over samples [b0, b0 + bc): g = d loss / d pooled output [bc][npoint][C3] (masked in place)
static int sa_backward_chunk(mpn_ctx* c, cudaStream_t s, int m, int b0, int bc, int N_in, const float* xyz, int stride,
                             const float* feats, int fstride, const float* new_xyz, const int32_t* ball, const uint8_t* arg,
                             float* g, const float* out, float* grads, float* dfeat_prev) {
  TrainWs& t = c->tw;
  const Linear* L = c->w.sa[m];
  static const int NPOINT[3] = {SA1_NPOINT, SA2_NPOINT, 1}, SLOTS[3] = {64, 128, 128}, CFEAT[3] = {1, 64, 256}, CINP[3] = {4, 68, 260};
  const int npoint = NPOINT[m], slots = SLOTS[m], C1 = L[0].out, C2 = L[1].out, C3 = L[2].out, CIN = L[0].in;
  const int G = bc * npoint;
  const long long R = (long long)G * slots;
  int r;
  // per-chunk views
  const float* xyz_c = xyz + (size_t)b0 * N_in * stride;
  const float* feats_c = feats + (size_t)b0 * N_in * fstride;
  const float* nx_c = new_xyz ? new_xyz + (size_t)b0 * npoint * 3 : nullptr;
  const int32_t* ball_c = ball ? ball + (size_t)b0 * npoint * NSAMPLE : nullptr;
  const uint8_t* arg_c = arg + (size_t)b0 * npoint * C3;
  float* g_c = g + (size_t)b0 * npoint * C3;
  const float* out_c = out + (size_t)b0 * npoint * C3;
  const bool saved = m == 2 && t.sa3_h1 && t.sa3_h2;   // bf16 mode: the forward kept the group-all level's hidden activations
  // 1. active rows -> slots
  {
    const int grid = (G + 7) / 8;
    if (m == 0) sa_prepare_kernel<64, 64><<<grid, 256, 0, s>>>(arg_c, g_c, out_c, ball_c, G, t.slot, t.src);
    else if (m == 1) sa_prepare_kernel<256, 128><<<grid, 256, 0, s>>>(arg_c, g_c, out_c, ball_c, G, t.slot, t.src);
    else if (saved) sa3_identity_prepare_kernel<<<(unsigned)(((size_t)G * 1024 + 255) / 256), 256, 0, s>>>(arg_c, g_c, out_c, G, t.slot, t.src);
    else sa_prepare_kernel<1024, 128><<<grid, 256, 0, s>>>(arg_c, g_c, out_c, nullptr, G, t.slot, t.src);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  // 2. operand rows, layers 1-2 recomputed for the active rows
  {
    const long long n = R * CINP[m];
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (m == 0) sa_gather_kernel<1, 4, true><<<grid, 256, 0, s>>>(t.src, R, slots, npoint, xyz_c, stride, N_in, feats_c, fstride, nx_c, t.X);
    else if (m == 1) sa_gather_kernel<64, 68, true><<<grid, 256, 0, s>>>(t.src, R, slots, npoint, xyz_c, stride, N_in, feats_c, fstride, nx_c, t.X);
    else sa_gather_kernel<256, 260, false><<<grid, 256, 0, s>>>(t.src, R, slots, npoint, xyz_c, stride, N_in, feats_c, fstride, nullptr, t.X);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  if (saved) {
    const long long n = R * 512;
    const unsigned grid = (unsigned)((n / 8 + 255) / 256);
    widen_bf16_kernel<<<grid, 256, 0, s>>>(t.sa3_h1 + (size_t)b0 * SA2_NPOINT * 512, n, t.H1);
    widen_bf16_kernel<<<grid, 256, 0, s>>>(t.sa3_h2 + (size_t)b0 * SA2_NPOINT * 512, n, t.H2);
    c->launches += 2;
    MPN_CHECK_CUDA(cudaGetLastError());
  } else {
    if ((r = launch_linear_ex(c, s, t.X, CINP[m], L[0].w, CIN, L[0].b, R, C1, CIN, t.H1, C1, 2))) return r;
    if ((r = launch_linear_ex(c, s, t.H1, C1, L[1].w, C1, L[1].b, R, C2, C1, t.H2, C2, 2))) return r;
  }
  // 3. layer 3 (sparse): H2 <- dZ2, gW3 / gb3
  if (m == 0) {
    if ((r = launch_sa_l3<64, 64, 64, float>(c, s, g_c, t.slot, L[2], t.H2, G, grads))) return r;
  } else if (m == 1) {
    if ((r = launch_sa_l3<128, 256, 128, float>(c, s, g_c, t.slot, L[2], t.H2, G, grads))) return r;
  } else {
    int splits = std::max(1, std::min(8, bc / 16));
    int per = (bc + splits - 1) / splits;
    splits = (bc + per - 1) / per;
    sa3_dw3_kernel<<<dim3(1024 / 8, splits), 512, 0, s>>>(g_c, t.slot, t.H2, bc, per, t.partial);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    if ((r = reduce_partials(c, s, t.partial, splits, 1024LL * 512, gw(c, grads, L[2]), 1))) return r;
    colsum_kernel<<<1024 / 32, 256, 0, s>>>(g_c, 1024, bc, 1024, gbias(c, grads, L[2]));
    c->launches++;
    sa3_dz2_kernel<<<bc, 512, 0, s>>>(g_c, t.slot, L[2].w, t.H2);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  // 4. layer 2: gW2 += dZ2^T H1 ; dZ1 = (dZ2 W2) * relu'(H1), in place over H1
  const bool wtc = saved && t.sa3_a3 && R >= 512 && getenv("MPN_TRAIN_SA3_SIMT_WGRAD") == nullptr;   // group-all level, bf16 mode: on tcgen05
  if (!wtc && (r = wgrad(c, s, t.H2, C2, t.H1, C1, R, C2, C1, gw(c, grads, L[1]), gbias(c, grads, L[1])))) return r;
  // bf16 mode, group-all level: the two data-gradient GEMMs (R x 512 x 512, R x 256 x 512) run on the TMA / tcgen05 GEMM with
  // transposed bf16 weight tiles; the saved activations' slots in the forward scratch hold the bf16 operands
  __nv_bfloat16* w2t = t.tcw + 8 * 128 * 128;
  __nv_bfloat16* w1t = w2t + 512 * 512;
  const float* zero_bias = t.b2dup + 256;
  if (saved) {
    __nv_bfloat16* dz_bf = const_cast<__nv_bfloat16*>(t.sa3_h2) + (size_t)b0 * SA2_NPOINT * 512;
    const long long n = R * 512;
    if (b0 == 0) {   // once per step: W2^T [in][out] and the feature rows of W1^T as K-major bf16
      pack_weight_kernel_f32<<<(512 * 512 + 255) / 256, 256, 0, s>>>(L[1].wt, 512 * 512, w2t);
      pack_weight_kernel_f32<<<(256 * 512 + 255) / 256, 256, 0, s>>>(L[0].wt + (size_t)3 * C1, 256 * 512, w1t);
      c->launches += 2;
    }
    narrow_bf16_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, s>>>(t.H2, n, dz_bf);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    if (wtc) {   // gW2 = dZ2^T H1 on the saved bf16 activations (16 tile pairs x row CTAs), gb2 = column sums of the fp32 dZ2
      if ((r = wgrad_tc2d_into(c, s, dz_bf, 512, 512, t.sa3_h1 + (size_t)b0 * SA2_NPOINT * 512, 512, 512, R, 512, 0, gw(c, grads, L[1])))) return r;
      if ((r = colsum_into(c, s, t.H2, 512, (int)R, 512, gbias(c, grads, L[1])))) return r;
    }
    if ((r = launch_gemm_tc(c, s, 1, dz_bf, 512, w2t, 512, zero_bias, (int)R, 512, t.H2, 512))) return r;   // dZ2 W2 -> H2 (fp32)
    relu_mask_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(t.H1, t.H2, n);                            // dZ1 over H1
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  } else {
    if ((r = launch_linear_ex(c, s, t.H2, C2, L[1].wt, C2, nullptr, R, C1, C2, t.H1, C1, 0, t.H1, C1, 2))) return r;
  }
  //    layer 1: gW1 += dZ1^T X
  if (!wtc && (r = wgrad(c, s, t.H1, C1, t.X, CINP[m], R, C1, CIN, gw(c, grads, L[0]), gbias(c, grads, L[0])))) return r;
  // 5. feature part of dX = dZ1 W1[:, 3:]  ->  previous level's feature gradient
  if (m == 2) {
    float* dst = dfeat_prev + (size_t)b0 * SA2_NPOINT * 256;   // group-all: rows are the SA2 centroids themselves
    // empty slots do not occur here only if every row is active; rows are addressed through src, so scatter (no collisions)
    if (saved) {
      __nv_bfloat16* dz1_bf = const_cast<__nv_bfloat16*>(t.sa3_h1) + (size_t)b0 * SA2_NPOINT * 512;
      const long long n = R * 512;
      narrow_bf16_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, s>>>(t.H1, n, dz1_bf);
      c->launches++;
      MPN_CHECK_CUDA(cudaGetLastError());
      if (wtc) {   // gW1 = dZ1^T X against SA2's bf16 output rows [256 features | x y z | pad] (column order rotated back on extraction)
        if ((r = wgrad_tc2d_into(c, s, dz1_bf, 512, 512, t.sa3_a3 + (size_t)b0 * SA2_NPOINT * 272, 272, 272, R, CIN, 256, gw(c, grads, L[0])))) return r;
        if ((r = colsum_into(c, s, t.H1, 512, (int)R, 512, gbias(c, grads, L[0])))) return r;
      }
      if ((r = launch_gemm_tc(c, s, 1, dz1_bf, 512, w1t, 512, zero_bias, (int)R, 256, t.H2, 256))) return r;
    } else if ((r = launch_linear_ex(c, s, t.H1, C1, L[0].wt + (size_t)3 * C1, C1, nullptr, R, CFEAT[m], C1, t.H2, CFEAT[m], 0))) return r;
    const long long n = R * CFEAT[m];
    sa_scatter_add_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t.H2, t.src, R, slots, npoint, SA2_NPOINT, CFEAT[m], dst);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  } else if (m == 1) {
    float* dst = dfeat_prev + (size_t)b0 * SA1_NPOINT * 64;
    if ((r = launch_linear_ex(c, s, t.H1, C1, L[0].wt + (size_t)3 * C1, C1, nullptr, R, CFEAT[m], C1, t.H2, CFEAT[m], 0))) return r;
    const long long n = R * CFEAT[m];
    sa_scatter_add_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t.H2, t.src, R, slots, npoint, SA1_NPOINT, CFEAT[m], dst);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  return MPN_OK;
}


// ---- tensor-core variant of sa_backward_chunk for SA1 (m = 0) and SA2 (m = 1)
int pack_w(mpn_ctx* c, cudaStream_t s, const float* src, int rows, int cols, int ld, int mode, int dst_rows, __nv_bfloat16* dst) {
  pack_train_weight_kernel<<<(dst_rows * 128 + 255) / 256, 256, 0, s>>>(src, rows, cols, ld, mode, dst_rows, dst);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

// dY^T X on tcgen05 -> gW (and the bias gradient from `bias_col` of the product, when >= 0)
static int wgrad_tc_into(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* dY, const __nv_bfloat16* X, long long R, int out, int in,
                         int paired, float* gW, int bias_col, float* gb, const int* rows_dev = nullptr, int rows_shift = 0) {
  TrainWs& t = c->tw;
  int n = 0, r;
  if ((r = launch_wgrad_tc(c, s, dY, X, R, t.partial, t.partial_floats, &n, 0, rows_dev, rows_shift))) return r;
  const int total = out * in + (bias_col >= 0 ? out : 0);
  wgrad_extract_kernel<<<(total + 255) / 256, 256, 0, s>>>(t.partial, n, out, in, paired, gW, bias_col, gb);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

static int colsum_bf16_into(mpn_ctx* c, cudaStream_t s, const __nv_bfloat16* Y, long long R, int out, int paired, float* gb,
                            const int* rows_dev = nullptr, int rows_shift = 0) {
  TrainWs& t = c->tw;
  long long ctas = std::max(1LL, std::min<long long>(4LL * c->sm_count, (R + 255) / 256));
  long long rpc = (R + ctas - 1) / ctas;
  ctas = (R + rpc - 1) / rpc;
  colsum_bf16_kernel<<<(unsigned)ctas, 256, 0, s>>>(Y, R, rpc, t.partial, rows_dev, rows_shift);
  bias_extract_kernel<<<1, 1024, 0, s>>>(t.partial, (int)ctas, out, paired, gb);
  c->launches += 2;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

static int sa_backward_chunk_tc(mpn_ctx* c, cudaStream_t s, int m, int b0, int bc, int N_in, const float* xyz, int stride,
                                const float* feats, int fstride, const float* new_xyz, const int32_t* ball, const uint8_t* arg,
                                float* g, const float* out, float* grads, float* dfeat_prev) {
  TrainWs& t = c->tw;
  const Linear* L = c->w.sa[m];
  const int npoint = m == 0 ? SA1_NPOINT : SA2_NPOINT, slots = m == 0 ? 64 : 128, C3 = L[2].out;
  const int G = bc * npoint;
  const long long R = (long long)G * slots;   // capacity: every slot of every group (the fixed-slot layout; grid sizing otherwise)
  int r;
  const float* xyz_c = xyz + (size_t)b0 * N_in * stride;
  const float* feats_c = feats + (size_t)b0 * N_in * fstride;
  const float* nx_c = new_xyz + (size_t)b0 * npoint * 3;
  const int32_t* ball_c = ball + (size_t)b0 * npoint * NSAMPLE;
  const uint8_t* arg_c = arg + (size_t)b0 * npoint * C3;
  float* g_c = g + (size_t)b0 * npoint * C3;
  const float* out_c = out + (size_t)b0 * npoint * C3;
  __nv_bfloat16* Xb = reinterpret_cast<__nv_bfloat16*>(t.X);
  __nv_bfloat16* H1b = reinterpret_cast<__nv_bfloat16*>(t.H1);
  __nv_bfloat16* H2b = reinterpret_cast<__nv_bfloat16*>(t.H2);
  __nv_bfloat16* Wa = t.tcw;                 // layer-2 weight            [128][128]
  __nv_bfloat16* Wb = t.tcw + 128 * 128;     // layer-2 weight transposed [128][128]
  __nv_bfloat16* Wc = t.tcw + 2 * 128 * 128; // SA2: layer-1 weight, K padded to 128
  __nv_bfloat16* Wd = t.tcw + 3 * 128 * 128; // SA2: feature columns of layer 1, transposed [64][128]
  __nv_bfloat16* We = t.tcw + (size_t)8 * 128 * 128 + 512 * 512 + 256 * 512;   // layer-3 weight transposed (SA2: two 128-channel halves)
  // dY3 (see sa_dy3_scatter_kernel) lives in the upper halves of the H1 / H2 allocations: they are sized for fp32 rows, the bf16 rows of
  // this path fill the lower halves even when every slot of every group is active
  __nv_bfloat16* dYa = H1b + (size_t)R * (m == 0 ? 64 : 128);
  __nv_bfloat16* dYb = H2b + (size_t)R * 128;
  // active rows compacted over the chunk's groups (default) or one fixed slot range per group (MPN_TRAIN_NOCOMPACT=1, the A/B reference)
  const bool compact = getenv("MPN_TRAIN_NOCOMPACT") == nullptr;
  const int32_t* row_grp = compact ? t.row_grp : nullptr;
  const int32_t* grp_off = compact ? t.grp_off : nullptr;
  const int* rows_exact = compact ? t.rows : nullptr;       // active rows
  const int* rows_pad = compact ? t.rows + 1 : nullptr;     // ... padded with empty rows to a multiple of 256
  {
    const int grid = (G + 7) / 8;
    if (compact) {
      if (m == 0) sa_rows_count_kernel<64><<<grid, 256, 0, s>>>(arg_c, g_c, out_c, G, t.grp_mask, t.grp_cnt);
      else sa_rows_count_kernel<256><<<grid, 256, 0, s>>>(arg_c, g_c, out_c, G, t.grp_mask, t.grp_cnt);
      sa_rows_scan_kernel<<<1, 1024, 0, s>>>(t.grp_cnt, G, t.grp_off, t.rows, t.src, t.row_grp, 256);
      if (m == 0) sa_rows_fill_kernel<64><<<grid, 256, 0, s>>>(arg_c, g_c, ball_c, G, t.grp_mask, t.grp_off, t.slot, t.src, t.row_grp);
      else sa_rows_fill_kernel<256><<<grid, 256, 0, s>>>(arg_c, g_c, ball_c, G, t.grp_mask, t.grp_off, t.slot, t.src, t.row_grp);
      c->launches += 3;
    } else {
      if (m == 0) sa_prepare_kernel<64, 64><<<grid, 256, 0, s>>>(arg_c, g_c, out_c, ball_c, G, t.slot, t.src);
      else sa_prepare_kernel<256, 128><<<grid, 256, 0, s>>>(arg_c, g_c, out_c, ball_c, G, t.slot, t.src);
      c->launches++;
    }
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  const unsigned stride_grid = (unsigned)(8 * c->sm_count);   // grid-stride kernels over a row count that lives on the device
  if (m == 1) {
    if ((r = pack_w(c, s, L[1].w, 128, 128, 128, 0, 128, Wa))) return r;
    if ((r = pack_w(c, s, L[1].wt, 128, 128, 128, 0, 128, Wb))) return r;
    if ((r = pack_w(c, s, L[0].w, 128, 67, 67, 0, 128, Wc))) return r;
    if ((r = pack_w(c, s, L[0].wt + (size_t)3 * 128, 64, 128, 128, 0, 64, Wd))) return r;
    sa2_gather_bf16_kernel<<<compact ? stride_grid : (unsigned)((R * 16 + 255) / 256), 256, 0, s>>>(t.src, R, slots, npoint, xyz_c, N_in, feats_c, nx_c,
                                                                                                    Xb, row_grp, rows_pad);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    if ((r = launch_rows_gemm_tc(c, s, 0, Xb, Wc, L[0].b, nullptr, R, 128, H1b, nullptr, nullptr, 0, rows_pad, 0))) return r;
    if ((r = launch_rows_gemm_tc(c, s, 0, H1b, Wa, L[1].b, nullptr, R, 128, H2b, nullptr, nullptr, 0, rows_pad, 0))) return r;
    if (compact) {   // layer 3 as dense products over the compacted rows (see sa_dy3_scatter_kernel)
      zero_rows_kernel<<<stride_grid, 256, 0, s>>>(reinterpret_cast<uint4*>(dYa), 16, rows_pad);
      zero_rows_kernel<<<stride_grid, 256, 0, s>>>(reinterpret_cast<uint4*>(dYb), 16, rows_pad);
      sa_dy3_scatter_kernel<256><<<stride_grid, 256, 0, s>>>(g_c, t.slot, t.grp_off, (long long)G * 256, dYa, dYb);
      c->launches += 3;
      MPN_CHECK_CUDA(cudaGetLastError());
      if ((r = pack_w(c, s, L[2].wt, 128, 128, 256, 0, 128, We))) return r;                        // W3^T [j][c], c = 0..127
      if ((r = pack_w(c, s, L[2].wt + 128, 128, 128, 256, 0, 128, We + 128 * 128))) return r;      // c = 128..255
      if ((r = wgrad_tc_into(c, s, dYa, H2b, R, 128, 128, 0, gw(c, grads, L[2]), -1, nullptr, rows_pad, 0))) return r;
      if ((r = wgrad_tc_into(c, s, dYb, H2b, R, 128, 128, 0, gw(c, grads, L[2]) + 128 * 128, -1, nullptr, rows_pad, 0))) return r;
      if ((r = colsum_bf16_into(c, s, dYa, R, 128, 0, gbias(c, grads, L[2]), rows_pad, 0))) return r;
      if ((r = colsum_bf16_into(c, s, dYb, R, 128, 0, gbias(c, grads, L[2]) + 128, rows_pad, 0))) return r;
      if ((r = launch_rows_gemm_tc(c, s, 1, dYa, We, nullptr, nullptr, R, 128, dYa, nullptr, nullptr, 0, rows_pad, 0))) return r;              // T1 in place
      if ((r = launch_rows_gemm_tc(c, s, 1, dYb, We + 128 * 128, nullptr, nullptr, R, 128, dYb, nullptr, nullptr, 0, rows_pad, 0))) return r;  // T2 in place
      add_mask_rows_kernel<<<stride_grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(dYa), reinterpret_cast<const uint4*>(dYb),
                                                       reinterpret_cast<uint4*>(H2b), rows_pad);                                            // dZ2 over H2
      c->launches++;
      MPN_CHECK_CUDA(cudaGetLastError());
    } else if ((r = launch_sa_l3<128, 256, 128, __nv_bfloat16>(c, s, g_c, t.slot, L[2], H2b, G, grads, grp_off))) return r;
    if ((r = wgrad_tc_into(c, s, H2b, H1b, R, 128, 128, 0, gw(c, grads, L[1]), -1, nullptr, rows_exact, 0))) return r;
    if ((r = colsum_bf16_into(c, s, H2b, R, 128, 0, gbias(c, grads, L[1]), rows_exact, 0))) return r;
    if ((r = launch_rows_gemm_tc(c, s, 2, H2b, Wb, nullptr, H1b, R, 128, H1b, nullptr, nullptr, 0, rows_pad, 0))) return r;   // dZ1 over H1
    if ((r = wgrad_tc_into(c, s, H1b, Xb, R, 128, 67, 0, gw(c, grads, L[0]), 127, gbias(c, grads, L[0]), rows_exact, 0))) return r;
    if ((r = launch_rows_gemm_tc(c, s, 1, H1b, Wd, nullptr, nullptr, R, 64, H2b, nullptr, nullptr, 0, rows_pad, 0))) return r;  // dX features [R][64]
    const long long n = R * 64;
    sa_scatter_add_kernel<__nv_bfloat16><<<compact ? stride_grid : (unsigned)((n + 255) / 256), 256, 0, s>>>(
        H2b, t.src, R, slots, npoint, SA1_NPOINT, 64, dfeat_prev + (size_t)b0 * SA1_NPOINT * 64, row_grp, rows_exact);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
  } else {
    if ((r = pack_w(c, s, L[1].w, 64, 64, 64, 1, 128, Wa))) return r;
    if ((r = pack_w(c, s, L[1].wt, 64, 64, 64, 1, 128, Wb))) return r;
    MPN_CHECK_CUDA(cudaMemcpyAsync(t.b2dup, L[1].b, 64 * 4, cudaMemcpyDeviceToDevice, s));
    MPN_CHECK_CUDA(cudaMemcpyAsync(t.b2dup + 64, L[1].b, 64 * 4, cudaMemcpyDeviceToDevice, s));
    float* X4 = t.X;
    sa1_gather_h1_kernel<<<compact ? stride_grid : (unsigned)((R * 8 + 255) / 256), 256, 0, s>>>(t.src, R, slots, npoint, xyz_c, N_in, nx_c, L[0].w,
                                                                                                 L[0].b, X4, H1b, row_grp, rows_pad);
    c->launches++;
    MPN_CHECK_CUDA(cudaGetLastError());
    const long long R2 = R / 2;                                                                          // rows in pairs: [R/2][128]
    if ((r = launch_rows_gemm_tc(c, s, 0, H1b, Wa, t.b2dup, nullptr, R2, 128, H2b, nullptr, nullptr, 0, rows_pad, 1))) return r;
    if (compact) {   // layer 3 as dense products over the compacted rows, two 64-wide rows per 128-wide GEMM row
      zero_rows_kernel<<<stride_grid, 256, 0, s>>>(reinterpret_cast<uint4*>(dYa), 8, rows_pad);
      sa_dy3_scatter_kernel<64><<<stride_grid, 256, 0, s>>>(g_c, t.slot, t.grp_off, (long long)G * 64, dYa, nullptr);
      c->launches += 2;
      MPN_CHECK_CUDA(cudaGetLastError());
      if ((r = pack_w(c, s, L[2].wt, 64, 64, 64, 1, 128, We))) return r;                            // diag(W3^T, W3^T)
      if ((r = wgrad_tc_into(c, s, dYa, H2b, R2, 64, 64, 1, gw(c, grads, L[2]), -1, nullptr, rows_pad, 1))) return r;
      if ((r = colsum_bf16_into(c, s, dYa, R2, 64, 1, gbias(c, grads, L[2]), rows_pad, 1))) return r;
      if ((r = launch_rows_gemm_tc(c, s, 2, dYa, We, nullptr, H2b, R2, 128, H2b, nullptr, nullptr, 0, rows_pad, 1))) return r;   // dZ2 over H2
    } else if ((r = launch_sa_l3<64, 64, 64, __nv_bfloat16>(c, s, g_c, t.slot, L[2], H2b, G, grads, grp_off))) return r;
    if ((r = wgrad_tc_into(c, s, H2b, H1b, R2, 64, 64, 1, gw(c, grads, L[1]), -1, nullptr, rows_pad, 1))) return r;
    if ((r = colsum_bf16_into(c, s, H2b, R2, 64, 1, gbias(c, grads, L[1]), rows_pad, 1))) return r;
    if ((r = launch_rows_gemm_tc(c, s, 2, H2b, Wb, nullptr, H1b, R2, 128, H1b, nullptr, nullptr, 0, rows_pad, 1))) return r;  // dZ1 over H1
    long long ctas = std::max(1LL, std::min<long long>(4LL * c->sm_count, (R + 1023) / 1024));
    long long rpc = (R + ctas - 1) / ctas;
    ctas = (R + rpc - 1) / rpc;
    sa1_wgrad1_kernel<<<(unsigned)ctas, 256, 0, s>>>(H1b, X4, R, rpc, t.partial, rows_pad);
    sa1_w1_extract_kernel<<<2, 256, 0, s>>>(t.partial, (int)ctas, gw(c, grads, L[0]), gbias(c, grads, L[0]));
    c->launches += 2;
    MPN_CHECK_CUDA(cudaGetLastError());
  }
  return MPN_OK;
}

int train_step_grads(mpn_ctx* c, cudaStream_t s, const mpn_scene& sc, int B, int N, const float* cloud, const float* q_norm,
                     const float* supervision, int n_loss_points, float margin, float w_collision, float w_bc, float* losses,
                     float* y_hat, float* grads, int precision) {
  int r;
  if ((r = ensure_train_ws(c, B, N))) return r;
  const bool tcp = precision == MPN_PREC_BF16;
  Workspace& w = c->ws;
  TrainWs& t = c->tw;
  const Weights& W = c->w;
  const int CAT = ENC_DIM + QF_DIM;
  if ((r = train_forward(c, s, cloud, q_norm, B, N, tcp))) return r;
  // y_hat = clamp(q + net(xyz, q), -1, 1) (model.py:202); losses + d(weighted loss) / d y_hat
  yhat_kernel<<<(B * 7 + 255) / 256, 256, 0, s>>>(q_norm, w.dq, B * 7, t.yhat);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  if (y_hat) MPN_CHECK_CUDA(cudaMemcpyAsync(y_hat, t.yhat, (size_t)B * 7 * 4, cudaMemcpyDeviceToDevice, s));
  if ((r = launch_bc_collision_losses(c, s, sc, B, t.yhat, supervision, n_loss_points, margin, w_collision, w_bc, losses, t.gy))) return r;
  if (!grads) return MPN_OK;
  clamp_bwd_kernel<<<(B * 7 + 255) / 256, 256, 0, s>>>(q_norm, w.dq, B * 7, t.gy);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  MPN_CHECK_CUDA(cudaMemsetAsync(grads, 0, (size_t)c->w.n_params * 4, s));
  // decoder (model.py:58-66): 2112 -> 512 -> 256 -> 128 -> 7, LeakyReLU between
  if ((r = dense_backward(c, s, W.dec[3], grads, t.gy, 7, t.d[2], 128, B, t.ga, 128, 1))) return r;
  if ((r = dense_backward(c, s, W.dec[2], grads, t.ga, 128, t.d[1], 256, B, t.gb, 256, 1))) return r;
  if ((r = dense_backward(c, s, W.dec[1], grads, t.gb, 256, t.d[0], 512, B, t.ga, 512, 1))) return r;
  const bool dtc = tcp && dense_tc_ok(W.fc[0], B);
  auto dense_bwd = [&](const Linear& L, const float* gY, int ldgy, const float* a_in, int lda, float* gX, int ldgx) {
    return dtc ? dense_backward_tc(c, s, L, grads, gY, ldgy, a_in, lda, B, gX, ldgx)
               : dense_backward(c, s, L, grads, gY, ldgy, a_in, lda, B, gX, ldgx, 0);
  };
  if ((r = dense_bwd(W.dec[0], t.ga, 512, w.cat, CAT, t.gcat, CAT))) return r;
  // feature_encoder (model.py:47-57): 7 -> 32 -> 64 -> 128 -> 128 -> 64
  if ((r = dense_backward(c, s, W.fe[4], grads, t.gcat + ENC_DIM, CAT, t.f[3], 128, B, t.ga, 128, 1))) return r;
  if ((r = dense_backward(c, s, W.fe[3], grads, t.ga, 128, t.f[2], 128, B, t.gb, 128, 1))) return r;
  if ((r = dense_backward(c, s, W.fe[2], grads, t.gb, 128, t.f[1], 64, B, t.ga, 64, 1))) return r;
  if ((r = dense_backward(c, s, W.fe[1], grads, t.ga, 64, t.f[0], 32, B, t.gb, 32, 1))) return r;
  if ((r = dense_backward(c, s, W.fe[0], grads, t.gb, 32, q_norm, 7, B, nullptr, 0, 0))) return r;
  // FC head (model.py:385-393)
  if ((r = dense_bwd(W.fc[2], t.gcat, CAT, t.a2, 2048, t.ga, 2048))) return r;
  if ((r = gn_backward(c, s, t.z2, t.st2, W.gn_w[1], W.gn_b[1], t.ga, B, 2048, grads + (W.gn_w[1] - W.params), grads + (W.gn_b[1] - W.params)))) return r;
  if ((r = dense_bwd(W.fc[1], t.ga, 2048, t.a1, 4096, t.gb, 4096))) return r;
  if ((r = gn_backward(c, s, t.z1, t.st1, W.gn_w[0], W.gn_b[0], t.gb, B, 4096, grads + (W.gn_w[0] - W.params), grads + (W.gn_b[0] - W.params)))) return r;
  if ((r = dense_bwd(W.fc[0], t.gb, 4096, w.feat3, 1024, t.gfeat3, 1024))) return r;
  // set abstraction levels, last to first, in chunks of samples
  MPN_CHECK_CUDA(cudaMemsetAsync(t.gfeat2, 0, (size_t)B * SA2_NPOINT * 256 * 4, s));
  MPN_CHECK_CUDA(cudaMemsetAsync(t.gfeat1, 0, (size_t)B * SA1_NPOINT * 64 * 4, s));
  const int chunk = t.chunk;
  for (int b0 = 0; b0 < B; b0 += chunk)
    if ((r = sa_backward_chunk(c, s, 2, b0, std::min(chunk, B - b0), SA2_NPOINT, w.xyz2, 3, w.feat2, 256, nullptr, nullptr, t.arg3,
                               t.gfeat3, w.feat3, grads, t.gfeat2))) return r;
  for (int b0 = 0; b0 < B; b0 += chunk)
    if ((r = (tcp ? sa_backward_chunk_tc : sa_backward_chunk)(c, s, 1, b0, std::min(chunk, B - b0), SA1_NPOINT, w.xyz1, 3, w.feat1, 64,
                                                              w.xyz2, t.ball2, t.arg2, t.gfeat2, w.feat2, grads, t.gfeat1))) return r;
  for (int b0 = 0; b0 < B; b0 += chunk)
    if ((r = (tcp ? sa_backward_chunk_tc : sa_backward_chunk)(c, s, 0, b0, std::min(chunk, B - b0), N, cloud, 4, cloud + 3, 4, w.xyz1,
                                                              t.ball1, t.arg1, t.gfeat1, w.feat1, grads, nullptr))) return r;
  return MPN_OK;
}

int adam_step(mpn_ctx* c, cudaStream_t s, const float* grads, float lr, float beta1, float beta2, float eps, float clip_norm,
              int step, float* grad_norm_out) {
  TrainWs& t = c->tw;
  const long long n = c->w.n_params;
  if (!t.adam_m) {
    int r = 0;
    r |= talloc(&t.adam_m, (size_t)n);
    r |= talloc(&t.adam_v, (size_t)n);
    r |= talloc(&t.norm, (size_t)NORM_CTAS + 1);
    if (r) return MPN_ERR_NOMEM;
    MPN_CHECK_CUDA(cudaMemsetAsync(t.adam_m, 0, (size_t)n * 4, s));
    MPN_CHECK_CUDA(cudaMemsetAsync(t.adam_v, 0, (size_t)n * 4, s));
  }
  sumsq_partial_kernel<<<NORM_CTAS, 256, 0, s>>>(grads, n, t.norm + 1);
  sumsq_final_kernel<<<1, NORM_CTAS, 0, s>>>(t.norm + 1, t.norm);
  c->launches += 2;
  if (grad_norm_out) MPN_CHECK_CUDA(cudaMemcpyAsync(grad_norm_out, t.norm, 4, cudaMemcpyDeviceToDevice, s));
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(c->w.params, grads, t.adam_m, t.adam_v, n, lr, beta1, beta2, eps, bc1,
                                                          sqrtf(bc2), clip_norm, t.norm);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  int r = refresh_transposes(c, s);
  if (r) return r;
  return tc_refresh_weights(c, s);   // bf16 operand copies of the tensor-core kernels (forward of the bf16 training mode, inference)
}

}  // namespace mpn
