// tc_common.cuh -- inline-PTX wrappers for the sm_100a tensor-core path: tcgen05.mma / alloc / ld / commit, mbarrier,
// proxy fences, and the shared-memory matrix / instruction descriptors.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded spin: a descriptor / protocol bug must not hang the GPU box (returns false on timeout)
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 22); ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// `count` arrivals at once (count >= 1)
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// ---- election
// one lane of a converged warp (the MMA issuer); the enclosing branch must be warp-uniform
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
  return p != 0;
}

// ---- fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors
// instruction descriptor, kind::f16, A/B = bf16 (K-major), D = fp32
__device__ __host__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4)                       // D format  : F32
         | (1u << 7)                     // A format  : BF16
         | (1u << 10)                    // B format  : BF16
         | (0u << 15) | (0u << 16)       // A, B K-major
         | ((uint32_t)(N >> 3) << 17)    // N / 8
         | ((uint32_t)(M >> 4) << 24);   // M / 16
}
enum { LAYOUT_NONE = 0, LAYOUT_SW128 = 2, LAYOUT_SW64 = 4, LAYOUT_SW32 = 6 };
// shared-memory matrix descriptor (sm_100: version field = 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same MMA with the two smem descriptors given as a precomputed 64-bit base plus a 16-byte-unit offset added to the
// start-address field (bits 0..13): building descriptors once and only adding per K-step keeps the single issuing
// thread's instruction chain short (a full descriptor rebuild per MMA costs ~200 cycles of dependent ALU work).
__device__ __forceinline__ void mma_bf16_ss_off(uint32_t d_tmem, uint64_t adesc, uint32_t aoff16, uint64_t bdesc, uint32_t boff16,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "add.u64 da, %1, %5;\n\t"
      "add.u64 db, %2, %6;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "l"((uint64_t)aoff16), "l"((uint64_t)boff16)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the A operand (M = 128 rows = TMEM lanes, K bf16 elements packed two per 32-bit column)
// is read from tensor memory, only B comes from shared memory; issued by ONE thread
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 columns (thread t of warp w reads lane 32*(w%4)+t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// ---- TMEM -> registers in the MMA accumulator-fragment layout: 16 lanes x 32 columns (16x256b.x4).  Thread t of the warp
// gets v[4*rep + 2*half + c] = (lane base + t/4 + 8*half, column base + 8*rep + 2*(t%4) + c): every thread holds 2 rows x
// 8 columns, so a reduction over ROWS is mostly in-thread (used by the max-pool epilogues)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// ---- registers -> TMEM: 32 lanes x 8 columns (thread t of warp w writes lane 32*(w%4)+t)
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// ---- split-bf16 ("bf16x3") operands: x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi); |x - hi - lo| <= 2^-16 |x|.
// A product a*w is then taken as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo on the bf16 tensor pipe (fp32 accumulate).
// Packed pairs: `first` in the low half-word (K element 2j), `second` in the high half-word (K element 2j + 1).
__device__ __forceinline__ void split_pack(float first, float second, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(second), "f"(first));
  const float r0 = first - __uint_as_float(hi << 16), r1 = second - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
// ReLU + split of an activation pair in 6 instructions: hi = bf16 TRUNCATION of relu(x) (cvt.rz.relu), so the remainder x - hi is >= 0
// for x >= 0 and equals x < 0 for x < 0 -- a second cvt with .relu therefore yields lo = bf16(relu(x) - hi) without a separate max.
// (|relu(x) - hi - lo| <= 2^-16 |x|: one bit less than the round-to-nearest split, two instructions fewer per pair; the split epilogues
// were half of the SA1 kernel's instructions.)
__device__ __forceinline__ void split_relu_pack(float first, float second, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(second), "f"(first));
  const float r0 = first - __uint_as_float(hi << 16), r1 = second - __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// K-major operand tile in the canonical no-swizzle ("interleaved") layout: 8x(16 B) core matrices, the core matrices of
// one 8-row group are contiguous along K (LBO = 128 B), row groups follow at SBO = (K/8)*128 B.
// byte offset of the 16-byte chunk holding K-elements [8*kc, 8*kc+8) of row r:
__device__ __forceinline__ uint32_t kmajor_chunk_off(int r, int kc, int kchunks) {
  return (uint32_t)(((r >> 3) * kchunks + kc) * 128 + (r & 7) * 16);
}

// ---- packed row tiles of the set-abstraction kernels.  pointnet2's ball query pads a neighbourhood of H < nsample hits with copies
// of its first hit, and the max-pool over a group ignores duplicates -- so only the H distinct rows of a group need to go through the
// shared MLP.  A round of (up to) 4 centroids is therefore packed into as few 128-row MMA tiles as possible at a granularity of one
// QUARTER (32 rows = one warp = 32 TMEM lanes): centroid c takes ceil(H_c / 32) consecutive quarters of one tile (first fit in centroid
// order, a centroid never straddles tiles); quarters left over at the end of a tile continue the last centroid's (padded) list, so every
// row of every tile is a valid row of the centroid that owns its quarter and the pooled result is bit-identical to the unpacked one.
struct TilePack {
  uint32_t tl;   // 2 bits per centroid: its tile
  uint32_t qs;   // 2 bits per centroid: its first quarter inside the tile
  int ntiles;
};
__device__ __forceinline__ TilePack pack_round(const int* hcnt, int nvalid, int pack) {
  TilePack p{0u, 0u, 1};
  int tile = 0, fill = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < nvalid) {
      const int q = pack ? (hcnt[c] + 31) >> 5 : 4;
      if (fill + q > 4) { ++tile; fill = 0; }
      p.tl |= (uint32_t)tile << (2 * c);
      p.qs |= (uint32_t)fill << (2 * c);
      fill += q;
    }
  p.ntiles = tile + 1;
  return p;
}
__device__ __forceinline__ int pack_tile(const TilePack& p, int c) { return (int)((p.tl >> (2 * c)) & 3u); }
__device__ __forceinline__ int pack_q0(const TilePack& p, int c) { return (int)((p.qs >> (2 * c)) & 3u); }
// one past the last quarter of centroid c inside its tile (a tile's last centroid also owns the left-over quarters)
__device__ __forceinline__ int pack_q1(const TilePack& p, int nvalid, int c) {
  return (c + 1 < nvalid && pack_tile(p, c + 1) == pack_tile(p, c)) ? pack_q0(p, c + 1) : 4;
}
// the centroid (0..3 within the round) that owns quarter q of tile `tile`
__device__ __forceinline__ int pack_owner(const TilePack& p, int nvalid, int tile, int q) {
  int mc = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (c < nvalid && pack_tile(p, c) == tile && pack_q0(p, c) <= q) mc = c;
  return mc;
}

// The same packing for rounds of EIGHT centroids at a granularity of one EIGHTH (16 rows) of a tile -- sa2x3h_tc_kernel, whose chains
// have eight warps (one ball query each) and pool the two 16-lane halves of a warp's accumulator fragment separately.
struct TilePack8 {
  uint32_t tl;   // 3 bits per centroid: its tile
  uint32_t es;   // 3 bits per centroid: its first eighth inside the tile
  int ntiles;
};
__device__ __forceinline__ TilePack8 pack_round8(const int* hcnt, int nvalid, int pack) {
  TilePack8 p{0u, 0u, 1};
  int tile = 0, fill = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < nvalid) {
      const int q = pack ? (hcnt[c] + 15) >> 4 : 8;
      if (fill + q > 8) { ++tile; fill = 0; }
      p.tl |= (uint32_t)tile << (3 * c);
      p.es |= (uint32_t)fill << (3 * c);
      fill += q;
    }
  p.ntiles = tile + 1;
  return p;
}
__device__ __forceinline__ int pack8_tile(const TilePack8& p, int c) { return (int)((p.tl >> (3 * c)) & 7u); }
__device__ __forceinline__ int pack8_e0(const TilePack8& p, int c) { return (int)((p.es >> (3 * c)) & 7u); }
__device__ __forceinline__ int pack8_e1(const TilePack8& p, int nvalid, int c) {
  return (c + 1 < nvalid && pack8_tile(p, c + 1) == pack8_tile(p, c)) ? pack8_e0(p, c + 1) : 8;
}
__device__ __forceinline__ int pack8_owner(const TilePack8& p, int nvalid, int tile, int e) {
  int mc = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < nvalid && pack8_tile(p, c) == tile && pack8_e0(p, c) <= e) mc = c;
  return mc;
}

// the context's 64-bit tile counters live behind its sticky error flag (tc_error_flag, sa_tc.cu): module 0 = SA1, 1 = SA2
__device__ __forceinline__ unsigned long long* sa_tile_counter(int* err, int module) {
  return reinterpret_cast<unsigned long long*>(err + 2) + module;
}

}  // namespace tc
}  // namespace mpn
