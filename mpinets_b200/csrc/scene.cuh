// scene.cuh -- per-problem primitive lists (TorchCuboids / TorchCylinders, geometry.py:126-568) staged as inverse frames in
// shared memory, shared by the SDF, sweep, evaluation and loss kernels.
#pragma once
#include "engine.h"
#include "spec_math.cuh"

namespace mpn {

// per-problem primitive list -> inverse frames in shared memory (<= M1+M2 <= 128 entries of 64 B)
constexpr int MAX_PRIMS = 128;

__device__ __forceinline__ PrimFrame prim_frame_of(const mpn_scene& sc, int b, int M1, int M2, int m, bool quirk) {
  PrimFrame f;
  if (m < M1) {
    const float* d = sc.cuboid_dims + ((size_t)b * M1 + m) * 3;
    float d0 = d[0], d1 = d[1], d2 = d[2];
    f.valid = (is_close0(d0) || is_close0(d1) || is_close0(d2)) ? 0.f : 1.f;
    make_inv_frame(sc.cuboid_centers + ((size_t)b * M1 + m) * 3, sc.cuboid_quats + ((size_t)b * M1 + m) * 4, quirk, f);
    f.h[0] = fdiv(d0, 2.0f); f.h[1] = fdiv(d1, 2.0f); f.h[2] = fdiv(d2, 2.0f);
  } else {
    int k = m - M1;
    float r = sc.cylinder_radii[(size_t)b * M2 + k], h = sc.cylinder_heights[(size_t)b * M2 + k];
    f.valid = (is_close0(r) || is_close0(h)) ? 0.f : 1.f;
    make_inv_frame(sc.cylinder_centers + ((size_t)b * M2 + k) * 3, sc.cylinder_quats + ((size_t)b * M2 + k) * 4, quirk, f);
    f.h[0] = r; f.h[1] = fdiv(h, 2.0f); f.h[2] = 0.f;
  }
  return f;
}

__device__ __forceinline__ void stage_scene(const mpn_scene& sc, int b, int M1, int M2, bool quirk, PrimFrame* fr) {
  for (int m = threadIdx.x; m < M1 + M2; m += blockDim.x) fr[m] = prim_frame_of(sc, b, M1, M2, m, quirk);
}

// Same, but only the valid (non zero-volume) primitives are kept, packed to the front in their original order: cuboids in
// fr[0, counts[0]), cylinders in fr[counts[0], counts[0] + counts[1]).  Padding rows (data_loader.py:198-215) make up most
// of the 2 x 40 slots, so consumers loop over ~10-20 entries instead of 80.  Needs blockDim.x >= 32 and a multiple of 32;
// `counts` and `wcnt` are shared-memory scratch ([2] and [blockDim.x / 32]); ends with a __syncthreads().
__device__ __forceinline__ void stage_scene_compact(const mpn_scene& sc, int b, int M1, int M2, bool quirk, PrimFrame* fr, int* counts,
                                                    int* wcnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int base = 0, ncub = 0;
  for (int m0 = 0; m0 < M1 + M2; m0 += blockDim.x) {
    const int m = m0 + threadIdx.x;
    PrimFrame f;
    bool keep = false;
    if (m < M1 + M2) { f = prim_frame_of(sc, b, M1, M2, m, quirk); keep = f.valid != 0.f; }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const unsigned balc = __ballot_sync(0xffffffffu, keep && m < M1);
    if (lane == 0) wcnt[warp] = __popc(bal) | (__popc(balc) << 16);
    __syncthreads();
    int off = base, tot = 0, totc = 0;
    for (int w = 0; w < nw; ++w) {
      const int v = wcnt[w];
      if (w < warp) off += v & 0xffff;
      tot += v & 0xffff; totc += v >> 16;
    }
    if (keep) fr[off + __popc(bal & ((1u << lane) - 1u))] = f;
    base += tot; ncub += totc;
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[0] = ncub; counts[1] = base - ncub; }
  __syncthreads();
}

// min over compacted ranges (no validity test)
__device__ __forceinline__ float scene_sdf_packed(const PrimFrame* fr, int c0, int c1, int y0, int y1, float px, float py, float pz) {
  float best = __int_as_float(0x7f800000);
  for (int m = c0; m < c1; ++m) best = fminf(best, sdf_cuboid(fr[m], px, py, pz));
  for (int m = y0; m < y1; ++m) best = fminf(best, sdf_cylinder(fr[m], px, py, pz));
  return best;
}

__device__ __forceinline__ float scene_sdf(const PrimFrame* fr, int c0, int c1, int y0, int y1, float px, float py, float pz) {
  float best = __int_as_float(0x7f800000);
  for (int m = c0; m < c1; ++m)
    if (fr[m].valid != 0.f) best = fminf(best, sdf_cuboid(fr[m], px, py, pz));
  for (int m = y0; m < y1; ++m)
    if (fr[m].valid != 0.f) best = fminf(best, sdf_cylinder(fr[m], px, py, pz));
  return best;
}


}  // namespace mpn
