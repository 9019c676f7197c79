// spec_math.cuh -- the "MPN-fp32 spec" arithmetic on the device.
//
// Everything whose result feeds a bit-exact output (collision flags, FPS / ball-query indices, cloud
// coordinates) is written with explicit round-to-nearest intrinsics, so the instruction sequence -- and
// therefore every rounding -- is fixed regardless of nvcc's -fmad setting.  DESIGN.md ("Arithmetic contract")
// states the same sequence in words; the CPU oracle restates it independently in C.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mpn {

#define MPN_NLINK 11

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

// dot3 with the spec's chain: fma(a2,b2, fma(a1,b1, a0*b0))
__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return ffma(a2, b2, ffma(a1, b1, fmul(a0, b0)));
}

// sin/cos: Cody-Waite reduction by pi/2 (3 constants) + degree-7/8 minimax polynomials; quadrant select.
__device__ __forceinline__ void spec_sincos(float x, float& so, float& co) {
  float k = rintf(fmul(x, 0.636619772367581343f));
  float r = ffma(k, -1.57079601287841796875f, x);
  r = ffma(k, -3.1391647326017846e-07f, r);
  r = ffma(k, -5.390302529957764e-15f, r);
  float s = fmul(r, r);
  float ps = ffma(s, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = ffma(s, ps, -1.6666654611e-1f);
  float sn = ffma(fmul(r, s), ps, r);
  float pc = ffma(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = ffma(s, pc, 4.166664568298827e-2f);
  float cs = ffma(fmul(s, s), pc, ffma(s, -0.5f, 1.0f));
  int n = ((int)k) & 3;
  if (n == 0) { so = sn; co = cs; }
  else if (n == 1) { so = cs; co = -sn; }
  else if (n == 2) { so = -sn; co = -cs; }
  else { so = -cs; co = sn; }
}

// ---- 3x4 rigid transforms, row-major float[12]
__device__ __forceinline__ void m34_mul(const float* A, const float* B, float* C) {
  float T[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) T[i * 4 + j] = dot3(A[i * 4], A[i * 4 + 1], A[i * 4 + 2], B[j], B[4 + j], B[8 + j]);
    T[i * 4 + 3] = fadd(dot3(A[i * 4], A[i * 4 + 1], A[i * 4 + 2], B[3], B[7], B[11]), A[i * 4 + 3]);
  }
#pragma unroll
  for (int i = 0; i < 12; ++i) C[i] = T[i];
}

__device__ __forceinline__ void m34_apply(const float* A, float px, float py, float pz, float& ox, float& oy, float& oz) {
  ox = fadd(dot3(A[0], A[1], A[2], px, py, pz), A[3]);
  oy = fadd(dot3(A[4], A[5], A[6], px, py, pz), A[7]);
  oz = fadd(dot3(A[8], A[9], A[10], px, py, pz), A[11]);
}

// Panda chain: frames[l*12..] for link0..7, hand, leftfinger, rightfinger; eef = right_gripper (optional)
__device__ inline void spec_fk(const float* q, float prismatic, float* frames, float* eef) {
  const float ox[7] = {0.f, 0.f, 0.f, 0.0825f, -0.0825f, 0.f, 0.088f};
  const float oy[7] = {0.f, 0.f, -0.316f, 0.f, 0.384f, 0.f, 0.f};
  const float oz[7] = {0.333f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int roll[7] = {0, -1, 1, 1, -1, 1, 1};
  const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
#pragma unroll
  for (int i = 0; i < 12; ++i) frames[i] = I[i];
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    float s, c;
    spec_sincos(q[j], s, c);
    float L[12];
    L[0] = c; L[1] = -s; L[2] = 0.f; L[3] = ox[j];
    if (roll[j] == 0) {
      L[4] = s; L[5] = c; L[6] = 0.f; L[8] = 0.f; L[9] = 0.f; L[10] = 1.f;
    } else if (roll[j] > 0) {
      L[4] = 0.f; L[5] = 0.f; L[6] = -1.f; L[8] = s; L[9] = c; L[10] = 0.f;
    } else {
      L[4] = 0.f; L[5] = 0.f; L[6] = 1.f; L[8] = -s; L[9] = -c; L[10] = 0.f;
    }
    L[7] = oy[j]; L[11] = oz[j];
    m34_mul(frames + 12 * j, L, frames + 12 * (j + 1));
  }
  const float r = 0.70710678118654752440f;
  const float H[12] = {r, r, 0.f, 0.f, -r, r, 0.f, 0.f, 0.f, 0.f, 1.f, 0.107f};
  m34_mul(frames + 12 * 7, H, frames + 12 * 8);
  const float LF[12] = {1, 0, 0, 0.f, 0, 1, 0, prismatic, 0, 0, 1, 0.0584f};
  const float RF[12] = {1, 0, 0, 0.f, 0, 1, 0, -prismatic, 0, 0, 1, 0.0584f};
  m34_mul(frames + 12 * 8, LF, frames + 12 * 9);
  m34_mul(frames + 12 * 8, RF, frames + 12 * 10);
  if (eef) {
    const float G[12] = {-1, 0, 0, 0.f, 0, -1, 0, 0.f, 0, 0, 1, 0.1f};
    m34_mul(frames + 12 * 8, G, eef);
  }
}

__device__ __forceinline__ float spec_unnormalize(float qn, float lo, float hi) {
  float range = fsub(hi, lo);
  float t = fsub(qn, -1.0f);
  t = fmul(t, range);
  t = fdiv(t, 2.0f);
  return fadd(t, lo);
}
__device__ __forceinline__ float spec_normalize(float q, float lo, float hi) {
  float range = fsub(hi, lo);
  float t = fdiv(fsub(q, lo), range);
  t = fmul(t, 2.0f);
  return fadd(t, -1.0f);
}

// ---- primitives (geometry.py:151-223 / 382-454)
struct PrimFrame {  // 16 floats
  float R[9];
  float Rt[3];
  float h[3];
  float valid;  // 1.0 valid, 0.0 masked
};

__device__ __forceinline__ void quat_normalize(const float* q, float* o) {
  float n = fsqrt(ffma(q[3], q[3], ffma(q[2], q[2], ffma(q[1], q[1], fmul(q[0], q[0])))));
  o[0] = fdiv(q[0], n); o[1] = fdiv(q[1], n); o[2] = fdiv(q[2], n); o[3] = fdiv(q[3], n);
}

// rotation terms shared by the inverse frame (sign = -1) and the forward rotation (sign = +1)
__device__ __forceinline__ void quat_rows(float w, float x, float y, float z, bool quirk, float* R) {
  float xx = fmul(2.0f, fmul(x, x)), yy = fmul(2.0f, fmul(y, y)), zz = fmul(2.0f, fmul(z, z));
  float wx = fmul(fmul(2.0f, w), x), wy = fmul(fmul(2.0f, w), y), wz = fmul(fmul(2.0f, w), z);
  float xy = fmul(fmul(2.0f, x), y), xz = fmul(fmul(2.0f, x), z), yz = fmul(fmul(2.0f, y), z);
  R[0] = fsub(fsub(1.0f, yy), zz); R[1] = fsub(xy, wz); R[2] = fadd(xz, wy);
  R[3] = fadd(xy, wz); R[4] = fsub(fsub(1.0f, xx), zz); R[5] = fsub(yz, wx);
  R[6] = fsub(xz, wy); R[7] = quirk ? fsub(yz, wx) : fadd(yz, wx); R[8] = fsub(fsub(1.0f, xx), yy);
}

__device__ __forceinline__ bool is_close0(float v) { return fabsf(v) <= 1e-8f; }

__device__ __forceinline__ void make_inv_frame(const float* center, const float* quat, bool quirk, PrimFrame& f) {
  float qn[4];
  quat_normalize(quat, qn);
  quat_rows(qn[0], -qn[1], -qn[2], -qn[3], quirk, f.R);
  float nx = -center[0], ny = -center[1], nz = -center[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) f.Rt[i] = dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], nx, ny, nz);
}

__device__ __forceinline__ float sdf_cuboid(const PrimFrame& f, float px, float py, float pz) {
  float d[3], m[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float l = fadd(dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], px, py, pz), f.Rt[i]);
    d[i] = fsub(fabsf(l), f.h[i]);
    m[i] = fmaxf(d[i], 0.0f);
  }
  float outside = fsqrt(ffma(m[2], m[2], ffma(m[1], m[1], fmul(m[0], m[0]))));
  float inside = fminf(fmaxf(d[0], fmaxf(d[1], d[2])), 0.0f);
  return fadd(outside, inside);
}

__device__ __forceinline__ float sdf_cylinder(const PrimFrame& f, float px, float py, float pz) {
  float l[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) l[i] = fadd(dot3(f.R[3 * i], f.R[3 * i + 1], f.R[3 * i + 2], px, py, pz), f.Rt[i]);
  float rho = fsqrt(ffma(l[1], l[1], fmul(l[0], l[0])));
  float d0 = fsub(fabsf(rho), f.h[0]);
  float d1 = fsub(fabsf(l[2]), f.h[1]);
  float m0 = fmaxf(d0, 0.0f), m1 = fmaxf(d1, 0.0f);
  float outside = fsqrt(ffma(m1, m1, fmul(m0, m0)));
  float inside = fminf(fmaxf(d0, d1), 0.0f);
  return fadd(outside, inside);
}

// ---- counter-based RNG
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                           uint32_t* o) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t r) { return fmul((float)(r >> 8), 5.9604644775390625e-08f); }

__device__ __forceinline__ uint32_t fmix(uint32_t v, uint32_t k) {
  v = (v + k) * 0x9E3779B1u; v ^= v >> 15; v *= 0x85EBCA6Bu; v ^= v >> 13;
  return v;
}
__device__ __forceinline__ uint32_t feistel_bits(uint32_t n) {
  uint32_t bits = 2;
  while (bits < 32 && (1u << bits) < n) bits += 2;
  return bits;
}
__device__ __forceinline__ uint32_t feistel_perm(uint32_t x, uint32_t n, uint32_t half, const uint32_t* key) {
  uint32_t mask = (1u << half) - 1u;
  do {
    uint32_t L = x >> half, R = x & mask;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint32_t t = L ^ (fmix(R, key[r]) & mask);
      L = R; R = t;
    }
    x = (L << half) | R;
  } while (x >= n);
  return x;
}

enum { STREAM_OBS_PERM = 1, STREAM_OBS_SAMPLE = 2, STREAM_ROBOT_PERM = 3, STREAM_TARGET_PERM = 4, STREAM_FIXED_PERM = 5 };

// ---- surface sampling (geometrout 0.0.3.4 semantics re-specified, DESIGN.md "RNG")
__device__ __forceinline__ void rot_apply(const float* R, const float* c, float lx, float ly, float lz, float* o) {
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = fadd(dot3(R[3 * i], R[3 * i + 1], R[3 * i + 2], lx, ly, lz), c[i]);
}

__device__ __forceinline__ void sample_cuboid(const float* c, const float* d, const float* quat, const uint32_t* r, float* o) {
  float hx = fdiv(d[0], 2.0f), hy = fdiv(d[1], 2.0f), hz = fdiv(d[2], 2.0f);
  float axy = fmul(d[0], d[1]), axz = fmul(d[0], d[2]), ayz = fmul(d[1], d[2]);
  float tot = fadd(fadd(axy, axz), ayz);
  float u = fmul(u01(r[0]), tot);
  float sgn = (r[1] & 0x80000000u) ? 1.0f : -1.0f;
  float a = ffma(2.0f, u01(r[2]), -1.0f), b = ffma(2.0f, u01(r[3]), -1.0f);
  float lx, ly, lz;
  if (u < axy) { lx = fmul(a, hx); ly = fmul(b, hy); lz = fmul(sgn, hz); }
  else if (u < fadd(axy, axz)) { lx = fmul(a, hx); ly = fmul(sgn, hy); lz = fmul(b, hz); }
  else { lx = fmul(sgn, hx); ly = fmul(a, hy); lz = fmul(b, hz); }
  float qn[4], R[9];
  quat_normalize(quat, qn);
  quat_rows(qn[0], qn[1], qn[2], qn[3], false, R);
  rot_apply(R, c, lx, ly, lz, o);
}

__device__ __forceinline__ void sample_cylinder(const float* c, float rad, float h, const float* quat, const uint32_t* r,
                                                float* o) {
  float aside = fmul(rad, h), acap = fmul(rad, rad);
  float u = fmul(u01(r[0]), fadd(aside, acap));
  float th = fmul(6.28318530717958647692f, u01(r[1]));
  float s, co;
  spec_sincos(th, s, co);
  float lx, ly, lz;
  if (u < aside) {
    lx = fmul(rad, co); ly = fmul(rad, s); lz = fmul(fsub(u01(r[2]), 0.5f), h);
  } else {
    float rho = fmul(rad, fsqrt(u01(r[2])));
    lx = fmul(rho, co); ly = fmul(rho, s);
    float hh = fdiv(h, 2.0f);
    lz = (r[3] & 0x80000000u) ? hh : -hh;
  }
  float qn[4], R[9];
  quat_normalize(quat, qn);
  quat_rows(qn[0], qn[1], qn[2], qn[3], false, R);
  rot_apply(R, c, lx, ly, lz, o);
}

// squared distance with the pointnet2_ops / nvcc-fmad chain: fma(dz,dz, fma(dy,dy, dx*dx))
__device__ __forceinline__ float dist2(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = fsub(ax, bx), dy = fsub(ay, by), dz = fsub(az, bz);
  return ffma(dz, dz, ffma(dy, dy, fmul(dx, dx)));
}

}  // namespace mpn
