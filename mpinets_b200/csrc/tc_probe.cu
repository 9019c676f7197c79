// tc_probe.cu -- single-CTA tcgen05 GEMM self-test: D[128][N] = A[128][K] * B[N][K]^T (bf16 in, fp32 out).
// Exercises the exact descriptor / layout conventions used by sa_tc.cu so that a convention error shows up as a
// numerical mismatch in a unit test instead of inside the fused kernels.
#include "engine.h"
#include "tc_common.cuh"

namespace mpn {
using namespace tc;

// stores a [rows][K] K-major bf16 matrix from global into smem in the layout selected by `mode`
__device__ void probe_stage(const __nv_bfloat16* __restrict__ g, int rows, int K, uint8_t* s, int mode) {
  const int KC = K / 8;
  for (int i = threadIdx.x; i < rows * KC; i += blockDim.x) {
    int r = i / KC, kc = i % KC;
    uint4 v = *reinterpret_cast<const uint4*>(g + (size_t)r * K + kc * 8);
    uint32_t off;
    if (mode == 0 || mode == 1) off = kmajor_chunk_off(r, kc, KC);
    else if (mode == 2) off = (uint32_t)((kc / 8) * rows * 128 + r * 128 + (((kc & 7) ^ (r & 7)) << 4));
    else off = (uint32_t)((kc * (rows / 8) + (r >> 3)) * 128 + (r & 7) * 16);
    *reinterpret_cast<uint4*>(s + off) = v;
  }
}

__device__ uint64_t probe_desc(uint32_t base, int rows, int K, int ks, int mode) {
  const int KC = K / 8;
  if (mode == 0) return make_smem_desc(base + ks * 256, 128, KC * 128, LAYOUT_NONE);
  if (mode == 1) return make_smem_desc(base + ks * 256, KC * 128, 128, LAYOUT_NONE);
  if (mode == 2) return make_smem_desc(base + (ks / 4) * rows * 128 + (ks % 4) * 32, 16, 1024, LAYOUT_SW128);
  return make_smem_desc(base + ks * 2 * (rows / 8) * 128, (rows / 8) * 128, 128, LAYOUT_NONE);
}

__global__ void __launch_bounds__(128) tc_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                       float* __restrict__ D, int N, int K, int mode, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * K * 2;
  probe_stage(A, 128, K, sA, mode);
  probe_stage(B, N, K, sB, mode);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    for (int ks = 0; ks < K / 16; ++ks)
      mma_bf16_ss(tmem, probe_desc(smem_u32(sA), 128, K, ks, mode), probe_desc(smem_u32(sB), N, K, ks, mode), idesc, ks > 0);
    mma_commit(&bar);
  }
  bool ok = mbar_wait(&bar, 0);
  tc_fence_after();
  if (!ok) {
    if (threadIdx.x == 0) *status = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(size_t)threadIdx.x * N + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int tc_probe(mpn_ctx* c, cudaStream_t s, const void* A, const void* B, float* D, int N, int K, int mode, int* status) {
  MPN_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && K >= 16 && K % 64 == 0 && K <= 256, "tc_probe: N in 16..256 (x16), K in 64..256 (x64)");
  MPN_REQUIRE(mode >= 0 && mode <= 3, "tc_probe: mode 0..3");
  size_t smem = (size_t)(128 + N) * K * 2 + 1024;
  MPN_CHECK_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MPN_CHECK_CUDA(cudaMemsetAsync(status, 0, sizeof(int), s));
  tc_probe_kernel<<<1, 128, smem, s>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, D, N, K, mode, status);
  c->launches++;
  MPN_CHECK_CUDA(cudaGetLastError());
  return MPN_OK;
}

}  // namespace mpn
