// Experiment for a tensor-core depthwise convolution (DESIGN.md, k4 outlook): can a tcgen05.mma A-operand descriptor
// start at a row that is NOT a multiple of 8 inside a 128B-swizzled K-major tile?  A Toeplitz formulation of the 7x7
// depthwise conv needs the SAME smem tile read at 7 row offsets (dy = 0..6).
// D[m][n] = sum_k A[m + dy][k] * I[n][k] must equal A[m + dy][n].  Variants: base_offset field (bits 49..51) = 0 or dy & 7.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I audioset-convnext-inf_b200/csrc -I include -o tools/ubench/umma_rowshift tools/ubench/umma_rowshift.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define ACX_MBAR_TIMEOUT_NS 60000000000ull
#include "ptx.cuh"
using namespace acx;

constexpr int ROWS = 144, BM = 128, BN = 64, BK = 64;

__global__ void __launch_bounds__(128) k(float* out, int dy, int use_base_offset) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 144 rows x 128 B
  uint8_t* sB = smem + ROWS * 128;          // 64 rows x 128 B   (ROWS * 128 = 18432 = 18 * 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + BN * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < ROWS * 8; i += 128) {          // 16-byte chunks
    const int r = i / 8, c = i % 8;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16((float)((r * 3 + c * 8 + j) % 17));
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  for (int i = tid; i < BN * 8; i += 128) {
    const int r = i / 8, c = i % 8;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16((c * 8 + j) == r ? 1.f : 0.f);
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    uint64_t da = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sA) + dy * 128);
    if (use_base_offset) da |= static_cast<uint64_t>(dy & 7) << 49;
    const uint64_t db = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sB));
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, BN);
    for (int kk = 0; kk < BK / 16; ++kk) ptx::umma_bf16(tmem, da + 2 * kk, db + 2 * kk, idesc, kk ? 1u : 0u);
    ptx::umma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  uint32_t r0[32], r1[32];
  const uint32_t ta = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  ptx::tmem_ld_32x32b_x32(ta, r0);
  ptx::tmem_ld_32x32b_x32(ta + 32, r1);
  ptx::tmem_ld_wait();
  for (int j = 0; j < 32; ++j) {
    out[tid * 64 + j] = __uint_as_float(r0[j]);
    out[tid * 64 + 32 + j] = __uint_as_float(r1[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 64);
}

int main() {
  float* out;
  cudaMallocManaged(&out, 128 * 64 * sizeof(float));
  const int smem = ROWS * 128 + BN * 128 + 64 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int dy = 0; dy < 9; ++dy) {
      k<<<1, 128, smem>>>(out, dy, ubo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("dy=%d base_offset=%d: CUDA error %s\n", dy, ubo, cudaGetErrorString(e));
        return 1;
      }
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) bad += out[m * 64 + n] != (float)(((m + dy) * 3 + n) % 17);
      printf("dy=%d base_offset_field=%s: %d mismatches of 8192\n", dy, ubo ? "dy&7" : "0", bad);
    }
  return 0;
}
