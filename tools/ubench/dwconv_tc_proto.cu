// Prototype of the tensor-core depthwise 7x7 convolution sketched in DESIGN.md (k4 outlook): numerics + cycle counts
// of the MMA core on one 128-row x 56-column tile, 8 channels (= all 512 TMEM columns).
//   D_c[h][w] = sum_dy sum_w' A_c[h + dy][w'] * T_{c,dy}[w][w'],   T_{c,dy}[w][w'] = tap_c[dy][w' - w + 3]  (0 <= . <= 6)
// A_c: channel-planar image rows (136 rows x 64 k, K-major, 128B swizzle, k >= 56 zero); the SAME tile is read at the 7
// row offsets dy (descriptor start + dy * 128 B).  T: 7 band matrices (64 x 64, K-major) rebuilt per channel.
// Integers small enough to be exact in bf16 / fp32, so the check against the CPU reference is exact.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I audioset-convnext-inf_b200/csrc
//      -I include -o tools/ubench/dwconv_tc_proto tools/ubench/dwconv_tc_proto.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define ACX_MBAR_TIMEOUT_NS 60000000000ull
#include "ptx.cuh"
using namespace acx;

constexpr int ROWS = 136, BM = 128, W = 56, CH = 8;
constexpr int A_TILE = ROWS * 128;            // 17408 B = 17 * 1024
constexpr int B_TILE = 64 * 128;              // 8 KB
__host__ __device__ inline int xval(int c, int r, int w) { return ((c * 7 + r * 3 + w * 5) % 9) - 4; }       // [-4, 4]
__host__ __device__ inline int tapval(int c, int dy, int dx) { return ((c * 5 + dy * 3 + dx) % 7) - 3; }      // [-3, 3]

template <int NMMA>
__global__ void __launch_bounds__(128) k(float* out, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                               // CH tiles
  uint8_t* sB = smem + CH * A_TILE;                 // 7 band tiles (one channel at a time)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 7 * B_TILE);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // planar A tiles (in the real kernel: in-SM transpose of an NHWC slab)
  for (int i = tid; i < CH * ROWS * 8; i += 128) {
    const int c = i / (ROWS * 8), r = (i / 8) % ROWS, ck = i % 8;
    __nv_bfloat16 v[8];
    for (int j = 0; j < 8; ++j) {
      const int w = ck * 8 + j;
      v[j] = __float2bfloat16(w < W ? (float)xval(c, r, w) : 0.f);
    }
    *reinterpret_cast<uint4*>(sA + c * A_TILE + r * 128 + ((ck ^ (r & 7)) << 4)) = *reinterpret_cast<uint4*>(v);
  }
  for (int i = tid; i < 7 * B_TILE / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);   // zero once
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = ptx::umma_idesc_bf16(BM, NMMA);   // NMMA < 64: timing only (narrower stages)
  long long t_bgen = 0, t_mma = 0;
  for (int c = 0; c < CH; ++c) {
    // ---- band matrices of channel c: only the 7 diagonals change (n = w_out row, k = w_in column) ------------------
    const long long t0 = clock64();
    for (int i = tid; i < 7 * 64 * 7; i += 128) {
      const int dy = i / (64 * 7), n = (i / 7) % 64, dx = i % 7;
      const int kcol = n + dx - 3;
      if (kcol >= 0 && kcol < 64 && n < W) {
        const int ck = kcol >> 3;
        *reinterpret_cast<__nv_bfloat16*>(sB + dy * B_TILE + n * 128 + ((ck ^ (n & 7)) << 4) + (kcol & 7) * 2) =
            __float2bfloat16((float)tapval(c, dy, dx));
      }
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) {
      ptx::tc_fence_after();
      const uint32_t d = tmem + c * 64;
      // descriptors differ only in the 14-bit start-address field (units of 16 B): +8 per row, +2 per K step, so the
      // 28 MMAs are issued back to back with one add each (built per MMA they cost ~65 cycles apiece -- issue-bound)
      const uint64_t da0 = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sA + c * A_TILE));
      const uint64_t db0 = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sB));
#pragma unroll
      for (int dy = 0; dy < 7; ++dy)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::umma_bf16(d, da0 + dy * 8 + 2 * kk, db0 + dy * (B_TILE / 16) + 2 * kk, idesc, (dy | kk) ? 1u : 0u);
      ptx::umma_commit(bar);
    }
    ptx::mbar_wait(bar, c & 1);          // B is single-buffered here: wait before the next channel rewrites it
    ptx::tc_fence_after();
    const long long t2 = clock64();
    t_bgen += t1 - t0;
    t_mma += t2 - t1;
  }
  // read back: lane = output row h, 64 columns per channel
  const long long t3 = clock64();
  const uint32_t ta = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < CH; ++c) {
    uint32_t r0[32], r1[32];
    ptx::tmem_ld_32x32b_x32(ta + c * 64, r0);
    ptx::tmem_ld_32x32b_x32(ta + c * 64 + 32, r1);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) {
      out[(c * 128 + tid) * 64 + j] = __uint_as_float(r0[j]);
      out[(c * 128 + tid) * 64 + 32 + j] = __uint_as_float(r1[j]);
    }
  }
  const long long t4 = clock64();
  if (tid == 0) {
    cyc[0] = t_bgen;
    cyc[1] = t_mma;
    cyc[2] = t4 - t3;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  float* out;
  long long* cyc;
  cudaMallocManaged(&out, CH * 128 * 64 * sizeof(float));
  cudaMallocManaged(&cyc, 3 * sizeof(long long));
  const int smem = CH * A_TILE + 7 * B_TILE + 64 + 1024;
  cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    k<32><<<1, 128, smem>>>(out, cyc);
    cudaDeviceSynchronize();
    if (rep) printf("timing only, M128 N32 K16: 28 MMAs issue->complete %lld cycles (%.1f per MMA)\n", cyc[1] / CH, cyc[1] / CH / 28.0);
    k<16><<<1, 128, smem>>>(out, cyc);
    cudaDeviceSynchronize();
    if (rep) printf("timing only, M128 N16 K16: 28 MMAs issue->complete %lld cycles (%.1f per MMA)\n", cyc[1] / CH, cyc[1] / CH / 28.0);
    k<64><<<1, 128, smem>>>(out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error %s\n", cudaGetErrorString(e));
      return 1;
    }
  }
  int bad = 0;
  for (int c = 0; c < CH; ++c)
    for (int h = 0; h < 128; ++h)
      for (int w = 0; w < W; ++w) {
        int ref = 0;
        for (int dy = 0; dy < 7; ++dy)
          for (int dx = 0; dx < 7; ++dx) {
            const int wi = w + dx - 3;
            if (wi >= 0 && wi < W) ref += xval(c, h + dy, wi) * tapval(c, dy, dx);
          }
        bad += out[(c * 128 + h) * 64 + w] != (float)ref;
      }
  printf("smem %d B; mismatches %d of %d\n", smem, bad, CH * 128 * W);
  printf("per channel (128 rows x 56 px): band build + sync %lld cycles, 28 MMAs issue->complete %lld cycles; TMEM read-out of 8 channels %lld cycles\n",
         cyc[0] / CH, cyc[1] / CH, cyc[2]);
  printf("=> %.2f output pixels/clk/SM for the MMA part alone (FP32-pipe kernel today: ~1.1)\n", 128.0 * W / (double)(cyc[1] / CH));
  return 0;
}
