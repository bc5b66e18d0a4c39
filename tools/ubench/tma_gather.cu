// Round-2 probes for the tensor-core depthwise conv:
//  (1) how fast can TMA gather / scatter 16-byte pieces (8 of C channels of an NHWC pixel)?  The LSU path measured
//      ~5 useful B/clk/SM for this pattern (one 32-byte sector request per pixel, half of it wasted).
//      box {8 ch, 56 w, 69 rows} load (62 KB) and box {8 ch, 28 w, 63 rows} store, one CTA or two CTAs per SM.
//  (2) mbarrier hand-off latency between two warps: hinted try_wait (10 ms suspend hint, what ptx::mbar_wait uses)
//      vs an unhinted try_wait spin.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I audioset-convnext-inf_b200/csrc -I include -o tools/ubench/tma_gather tools/ubench/tma_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define ACX_MBAR_TIMEOUT_NS 60000000000ull
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
namespace acx { void set_error(const char*, ...) {} const char* last_error() { return ""; } }
using namespace acx;

constexpr int H = 252, W = 56, C = 96, B = 64;

__global__ void __launch_bounds__(128) gather_kernel(const __grid_constant__ CUtensorMap tmIn,
                                                     const __grid_constant__ CUtensorMap tmOut, int units, int mode,
                                                     long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t ph = 0;
    const long long t0 = clock64();
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int g = u % 12, t = (u / 12) % 4, n = u / 48;
      if (mode == 0) {          // load: 69 rows x 56 px x 16 B
        ptx::mbar_arrive_expect_tx(bar, 69 * 56 * 16);
        ptx::tma_load_4d(smem, &tmIn, bar, g * 8, 0, t * 63 - 3, n);
        ptx::mbar_wait(bar, ph);
        ph ^= 1;
      } else {                  // store: 63 rows x 28 px x 16 B, two halves
        for (int half = 0; half < 2; ++half) {
          asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmOut)),
                       "r"(ptx::smem_u32(smem)), "r"(g * 8), "r"(half * 28), "r"(t * 63), "r"(n)
                       : "memory");
          ptx::tma_store_commit();
          ptx::tma_store_wait_read<0>();
        }
      }
    }
    if (mode == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    cyc[blockIdx.x] = clock64() - t0;
  }
}

// ping-pong between warp 0 and warp 1 through two mbarriers; mode 0 = ptx::mbar_wait (hinted), 1 = unhinted spin
__global__ void __launch_bounds__(64) pingpong(int iters, int mode, long long* cyc) {
  __shared__ uint64_t bars[2];
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto wait = [&](uint64_t* b, uint32_t par) {
    if (mode == 0) ptx::mbar_wait(b, par);
    else {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(ptx::smem_u32(b)), "r"(par) : "memory");
    }
  };
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (warp == 0) {
      if (lane == 0) ptx::mbar_arrive(&bars[0]);
      wait(&bars[1], i & 1);
    } else {
      wait(&bars[0], i & 1);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bars[1]);
    }
  }
  if (threadIdx.x == 0) cyc[0] = clock64() - t0;
}

int main() {
  const size_t n = (size_t)B * H * W * C;
  __nv_bfloat16 *x, *v;
  cudaMalloc(&x, n * 2);
  cudaMalloc(&v, n * 2);
  cudaMemset(x, 0, n * 2);
  long long* cyc;
  cudaMallocManaged(&cyc, 1024 * sizeof(long long));
  CUtensorMap tmIn, tmOut;
  {
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {8, 56, 69, 1};
    if (make_tmap_bf16(&tmIn, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE) != ACX_OK) { printf("tmap in failed\n"); return 1; }
    cuuint32_t box2[4] = {8, 28, 63, 1};
    if (make_tmap_bf16(&tmOut, v, 4, dims, strides, box2, CU_TENSOR_MAP_SWIZZLE_NONE) != ACX_OK) { printf("tmap out failed\n"); return 1; }
  }
  const int smem = 64 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int units = B * 4 * 12;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {148, 296, 444}) {
      gather_kernel<<<grid, 128, smem>>>(tmIn, tmOut, units, mode, cyc);
      cudaEventRecord(e0);
      gather_kernel<<<grid, 128, smem>>>(tmIn, tmOut, units, mode, cyc);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = mode == 0 ? (double)units * 69 * 56 * 16 : (double)units * 63 * 56 * 16;
      printf("TMA %s of 16-byte NHWC pieces, %d CTAs (%d per SM): %.1f us per stage-0 layer, %.0f GB/s useful, %.1f useful B/clk/SM, "
             "%lld clk per unit (CTA 0)\n", mode == 0 ? "gather (load)" : "scatter (store)", grid, grid / 148, ms * 1e3,
             bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 148 / 1.9e9, cyc[0] / ((units + grid - 1) / grid));
    }
  for (int mode = 0; mode < 2; ++mode) {
    pingpong<<<1, 64>>>(2000, mode, cyc);
    pingpong<<<1, 64>>>(2000, mode, cyc);
    cudaDeviceSynchronize();
    printf("mbarrier ping-pong (%s): %.0f cycles per round trip (2 hand-offs)\n", mode == 0 ? "hinted try_wait, ptx::mbar_wait" : "unhinted try_wait spin",
           (double)cyc[0] / 2000);
  }
  return 0;
}
