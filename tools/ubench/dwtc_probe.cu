// Probes for the tensor-core depthwise conv kernel (round 2):
//  (1) where do the rows of an M = 64 accumulator land in TMEM (lane map), and do the operand tricks the kernel relies
//      on hold: A tile = 128B-swizzled rows read at a row offset dy, at a K offset of 48 B (the right half-window of a
//      56-wide image row), from a tile base that is NOT 1024-byte aligned (base + c * 128); B tile = 64B-swizzled 32 x 32.
//  (2) issue-to-complete cost of the MMA shapes under consideration (M 128 / 64, N 64 / 32 / 16, K 16), one CTA per SM
//      and two CTAs per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I audioset-convnext-inf_b200/csrc -I include -o tools/ubench/dwtc_probe tools/ubench/dwtc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define ACX_MBAR_TIMEOUT_NS 60000000000ull
#include "ptx.cuh"
using namespace acx;

constexpr int AROWS = 160;

// A[r][k] = r (k-independent) + 0.25 * (k % 4)  -> D[m][n] = A[m + dy][koff + n]; B = identity on a 32-wide K window
__global__ void __launch_bounds__(128) probe(float* out, int M, int dy, int koff_elems, int base_rows) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem + base_rows * 128;     // tile base deliberately off the 1024 B grid when base_rows % 8 != 0
  uint8_t* sB = smem + 24 * 1024;           // 32 rows x 64 B, SWIZZLE_64B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4096);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < AROWS * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    const uint32_t rabs = (ptx::smem_u32(sA) >> 7) + r;                      // swizzle = f(absolute address)
    *reinterpret_cast<__nv_bfloat16*>(sA + r * 128 + ((((k >> 3) ^ (rabs & 7)) << 4) + ((k & 7) << 1))) =
        __float2bfloat16((float)r + 0.25f * (k % 4));
  }
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i / 32, k = i % 32;
    *reinterpret_cast<__nv_bfloat16*>(sB + n * 64 + ((((k >> 3) ^ ((n >> 1) & 3)) << 4) + ((k & 7) << 1))) =
        __float2bfloat16(k == n ? 1.f : 0.f);
  }
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(slot, 32);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  // clear the accumulator lanes first (so untouched lanes read as 0 rather than stale data)
  {
    const uint32_t ta = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
                 "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(ta), "r"(0xbf800000u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (tid == 0) {
    const uint64_t da = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sA) + dy * 128 + koff_elems * 2);
    const uint64_t db = ptx::umma_desc_sw64_kmajor(ptx::smem_u32(sB));
    const uint32_t idesc = ptx::umma_idesc_bf16(M, 32);
    for (int kk = 0; kk < 2; ++kk) ptx::umma_bf16(tmem, da + 2 * kk, db + 2 * kk, idesc, kk ? 1u : 0u);
    ptx::umma_commit(bar);
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  uint32_t r0[32];
  const uint32_t ta = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  ptx::tmem_ld_32x32b_x32(ta, r0);
  ptx::tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[tid * 32 + j] = __uint_as_float(r0[j]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 32);
}

// timing: REPS x (14 MMAs of shape M x N x 16 over a 7-row-offset A tile), issue-to-complete, per CTA.  Fully unrolled
// with compile-time descriptor offsets: the first version computed `i % 7` per MMA in the issuing thread and measured
// that thread's instruction stream (91 cycles per MMA whatever the shape), not the tensor core.
template <int M, int N>
__global__ void __launch_bounds__(128) timing(long long* cyc, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 24 * 1024;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 16 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (40 * 1024) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(slot, 256);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1 && ptx::elect_one()) {
    const uint64_t da = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sA));
    const uint64_t db = ptx::umma_desc_sw64_kmajor(ptx::smem_u32(sB));
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(M, N);
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      const uint32_t d = tmem + (rep & 1) * 64;
#pragma unroll
      for (int dy = 0; dy < 7; ++dy)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          ptx::umma_bf16(d, da + dy * 8 + 2 * kk, db + dy * 128 + 2 * kk, idesc, (dy | kk) ? 1u : 0u);
    }
    ptx::umma_commit(bar);
    const long long t1 = clock64();
    ptx::mbar_wait(bar, 0);
    cyc[blockIdx.x] = clock64() - t0;
    cyc[512 + blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

template <int M, int N>
void run_timing(long long* cyc) {
  const int tsmem = 40 * 1024 + 64 + 1024 + 1024;
  cudaFuncSetAttribute(timing<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, tsmem);
  for (int grid : {148, 296}) {
    const int reps = 64;
    timing<M, N><<<grid, 128, tsmem>>>(cyc, reps);
    timing<M, N><<<grid, 128, tsmem>>>(cyc, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("timing: CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
    double s = 0, si = 0;
    for (int i = 0; i < grid; ++i) s += (double)cyc[i], si += (double)cyc[512 + i];
    printf("grid %3d (%d CTA/SM)  M=%3d N=%2d K=16: %.1f cycles per MMA per CTA issue-to-complete (issue alone %.1f), %d MMAs\n",
           grid, grid / 148, M, N, s / grid / (14 * reps), si / grid / (14 * reps), 14 * reps);
  }
}

int main() {
  float* out;
  cudaMallocManaged(&out, 128 * 32 * sizeof(float));
  const int smem = 24 * 1024 + 4096 + 64 + 1024 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Case { int M, dy, koff, base; } cases[] = {{128, 0, 0, 0}, {128, 5, 24, 0}, {128, 3, 24, 3}, {128, 6, 0, 13},
                                                    {64, 0, 0, 0}, {64, 4, 24, 5}};
  for (auto cs : cases) {
    probe<<<1, 128, smem>>>(out, cs.M, cs.dy, cs.koff, cs.base);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    // expected: D[m][n] = (m + dy) + 0.25 * ((koff + n) % 4) at lane(m)
    int bad_id = 0;
    for (int m = 0; m < cs.M && cs.M == 128; ++m)
      for (int n = 0; n < 32; ++n) bad_id += out[m * 32 + n] != (float)(m + cs.dy) + 0.25f * ((cs.koff + n) % 4);
    printf("M=%d dy=%d koff=%d base_rows=%d:", cs.M, cs.dy, cs.koff, cs.base);
    if (cs.M == 128) printf(" %d mismatches of 4096 (identity lane map)\n", bad_id);
    else {
      printf(" lane -> row map (col 0; -1 = untouched):\n   ");
      for (int l = 0; l < 128; ++l) {
        const float v = out[l * 32];
        printf("%d ", v == -1.f ? -1 : (int)(v - cs.dy));
        if (l % 32 == 31) printf("\n   ");
      }
      int bad = 0;   // check columns under the discovered map
      for (int l = 0; l < 128; ++l) {
        if (out[l * 32] == -1.f) continue;
        const int m = (int)(out[l * 32] - cs.dy);
        for (int n = 0; n < 32; ++n) bad += out[l * 32 + n] != (float)(m + cs.dy) + 0.25f * ((cs.koff + n) % 4);
      }
      printf("column check under that map: %d mismatches\n", bad);
    }
  }
  long long* cyc;
  cudaMallocManaged(&cyc, 1024 * sizeof(long long));
  run_timing<128, 64>(cyc);
  run_timing<128, 32>(cyc);
  run_timing<128, 16>(cyc);
  run_timing<64, 64>(cyc);
  run_timing<64, 32>(cyc);
  run_timing<64, 16>(cyc);
  return 0;
}
