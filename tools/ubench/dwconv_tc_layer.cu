// Second prototype for the tensor-core depthwise 7x7 convolution (DESIGN.md, k4 outlook): a WHOLE stage-0 layer
// (B x 252 x 56 x 96, NHWC bf16 in -> conv + bias out, NHWC bf16) on the full GPU, deliberately simple (one CTA per SM,
// phases of a unit run one after the other) to measure what the straightforward version reaches before any overlap work.
//   unit = (clip, 126-row half, 8-channel group): NHWC -> channel-planar staging (one 16-byte load + eight 2-byte smem
//   stores per pixel), then per channel: band build (7 matrices, only the diagonals are rewritten), 28 tcgen05.mma
//   (M128 N64 K16, A tile read at 7 row offsets), and finally TMEM -> NHWC write-out (16 bytes = 8 channels per pixel).
// Validated against a CPU reference on sampled pixels; timed with CUDA events.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I audioset-convnext-inf_b200/csrc
//      -I include -o tools/ubench/dwconv_tc_layer tools/ubench/dwconv_tc_layer.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define ACX_MBAR_TIMEOUT_NS 60000000000ull
#include "ptx.cuh"
using namespace acx;

constexpr int H = 252, W = 56, C = 96, CG = 8, RT = 126;          // rows per tile (2 tiles per clip), channels per group
constexpr int AROWS = 136, A_TILE = AROWS * 128, B_TILE = 64 * 128;
constexpr int THREADS = 256;

__host__ __device__ inline float xval(long long i) { return (float)((int)((i * 2654435761ull >> 7) % 33) - 16) * 0.0625f; }   // bf16-exact
__host__ __device__ inline float tapval(int t, int c) { return (float)((int)(((t * 131 + c * 17) * 2654435761u >> 9) % 17) - 8) * 0.03125f; }
__host__ __device__ inline float biasval(int c) { return (float)(c % 7 - 3) * 0.25f; }

__global__ void fill_x(__nv_bfloat16* x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = __float2bfloat16(xval(i));
}

__global__ void __launch_bounds__(THREADS, 1)
    dwtc_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ taps /*[49][C]*/,
                const float* __restrict__ bias, __nv_bfloat16* __restrict__ v, int B, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + CG * A_TILE;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 7 * B_TILE);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // zero everything once: K padding (k >= 56) of the A tiles and the off-diagonal part of the bands never change
  for (int i = tid; i < (CG * A_TILE + 7 * B_TILE) / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
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
  constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 64);
  const uint32_t sA_u = ptx::smem_u32(sA), sB_u = ptx::smem_u32(sB);
  const int units = B * 2 * (C / CG);
  uint32_t phase = 0;
  long long t_stage = 0, t_band = 0, t_mma = 0, t_epi = 0;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int cg = unit % (C / CG), rt = (unit / (C / CG)) & 1, n = unit / (2 * (C / CG));
    const int h0 = rt * RT, c0 = cg * CG;
    long long t0 = clock64();
    // ---- (a) NHWC -> planar: A row r holds image row h0 - 3 + r, k = w ----------------------------------------------
    for (int p0 = tid; p0 < (RT + 6) * W; p0 += 4 * THREADS) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * THREADS;
        const int r = p / W, w = p - r * W, h = h0 - 3 + r;
        q[u] = make_uint4(0, 0, 0, 0);
        if (p < (RT + 6) * W && h >= 0 && h < H) q[u] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)n * H + h) * W + w) * C + c0));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * THREADS;
        if (p < (RT + 6) * W) {
          const int r = p / W, w = p - r * W;
          uint8_t* dst = sA + r * 128 + ((((w >> 3) ^ (r & 7)) << 4) + ((w & 7) << 1));
          const uint32_t ww[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint16_t*>(dst + (2 * j) * A_TILE) = (uint16_t)(ww[j] & 0xffffu);
            *reinterpret_cast<uint16_t*>(dst + (2 * j + 1) * A_TILE) = (uint16_t)(ww[j] >> 16);
          }
        }
      }
    }
    {
      const long long t1 = clock64();
      t_stage += t1 - t0;
      t0 = t1;
    }
    // ---- (b) per channel: bands, MMAs --------------------------------------------------------------------------------
    for (int c = 0; c < CG; ++c) {
      for (int i = tid; i < 7 * W * 7; i += THREADS) {          // (dy, n = w_out, dx): band element at k = n + dx - 3
        const int dy = i / (W * 7), rem = i - dy * (W * 7), nn = rem / 7, dx = rem - nn * 7;
        const int kcol = nn + dx - 3;
        if (kcol >= 0 && kcol < W)
          *reinterpret_cast<__nv_bfloat16*>(sB + dy * B_TILE + nn * 128 + ((((kcol >> 3) ^ (nn & 7)) << 4) + ((kcol & 7) << 1))) =
              taps[(dy * 7 + dx) * C + c0 + c];
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncthreads();
      {
        const long long t1 = clock64();
        t_band += t1 - t0;
        t0 = t1;
      }
      if (tid == 0) {
        ptx::tc_fence_after();
        const uint32_t d = tmem + c * 64;
        const uint64_t da0 = ptx::umma_desc_sw128_kmajor(sA_u + c * A_TILE);
        const uint64_t db0 = ptx::umma_desc_sw128_kmajor(sB_u);
#pragma unroll
        for (int dy = 0; dy < 7; ++dy)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::umma_bf16(d, da0 + dy * 8 + 2 * kk, db0 + dy * (B_TILE / 16) + 2 * kk, idesc, (dy | kk) ? 1u : 0u);
        ptx::umma_commit(bar);
      }
      ptx::mbar_wait(bar, phase);        // bands are single-buffered: wait before the next channel rewrites them
      phase ^= 1;
      ptx::tc_fence_after();
      {
        const long long t1 = clock64();
        t_mma += t1 - t0;
        t0 = t1;
      }
    }
    // ---- (c) TMEM -> NHWC: thread = output row (TMEM lane), warps 0-3 take w-blocks 0..3, warps 4-7 blocks 4..6 ----------
    {
      const int row = (warp & 3) * 32 + lane;
      const uint32_t ta = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
      const int wb0 = warp < 4 ? 0 : 4, wb1 = warp < 4 ? 4 : 7;
      float bs[CG];
#pragma unroll
      for (int c = 0; c < CG; ++c) bs[c] = bias[c0 + c];
      for (int wb = wb0; wb < wb1; ++wb) {
        uint32_t r[CG][8];
#pragma unroll
        for (int c = 0; c < CG; ++c)
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[c][0]), "=r"(r[c][1]), "=r"(r[c][2]), "=r"(r[c][3]), "=r"(r[c][4]), "=r"(r[c][5]),
                         "=r"(r[c][6]), "=r"(r[c][7])
                       : "r"(ta + c * 64 + wb * 8)
                       : "memory");
        ptx::tmem_ld_wait();
        if (row < RT) {
          __nv_bfloat16* dst = v + (((size_t)n * H + h0 + row) * W + wb * 8) * C + c0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint32_t o[4];
#pragma unroll
            for (int c = 0; c < CG; c += 2) {
              const __nv_bfloat162 pk = __floats2bfloat162_rn(__uint_as_float(r[c][j]) + bs[c], __uint_as_float(r[c + 1][j]) + bs[c + 1]);
              o[c / 2] = *reinterpret_cast<const uint32_t*>(&pk);
            }
            *reinterpret_cast<uint4*>(dst + (size_t)j * C) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();                    // next unit overwrites the A tiles and D
    ptx::tc_fence_after();
    t_epi += clock64() - t0;
  }
  if (blockIdx.x == 0 && tid == 0) {
    cyc[0] = t_stage;
    cyc[1] = t_band;
    cyc[2] = t_mma;
    cyc[3] = t_epi;
  }
  if (warp == 0) ptx::tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64;
  const long long n = (long long)B * H * W * C;
  __nv_bfloat16 *x, *v, *taps;
  float* bias;
  long long* cyc;
  cudaMallocManaged(&cyc, 4 * sizeof(long long));
  cudaMalloc(&x, n * 2);
  cudaMalloc(&v, n * 2);
  cudaMallocManaged(&taps, 49 * C * 2);
  cudaMallocManaged(&bias, C * 4);
  for (int t = 0; t < 49; ++t)
    for (int c = 0; c < C; ++c) taps[t * C + c] = __float2bfloat16(tapval(t, c));
  for (int c = 0; c < C; ++c) bias[c] = biasval(c);
  fill_x<<<1184, 256>>>(x, n);
  const int smem = CG * A_TILE + 7 * B_TILE + 64 + 1024;
  cudaFuncSetAttribute(dwtc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    dwtc_kernel<<<sms, THREADS, smem>>>(x, taps, bias, v, B, cyc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error: %s\n", cudaGetErrorString(e));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep) best = fminf(best, ms);
  }
  // sampled check
  std::vector<__nv_bfloat16> hv(n);
  cudaMemcpy(hv.data(), v, n * 2, cudaMemcpyDeviceToHost);
  int bad = 0, checked = 0;
  double maxerr = 0;
  for (int s = 0; s < 20000; ++s) {
    const long long i = ((long long)s * 7919 * 104729) % n;
    const int c = i % C, w = (i / C) % W, h = (i / ((long long)C * W)) % H, nn = i / ((long long)C * W * H);
    double ref = biasval(c);
    for (int dy = 0; dy < 7; ++dy)
      for (int dx = 0; dx < 7; ++dx) {
        const int hi = h + dy - 3, wi = w + dx - 3;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W)
          ref += (double)xval((((long long)nn * H + hi) * W + wi) * C + c) * (double)tapval(dy * 7 + dx, c);
      }
    const double got = __bfloat162float(hv[i]);
    const double err = fabs(got - ref);
    maxerr = fmax(maxerr, err / fmax(1.0, fabs(ref)));
    bad += err > 0.01 * fmax(1.0, fabs(ref));
    ++checked;
  }
  printf("B=%d: %d of %d sampled outputs off by more than 1%% (max rel err %.2e; bf16 output rounding = 3.9e-3)\n", B, bad, checked, maxerr);
  printf("tensor-core depthwise conv + bias, serial-phase prototype: %.1f us per layer (FP32-pipe dwconv+LN kernel: ~260 us at B=64)\n", best * 1e3);
  {
    const int units0 = (B * 2 * (C / CG) + sms - 1) / sms;
    printf("CTA 0, cycles per unit (8 channels x 126 rows): staging %lld, band builds %lld, MMAs %lld, write-out %lld\n", cyc[0] / units0,
           cyc[1] / units0, cyc[2] / units0, cyc[3] / units0);
  }
  return bad != 0;
}
