// Micro-benchmark: sustained throughput (lanes / clk / SM) of MUFU.TANH, MUFU.EX2, FFMA, FFMA2 on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes tools/ubench/pipes.cu && /tmp/pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cycles) {
  float v[8];
  float2 w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = 0.001f * (threadIdx.x + i);
    w[i] = make_float2(v[i], v[i] + 1.f);
  }
  const float2 c = make_float2(0.999f, 1.001f), d = make_float2(1e-3f, -1e-3f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(c.x), "f"(d.x));
      if (OP == 3) w[i] = __ffma2_rn(w[i], c, d);
      if (OP == 5 || OP == 6) {   // mixed: 1 MUFU.TANH + 3 (OP 5) or 1 (OP 6) FFMA2 per slot, independent chains
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
        w[i] = __ffma2_rn(w[i], c, d);
        if (OP == 5) {
          w[(i + 3) & 7] = __ffma2_rn(w[(i + 3) & 7], c, d);
          w[(i + 5) & 7] = __ffma2_rn(w[(i + 5) & 7], c, d);
        }
      }
      if (OP == 4) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + w[i].x + w[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// legacy tensor path (HMMA): mma.sync m16n8k16 / m16n8k8, bf16 inputs, fp32 accumulate; 8 independent accumulator
// fragments per warp.  Reported as warp-level mma instructions / clk / SM.
template <int K>
__global__ void __launch_bounds__(512) kmma(float* out, int iters, long long* cycles) {
  float d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  unsigned a0 = 0x3f803f80u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3c003c00u, b1 = 0x3c003c01u;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (K == 16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                     : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int K>
void run_mma(const char* name) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMallocManaged(&cyc, sizeof(long long));
  const int iters = 4096;
  kmma<K><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  kmma<K><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  const double warp_instr = 16.0 * iters * 8;       // per SM
  printf("%-22s %8.3f warp-mma/clk/SM  (%7.1f dense MAC/clk/SM)\n", name, warp_instr / *cyc,
         warp_instr * 16 * 8 * K / *cyc);
}

template <int OP>
void run(const char* name, int per_instr_elems) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMallocManaged(&cyc, sizeof(long long));
  const int iters = 4096;
  k<OP><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<OP><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  const double instr_lanes = 512.0 * iters * 8;     // per SM
  printf("%-22s %8.2f lane-instr/clk/SM  (%6.2f elements/clk/SM)\n", name, instr_lanes / *cyc,
         instr_lanes * per_instr_elems / *cyc);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<4>("tanh.approx.f16x2", 2);
  run<1>("ex2.approx.ftz.f32", 1);
  run<2>("fma.rn.f32", 1);
  run<3>("fma.rn.f32x2 (FFMA2)", 2);
  run<5>("mix 1 tanh + 3 FFMA2", 1);   // lane-instr counts the tanh only: 16 = XU-bound, less = pipes do not overlap
  run<6>("mix 1 tanh + 1 FFMA2", 1);
  run_mma<16>("mma.sync m16n8k16 bf16");
  run_mma<8>("mma.sync m16n8k8 bf16");
  return 0;
}
