// Micro-benchmark (round 2): sustained throughput of the packed / mixed half-precision FMA forms on this GPU,
// to decide which arithmetic unit the depthwise 7x7 kernel and the GELU epilogue should use.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes_half tools/ubench/pipes_half.cu && /tmp/pipes_half
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512) k(float* out, int iters, long long* cycles) {
  unsigned h[8];
  float v[8];
  float2 w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = 0x3c003c00u + threadIdx.x + i;
    v[i] = 0.001f * (threadIdx.x + i);
    w[i] = make_float2(v[i], v[i] + 1.f);
  }
  const unsigned ch = 0x3bff3bffu, dh = 0x14001400u;   // f16x2 constants (~0.9995, small)
  const float2 c = make_float2(0.999f, 1.001f), d = make_float2(1e-3f, -1e-3f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(ch), "r"(dh));
      if (OP == 1) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(ch), "r"(dh));
      if (OP == 2)   // FHFMA.BF16: bf16 x bf16 + f32 -> f32
        asm volatile("{.reg .b16 l, u;\n mov.b32 {l,u}, %1;\n fma.rn.f32.bf16 %0, l, u, %0;}" : "+f"(v[i]) : "r"(ch));
      if (OP == 3)   // FHFMA (f16)
        asm volatile("{.reg .b16 l, u;\n mov.b32 {l,u}, %1;\n fma.rn.f32.f16 %0, l, u, %0;}" : "+f"(v[i]) : "r"(ch));
      if (OP == 4)   // FHADD: f16 + f32 -> f32
        asm volatile("{.reg .b16 l, u;\n mov.b32 {l,u}, %1;\n add.rn.f32.f16 %0, u, %0;}" : "+f"(v[i]) : "r"(ch));
      if (OP == 5) {  // mix 1 HFMA2 : 1 FFMA2 (do they share a pipe?)
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(ch), "r"(dh));
        w[i] = __ffma2_rn(w[i], c, d);
      }
      if (OP == 6) {  // mix 1 HFMA2 : 1 FFMA
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(ch), "r"(dh));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(c.x), "f"(d.x));
      }
      if (OP == 7) {  // 7 HFMA2 + 2 FHADD (a 7-tap row with its fp32 flush)
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[i]) : "r"(ch), "r"(dh));
        if (i == 7) {
          asm volatile("{.reg .b16 l, u;\n mov.b32 {l,u}, %2;\n add.rn.f32.f16 %0, l, %0;\n add.rn.f32.f16 %1, u, %1;}"
                       : "+f"(v[0]), "+f"(v[1]) : "r"(h[0]));
        }
      }
      if (OP == 8) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 9) {  // mix 1 tanh.f16x2 : 3 HFMA2
        asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[(i + 1) & 7]) : "r"(ch), "r"(dh));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[(i + 3) & 7]) : "r"(ch), "r"(dh));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h[(i + 5) & 7]) : "r"(ch), "r"(dh));
      }
      if (OP == 10) {  // cvt.rn.f16x2.f32 (pack two fp32 into f16x2)
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[i]), "f"(v[(i + 1) & 7]));
        v[i] = __uint_as_float(h[i]);
      }
      if (OP == 11) {  // cvt.rn.bf16x2.f32
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[i]), "f"(v[(i + 1) & 7]));
        v[i] = __uint_as_float(h[i]);
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i] + w[i].x + w[i].y + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, double instr_per_slot) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMallocManaged(&cyc, sizeof(long long));
  const int iters = 4096;
  k<OP><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  k<OP><<<148, 512>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  const double slots = 512.0 * iters * 8;     // per SM
  printf("%-34s %8.2f slot-lanes/clk/SM  (%7.2f instr-lanes/clk/SM)\n", name, slots / *cyc,
         slots * instr_per_slot / *cyc);
}

int main() {
  run<0>("fma.rn.f16x2 (HFMA2)", 1);
  run<1>("fma.rn.bf16x2 (HFMA2.BF16)", 1);
  run<2>("fma.rn.f32.bf16 (FHFMA.BF16)", 1);
  run<3>("fma.rn.f32.f16 (FHFMA)", 1);
  run<4>("add.rn.f32.f16 (FHADD)", 1);
  run<5>("mix 1 HFMA2 + 1 FFMA2", 2);
  run<6>("mix 1 HFMA2 + 1 FFMA", 2);
  run<7>("7 HFMA2 + 2 FHADD", 1.25);
  run<8>("tanh.approx.f16x2", 1);
  run<9>("mix 1 tanh.f16x2 + 3 HFMA2", 4);
  run<10>("cvt.rn.f16x2.f32", 1);
  run<11>("cvt.rn.bf16x2.f32", 1);
  return 0;
}
