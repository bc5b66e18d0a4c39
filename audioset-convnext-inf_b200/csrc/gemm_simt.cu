// fp32 SIMT GEMM of the fp32-accurate path:  out[M,N] = epi(A[M,K] . W[N,K]^T), fp32 FMA.
// Also the on-device cross-check for the tcgen05 kernels.  A rows may overlap (STFT framing):
// row m starts at A + (m / rows_per_batch) * batch_stride + (m % rows_per_batch) * row_stride.
#include "common.cuh"

namespace acx {

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <int EPI>
__global__ void __launch_bounds__(SG_THREADS)
    gemm_f32_kernel(const float* __restrict__ A, long long batch_stride, int rows_per_batch, int row_stride,
                    const float* __restrict__ Wt, float* __restrict__ out, int ldo, int M, int N, int K,
                    const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ resid) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tx = tid % 16, ty = tid / 16;  // 16 x 16 threads, each 8 (m) x 4 (n)

  // loader mapping: A tile 128 rows x 16 k = 512 float4 -> 2 per thread; B tile 64 x 16 = 256 float4 -> 1
  const float* a_ptr[2];
  bool a_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int f = tid + i * SG_THREADS;
    const int r = f / 4;
    const int m = m0 + r;
    a_ok[i] = m < M;
    const int mm = a_ok[i] ? m : 0;
    a_ptr[i] = A + (long long)(mm / rows_per_batch) * batch_stride + (long long)(mm % rows_per_batch) * row_stride +
               (f % 4) * 4;
  }
  const int br = tid / 4;
  const bool b_ok = (n0 + br) < N;
  const float* b_ptr = Wt + (size_t)(b_ok ? n0 + br : 0) * K + (tid % 4) * 4;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * SG_THREADS;
      const int r = f / 4, kc = (f % 4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i]) v = *reinterpret_cast<const float4*>(a_ptr[i] + k0);
      As[kc + 0][r] = v.x;
      As[kc + 1][r] = v.y;
      As[kc + 2][r] = v.z;
      As[kc + 3][r] = v.w;
    }
    {
      const int kc = (tid % 4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b_ok) v = *reinterpret_cast<const float4*>(b_ptr + k0);
      Bs[kc + 0][br] = v.x;
      Bs[kc + 1][br] = v.y;
      Bs[kc + 2][br] = v.z;
      Bs[kc + 3][br] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[8], b[4];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (EPI == ACX_EPI_BIAS_GELU) v = gelu_erf(v);
      if (EPI == ACX_EPI_BIAS_SCALE_RESID) v = resid[(size_t)m * ldo + n] + gamma[n] * v;
      out[(size_t)m * ldo + n] = v;
    }
  }
}

}  // namespace acx

using namespace acx;

extern "C" int acx_gemm_f32(const float* A, long long batch_stride, int rows_per_batch, int row_stride, const float* W,
                            float* out, int ldo, int M, int N, int K, int epilogue, const float* bias,
                            const float* gamma, const float* resid, void* stream) {
  ACX_CHECK(A && W && out, ACX_ERR_ARG, "gemm_f32: null pointer");
  ACX_CHECK(M > 0 && N > 0 && K > 0 && K % 16 == 0, ACX_ERR_ARG, "gemm_f32: K=%d must be a positive multiple of 16", K);
  ACX_CHECK(rows_per_batch > 0 && row_stride % 4 == 0 && batch_stride % 4 == 0, ACX_ERR_ARG,
            "gemm_f32: row/batch strides must be multiples of 4 floats");
  ACX_CHECK(ldo >= N, ACX_ERR_ARG, "gemm_f32: ldo < N");
  if (epilogue == ACX_EPI_BIAS_SCALE_RESID)
    ACX_CHECK(gamma && resid && bias, ACX_ERR_ARG, "gemm_f32: scale+residual epilogue needs bias, gamma and resid");
  dim3 grid(ceil_div(N, SG_BN), ceil_div(M, SG_BM));
  ACX_CHECK(grid.y <= 65535, ACX_ERR_ARG, "gemm_f32: M=%d too large for one launch", M);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (epilogue) {
    case ACX_EPI_BIAS:
      gemm_f32_kernel<ACX_EPI_BIAS><<<grid, SG_THREADS, 0, st>>>(A, batch_stride, rows_per_batch, row_stride, W, out,
                                                                 ldo, M, N, K, bias, gamma, resid);
      break;
    case ACX_EPI_BIAS_GELU:
      gemm_f32_kernel<ACX_EPI_BIAS_GELU><<<grid, SG_THREADS, 0, st>>>(A, batch_stride, rows_per_batch, row_stride, W,
                                                                      out, ldo, M, N, K, bias, gamma, resid);
      break;
    case ACX_EPI_BIAS_SCALE_RESID:
      gemm_f32_kernel<ACX_EPI_BIAS_SCALE_RESID><<<grid, SG_THREADS, 0, st>>>(A, batch_stride, rows_per_batch,
                                                                             row_stride, W, out, ldo, M, N, K, bias,
                                                                             gamma, resid);
      break;
    default:
      set_error("gemm_f32: unknown epilogue %d", epilogue);
      return ACX_ERR_ARG;
  }
  ACX_CUDA(cudaGetLastError());
  return ACX_OK;
}
