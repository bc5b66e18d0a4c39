// Fused front end (tensor-core path):  padded waveform -> framing x windowed-DFT -> |.|^2 -> x mel -> 10 log10 -> bn0
// in ONE kernel; the complex spectrogram and the power spectrum never leave the SM.
// Replaces torchlibrosa Spectrogram + LogmelFilterBank + bn0 (reference convnext.py:298-306).
//
// One CTA = 128 consecutive frames of one clip.  Frequency bins are processed in chunks of 64:
//   GEMM1(c): D1[128 x 128] = frames[128 x 1024] . dft_chunk_c[128 x 1024]^T      (cols 0..63 real, 64..127 imag)
//             split precision x3: Ahi.Bhi + Ahi.Blo + Alo.Bhi, fp32 accumulation in TMEM (single-pass bf16 is 35 dB off
//             on band-limited audio).  The pairs are FP16 (11-bit mantissas, operands pre-scaled by
//             2^8 so the lo parts stay normal): 22 operand bits instead of the 16 of a bf16 pair at the same MMA cost --
//             the leakage floor of the first (bf16-pair) version, -100 dB below the strongest partial, drops by 36 dB
//   epilogue: P = re^2 + im^2 (fp32) -> bf16 hi/lo -> smem as the K-major, 128B-swizzled A operand of
//   GEMM2(c): D2[128 x 224] += P[128 x 64] . mel_chunk_c[224 x 64]^T              (again split x3)
// and after the last chunk D2 -> 10 log10(max(., 1e-10)) * bn_scale + bn_shift -> (B, T, 224) fp32.
// Framing needs no im2col: sample 320 t + 64 kb + kk is element (kk, kb % 5, t + kb / 5) of a 4-D TMA view
// (64, 5, hops, clips) of the padded waveform, so a 128-frame x 64-sample A tile is one box {64, 1, 128, 1}.
// Only bins that feed a non-zero mel weight are computed (n_chunks = 7 for fmax = 14 kHz).
//
// Warps: 0 = TMA ring producer, 1 = MMA issuer, 2 = TMEM alloc, 3 = mel-chunk TMA producer, 4..11 = epilogue.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace acx {

struct FeArgs {
  float* out;
  const float* bn_scale;
  const float* bn_shift;
  int T, n_chunks, tiles_per_clip, n_mels;
};

namespace fe {
constexpr int BM = 128, BK = 64, NK = 16 /* 1024 / 64 */;
constexpr int TILE = BM * BK * 2;            // 16 KB: one 128 x 64 bf16 operand tile
constexpr int STAGE = 4 * TILE;              // Ahi, Alo, Bhi, Blo
constexpr int STAGES = 2;
constexpr int MEL_ROWS = 224;
constexpr int MEL_TILE = MEL_ROWS * BK * 2;  // 28 KB
constexpr int OFF_P = STAGES * STAGE;        // P_hi, P_lo
constexpr int OFF_MEL = OFF_P + 2 * TILE;    // mel_hi, mel_lo
constexpr int OFF_BAR = OFF_MEL + 2 * MEL_TILE;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int NEPI = 8;
constexpr int THREADS = 128 + 32 * NEPI;
constexpr int D1_COLS = 128, D2_COL0 = 256, TMEM_COLS = 512;
static_assert(SMEM_BYTES <= 227 * 1024, "front-end smem budget");
static_assert(MEL_TILE % 1024 == 0, "swizzle alignment");
}  // namespace fe

__global__ void __launch_bounds__(fe::THREADS, 1)
    frontend_fused_kernel(const __grid_constant__ CUtensorMap tmWavHi, const __grid_constant__ CUtensorMap tmWavLo,
                          const __grid_constant__ CUtensorMap tmDftHi, const __grid_constant__ CUtensorMap tmDftLo,
                          const __grid_constant__ CUtensorMap tmMelHi, const __grid_constant__ CUtensorMap tmMelLo,
                          FeArgs a) {
  using namespace fe;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* d1_full = empty_bar + STAGES;
  uint64_t* d1_empty = d1_full + 2;
  uint64_t* p_full = d1_empty + 2;
  uint64_t* g2_done = p_full + 1;
  uint64_t* mel_full = g2_done + 1;
  uint64_t* d2_full = mel_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int clip = blockIdx.x / a.tiles_per_clip;
  const int t0 = (blockIdx.x % a.tiles_per_clip) * BM;

  if (warp == 0 && ptx::elect_one()) {
    ptx::prefetch_tensormap(&tmWavHi);
    ptx::prefetch_tensormap(&tmWavLo);
    ptx::prefetch_tensormap(&tmDftHi);
    ptx::prefetch_tensormap(&tmDftLo);
    ptx::prefetch_tensormap(&tmMelHi);
    ptx::prefetch_tensormap(&tmMelLo);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&d1_full[i], 1);
      ptx::mbar_init(&d1_empty[i], NEPI);
    }
    ptx::mbar_init(p_full, NEPI);
    ptx::mbar_init(g2_done, 1);
    ptx::mbar_init(mel_full, 1);
    ptx::mbar_init(d2_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();       // programmatic dependent launch (common.cuh): resources acquired, dependents may be scheduled
  pdl_wait();          // the split waveform is wave_prep's output
  uint8_t* sP = smem + OFF_P;
  uint8_t* sMel = smem + OFF_MEL;

  if (warp == 0) {
    // ===================== ring producer: frames (hi, lo) + DFT chunk rows (hi, lo) =====================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < a.n_chunks; ++c) {
        for (int kb = 0; kb < NK; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = smem + stage * STAGE;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], STAGE);
          ptx::tma_load_4d(s, &tmWavHi, &full_bar[stage], 0, kb % 5, t0 + kb / 5, clip);
          ptx::tma_load_4d(s + TILE, &tmWavLo, &full_bar[stage], 0, kb % 5, t0 + kb / 5, clip);
          ptx::tma_load_2d(s + 2 * TILE, &tmDftHi, &full_bar[stage], kb * BK, c * 128);
          ptx::tma_load_2d(s + 3 * TILE, &tmDftLo, &full_bar[stage], kb * BK, c * 128);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== mel-chunk producer ============================================================
    if (ptx::elect_one()) {
      for (int c = 0; c < a.n_chunks; ++c) {
        if (c > 0) ptx::mbar_wait(g2_done, (c - 1) & 1);  // GEMM2(c-1) finished reading the mel buffer
        ptx::mbar_arrive_expect_tx(mel_full, 2 * MEL_TILE);
        ptx::tma_load_2d(sMel, &tmMelHi, mel_full, 0, c * 256);
        ptx::tma_load_2d(sMel + MEL_TILE, &tmMelLo, mel_full, 0, c * 256);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ====================================================================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc1 = ptx::umma_idesc_f16(BM, 128);        // fp16 x fp16 -> fp32 (scaled split operands)
      constexpr uint32_t idesc2 = ptx::umma_idesc_bf16(BM, MEL_ROWS);
      const uint32_t d2 = tmem_base + D2_COL0;
      const uint64_t dPhi = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sP));
      const uint64_t dPlo = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sP + TILE));
      const uint64_t dMhi = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sMel));
      const uint64_t dMlo = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(sMel + MEL_TILE));
      auto gemm2 = [&](int cc) {
        ptx::mbar_wait(p_full, cc & 1);
        ptx::mbar_wait(mel_full, cc & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          ptx::umma_bf16(d2, dPhi + 2 * k, dMhi + 2 * k, idesc2, (cc | k) != 0 ? 1u : 0u);
          ptx::umma_bf16(d2, dPhi + 2 * k, dMlo + 2 * k, idesc2, 1u);
          ptx::umma_bf16(d2, dPlo + 2 * k, dMhi + 2 * k, idesc2, 1u);
        }
        ptx::umma_commit(g2_done);
      };
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < a.n_chunks; ++c) {
        const int buf = c & 1;
        ptx::mbar_wait(&d1_empty[buf], ((c >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t d1 = tmem_base + buf * D1_COLS;
        for (int kb = 0; kb < NK; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t s = ptx::smem_u32(smem + stage * STAGE);
          const uint64_t dAhi = ptx::umma_desc_sw128_kmajor(s);
          const uint64_t dAlo = ptx::umma_desc_sw128_kmajor(s + TILE);
          const uint64_t dBhi = ptx::umma_desc_sw128_kmajor(s + 2 * TILE);
          const uint64_t dBlo = ptx::umma_desc_sw128_kmajor(s + 3 * TILE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            ptx::umma_bf16(d1, dAhi + 2 * k, dBhi + 2 * k, idesc1, (kb | k) != 0 ? 1u : 0u);
            ptx::umma_bf16(d1, dAhi + 2 * k, dBlo + 2 * k, idesc1, 1u);
            ptx::umma_bf16(d1, dAlo + 2 * k, dBhi + 2 * k, idesc1, 1u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        ptx::umma_commit(&d1_full[buf]);
        if (c >= 1) gemm2(c - 1);  // issued behind GEMM1(c): the tensor pipe stays busy while P(c-1) is formed
      }
      gemm2(a.n_chunks - 1);
      ptx::umma_commit(d2_full);
    }
  } else if (warp >= 4) {
    // ===================== epilogue ======================================================================
    const int quad = warp & 3;
    const int group = (warp - 4) >> 2;       // bins [32 group, 32 group + 32) of the chunk
    const int row = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    for (int c = 0; c < a.n_chunks; ++c) {
      const int buf = c & 1;
      ptx::mbar_wait(&d1_full[buf], (c >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t re[32], im[32];
      ptx::tmem_ld_32x32b_x32(lane_base + buf * D1_COLS + group * 32, re);
      ptx::tmem_ld_32x32b_x32(lane_base + buf * D1_COLS + 64 + group * 32, im);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&d1_empty[buf]);
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float r0 = __uint_as_float(re[2 * j]), i0 = __uint_as_float(im[2 * j]);
        const float r1 = __uint_as_float(re[2 * j + 1]), i1 = __uint_as_float(im[2 * j + 1]);
        // operands carried 2^ACX_FE_SCALE_LOG2 each -> the power carries 2^(4 * ACX_FE_SCALE_LOG2): undo it exactly
        constexpr float kUnscale = 1.0f / (float)(1ull << (4 * ACX_FE_SCALE_LOG2));
        const float p0 = fmaf(r0, r0, i0 * i0) * kUnscale, p1 = fmaf(r1, r1, i1 * i1) * kUnscale;
        const __nv_bfloat162 h = __floats2bfloat162_rn(p0, p1);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(p0 - hf.x, p1 - hf.y);
        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
      }
      if (c >= 1) ptx::mbar_wait(g2_done, (c - 1) & 1);  // GEMM2(c-1) no longer reads P
      // K-major SWIZZLE_128B: row r at r*128 B, 16-byte chunk j stored at chunk (j ^ (r & 7))
      uint8_t* prow_hi = sP + row * 128;
      uint8_t* prow_lo = prow_hi + TILE;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = ((group * 4 + q) ^ (row & 7)) * 16;
        *reinterpret_cast<uint4*>(prow_hi + chunk) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
        *reinterpret_cast<uint4*>(prow_lo + chunk) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    // ---- log-mel epilogue --------------------------------------------------------------------------------
    ptx::mbar_wait(d2_full, 0);
    ptx::tc_fence_after();
    const int t = t0 + row;
    const bool ok = t < a.T;
    float* orow = a.out + ((size_t)clip * a.T + (ok ? t : 0)) * a.n_mels;
    constexpr int COLS = MEL_ROWS / 2;  // 112 per column group
#pragma unroll 1
    for (int c0 = group * COLS; c0 < (group + 1) * COLS; c0 += 16) {
      uint32_t r[16];
      ptx::tmem_ld_32x32b_x16(lane_base + D2_COL0 + c0, r);
      ptx::tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(a.bn_scale + c0 + j));
          const float4 sh = __ldg(reinterpret_cast<const float4*>(a.bn_shift + c0 + j));
          float4 o;
          // 10 log10(x) = 3.0102999566 * log2(x); power_to_db clamps at amin = 1e-10 (reference CX:165)
          o.x = fmaf(3.01029995664f * __log2f(fmaxf(__uint_as_float(r[j + 0]), 1e-10f)), sc.x, sh.x);
          o.y = fmaf(3.01029995664f * __log2f(fmaxf(__uint_as_float(r[j + 1]), 1e-10f)), sc.y, sh.y);
          o.z = fmaf(3.01029995664f * __log2f(fmaxf(__uint_as_float(r[j + 2]), 1e-10f)), sc.z, sh.z);
          o.w = fmaf(3.01029995664f * __log2f(fmaxf(__uint_as_float(r[j + 3]), 1e-10f)), sc.w, sh.w);
          *reinterpret_cast<float4*>(orow + c0 + j) = o;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, fe::TMEM_COLS);
  }
}

}  // namespace acx

using namespace acx;

extern "C" int acx_frontend_fused(const void* hi, const void* lo, int ld_pad, const void* dft_hi, const void* dft_lo,
                                  const void* mel_hi, const void* mel_lo, int n_chunks, const float* bn_scale,
                                  const float* bn_shift, float* out, int B, int T, int n_fft, int hop, int n_mels,
                                  void* stream) {
  ACX_CHECK(hi && lo && dft_hi && dft_lo && mel_hi && mel_lo && bn_scale && bn_shift && out, ACX_ERR_ARG,
            "frontend_fused: null pointer");
  ACX_CHECK(n_fft == 1024 && hop == 320 && n_mels == 224, ACX_ERR_UNSUPPORTED,
            "frontend_fused: built for n_fft=1024, hop=320, 224 mel bins (reference convnext.py:161-174)");
  ACX_CHECK(B > 0 && T > 0 && n_chunks >= 1 && n_chunks <= 9, ACX_ERR_ARG, "frontend_fused: bad sizes");
  const int hops = ld_pad / hop;
  ACX_CHECK(ld_pad % 8 == 0 && hops >= T + 3, ACX_ERR_ARG,
            "frontend_fused: ld_pad=%d must be a multiple of 8 and cover %d hops of %d samples", ld_pad, T + 3, hop);
  CUtensorMap tmWavHi, tmWavLo, tmDftHi, tmDftLo, tmMelHi, tmMelLo;
  {
    // (kk: 64 samples) x (j: 5 sub-blocks of a hop) x (hop index) x (clip)
    cuuint64_t dims[4] = {64, 5, (cuuint64_t)hops, (cuuint64_t)B};
    cuuint64_t strides[3] = {128, (cuuint64_t)hop * 2, (cuuint64_t)ld_pad * 2};
    cuuint32_t box[4] = {64, 1, 128, 1};
    int rc = make_tmap_bf16(&tmWavHi, hi, 4, dims, strides, box);
    if (rc != ACX_OK) return rc;
    rc = make_tmap_bf16(&tmWavLo, lo, 4, dims, strides, box);
    if (rc != ACX_OK) return rc;
  }
  int rc = make_tmap_2d_bf16(&tmDftHi, dft_hi, (uint64_t)n_fft, (uint64_t)n_chunks * 128, (uint64_t)n_fft * 2, 64, 128);
  if (rc != ACX_OK) return rc;
  rc = make_tmap_2d_bf16(&tmDftLo, dft_lo, (uint64_t)n_fft, (uint64_t)n_chunks * 128, (uint64_t)n_fft * 2, 64, 128);
  if (rc != ACX_OK) return rc;
  rc = make_tmap_2d_bf16(&tmMelHi, mel_hi, 64, (uint64_t)n_chunks * 256, 128, 64, fe::MEL_ROWS);
  if (rc != ACX_OK) return rc;
  rc = make_tmap_2d_bf16(&tmMelLo, mel_lo, 64, (uint64_t)n_chunks * 256, 128, 64, fe::MEL_ROWS);
  if (rc != ACX_OK) return rc;

  ACX_SET_MAX_SMEM(frontend_fused_kernel, fe::SMEM_BYTES);
  FeArgs a;
  a.out = out;
  a.bn_scale = bn_scale;
  a.bn_shift = bn_shift;
  a.T = T;
  a.n_chunks = n_chunks;
  a.tiles_per_clip = ceil_div(T, fe::BM);
  a.n_mels = n_mels;
  const int grid = B * a.tiles_per_clip;
  ACX_CUDA(launch_pdl(frontend_fused_kernel, dim3(grid), dim3(fe::THREADS), fe::SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream), 1, PDL_SMALL,
                      tmWavHi, tmWavLo, tmDftHi, tmDftLo, tmMelHi, tmMelLo, a));
  return ACX_OK;
}
