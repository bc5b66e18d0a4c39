// Entry points whose tcgen05 kernels are still being brought up; they fail loudly (no fallback).
#include "common.cuh"
using namespace acx;
extern "C" {
int acx_mlp_fused(const void*, void*, const void*, const float*, const void*, const float*, const float*, int, int,
                  void*) {
  set_error("acx_mlp_fused: not available in this build");
  return ACX_ERR_UNSUPPORTED;
}
}
