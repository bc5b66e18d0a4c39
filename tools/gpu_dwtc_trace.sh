ACX_NVCC_EXTRA="-DACX_ENABLE_TRACE" ACX_LIBACX=$PWD/audioset-convnext-inf_b200/libacx_trace.so timeout 120 python tools/time_dwtc.py 2>&1 | tail -14
