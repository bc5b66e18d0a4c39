timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_umma.py -m gpu -x -q -k "dwconv_tc or mlp_fused" 2>&1 | tail -5
timeout 120 python tools/time_dwtc.py 2>&1 | grep -E "planar|stage [01]"
timeout 120 python tools/time_mlp.py
