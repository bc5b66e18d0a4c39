mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "downsample_fused" 2>&1 | tail -5
ACX_DS_FUSED=1 python tools/time_stages.py 64 2>&1 | grep -i "ds_\|patchify\|clips/s"
