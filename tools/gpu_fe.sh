mkdir -p gpurun_out
ACX_FRONTEND=fused timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "frontend_fused_logmel and fused" 2>&1 | tail -4
for cl in 1 2 4; do ACX_FE_CLUSTER=$cl ACX_FRONTEND=fused python tools/time_stages.py 64 2>&1 | grep -i "front\|clips/s" | tr '\n' ' '; echo; done
