# one GPU round of the group-planar tensor-core depthwise-conv work: kernel tests, model parity, timings
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_umma.py -m gpu -x -q -k "dwconv_tc or mlp_fused" 2>&1 | tail -8
timeout 120 python tools/time_dwtc.py 2>&1 | grep -v CTA | tail -8
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_b.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for r in d["roofline_all"][:14]: print(r["kernel"], r["launches"], r["avg_launch_ms"], r["frac"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_bench_b.err").read()[-1500:])
P
