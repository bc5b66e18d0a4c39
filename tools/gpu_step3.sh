timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_umma.py -m gpu -x -q -k "dwconv_tc or stage2 or planar" 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/r2_bench_d.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for r in d['roofline_all']: print(r['kernel'], r['launches'], r['avg_launch_ms'], r['frac'])"
tail -3 gpurun_out/r2_bench_d.err
