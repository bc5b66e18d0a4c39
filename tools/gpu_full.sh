mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/bench_full.err | tee gpurun_out/bench_full.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches'])
print(d['roofline'])
for r in d['roofline_all']: print(r['kernel'], r['launches'], r['avg_launch_ms'], r['frac'])"
tail -3 gpurun_out/bench_full.err
