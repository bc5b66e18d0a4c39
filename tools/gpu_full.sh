mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/bench_full.err | tee gpurun_out/bench_full.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['gpu_launches'])"
tail -3 gpurun_out/bench_full.err
python -c "import __graft_entry__ as g; g.smoke()"
