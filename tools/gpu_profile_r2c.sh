# round-2 (third pass) evidence: launch list of one forward, --set full on the new downsample kernel, sanitizers on its tests
# and on a whole forward with programmatic dependent launch on
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches.csv python tools/run_once.py 64 > gpurun_out/r2c_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"ds_fused" -o /tmp/r2c_prof -f python tools/run_once.py 64 > gpurun_out/r2c_prof.log 2>&1
ncu -i /tmp/r2c_prof.ncu-rep --page raw --csv > gpurun_out/r2c_prof_raw.csv 2>/dev/null
ls -la /tmp/r2c_prof.ncu-rep gpurun_out/ | tail -5
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2c_racecheck.log python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "downsample_fused" > gpurun_out/r2c_racecheck_pytest.log 2>&1; tail -2 gpurun_out/r2c_racecheck_pytest.log; grep -c "Error: Race" gpurun_out/r2c_racecheck.log; tail -3 gpurun_out/r2c_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2c_memcheck.log python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q -k "downsample_fused or routes_agree" > gpurun_out/r2c_memcheck_pytest.log 2>&1; tail -2 gpurun_out/r2c_memcheck_pytest.log; tail -3 gpurun_out/r2c_memcheck.log
