# round-2 (third pass) evidence: launch list of one forward, --set full over ALL its launches (raw CSV made on the box),
# sanitizers on the new kernel's tests, on the tcgen05 kernel tests (programmatic dependent launch on) and on a whole forward
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c_launches.csv python tools/run_once.py 64 > gpurun_out/r2c_launches.log 2>&1
ncu --set full --clock-control none --profile-from-start off -o /tmp/r2c_all -f python tools/run_once.py 64 > gpurun_out/r2c_all.log 2>&1
ncu -i /tmp/r2c_all.ncu-rep --page raw --csv > gpurun_out/r2c_all_raw.csv 2>/dev/null
ls -la /tmp/r2c_all.ncu-rep gpurun_out/r2c_all_raw.csv
timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/r2c_racecheck_all.log python -m pytest tests/test_gpu_kernels.py tests/test_gpu_umma.py -m gpu -x -q -k "dwconv_tc or mlp_fused or downsample_fused or stage2 or gemm" > gpurun_out/r2c_racecheck_all_pytest.log 2>&1; tail -2 gpurun_out/r2c_racecheck_all_pytest.log; grep -c "Error: Race" gpurun_out/r2c_racecheck_all.log; tail -3 gpurun_out/r2c_racecheck_all.log
