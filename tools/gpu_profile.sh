# round-2 evidence run: launch list, --set full on the kernels that changed (exported to CSV ON the box: the .ncu-rep
# with sources is larger than the 64 MiB that travel back), sanitizers on the new kernels' tests
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python tools/run_once.py 64 > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"dwconv_tc3|mlp_fused|dwconv_ln" -o /tmp/r2_prof -f python tools/run_once.py 64 > gpurun_out/r2_prof.log 2>&1
ncu -i /tmp/r2_prof.ncu-rep --page raw --csv > gpurun_out/r2_prof_raw.csv 2>/dev/null
ncu -i /tmp/r2_prof.ncu-rep --page source --csv -k regex:"dwconv_tc3_kernel<56>" -c 1 > gpurun_out/r2_src_dwtc3.csv 2>/dev/null
ls -la /tmp/r2_prof.ncu-rep gpurun_out/
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2b_racecheck.log python -m pytest tests/test_gpu_kernels.py tests/test_gpu_umma.py -m gpu -x -q -k "dwconv_tc or mlp_fused" > gpurun_out/r2b_racecheck_pytest.log 2>&1; tail -2 gpurun_out/r2b_racecheck_pytest.log; grep -c "Error: Race" gpurun_out/r2b_racecheck.log
timeout 120 python tools/time_mlp.py
