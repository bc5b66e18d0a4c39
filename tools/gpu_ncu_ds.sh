mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:ds_fused -s 2 -c 1 -o /tmp/ds_prof -f python tools/trace_ds.py 0 64 > gpurun_out/ds_prof.log 2>&1
ncu -i /tmp/ds_prof.ncu-rep --page raw --csv > gpurun_out/ds_prof_raw.csv 2>/dev/null
ncu -i /tmp/ds_prof.ncu-rep --page source --csv > gpurun_out/ds_prof_src.csv 2>/dev/null
ls -la gpurun_out/ds_prof*
tail -3 gpurun_out/ds_prof.log
