for mask in 0 31 1 2 4 8 16 0 31 3 24 28; do ACX_PDL=$mask python tools/time_step.py 64 30 3 2>&1 | tail -1; done
