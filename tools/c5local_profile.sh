python tools/slab_local_c5.py 8 > gpurun_out/r02_c5local_plain.log 2>&1; tail -2 gpurun_out/r02_c5local_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/r02_c5local_launches.csv python tools/slab_local_c5.py 8 > gpurun_out/r02_c5local.log 2>&1; tail -3 gpurun_out/r02_c5local.log; wc -l gpurun_out/r02_c5local_launches.csv
