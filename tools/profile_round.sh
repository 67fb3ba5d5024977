#!/bin/bash
# The measurement recipe of a round (run on the GPU box: gpurun -- 'bash tools/profile_round.sh r02'): ncu launch list of the C2
# bench state, full captures of the dominant kernels (force_kernel = Verlet rows at C2; tile_force_kernel, tile_build_kernel and
# integrate_kernel at C5; build_list_kernel at C2), each exported as raw-metric CSV; outputs under gpurun_out/.
tag=${1:-round}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# every launch of 2 rebuild windows at the bench state (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 400 -c 150 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_run.py 150 60 > gpurun_out/${tag}_launches.log 2>&1
# C2 bench state: the fused force launch (rows), the row build, the integrator
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"force_kernel|build_list_kernel|integrate_kernel" -s 30 -c 6 -f -o gpurun_out/${tag}_c2 python tools/profile_run.py 60 40 > gpurun_out/${tag}_c2.log 2>&1
ncu -i gpurun_out/${tag}_c2.ncu-rep --page raw --csv > gpurun_out/${tag}_c2_raw.csv 2>/dev/null
# C5 (1 043 200 beads): the tile force launch, the tile list build, the integrator at a bandwidth-bound size
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"tile_force_kernel|tile_build_kernel|integrate_kernel" -s 20 -c 4 -f -o gpurun_out/${tag}_c5 python tools/c5_profile_run.py > gpurun_out/${tag}_c5.log 2>&1
ncu -i gpurun_out/${tag}_c5.ncu-rep --page raw --csv > gpurun_out/${tag}_c5_raw.csv 2>/dev/null
