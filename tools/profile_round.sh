#!/bin/bash
# The measurement recipe of a round (run on the GPU box: gpurun -- 'bash tools/profile_round.sh r01d'):
# parity suite, bench (both arms), ncu launch list and one full capture of the force kernel; outputs under gpurun_out/.
tag=${1:-round}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 13500 -c 300 --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_run.py 6000 400 > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:force_kernel -s 6000 -c 1 -o gpurun_out/${tag}_force -f python tools/profile_run.py 6000 100 > gpurun_out/${tag}_force.log 2>&1
ncu -i gpurun_out/${tag}_force.ncu-rep --page raw --csv > gpurun_out/${tag}_force_raw.csv 2>/dev/null
# (launch list: -s 13500 skips the ~13 400 launches of the 6 000 equilibration steps; full capture -s 6000: the launch after the equilibration, i.e. the compacted condensate bench.py times; r01c/r01d used -s 400, a lighter state)
python tools/sass_budget.py gpurun_out/${tag}_force.ncu-rep > gpurun_out/${tag}_force_sass_budget.txt 2>/dev/null
