#!/bin/bash
# usage (on the GPU box): tools/gpu_prof_ccmc.sh <tag> [excips] [launchlist]
# ncu --set full capture (SASS-level source counters) of one k_ccmc_cluster launch; optional launch list of two cycles
tag=$1; n=${2:-3e6}
mkdir -p gpurun_out
if [ -n "$3" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^(k_|void k_)' -c 300 --csv --log-file gpurun_out/launches_ccmc_${tag}.csv \
    python tools/bench_ccmc.py --excips $n --steps 2 --warmup 1 > gpurun_out/launches_ccmc_${tag}.log 2>&1
fi
ncu --set full --import-source on --clock-control none -k regex:k_ccmc_cluster -s 1 -c 1 -f -o gpurun_out/prof_ccmc_${tag} \
    python tools/bench_ccmc.py --excips $n --steps 1 --warmup 1 > gpurun_out/prof_ccmc_${tag}.log 2>&1
ncu -i gpurun_out/prof_ccmc_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_ccmc_${tag}.csv
ncu -i gpurun_out/prof_ccmc_${tag}.ncu-rep --page source --csv > gpurun_out/src_ccmc_${tag}.csv
rm -f gpurun_out/prof_ccmc_${tag}.ncu-rep
tail -2 gpurun_out/prof_ccmc_${tag}.log | cut -c1-300
