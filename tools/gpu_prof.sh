#!/bin/bash
# usage (on the GPU box): tools/gpu_prof.sh <tag> <excit_gen> [walkers]
# ncu --set full capture (with SASS-level source counters) of ONE k_spawn_death launch; exports raw + source CSV.
tag=$1; gen=${2:-heat_bath}; n=${3:-1e7}; extra=${4:-}
mkdir -p gpurun_out
cp hande_b200/libhande_b200.so gpurun_out/lib_${tag}.so
ncu --set full --import-source on --clock-control none -k regex:k_spawn_death -s 1 -c 1 -f -o gpurun_out/prof_${tag} \
    python bench.py $extra --excit-gen $gen --walkers $n --steps 1 --warmup 1 --tau 5.3e-7 --no-cpu-baseline --no-e2e > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}.csv
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv > gpurun_out/src_${tag}.csv
rm -f gpurun_out/prof_${tag}.ncu-rep
tail -2 gpurun_out/prof_${tag}.log | cut -c1-300
