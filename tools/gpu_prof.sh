#!/bin/bash
# usage (on the GPU box): tools/gpu_prof.sh <tag> <excit_gen> [walkers] [kernel regex] [object file]
# ncu --set full capture (with SASS-level source counters) of ONE spawn-kernel launch; exports raw + source CSV.
tag=$1; gen=${2:-heat_bath}; n=${3:-1e7}; kre=${4:-k_spawn}; obj=${5:-hande_b200/build/hb_spawn_w2_g0.o}
mkdir -p gpurun_out
cp $obj gpurun_out/obj_${tag}.o
ncu --set full --import-source on --clock-control none -k regex:$kre -s 1 -c 1 -f -o gpurun_out/prof_${tag} \
    python bench.py --excit-gen $gen --walkers $n --steps 1 --warmup 1 --tau 5.3e-7 --no-cpu-baseline --no-e2e > gpurun_out/prof_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/raw_${tag}.csv
ncu -i gpurun_out/prof_${tag}.ncu-rep --page source --csv > gpurun_out/src_${tag}.csv
rm -f gpurun_out/prof_${tag}.ncu-rep
tail -2 gpurun_out/prof_${tag}.log | cut -c1-300
