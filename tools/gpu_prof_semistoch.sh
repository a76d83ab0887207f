#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_prof_semistoch.sh <tag>
# ncu --set full of the semi-stochastic kernels (k_ss_hamil, k_ss_locate, k_ss_project) in a bench.py --semi-stoch run:
# 1e7 walkers, a 20,000-determinant CISD-like space (4.85e7 stored Hamiltonian elements).  Key counters are exported to
# gpurun_out/<tag>_ncu_semistoch.csv (copied to profiles/ by hand).
tag=${1:-r2}
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'k_ss_' -c 12 -f -o gpurun_out/${tag}_prof_ss \
    python bench.py --walkers 1e7 --semi-stoch 20000 --steps 2 --warmup 1 --tau 5.3e-7 --no-cpu-baseline > gpurun_out/${tag}_prof_ss.log 2>&1
ncu -i gpurun_out/${tag}_prof_ss.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active \
    > gpurun_out/${tag}_ncu_semistoch.csv
rm -f gpurun_out/${tag}_prof_ss.ncu-rep
cat gpurun_out/${tag}_ncu_semistoch.csv | cut -c1-400 | head -20
