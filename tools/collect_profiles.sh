#!/bin/bash
# usage (on the GPU box, via gpurun): tools/collect_profiles.sh <tag>
# Everything profiles/<tag>_summary.md is built from (tools/summarize_profiles.py <tag> afterwards, in the build container):
#   <tag>_bench.json / <tag>_bench_reference.json  bench.py default line and the reference arm
#   <tag>_launches.csv                             ncu launch list (gpu__time_duration) of a 1e7-walker run
#   <tag>_raw_1e8.csv, <tag>_src_spawn.csv         ncu --set full (+SASS source counters) of one cycle's hot kernels at 1e8
#                                                  (line attribution uses the in-tree libhande_b200.so that was profiled)
tag=${1:-r1}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --walkers 1e7 --steps 2 --warmup 1 --tau 5.3e-7 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'k_spawn_death|k_merge|k_round_count|k_radix_scatter|k_annihilate' \
    -s 18 -c 18 -f -o gpurun_out/${tag}_prof_1e8 \
    python bench.py --walkers 1e8 --steps 1 --warmup 1 --tau 5.3e-7 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_prof_1e8.log 2>&1
# export what the summary needs and drop the (large) report: gpurun brings back at most 64 MiB
ncu -i gpurun_out/${tag}_prof_1e8.ncu-rep --page raw --csv > gpurun_out/${tag}_raw_1e8.csv
ncu -i gpurun_out/${tag}_prof_1e8.ncu-rep --page source --csv --kernel-name regex:k_spawn_death > gpurun_out/${tag}_src_spawn.csv
rm -f gpurun_out/${tag}_prof_1e8.ncu-rep
tail -c 400 gpurun_out/${tag}_bench.json; echo; tail -3 gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_prof_1e8.log | cut -c1-200
