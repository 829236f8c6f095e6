#!/bin/bash
# profile_round.sh without the primary-stage capture (the two .ncu-rep files together are at gpurun's 64 MiB pull limit)
set -x
mkdir -p gpurun_out
B="--no-e2e --no-cpu-baseline --no-other-configs"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 $B $BENCH_ARGS > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_shade_wf -c 1 -f -o gpurun_out/k_shade_wf_full \
    python bench.py --res 256 --spp 256 --steps 1 --warmup 0 $B $BENCH_ARGS > gpurun_out/full_bench.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum \
    --clock-control none -k regex:k_shade_wf -c 1 --csv --log-file gpurun_out/traffic_full.csv \
    python bench.py --steps 1 --warmup 0 $B $BENCH_ARGS > gpurun_out/traffic_bench.log 2>&1
