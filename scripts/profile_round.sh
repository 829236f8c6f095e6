#!/bin/bash
# ncu evidence of one round (run under gpurun, 1 GPU): launch list of a full-size step, --set full of the dominant
# kernel at 256^2 x 256 spp, DRAM / L2 / L1 traffic of the dominant kernel at full size.  Outputs in gpurun_out/.
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
ncu --set full --clock-control none --import-source on -k regex:k_prim_shade -c 1 -f -o gpurun_out/k_prim_shade_full \
    python bench.py --config 1 --steps 1 --warmup 0 $B > gpurun_out/prim_bench.log 2>&1
ls -la gpurun_out/
