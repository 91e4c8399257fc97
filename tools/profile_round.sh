#!/bin/bash
# Round-end measurements on the GPU box (run through gpurun): bench lines of the three workloads and of the reference arm,
# the ncu launch list (+ DRAM bytes per launch) of the bench command, and one full ncu capture of a steady-state launch.
# Numbers printed by runs under ncu are never bench values.
O=gpurun_out; T=${1:-r02b}
python bench.py > $O/${T}_bench_cp20.log 2>&1
python bench.py --workload cp40 > $O/${T}_bench_cp40.log 2>&1
python bench.py --workload syn30 > $O/${T}_bench_syn30.log 2>&1
python bench.py --impl reference > $O/${T}_bench_reference.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv \
    --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-mailbox > $O/${T}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:closed_loop_kernel_m -s 4 -c 1 -f -o $O/${T}_loop \
    python tools/loop_timing.py 512 6 20 > $O/${T}_ncu_full.log 2>&1
python tools/bench_summary.py $O/${T}_bench_cp20.log $O/${T}_bench_cp40.log $O/${T}_bench_syn30.log
tail -c 600 $O/${T}_bench_reference.log
