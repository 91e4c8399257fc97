O=gpurun_out; T=r02c
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-mailbox > $O/${T}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:closed_loop_kernel_m -s 4 -c 1 -f -o $O/${T}_loop python tools/loop_timing.py 512 6 20 > $O/${T}_ncu_full.log 2>&1
tail -1 $O/${T}_ncu_full.log
