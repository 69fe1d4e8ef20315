#!/bin/bash
# DRAM traffic per launch (ncu) for the lines whose roofline.traffic was null: 4-source merge kernel, whole TIES plan run
mkdir -p gpurun_out
{
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ties_ -s 33 -c 11 --csv \
   --log-file gpurun_out/r2_ties_traffic.csv python bench.py --workload ties > gpurun_out/r2_prof9_a.log 2>&1
tail -1 gpurun_out/r2_prof9_a.log | cut -c1-200
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:merge_kernel -s 3 -c 1 --csv \
   --log-file gpurun_out/r2_merge_n4_traffic.csv python bench.py --workload merge --merge-config n4 --no-e2e --steps 2 --warmup 3 > gpurun_out/r2_prof9_b.log 2>&1
tail -1 gpurun_out/r2_prof9_b.log | cut -c1-200
grep -v "^==" gpurun_out/r2_ties_traffic.csv | tail -40 | cut -c1-200
grep -v "^==" gpurun_out/r2_merge_n4_traffic.csv | tail -5 | cut -c1-200
} > gpurun_out/r2_prof9.log 2>&1
tail -c 7000 gpurun_out/r2_prof9.log
