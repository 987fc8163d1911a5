for v in 0 1; do
ncu --metrics gpu__time_duration.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.per_cycle_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,dram__bytes_write.sum,dram__bytes_read.sum --clock-control none -c 40 --csv --log-file gpurun_out/r3_launches_c2_lean$v.csv python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra --option dfa_lean=$v > /dev/null 2>&1
done
ls -la gpurun_out/r3_launches*
