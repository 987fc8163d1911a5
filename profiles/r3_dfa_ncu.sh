ncu --set full --clock-control none --import-source on -k regex:dfa_scan_tma_lean2 -s 1 -c 1 -o gpurun_out/r3_dfa_lean2_c2 python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
