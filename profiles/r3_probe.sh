for r in 2; do
ACM_B200_DFA_PROBE=$r ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dfa_scan_tma -c 2 --csv --log-file gpurun_out/r3_probe_p$r.csv python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
grep dfa_scan gpurun_out/r3_probe_p$r.csv | awk -F'","' '{print $5, $NF}'
done
