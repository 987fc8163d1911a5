# round-2 (session 4): dense mode with byte-sized stage entries (two-row tiles for 32-bit symbols) + one 16-byte record per output state for the event kernels
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_blob.py -x -q -m gpu -k "not config2_full and not config4_full and not config3_full" 2>&1 | tail -4
for c in c5 c2; do
  python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>gpurun_out/r4_$c.err | tail -1 > gpurun_out/r4e_bench_$c.json
  python -c "import json; d=json.load(open('gpurun_out/r4e_bench_$c.json')); print('$c', round(d['value'],1), d['ms_per_step'], d['kernel_ms'], round(d['roofline']['frac'],4), d['matches_per_step'], d['candidates_per_step_rank0'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r4e_launches_c5.csv python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
grep -E "filter_|scan_" gpurun_out/r4e_launches_c5.csv | awk -F'","' '{print $5, $NF}' | sed -n 7,14p
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r4e_launches_c2.csv python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
grep -E "dfa_|scan_" gpurun_out/r4e_launches_c2.csv | awk -F'","' '{print $5, $NF}' | tail -6
