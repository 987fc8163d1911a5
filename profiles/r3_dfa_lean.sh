# round-2 (session 3): lean DFA count pass + read-ahead event expansion -- parity, then the config-2 bench line and its launch list
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_fuzz.py -x -q -m gpu -k "dfa or config1 or config2 or fuzz_block or carried or generic or streaming" 2>&1 | tail -4
for v in "dfa_lean=0" "dfa_lean=1"; do
  python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --option $v 2>gpurun_out/r3_c2.err | tail -1 > gpurun_out/r3_c2_$v.json
  python -c "import json; d=json.load(open('gpurun_out/r3_c2_$v.json')); print('$v', d['ms_per_step'], d['kernel_ms'], round(d['roofline']['frac'],4), d['matches_per_step'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r3_launches_c2_1gib.csv python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
grep -c dfa_ gpurun_out/r3_launches_c2_1gib.csv
