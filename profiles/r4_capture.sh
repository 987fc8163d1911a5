# round-2 (session 4) evidence: the whole GPU suite, the default bench line with its extra configs, the reference arm, launch lists and
# ncu --set full captures of the kernels touched this session (each capture after the plain command exited 0)
set -x
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r4_pytest.log 2>&1; tail -3 gpurun_out/r4_pytest.log
python bench.py > gpurun_out/r4_bench_c3.json 2> gpurun_out/r4_bench_c3.err || { tail -5 gpurun_out/r4_bench_c3.err; exit 1; }
python -c "
import json; d=json.load(open('gpurun_out/r4_bench_c3.json'))
print('c3', round(d['value'],1), d['ms_per_step'], d['kernel_ms'], round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), 'cpu', d['cpu_baseline']['value'])
for e in d['extra_configs']: print(e['config']['name'], round(e['value'],1), e['ms_per_step'], e['kernel_ms'], round(e['roofline']['frac'],4), 'e2e', round(e['e2e']['value'],1), e.get('meyer',{}).get('table_update_ms_mean'))
"
python bench.py --impl reference > gpurun_out/r4_bench_c3_reference.json 2>/dev/null; cat gpurun_out/r4_bench_c3_reference.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r4_launches_c3_8gib.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r4_launches_c2_1gib.csv python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"filter_scan_kernel|filter_verify|filter_tile_totals" -s 6 -c 4 -o gpurun_out/r4_c5_kernels python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ncu --set full --clock-control none -k regex:filter_scan_s2 -s 1 -c 1 -o gpurun_out/r4_f1s_c4s_2gib python bench.py --config c4s --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"filter_scan_s2|filter_verify" -s 2 -c 2 -o gpurun_out/r4_f1s_f2_c3_8gib python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ls -la gpurun_out/ | tail -12
