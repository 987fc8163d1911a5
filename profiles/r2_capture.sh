# round-2 evidence: launch lists and ncu --set full captures of the dominant kernels (each after the plain command exited 0)
set -x
python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/plain_c3.log 2>&1 || exit 1
for b in 1 2 3; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --option s2_batches=$b 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('s2_batches', $b, d['kernel_ms'], round(d['roofline']['frac'],4))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c3_8gib.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:filter_scan_s2 -s 1 -c 1 -o gpurun_out/r2_f1s_c3_8gib python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
python bench.py --config c2 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/plain_c2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c2_1gib.csv python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:dfa_ -s 4 -c 4 -o gpurun_out/r2_dfa_c2_1gib python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
python bench.py --config c4s --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/plain_c4s.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:filter_scan_s2 -s 1 -c 1 -o gpurun_out/r2_f1s_c4s_2gib python bench.py --config c4s --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ls -la gpurun_out/
