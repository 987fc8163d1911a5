# round-2 (session 4): config-5 slice -- where the 1.0 ms outside the streaming kernel goes (launch list + full sets of the verification kernels)
python bench.py --config c5 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > gpurun_out/r4_plain_c5.log 2>&1 || { tail -5 gpurun_out/r4_plain_c5.log; exit 1; }
tail -1 gpurun_out/r4_plain_c5.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms'], d['matches_per_step'], d['candidates_per_step_rank0'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r4_launches_c5.csv python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"filter_verify|filter_tile_totals" -s 3 -c 3 -o gpurun_out/r4_verify_c5 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
ls -la gpurun_out/
