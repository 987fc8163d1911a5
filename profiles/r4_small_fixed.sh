# round-2 (session 4): 16-byte filter copy into shared memory + one-block scan for up to 32 Ki tiles -- parity, then the per-scan fixed costs at shard sizes
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_concurrent.py tests/test_blob.py -x -q -m gpu -k "not config2_full and not config4_full" 2>&1 | tail -3
for g in 8 1 0.25; do
  python bench.py --gib $g --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('gib', $g, round(d['value'],1), 'GB/s  ms/step', round(d['ms_per_step'],4), d['kernel_ms'], 'launches/step', d['gpu_launches']//20)"
done
python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c2', round(d['value'],1), d['ms_per_step'], d['kernel_ms'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/r4_launches_c3_1gib.csv python bench.py --gib 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extra > /dev/null 2>&1
grep -E "filter_|scan_" gpurun_out/r4_launches_c3_1gib.csv | awk -F'","' '{print $5, $NF}' | tail -7
