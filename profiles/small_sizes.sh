# per-scan fixed costs: the default config at the shard sizes of a strong-scaling run (8 GiB / N)
for g in 8 4 2 1; do
  python bench.py --gib $g --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('gib', $g, round(d['value'],1), 'GB/s  ms/step', round(d['ms_per_step'],4), d['kernel_ms'], 'launches/step', d['gpu_launches']//20)"
done
