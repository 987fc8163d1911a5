for b in 2 3; do
  python bench.py --config c4s --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --no-extra --option s2_batches=$b 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batches', $b, round(d['value'],1), d['kernel_ms'], round(d['roofline']['frac'],4))"
done
