for lg in 22 23 24; do
  python bench.py --config c4s --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --no-extra --option s2_dist_log2=$lg 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lg', $lg, round(d['value'],1), d['kernel_ms'], round(d['roofline']['frac'],4), d['config']['filter_hit_rate'], d['matches_per_step'], d['candidates_per_step_rank0'], d['config']['finalise_ms'])"
done
