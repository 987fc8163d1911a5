for kb in 180 196 212 228; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --option s2_smem_kb=$kb 2>/dev/null | tail -1 > gpurun_out/sw_$kb.json
  python -c "import json; d=json.load(open('gpurun_out/sw_$kb.json')); print('kb',$kb, d['config']['smem_bytes'], d['config']['filter_hit_rate'], d['kernel_ms'], round(d['roofline']['frac'],4))"
done
