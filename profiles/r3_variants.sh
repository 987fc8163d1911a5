# round-2 (session 3): stride-2 kernel variants -- parity under each variant, then the config-3 bench line of each
set -x
for v in 1 2; do
  ACM_B200_S2_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -x -q -m gpu -k "stride2 or streaming or carried" 2>&1 | tail -3
done
for v in 0 1 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --option s2_variant=$v 2>gpurun_out/r3_v$v.err | tail -1 > gpurun_out/r3_v$v.json
  python -c "import json; d=json.load(open('gpurun_out/r3_v$v.json')); print('variant',$v, d['kernel_ms'], round(d['roofline']['frac'],4), d['matches_per_step'], d['candidates_per_step_rank0'])"
done
