# round-2 (session 4): config 2 on one box -- the lean count pass (events only + count from the event lists) against the counting pass
for v in "dfa_lean=1" "dfa_lean=0" "dfa_lean=1" "dfa_lean=0"; do
  python bench.py --config c2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extra --option $v 2>gpurun_out/r4_c2.err | tail -1 > gpurun_out/r4f_c2.json
  python -c "import json; d=json.load(open('gpurun_out/r4f_c2.json')); print('$v', round(d['value'],1), d['ms_per_step'], d['kernel_ms'], round(d['roofline']['frac'],4), d['matches_per_step'])"
done
