#!/bin/bash
# usage: ab.sh  (expects libac75_base.tmp.so in repo root = baseline; current build = new)
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["kernel_ms"], d["ms_per_step"], d["matches_per_step"], d["config"]["fallback_count"])'
cp aho-corasick-1975_b200/libac75.so new.tmp.so
for i in 1 2; do
  cp libac75_base.tmp.so aho-corasick-1975_b200/libac75.so; echo base; $B 2>/dev/null | tail -1 | python -c "$P"
  cp new.tmp.so aho-corasick-1975_b200/libac75.so; echo new; $B 2>/dev/null | tail -1 | python -c "$P"
done
