import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from helpers import ac75, random_patterns
rng = np.random.default_rng(1002)
flat, offsets = random_patterns(300, lmin=1, lmax=8, seed=2, alphabet=2, width=1)
text = rng.integers(0, 2, size=200_000).astype(np.uint8)
for tma in (1, 0, 1):
    m = ac75().Machine(1); m.insert_many(flat=flat, offsets=offsets); m.set_option("dfa_tma", tma)
    tot = [m.scan(text, count_only=True) for _ in range(30)]
    print("tma", tma, sorted(set(tot)), m.stats()["dfa_tma_scans"])
    m.close()
# sparse text (no event overflow)
flat, offsets = random_patterns(300, lmin=3, lmax=8, seed=3, alphabet=26, width=1)
text = rng.integers(0, 26, size=3_000_000).astype(np.uint8)
for tma in (1, 0):
    m = ac75().Machine(1); m.insert_many(flat=flat, offsets=offsets); m.set_option("dfa_tma", tma)
    tot = [m.scan(text, count_only=True) for _ in range(30)]
    print("sparse tma", tma, sorted(set(tot)), m.stats()["engine"], m.stats()["dfa_tma_scans"])
    m.close()
