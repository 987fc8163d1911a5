"""N>1 path on CPU: world_size-2 and -3 gloo process groups run the sharding logic (plan, per-shard scan with lead, all-gather of
the match counts, host concatenation).  The per-shard scanner here is the oracle (tests may use it); on the GPU box bench.py
plugs the CUDA scan into the very same plan."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ac75, random_patterns
from oracle import pyoracle


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    flat, offsets = random_patterns(300, lmin=2, lmax=12, seed=7, alphabet=4)
    text = np.random.default_rng(3).integers(0, 4, size=50_000).astype(np.uint8)
    o = pyoracle.Oracle("port", 1)
    o.insert_many(flat=flat, offsets=offsets)

    def scan_fn(first, nb, lead, base):
        return o.scan_lead(text[first:first + nb], lead, base=base, cap=1 << 22)

    def all_gather_counts(local):
        t = torch.tensor([local], dtype=torch.int64)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [int(x.item()) for x in out]

    recs, offset, total = ac75().sharded_scan(scan_fn, len(text), o.lmax, rank, world, all_gather_counts)
    np.save(os.path.join(tmpdir, f"recs{rank}.npy"), recs)
    np.save(os.path.join(tmpdir, f"meta{rank}.npy"), np.array([offset, total]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_equals_single_scan(tmp_path, world):
    port = 29500 + world + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    flat, offsets = random_patterns(300, lmin=2, lmax=12, seed=7, alphabet=4)
    text = np.random.default_rng(3).integers(0, 4, size=50_000).astype(np.uint8)
    o = pyoracle.Oracle("port", 1)
    o.insert_many(flat=flat, offsets=offsets)
    want = o.scan(text, cap=1 << 22)
    parts = [np.load(os.path.join(tmp_path, f"recs{r}.npy")) for r in range(world)]
    metas = [np.load(os.path.join(tmp_path, f"meta{r}.npy")) for r in range(world)]
    got = np.concatenate(parts)
    assert np.array_equal(got, want)
    assert all(int(m[1]) == len(want) for m in metas)
    assert [int(m[0]) for m in metas] == list(np.cumsum([0] + [len(p) for p in parts[:-1]]))


def test_plan_shards_tiles_the_text():
    for n, world, lmax in ((1000, 3, 17), (1 << 20, 8, 32), (5, 8, 4), (0, 2, 3)):
        plan = ac75().plan_shards(n, world, lmax)
        assert plan[0][0] == 0 and plan[-1][1] == n
        for (a, b, lead), (a2, _, _) in zip(plan, plan[1:]):
            assert b == a2 and lead <= a and (lead >= min(a, lmax - 1))
