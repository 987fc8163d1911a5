"""Randomized differential test: GPU batch scan == oracle, over random dictionaries / texts / engines / leads / carried cursors
with insertions between scans / streaming chunk sizes.  Seeds are fixed; every case prints enough to be replayed."""
import numpy as np
import pytest

from helpers import ac75, best_oracle_kind
from oracle import pyoracle

pytestmark = pytest.mark.gpu
DT = {1: np.uint8, 2: np.uint16, 4: np.uint32}


STRIDE2_SCANS = [0]  # scans of the stride-2 fuzz that really ran the stride-2 kernel (a dense text falls back to the dense mode)


def one_case(seed, stride2=False):
    rng = np.random.default_rng(seed)
    width = 1 if stride2 else int(rng.choice([1, 1, 1, 2, 4]))
    alpha_size = int(rng.choice([64, 200, 256])) if stride2 else int(rng.choice([2, 3, 4, 16, 200]))
    top = {1: 256, 2: 65536, 4: 2**32}[width]
    alphabet = rng.choice(top, size=min(alpha_size, top), replace=False) if top <= 65536 else rng.integers(0, top, size=alpha_size)
    alphabet = np.unique(np.asarray(alphabet)).astype(DT[width])
    lmin = int(rng.choice([4, 4, 5, 6])) if stride2 else int(rng.choice([1, 1, 2, 3, 4, 5]))
    lmax = lmin + int(rng.integers(0, 40))
    engines = ["auto", "filter"] + (["dfa_smem", "dfa_global"] if width == 1 else [])
    engine = "filter" if stride2 else str(rng.choice(engines))
    o = pyoracle.Oracle("port", width)
    m = ac75().Machine(width)
    m.set_option("engine", engine)
    if rng.random() < 0.3:
        m.set_option("stream_bytes", int(rng.choice([4096, 16384, 65536])))
    pos = 0
    desc = dict(seed=seed, width=width, alpha=len(alphabet), lmin=lmin, lmax=lmax, engine=engine)
    for rnd in range(int(rng.integers(1, 5))):
        nk = int(rng.integers(1, 400))
        kws = [rng.choice(alphabet, size=int(rng.integers(lmin, lmax + 1))).astype(DT[width]) for _ in range(nk)]
        if rng.random() < 0.5:  # keywords taken from a common stem: deep shared suffixes / prefixes
            stem = rng.choice(alphabet, size=lmax).astype(DT[width])
            kws += [stem[int(rng.integers(0, lmax - lmin + 1)):][:int(rng.integers(lmin, lmax + 1))] for _ in range(20)]
            kws = [k for k in kws if len(k) >= 1]
        assert np.array_equal(o.insert_many(kws), m.insert_many(kws)), desc
        n = int(rng.choice([0, 1, 7, 100, 5000, 70000])) if not stride2 else int(rng.choice([3, 4, 2047, 2049, 5000, 70001, 300000]))
        text = rng.choice(alphabet, size=n).astype(DT[width])
        for _ in range(int(rng.integers(0, 30))):  # plant keywords
            k = kws[int(rng.integers(0, len(kws)))]
            if len(k) <= n:
                at = int(rng.integers(0, n - len(k) + 1))
                text[at:at + len(k)] = k
        carry = bool(rng.random() < 0.6)
        if not carry:
            o.reset_cursor()
            m.reset_cursor()
        lead = int(rng.integers(0, n + 1)) if (not carry and rng.random() < 0.3) else 0
        want = o.scan(text, base=pos, cap=1 << 24)
        want = want[want["end"] >= pos + lead]
        got = m.scan(text, base=pos, lead=lead, carry=True, capacity=1 << 24)
        assert np.array_equal(got, want), (desc, dict(round=rnd, n=n, carry=carry, lead=lead, got=len(got), want=len(want), engine_used=m.stats()["engine"]))
        if stride2 and n and m.stats()["filter_stride"] == 2:
            STRIDE2_SCANS[0] += 1
        pos += n
    m.close(), o.close()


@pytest.mark.parametrize("block", range(16))
def test_fuzz_block(block):
    for seed in range(block * 40, block * 40 + 40):
        one_case(1000 + seed)


@pytest.mark.parametrize("block", range(4))
def test_fuzz_stride2_kernel(block):
    """The same differential test pinned to the conditions of the stride-2 filter kernel (byte alphabet, shortest keyword >= 4, large
    alphabet): text lengths around the 2 KiB tile size, carried cursors whose prefix reaches into the text, leads, streaming."""
    STRIDE2_SCANS[0] = 0
    for seed in range(block * 30, block * 30 + 30):
        one_case(5000 + seed, stride2=True)
    assert STRIDE2_SCANS[0] >= 20, STRIDE2_SCANS
