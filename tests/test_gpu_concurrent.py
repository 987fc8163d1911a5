"""Several host threads scanning ONE machine at the same time (the reference's scan is lock-free with a caller-owned cursor,
README.md:266,364): every thread gets a scan context of its own, the machine lock is not held while the GPU works, and each
thread's records equal the sequential result."""
import threading

import numpy as np
import pytest

from helpers import ac75, generate_text, random_patterns

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine", ["auto", "dfa_global"])
def test_threads_share_one_machine(engine):
    flat, offsets = random_patterns(5000, seed=11)
    m = ac75().Machine(1)
    m.insert_many(flat=flat, offsets=offsets)
    if engine != "auto":
        m.set_option("engine", engine)
    texts = [generate_text((3 << 20) + 4096 * i, first=i << 24, kind=0, plant_period=1024, dict_flat=flat, dict_offsets=offsets) for i in range(6)]
    want = [m.scan(t, base=i << 24) for i, t in enumerate(texts)]
    assert all(len(w) > 1000 for w in want)
    got, errors = [None] * len(texts), []

    def work(i):
        try:
            for _ in range(3):
                got[i] = m.scan(texts[i], base=i << 24)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(texts))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    m.close()


def test_insertion_between_concurrent_scans():
    """A scan that starts after an insertion sees the new keyword; scans in flight keep the tables they started with."""
    flat, offsets = random_patterns(3000, seed=12)
    m = ac75().Machine(1)
    m.insert_many(flat=flat, offsets=offsets)
    text = generate_text(4 << 20, kind=0, plant_period=2048, dict_flat=flat, dict_offsets=offsets)
    before = m.scan(text)
    results = []

    def work():
        for _ in range(4):
            results.append(len(m.scan(text)))

    t = threading.Thread(target=work)
    t.start()
    new = bytes(text[5000:5009])
    m.insert_many([new])
    t.join()
    after = m.scan(text)
    assert len(after) >= len(before) + 1 and set(results) <= {len(before), len(after)}
    m.close()
